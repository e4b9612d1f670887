"""TEST INFRASTRUCTURE ONLY.  ctypes loader for oracle/_ref/libref_{strict,shipped}.so, the
UNMODIFIED reference compiled behind oracle/ref_shim.cpp.

May be imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package never imports this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BL, RM, CDDT, PCDDT, GLT = 0, 1, 2, 3, 4
RMGPU = 5  # ranges::RayMarchingGPU; only in flavor "cuda" (libref_cuda.so: the reference's kernels.cu for sm_100a)

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int)


def lib_path(flavor="strict"):
    return os.path.join(_HERE, "_ref", "libref_%s.so" % flavor)


def available(flavor="strict"):
    return os.path.exists(lib_path(flavor))


_libs = {}


def _load(flavor):
    if flavor in _libs:
        return _libs[flavor]
    L = C.CDLL(lib_path(flavor))
    L.ref_map_create.restype = C.c_void_p
    L.ref_map_create.argtypes = [_u8p, C.c_int, C.c_int]
    L.ref_map_load_png.restype = C.c_void_p
    L.ref_map_load_png.argtypes = [C.c_char_p, C.c_float]
    L.ref_map_width.argtypes = [C.c_void_p]
    L.ref_map_height.argtypes = [C.c_void_p]
    L.ref_map_get.argtypes = [C.c_void_p, _u8p]
    L.ref_map_edge.argtypes = [C.c_void_p, _u8p]
    L.ref_map_set_world.argtypes = [C.c_void_p] + [C.c_float] * 6
    L.ref_map_destroy.argtypes = [C.c_void_p]
    L.ref_method_create.restype = C.c_void_p
    L.ref_method_create.argtypes = [C.c_int, C.c_void_p, C.c_float, C.c_uint]
    L.ref_method_destroy.argtypes = [C.c_void_p]
    L.ref_prune.argtypes = [C.c_void_p, C.c_float]
    L.ref_calc_range.restype = C.c_float
    L.ref_calc_range.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    L.ref_calc_range_many.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int]
    L.ref_numpy_calc_range.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int]
    L.ref_numpy_calc_range_angles.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
    L.ref_set_sensor_model.argtypes = [C.c_void_p, _f64p, C.c_int]
    L.ref_eval_sensor_model.argtypes = [C.c_void_p, _f32p, _f32p, _f64p, C.c_int, C.c_int, C.c_int]
    L.ref_calc_range_repeat_angles_eval_sensor_model.argtypes = [
        C.c_void_p, _f32p, _f32p, _f32p, _f64p, C.c_int, C.c_int, C.c_int]
    L.ref_calc_range_many_radial_optimized.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int, C.c_float, C.c_float]
    L.ref_calc_range_pair.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, _f32p, _f32p]
    L.ref_get_dt.argtypes = [C.c_void_p, _f32p]
    L.ref_cddt_dims.restype = C.c_int64
    L.ref_cddt_dims.argtypes = [C.c_void_p, _i64p, _i32p, _f32p]
    L.ref_cddt_dump.argtypes = [C.c_void_p, _i64p, _f32p]
    L.ref_glt_dump.argtypes = [C.c_void_p, C.POINTER(C.c_uint16), C.c_int]
    _libs[flavor] = L
    return L


def _p(a, t):
    return a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RefMap:
    """occ: uint8/bool array [W, H] x-major (occ[x, y]); or a PNG path through the reference's loader."""

    def __init__(self, occ=None, png=None, threshold=128.0, flavor="strict"):
        self.L = _load(flavor)
        self.flavor = flavor
        if png is not None:
            self.h = self.L.ref_map_load_png(os.fsencode(png), float(threshold))
            if not self.h:
                raise IOError("reference failed to load %s" % png)
        else:
            occ = np.ascontiguousarray(occ, dtype=np.uint8)
            self.h = self.L.ref_map_create(_p(occ, _u8p), occ.shape[0], occ.shape[1])
        self.width = self.L.ref_map_width(self.h)
        self.height = self.L.ref_map_height(self.h)

    def occ(self):
        out = np.zeros((self.width, self.height), np.uint8)
        self.L.ref_map_get(self.h, _p(out, _u8p))
        return out

    def edge(self):
        out = np.zeros((self.width, self.height), np.uint8)
        self.L.ref_map_edge(self.h, _p(out, _u8p))
        return out

    def set_world(self, scale=1.0, angle=0.0, ox=0.0, oy=0.0, sin_a=0.0, cos_a=1.0):
        self.L.ref_map_set_world(self.h, scale, angle, ox, oy, sin_a, cos_a)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_map_destroy(self.h)
            self.h = None


class RefMethod:
    def __init__(self, kind, rmap, max_range, theta_disc=108, threads=1):
        self.L = rmap.L
        self.map = rmap
        self.kind = kind
        self.threads = threads
        self.h = self.L.ref_method_create(kind, rmap.h, float(max_range), int(theta_disc))
        if not self.h:
            raise ValueError("bad kind")

    def prune(self, max_range):
        self.L.ref_prune(self.h, float(max_range))

    def calc_range(self, x, y, th):
        return self.L.ref_calc_range(self.h, x, y, th)

    def calc_range_many(self, ins):
        ins = _f32(ins)
        outs = np.empty(ins.shape[0], np.float32)
        self.L.ref_calc_range_many(self.h, _p(ins, _f32p), _p(outs, _f32p), ins.shape[0], self.threads)
        return outs

    def numpy_calc_range(self, ins):
        ins = _f32(ins)
        outs = np.empty(ins.shape[0], np.float32)
        self.L.ref_numpy_calc_range(self.h, _p(ins, _f32p), _p(outs, _f32p), ins.shape[0], self.threads)
        return outs

    def numpy_calc_range_angles(self, ins, angles):
        ins, angles = _f32(ins), _f32(angles)
        outs = np.empty(ins.shape[0] * angles.shape[0], np.float32)
        self.L.ref_numpy_calc_range_angles(self.h, _p(ins, _f32p), _p(angles, _f32p), _p(outs, _f32p),
                                           ins.shape[0], angles.shape[0], self.threads)
        return outs

    def set_sensor_model(self, table):
        table = np.ascontiguousarray(table, dtype=np.float64)
        assert table.shape[0] == table.shape[1]
        self.L.ref_set_sensor_model(self.h, _p(table, _f64p), table.shape[0])

    def eval_sensor_model(self, obs, ranges, num_rays, num_particles):
        obs, ranges = _f32(obs), _f32(ranges)
        outs = np.empty(num_particles, np.float64)
        self.L.ref_eval_sensor_model(self.h, _p(obs, _f32p), _p(ranges, _f32p), _p(outs, _f64p), num_rays,
                                     num_particles, self.threads)
        return outs

    def calc_range_repeat_angles_eval_sensor_model(self, ins, angles, obs):
        ins, angles, obs = _f32(ins), _f32(angles), _f32(obs)
        w = np.empty(ins.shape[0], np.float64)
        self.L.ref_calc_range_repeat_angles_eval_sensor_model(
            self.h, _p(ins, _f32p), _p(angles, _f32p), _p(obs, _f32p), _p(w, _f64p), ins.shape[0],
            angles.shape[0], self.threads)
        return w

    def calc_range_many_radial_optimized(self, num_rays, min_angle, max_angle, ins, outs):
        """outs f32[N*num_rays] is updated in place (the reference leaves some beams unwritten)."""
        ins = _f32(ins)
        assert outs.dtype == np.float32 and outs.flags.c_contiguous and outs.size >= ins.shape[0] * num_rays
        self.L.ref_calc_range_many_radial_optimized(self.h, _p(ins, _f32p), _p(outs, _f32p), ins.shape[0], num_rays,
                                                    min_angle, max_angle)
        return outs

    def calc_range_pair(self, x, y, th):
        r, ri = C.c_float(), C.c_float()
        self.L.ref_calc_range_pair(self.h, x, y, th, C.byref(r), C.byref(ri))
        return r.value, ri.value

    def dt(self):
        out = np.empty((self.map.width, self.map.height), np.float32)
        if self.L.ref_get_dt(self.h, _p(out, _f32p)) != 0:
            raise ValueError("not an RM method")
        return out

    def cddt_table(self, theta_disc):
        nv = C.c_int64(0)
        widths = np.zeros(theta_disc, np.int32)
        trans = np.zeros(theta_disc, np.float32)
        nb = self.L.ref_cddt_dims(self.h, C.byref(nv), _p(widths, _i32p), _p(trans, _f32p))
        offsets = np.zeros(nb + 1, np.int64)
        values = np.zeros(max(nv.value, 1), np.float32)
        self.L.ref_cddt_dump(self.h, _p(offsets, _i64p), _p(values, _f32p))
        return widths, trans, offsets, values[: nv.value]

    def glt_table(self, theta_disc):
        out = np.zeros((self.map.width, self.map.height, theta_disc), np.uint16)
        if self.L.ref_glt_dump(self.h, out.ctypes.data_as(C.POINTER(C.c_uint16)), theta_disc) != 0:
            raise ValueError("not a GLT method")
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_method_destroy(self.h)
            self.h = None
