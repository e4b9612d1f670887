"""TEST INFRASTRUCTURE ONLY.  ctypes loader for oracle/liboracle.so (rangelib_oracle.c, our CPU
restatement of the reference's hot path).

May be imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BL, RM, CDDT, PCDDT, GLT = 0, 1, 2, 3, 4

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int)
_u64p = C.POINTER(C.c_uint64)

_lib = None


def lib_path():
    return os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "rangelib_oracle.c")
    if force or not os.path.exists(lib_path()) or os.path.getmtime(lib_path()) < os.path.getmtime(src):
        subprocess.check_call(["make", "-f", os.path.join(_HERE, "Makefile"), lib_path()])


def _load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(lib_path())
    L.orc_sinf.restype = C.c_float
    L.orc_sinf.argtypes = [C.c_float]
    L.orc_cosf.restype = C.c_float
    L.orc_cosf.argtypes = [C.c_float]
    L.orc_trig_compare.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, _u64p]
    L.orc_edt.argtypes = [_u8p, C.c_int, C.c_int, _f32p]
    L.orc_edge_map.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int, _u8p, C.c_int, C.c_int, C.c_float, C.c_uint]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_world.argtypes = [C.c_void_p] + [C.c_float] * 6
    L.orc_prune.argtypes = [C.c_void_p, C.c_float]
    L.orc_prune_unassigned.restype = C.c_int64
    L.orc_prune_unassigned.argtypes = [C.c_void_p]
    L.orc_calc_range.restype = C.c_float
    L.orc_calc_range.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    L.orc_rm_step_counts.argtypes = [C.c_void_p, _f32p, _i32p, C.c_int]
    L.orc_glt.restype = C.POINTER(C.c_uint16)
    L.orc_glt.argtypes = [C.c_void_p]
    L.orc_dt.restype = _f32p
    L.orc_dt.argtypes = [C.c_void_p]
    for name, rt in [("orc_cddt_nbins", C.c_int64), ("orc_cddt_nvalues", C.c_int64), ("orc_cddt_widths", _i32p),
                     ("orc_cddt_trans", _f32p), ("orc_cddt_offsets", _i64p), ("orc_cddt_values", _f32p)]:
        getattr(L, name).restype = rt
        getattr(L, name).argtypes = [C.c_void_p]
    L.orc_calc_range_many.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int]
    L.orc_numpy_calc_range.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int]
    L.orc_numpy_calc_range_angles.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
    L.orc_set_sensor_model.argtypes = [C.c_void_p, _f64p, C.c_int]
    L.orc_calc_range_many_radial_optimized.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int, C.c_float, C.c_float]
    L.orc_eval_sensor_model.argtypes = [C.c_void_p, _f32p, _f32p, _f64p, C.c_int, C.c_int, C.c_int]
    L.orc_calc_range_repeat_angles_eval_sensor_model.argtypes = [
        C.c_void_p, _f32p, _f32p, _f32p, _f64p, C.c_int, C.c_int, C.c_int]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def sinf(x):
    return _load().orc_sinf(float(x))


def cosf(x):
    return _load().orc_cosf(float(x))


def trig_compare(lo, hi, step):
    out = np.zeros(2, np.uint64)
    _load().orc_trig_compare(lo, hi, step, _p(out, _u64p))
    return int(out[0]), int(out[1])


def edt(occ):
    occ = np.ascontiguousarray(occ, dtype=np.uint8)
    out = np.empty(occ.shape, np.float32)
    _load().orc_edt(_p(occ, _u8p), occ.shape[0], occ.shape[1], _p(out, _f32p))
    return out


def edge_map(occ):
    occ = np.ascontiguousarray(occ, dtype=np.uint8)
    out = np.empty(occ.shape, np.uint8)
    _load().orc_edge_map(_p(occ, _u8p), occ.shape[0], occ.shape[1], _p(out, _u8p))
    return out


class Oracle:
    """occ: uint8 [W, H] x-major (occ[x, y]), the layout of the reference's OMap::grid[x][y]."""

    def __init__(self, kind, occ, max_range, theta_disc=108, threads=1):
        self.L = _load()
        occ = np.ascontiguousarray(occ, dtype=np.uint8)
        self.width, self.height = occ.shape
        self.kind = kind
        self.td = theta_disc
        self.threads = threads
        self.h = self.L.orc_create(kind, _p(occ, _u8p), occ.shape[0], occ.shape[1], float(max_range), theta_disc)

    def set_world(self, scale=1.0, angle=0.0, ox=0.0, oy=0.0, sin_a=0.0, cos_a=1.0):
        self.L.orc_set_world(self.h, scale, angle, ox, oy, sin_a, cos_a)

    def prune(self, max_range):
        self.L.orc_prune(self.h, float(max_range))

    def prune_unassigned(self):
        return self.L.orc_prune_unassigned(self.h)

    def calc_range(self, x, y, th):
        return self.L.orc_calc_range(self.h, x, y, th)

    def calc_range_many(self, ins):
        ins = _f32(ins)
        outs = np.empty(ins.shape[0], np.float32)
        self.L.orc_calc_range_many(self.h, _p(ins, _f32p), _p(outs, _f32p), ins.shape[0], self.threads)
        return outs

    def numpy_calc_range(self, ins):
        ins = _f32(ins)
        outs = np.empty(ins.shape[0], np.float32)
        self.L.orc_numpy_calc_range(self.h, _p(ins, _f32p), _p(outs, _f32p), ins.shape[0], self.threads)
        return outs

    def numpy_calc_range_angles(self, ins, angles):
        ins, angles = _f32(ins), _f32(angles)
        outs = np.empty(ins.shape[0] * angles.shape[0], np.float32)
        self.L.orc_numpy_calc_range_angles(self.h, _p(ins, _f32p), _p(angles, _f32p), _p(outs, _f32p),
                                           ins.shape[0], angles.shape[0], self.threads)
        return outs

    def set_sensor_model(self, table):
        table = np.ascontiguousarray(table, dtype=np.float64)
        assert table.shape[0] == table.shape[1]
        self.L.orc_set_sensor_model(self.h, _p(table, _f64p), table.shape[0])

    def eval_sensor_model(self, obs, ranges, num_rays, num_particles):
        obs, ranges = _f32(obs), _f32(ranges)
        outs = np.empty(num_particles, np.float64)
        self.L.orc_eval_sensor_model(self.h, _p(obs, _f32p), _p(ranges, _f32p), _p(outs, _f64p), num_rays,
                                     num_particles, self.threads)
        return outs

    def calc_range_repeat_angles_eval_sensor_model(self, ins, angles, obs):
        ins, angles, obs = _f32(ins), _f32(angles), _f32(obs)
        w = np.empty(ins.shape[0], np.float64)
        self.L.orc_calc_range_repeat_angles_eval_sensor_model(
            self.h, _p(ins, _f32p), _p(angles, _f32p), _p(obs, _f32p), _p(w, _f64p), ins.shape[0],
            angles.shape[0], self.threads)
        return w

    def calc_range_many_radial_optimized(self, num_rays, min_angle, max_angle, ins, outs):
        """RangeLib.h:616-676; outs f32[N*num_rays] updated in place."""
        ins = _f32(ins)
        assert outs.dtype == np.float32 and outs.flags.c_contiguous and outs.size >= ins.shape[0] * num_rays
        self.L.orc_calc_range_many_radial_optimized(self.h, _p(ins, _f32p), _p(outs, _f32p), ins.shape[0], num_rays,
                                                    min_angle, max_angle)
        return outs

    def rm_step_counts(self, ins_grid):
        ins = _f32(ins_grid)
        out = np.zeros(ins.shape[0], np.int32)
        self.L.orc_rm_step_counts(self.h, _p(ins, _f32p), _p(out, _i32p), ins.shape[0])
        return out

    def dt(self):
        ptr = self.L.orc_dt(self.h)
        return np.ctypeslib.as_array(ptr, shape=(self.width, self.height)).copy()

    def glt_table(self):
        ptr = self.L.orc_glt(self.h)
        return np.ctypeslib.as_array(ptr, shape=(self.width, self.height, self.td)).copy()

    def cddt_table(self):
        nb = self.L.orc_cddt_nbins(self.h)
        nv = self.L.orc_cddt_nvalues(self.h)
        widths = np.ctypeslib.as_array(self.L.orc_cddt_widths(self.h), shape=(self.td,)).copy()
        trans = np.ctypeslib.as_array(self.L.orc_cddt_trans(self.h), shape=(self.td,)).copy()
        offsets = np.ctypeslib.as_array(self.L.orc_cddt_offsets(self.h), shape=(nb + 1,)).copy()
        values = np.ctypeslib.as_array(self.L.orc_cddt_values(self.h), shape=(max(nv, 1),)).copy()[:nv]
        return widths, trans, offsets, values

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None
