"""Python mirror of the reference's ``range_libc`` classes on top of the C ABI (cabi.py)."""
import ctypes as C

import numpy as np

from . import cabi
from .cabi import check, lib


def kernel_launches():
    return int(lib().rl_stat_kernel_launches())


def _is_torch(a):
    return hasattr(a, "data_ptr") and hasattr(a, "is_cuda")


def _buf(a, dtype, ndim, name):
    """Pointer + shape of a C-contiguous array of the exact dtype/ndim, numpy or torch.  Mirrors the
    typed-memoryview checks of the reference's Cython signatures (ValueError on mismatch)."""
    if _is_torch(a):
        import torch
        want = {np.float32: torch.float32, np.float64: torch.float64, np.uint8: torch.uint8, np.int8: torch.int8}[dtype]
        if a.dtype != want:
            raise ValueError("%s: expected dtype %s, got %s" % (name, want, a.dtype))
        if a.dim() != ndim:
            raise ValueError("%s: expected %d dimensions, got %d" % (name, ndim, a.dim()))
        if not a.is_contiguous():
            raise ValueError("%s: tensor is not contiguous" % name)
        return C.c_void_p(a.data_ptr()), tuple(a.shape)
    if not isinstance(a, np.ndarray):
        raise ValueError("%s: expected a numpy array or torch tensor" % name)
    if a.dtype != dtype:
        raise ValueError("%s: Buffer dtype mismatch, expected %s but got %s" % (name, np.dtype(dtype), a.dtype))
    if a.ndim != ndim:
        raise ValueError("%s: Buffer has wrong number of dimensions (expected %d, got %d)" % (name, ndim, a.ndim))
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("%s: ndarray is not C-contiguous" % name)
    return C.c_void_p(a.ctypes.data), a.shape


class PyOMap:
    """PyOMap(np.ndarray bool[H, W]) | PyOMap(width, height) | PyOMap(OccupancyGrid-like msg)
    (RangeLibc.pyx:130-198).  PNG paths are not decoded here: PNG ingest is outside the hot path
    (SURVEY.md section 8f); load the image with any decoder and pass the boolean array.

    Grid convention of the reference: grid[x][y] with x the image column and y the image row, so a
    numpy image ``arr[row, col]`` maps to grid[col][row]."""

    def __init__(self, arg1, arg2=None):
        self._h = C.c_void_p()
        self._L = lib()  # kept so that __del__ works during interpreter shutdown
        world = None
        if isinstance(arg1, (int, np.integer)) and isinstance(arg2, (int, np.integer)):
            occ = np.zeros((int(arg1), int(arg2)), np.uint8)
        elif isinstance(arg1, np.ndarray):
            if arg1.ndim != 2:
                raise ValueError("occupancy array must be 2-D")
            occ = np.ascontiguousarray(arg1.T != 0, dtype=np.uint8)  # [W, H] x-major
        elif hasattr(arg1, "info") and hasattr(arg1, "data"):
            # ROS nav_msgs/OccupancyGrid: 0 permissible, -1 unmapped, 100 blocked (pyx:146-167).
            # The reference builds OMap(height, width) and fills grid[x][y] for x in range(height),
            # y in range(width) from data.reshape(height, width)[x, y] > 10.
            info = arg1.info
            arr = np.asarray(arg1.data).reshape((info.height, info.width))
            occ = np.ascontiguousarray(arr > 10, dtype=np.uint8)  # grid[x=row][y=col]
            q = info.origin.orientation
            yaw = np.arctan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z))
            angle = -1.0 * yaw
            world = (info.resolution, angle, info.origin.position.x, info.origin.position.y, np.sin(angle),
                     np.cos(angle))
        else:
            raise ValueError("Failed to construct PyOMap, check argument types "
                             "(PNG decoding is not part of this backend; pass a boolean array).")
        self._w, self._hgt = occ.shape
        check(lib().rl_map_create(C.c_void_p(occ.ctypes.data), occ.shape[0], occ.shape[1], C.byref(self._h)))
        if world is not None:
            self.set_world(*world)

    def set_world(self, scale=1.0, angle=0.0, origin_x=0.0, origin_y=0.0, sin_angle=0.0, cos_angle=1.0):
        check(lib().rl_map_set_world(self._h, scale, angle, origin_x, origin_y, sin_angle, cos_angle))

    def isOccupied(self, x, y):
        return bool(lib().rl_map_is_occupied(self._h, int(x), int(y)))

    def error(self):
        return False

    def width(self):
        return self._w

    def height(self):
        return self._hgt

    def grid(self):
        out = np.empty((self._w, self._hgt), np.uint8)
        check(lib().rl_map_get(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def save(self, fn):
        """OMap::save (RangeLib.h:264-291): RGBA PNG, occupied cells black, free cells white.  Returns False on
        success, like the reference (whose return value is lodepng's error code)."""
        from .mapio import save_png
        return save_png(fn, self.grid())

    def update(self, patch_xmajor, x0, y0):
        """Dynamic maps: overwrite a [w, h] x-major patch of the HOST map (see RangeMethod.update_map)."""
        p = np.ascontiguousarray(patch_xmajor, dtype=np.uint8)
        check(lib().rl_map_update(self._h, C.c_void_p(p.ctypes.data), int(x0), int(y0), p.shape[0], p.shape[1]))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value and getattr(self, "_L", None) is not None:
            self._L.rl_map_destroy(self._h)
            self._h.value = None


class _RangeMethod:
    _kind = None

    def __init__(self, Map, max_range, theta_disc=0, device=-1):
        self._h = C.c_void_p()
        self._L = lib()
        self.max_range = float(max_range)
        self._map = Map
        check(lib().rl_method_create(self._kind, Map._h, float(max_range), int(theta_disc), int(device),
                                     C.byref(self._h)))

    # -- reference API ---------------------------------------------------------------------
    def calc_range(self, x, y, heading):
        out = C.c_float()
        check(lib().rl_calc_range(self._h, float(x), float(y), float(heading), C.byref(out)))
        return out.value

    def calc_range_many(self, ins, outs):
        """ins f32[N,3] WORLD poses, outs f32[N]; N is taken from outs (RangeLibc.pyx:208-209)."""
        pi, si = _buf(ins, np.float32, 2, "ins")
        po, so = _buf(outs, np.float32, 1, "outs")
        if si[1] != 3 or si[0] < so[0]:
            raise ValueError("ins must be [N,3] with N >= len(outs)")
        check(lib().rl_numpy_calc_range(self._h, pi, po, so[0]))

    def calc_range_many_grid(self, ins, outs):
        """RayMarchingGPU::calc_range_many (RangeLib.h:819-831): GRID coordinates, no conversion."""
        pi, si = _buf(ins, np.float32, 2, "ins")
        po, so = _buf(outs, np.float32, 1, "outs")
        if si[1] != 3 or si[0] < so[0]:
            raise ValueError("ins must be [N,3] with N >= len(outs)")
        check(lib().rl_calc_range_many(self._h, pi, po, so[0]))

    def calc_range_repeat_angles(self, ins, angles, outs):
        pi, si = _buf(ins, np.float32, 2, "ins")
        pa, sa = _buf(angles, np.float32, 1, "angles")
        po, so = _buf(outs, np.float32, 1, "outs")
        if si[1] != 3 or so[0] < si[0] * sa[0]:
            raise ValueError("outs must hold N*M floats")
        check(lib().rl_numpy_calc_range_angles(self._h, pi, pa, po, si[0], sa[0]))

    def calc_range_repeat_angles_eval_sensor_model(self, ins, angles, obs, weights):
        pi, si = _buf(ins, np.float32, 2, "ins")
        pa, sa = _buf(angles, np.float32, 1, "angles")
        pb, sb = _buf(obs, np.float32, 1, "obs")
        pw, sw = _buf(weights, np.float64, 1, "weights")
        if si[1] != 3 or sb[0] < sa[0] or sw[0] < si[0]:
            raise ValueError("shape mismatch")
        check(lib().rl_calc_range_repeat_angles_eval_sensor_model(self._h, pi, pa, pb, pw, si[0], sa[0]))

    # -- particle-filter steps either side of the sensor update (not in the reference; rl_pf.cu) ----------
    def normalize_weights(self, weights, inv_squash=1.0, want_sum=False):
        """weights f64[N] in place: w <- pow(w, inv_squash) / sum.  Returns the sum of the squashed weights when
        want_sum (blocking)."""
        pw, sw = _buf(weights, np.float64, 1, "weights")
        s = C.c_double()
        check(lib().rl_pf_normalize_weights(self._h, pw, sw[0], float(inv_squash), C.byref(s) if want_sum else None))
        return s.value if want_sum else None

    def resample(self, particles, weights, out_particles, u0):
        """Systematic resampling of particles f32[N,3] by (normalised) weights f64[N] into out_particles, one uniform
        draw u0 in [0, 1)."""
        pp, sp = _buf(particles, np.float32, 2, "particles")
        pw, sw = _buf(weights, np.float64, 1, "weights")
        po, so = _buf(out_particles, np.float32, 2, "out_particles")
        if sp[1] != 3 or so[1] != 3 or sw[0] < sp[0] or so[0] < sp[0]:
            raise ValueError("shape mismatch")
        check(lib().rl_pf_resample(self._h, pp, pw, po, sp[0], float(u0)))

    def motion_update(self, particles, dx, dy, dtheta, noise=None):
        """Odometry step of particles f32[N,3] in place; noise f32[N,3] or None."""
        pp, sp = _buf(particles, np.float32, 2, "particles")
        pn = None
        if noise is not None:
            pn, sn = _buf(noise, np.float32, 2, "noise")
            if sn[0] < sp[0] or sn[1] != 3:
                raise ValueError("noise must be [N,3]")
        if sp[1] != 3:
            raise ValueError("particles must be [N,3]")
        check(lib().rl_pf_motion_update(self._h, pp, sp[0], float(dx), float(dy), float(dtheta), pn))

    def calc_range_many_radial_optimized(self, num_rays, min_angle, max_angle, ins, outs):
        """RangeLibc.pyx:274-276 (argument order as there).  outs f32[N*num_rays], updated in place: beams the
        reference does not write keep their previous content."""
        pi, si = _buf(ins, np.float32, 2, "ins")
        po, so = _buf(outs, np.float32, 1, "outs")
        if si[1] != 3 or so[0] < si[0] * int(num_rays):
            raise ValueError("shape mismatch")
        check(lib().rl_calc_range_many_radial_optimized(self._h, pi, po, si[0], int(num_rays), float(min_angle),
                                                        float(max_angle)))

    def calc_range_repeat_angles_eval_sensor_model_peers(self, ins, angles, obs, peer_ptrs, offset):
        """Multi-GPU fused update: weights of the local particles `ins` are stored by the kernel into every
        rank's gathered array (peer_ptrs: one peer-mapped device pointer per rank) at `offset`."""
        pi, si = _buf(ins, np.float32, 2, "ins")
        pa, sa = _buf(angles, np.float32, 1, "angles")
        pb, sb = _buf(obs, np.float32, 1, "obs")
        if si[1] != 3 or sb[0] < sa[0]:
            raise ValueError("shape mismatch")
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        check(lib().rl_calc_range_repeat_angles_eval_sensor_model_peers(self._h, pi, pa, pb, arr, len(peer_ptrs),
                                                                        int(offset), si[0], sa[0]))

    def peers_init(self, weights0_ptrs, weights1_ptrs, flags_ptrs, rank):
        """Signalled multi-GPU mode: peer-mapped pointers of every rank's two weight buffers and flag array."""
        n = len(weights0_ptrs)
        mk = lambda ps: (C.c_void_p * n)(*[int(p) for p in ps])  # noqa: E731
        check(lib().rl_method_peers_init(self._h, mk(weights0_ptrs), mk(weights1_ptrs), mk(flags_ptrs), n, int(rank)))

    def calc_range_repeat_angles_eval_sensor_model_signalled(self, ins, angles, obs, offset):
        """One kernel per rank: compute + peer-store all-gather + completion flags.  Returns the buffer index."""
        pi, si = _buf(ins, np.float32, 2, "ins")
        pa, sa = _buf(angles, np.float32, 1, "angles")
        pb, sb = _buf(obs, np.float32, 1, "obs")
        if si[1] != 3 or sb[0] < sa[0]:
            raise ValueError("shape mismatch")
        buf = C.c_int()
        check(lib().rl_calc_range_repeat_angles_eval_sensor_model_signalled(self._h, pi, pa, pb, int(offset), si[0], sa[0],
                                                                            C.byref(buf)))
        return buf.value

    def peers_wait(self):
        check(lib().rl_method_peers_wait(self._h))

    def calc_range_repeat_angles_eval_sensor_model_sharded(self, ins, angles, obs, weights_all, offset):
        """The sharded update as one call: `ins` are this rank's particles (the slice of the cloud starting at
        particle `offset`), `weights_all` (f64[n_total]) receives the weights of ALL particles.  Host arrays:
        blocking; device tensors: asynchronous on the handle's stream.  Needs peers_init."""
        pi, si = _buf(ins, np.float32, 2, "ins")
        pa, sa = _buf(angles, np.float32, 1, "angles")
        pb, sb = _buf(obs, np.float32, 1, "obs")
        pw, sw = _buf(weights_all, np.float64, 1, "weights_all")
        if si[1] != 3 or sb[0] < sa[0] or sw[0] < int(offset) + si[0]:
            raise ValueError("shape mismatch")
        check(lib().rl_calc_range_repeat_angles_eval_sensor_model_sharded(self._h, pi, pa, pb, pw, int(offset), si[0],
                                                                          sa[0], sw[0]))

    def eval_sensor_model(self, observation, ranges, outs, num_rays, num_particles):
        pb, sb = _buf(observation, np.float32, 1, "observation")
        pr, sr = _buf(ranges, np.float32, 1, "ranges")
        po, so = _buf(outs, np.float64, 1, "outs")
        if sb[0] < num_rays or sr[0] < num_rays * num_particles or so[0] < num_particles:
            raise ValueError("shape mismatch")
        check(lib().rl_eval_sensor_model(self._h, pb, pr, po, int(num_rays), int(num_particles)))

    def set_sensor_model(self, table):
        pt, st = _buf(table, np.float64, 2, "table")
        if st[0] != st[1]:
            print("Sensor model must have equal matrix dimensions, failing!")  # RangeLibc.pyx:222-224
            return
        check(lib().rl_set_sensor_model(self._h, pt, st[0]))

    def saveTrace(self, path):
        print("WARNING: trace map not generated, must compile with trace support enabled.")  # RangeLib.h:430

    # -- extensions ------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream; 0 is the
        default stream); None goes back to the method's private stream."""
        if cuda_stream is None:
            check(lib().rl_method_use_own_stream(self._h))
        else:
            check(lib().rl_method_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def synchronize(self):
        check(lib().rl_method_synchronize(self._h))

    def update_map(self, patch_xmajor, x0, y0):
        """Dynamic maps: apply a [w, h] x-major uint8 patch on the device and refresh the structures."""
        if _is_torch(patch_xmajor):
            pp, sp = _buf(patch_xmajor, np.uint8, 2, "patch")
        else:
            patch_xmajor = np.ascontiguousarray(patch_xmajor, dtype=np.uint8)
            pp, sp = _buf(patch_xmajor, np.uint8, 2, "patch")
        check(lib().rl_method_update_map(self._h, pp, int(x0), int(y0), sp[0], sp[1]))

    def set_map_occupancy_grid(self, data):
        """Replace the whole map on the device from ROS OccupancyGrid data: int8 [rows, cols] (numpy or torch CUDA
        tensor), occupied iff value > 10, rows = map width (RangeLibc.pyx:146-157)."""
        if not _is_torch(data):
            data = np.ascontiguousarray(data, dtype=np.int8)
        pd, sd = _buf(data, np.int8, 2, "data")
        check(lib().rl_method_set_map_occupancy_grid(self._h, pd, sd[0], sd[1]))

    def set_map_rgba(self, rgba, threshold=128.0):
        """Replace the whole map on the device from an RGBA8 image [rows, cols, 4] (numpy or torch CUDA tensor) with
        the reference's OMap(png, threshold) conversion (RangeLib.h:189-199); cols = map width, rows = map height."""
        if not _is_torch(rgba):
            rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        pr, sr = _buf(rgba, np.uint8, 3, "rgba")
        if sr[2] != 4:
            raise ValueError("rgba must be [rows, cols, 4]")
        check(lib().rl_method_set_map_rgba(self._h, pr, sr[1], sr[0], float(threshold)))

    def update_map_batch(self, patches, rects):
        """Dynamic maps: n non-overlapping patches in one launch.  rects: int32 [n, 4] (x0, y0, w, h) host array;
        patches: uint8 1-D, the patches' bytes concatenated (numpy or torch CUDA tensor)."""
        rects = np.ascontiguousarray(rects, dtype=np.int32)
        if not _is_torch(patches):
            patches = np.ascontiguousarray(patches, dtype=np.uint8)
        pp, sp = _buf(patches, np.uint8, 1, "patches")
        if rects.ndim != 2 or rects.shape[1] != 4 or int((rects[:, 2].astype(np.int64) * rects[:, 3]).sum()) != sp[0]:
            raise ValueError("rects must be [n,4] and patches must hold sum(w*h) bytes")
        check(lib().rl_method_update_map_batch(self._h, pp, C.c_void_p(rects.ctypes.data), rects.shape[0]))

    def occupancy(self):
        """The occupancy resident on the device, uint8 [W, H] x-major."""
        out = np.empty((self._map.width(), self._map.height()), np.uint8)
        check(lib().rl_debug_get_occ(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def memory(self):
        return int(lib().rl_method_memory(self._h))

    def set_persistent(self, on):
        """Tuning knob (RM): persistent-warp kernel with lane re-queuing for large batches."""
        check(lib().rl_debug_set_persistent(self._h, int(on)))

    def set_spatial_sort(self, on):
        """Tuning knob: tile-ordered processing of big clouds on > L2 structures (default on); CDDT: False = no query
        index, 2 = query index whatever the table size.  Results are identical."""
        check(lib().rl_debug_set_spatial_sort(self._h, 2 if on == 2 else (1 if on else 0)))

    def set_coop_threshold(self, rays):
        """Tuning knob (RM, small launches): a CTA with <= rays live rays finishes them cooperatively (0 = off)."""
        check(lib().rl_debug_set_coop_threshold(self._h, int(rays)))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value and getattr(self, "_L", None) is not None:
            self._L.rl_method_destroy(self._h)
            self._h.value = None


class PyBresenhamsLine(_RangeMethod):
    _kind = cabi.RL_BL

    def __init__(self, Map, max_range, device=-1):
        super().__init__(Map, max_range, 0, device)


class PyRayMarching(_RangeMethod):
    _kind = cabi.RL_RM

    def __init__(self, Map, max_range, device=-1):
        super().__init__(Map, max_range, 0, device)

    def distance_transform(self):
        out = np.empty((self._map.width(), self._map.height()), np.float32)
        check(lib().rl_debug_get_dt(self._h, C.c_void_p(out.ctypes.data)))
        return out


class PyRayMarchingGPU(PyRayMarching):
    """Same backend as PyRayMarching; kept because callers of the reference select the GPU caster
    by this name (RangeLibc.pyx:313-342)."""


class PyCDDTCast(_RangeMethod):
    _kind = cabi.RL_CDDT

    def __init__(self, Map, max_range, theta_disc, device=-1):
        self.theta_disc = int(theta_disc)
        super().__init__(Map, max_range, theta_disc, device)

    def prune(self, max_range=-1.0):
        check(lib().rl_method_prune(self._h, self.max_range if max_range < 0.0 else float(max_range)))

    def save(self, path):
        """Binary checkpoint of the (possibly pruned) table; see rl_method_save_cddt.  (The reference only dumps YAML /
        JSON text for its viewer, RangeLib.h:1652-1735, and cannot load it back.)"""
        check(lib().rl_method_save_cddt(self._h, str(path).encode()))

    @classmethod
    def load(cls, Map, path, device=-1):
        """A CDDT / PCDDT method for `Map` from a checkpoint written by save(): no rebuild, no prune."""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        self._L = lib()
        self._map = Map
        check(lib().rl_method_create_from_cddt(Map._h, str(path).encode(), int(device), C.byref(self._h)))
        mr, td, pr = C.c_float(), C.c_uint(), C.c_int()
        check(lib().rl_method_get_params(self._h, C.byref(mr), C.byref(td), C.byref(pr)))
        self.max_range, self.theta_disc, self.pruned = mr.value, int(td.value), bool(pr.value)
        return self

    def table(self):
        nb, nv = C.c_int64(), C.c_int64()
        widths = np.zeros(self.theta_disc, np.int32)
        trans = np.zeros(self.theta_disc, np.float32)
        check(lib().rl_debug_cddt_dims(self._h, C.byref(nb), C.byref(nv), C.c_void_p(widths.ctypes.data),
                                       C.c_void_p(trans.ctypes.data)))
        offsets = np.zeros(nb.value + 1, np.int64)
        values = np.zeros(max(nv.value, 1), np.float32)
        check(lib().rl_debug_cddt_dump(self._h, C.c_void_p(offsets.ctypes.data), C.c_void_p(values.ctypes.data)))
        return widths, trans, offsets, values[: nv.value]


class PyGiantLUTCast(_RangeMethod):
    """GiantLUTCast (RangeLib.h:1772-1904): uint16 range for every (x, y, theta bin); the W*H*td table is
    filled on the device by the RM kernel and a query is one gather."""
    _kind = cabi.RL_GLT

    def __init__(self, Map, max_range, theta_disc, device=-1):
        self.theta_disc = int(theta_disc)
        super().__init__(Map, max_range, theta_disc, device)

    def table(self):
        out = np.empty((self._map.width(), self._map.height(), self.theta_disc), np.uint16)
        check(lib().rl_debug_glt_dump(self._h, C.c_void_p(out.ctypes.data)))
        return out


def device_sincosf(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    s = np.empty_like(x)
    c = np.empty_like(x)
    check(lib().rl_debug_sincosf(C.c_void_p(x.ctypes.data), C.c_void_p(s.ctypes.data), C.c_void_p(c.ctypes.data),
                                 x.size))
    return s, c
