# cython: language_level=3, boundscheck=False, wraparound=False
"""range_libc -- drop-in replacement of the reference's Cython module of the same name
(/root/reference/pywrapper/RangeLibc.pyx), re-pointed at the B200 C ABI (include/rangelib_b200.h).

Same classes and methods a particle filter written against the reference uses:

    PyOMap(bool_array | (w, h) | OccupancyGrid | png_path[, threshold])
    PyBresenhamsLine / PyRayMarching / PyRayMarchingGPU (omap, max_range)
    PyCDDTCast(omap, max_range, theta_disc)  (+ .prune())
    PyGiantLUTCast(omap, max_range, theta_disc)
      .calc_range(x, y, heading)
      .calc_range_many(ins[N,3], outs[N])
      .calc_range_repeat_angles(ins[N,3], angles[M], outs[N*M])
      .calc_range_repeat_angles_eval_sensor_model(ins, angles, obs[M], weights[N])
      .calc_range_many_radial_optimized(num_rays, min_angle, max_angle, ins, outs[N*num_rays])
      .eval_sensor_model(obs, ranges, outs, num_rays, num_particles)
      .set_sensor_model(table[K,K])

Every method runs on the GPU; there is no CPU implementation behind this module.
Typed memoryviews enforce dtype / ndim / C-contiguity exactly like the reference's buffer
signatures (ValueError otherwise).
"""
import numpy as np

from libc.stdint cimport uint8_t, int64_t, uint64_t

cdef extern from "rangelib_b200.h":
    ctypedef struct rl_map:
        pass
    ctypedef struct rl_method:
        pass
    enum:
        RL_BL
        RL_RM
        RL_CDDT
        RL_PCDDT
        RL_GLT
    const char* rl_last_error()
    uint64_t rl_stat_kernel_launches()
    int rl_map_create(const uint8_t* occ, int w, int h, rl_map** out)
    int rl_map_set_world(rl_map* m, float scale, float angle, float ox, float oy, float s, float c)
    int rl_map_is_occupied(const rl_map* m, int x, int y)
    int rl_map_get(const rl_map* m, uint8_t* out)
    void rl_map_destroy(rl_map* m)
    int rl_method_create(int kind, const rl_map* m, float max_range, unsigned td, int device, rl_method** out)
    void rl_method_destroy(rl_method* m)
    int rl_method_prune(rl_method* m, float max_range)
    int rl_method_save_cddt(rl_method* m, const char* path)
    int rl_method_create_from_cddt(const rl_map* m, const char* path, int device, rl_method** out)
    int rl_method_get_params(const rl_method* m, float* max_range, unsigned* td, int* pruned)
    int rl_pf_normalize_weights(rl_method* m, double* weights, int n, double inv_squash, double* sum_out)
    int rl_pf_resample(rl_method* m, const float* particles, const double* weights, float* out_particles, int n, double u0)
    int rl_pf_motion_update(rl_method* m, float* particles, int n, float dx, float dy, float dtheta, const float* noise)
    int rl_calc_range(rl_method* m, float x, float y, float heading, float* out)
    int rl_calc_range_many(rl_method* m, const float* ins, float* outs, int n)
    int rl_numpy_calc_range(rl_method* m, const float* ins, float* outs, int n)
    int rl_numpy_calc_range_angles(rl_method* m, const float* ins, const float* angles, float* outs, int n, int k)
    int rl_set_sensor_model(rl_method* m, const double* table, int k)
    int rl_eval_sensor_model(rl_method* m, const float* obs, const float* ranges, double* outs, int k, int n)
    int rl_calc_range_repeat_angles_eval_sensor_model(rl_method* m, const float* ins, const float* angles,
                                                      const float* obs, double* weights, int n, int k)
    int rl_calc_range_many_radial_optimized(rl_method* m, const float* ins, float* outs, int n, int num_rays,
                                            float min_angle, float max_angle)
    int rl_method_peers_init(rl_method* m, double** w0, double** w1, int64_t** flags, int n_peers, int rank)
    int rl_calc_range_repeat_angles_eval_sensor_model_sharded(rl_method* m, const float* ins, const float* angles,
                                                              const float* obs, double* weights_all, int64_t offset,
                                                              int n, int k, int64_t n_total)

# the reference exports its compile-time switches; keep the names importable
USE_CACHED_TRIG = False
USE_ALTERNATE_MOD = True
USE_CACHED_CONSTANTS = True
USE_FAST_ROUND = False
NO_INLINE = False
USE_LRU_CACHE = False
LRU_CACHE_SIZE = 1000000
SHOULD_USE_CUDA = True


class RangeLibError(RuntimeError):
    pass


cdef int _ck(int rc) except -1:
    if rc != 0:
        raise RangeLibError("rangelib_b200 error %d: %s" % (rc, rl_last_error().decode("utf-8", "replace")))
    return 0


def kernel_launches():
    return int(rl_stat_kernel_launches())


def _decode_png(path, threshold):
    """The reference's OMap(filename, threshold) (RangeLib.h:159-201): RGBA8 decode, then
    gray = (int)(0.229*B + 0.587*G + 0.114*R) -- the reference reads R and B from swapped byte
    positions -- occupied iff gray < threshold.  Returns uint8 [W, H] x-major."""
    from PIL import Image
    if isinstance(path, bytes):
        path = path.decode()
    img = np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8)
    # all three products and both sums in double, narrowed to float on return and truncated (RangeUtils.h:30-32),
    # exactly as range_libc_b200.mapio.occupancy_from_rgba and the device ingest kernel do
    r = img[:, :, 2].astype(np.float64)
    g = img[:, :, 1].astype(np.float64)
    b = img[:, :, 0].astype(np.float64)
    gray = ((0.229 * r + 0.587 * g) + 0.114 * b).astype(np.float32).astype(np.int32)
    return np.ascontiguousarray((gray < threshold).T, dtype=np.uint8)


cdef class PyOMap:
    cdef rl_map* ptr
    cdef int _w, _h
    cdef bint _err

    def __cinit__(self, arg1, arg2=None):
        self.ptr = NULL
        self._err = False
        world = None
        if isinstance(arg1, (int, np.integer)) and isinstance(arg2, (int, np.integer)):
            occ = np.zeros((int(arg1), int(arg2)), np.uint8)
        elif isinstance(arg1, np.ndarray):
            occ = np.ascontiguousarray(arg1.T != 0, dtype=np.uint8)
        elif isinstance(arg1, (str, bytes)):
            try:
                occ = _decode_png(arg1, 128 if arg2 is None else float(arg2))
            except Exception as ex:
                print("ERROR loading map: %s" % ex)
                self._err = True
                occ = np.zeros((1, 1), np.uint8)
        elif hasattr(arg1, "info") and hasattr(arg1, "data"):
            info = arg1.info
            arr = np.asarray(arg1.data).reshape((info.height, info.width))
            occ = np.ascontiguousarray(arr > 10, dtype=np.uint8)
            q = info.origin.orientation
            yaw = np.arctan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z))
            ang = -1.0 * yaw
            world = (info.resolution, ang, info.origin.position.x, info.origin.position.y, np.sin(ang), np.cos(ang))
        else:
            print("Failed to construct PyOMap, check argument types.")
            occ = np.zeros((1, 1), np.uint8)
        cdef uint8_t[:, ::1] view = occ
        self._w = occ.shape[0]
        self._h = occ.shape[1]
        _ck(rl_map_create(&view[0, 0], self._w, self._h, &self.ptr))
        if world is not None:
            _ck(rl_map_set_world(self.ptr, world[0], world[1], world[2], world[3], world[4], world[5]))

    def __dealloc__(self):
        if self.ptr != NULL:
            rl_map_destroy(self.ptr)

    def set_world(self, float scale=1.0, float angle=0.0, float origin_x=0.0, float origin_y=0.0,
                  float sin_angle=0.0, float cos_angle=1.0):
        _ck(rl_map_set_world(self.ptr, scale, angle, origin_x, origin_y, sin_angle, cos_angle))

    cpdef bint isOccupied(self, int x, int y):
        return rl_map_is_occupied(self.ptr, x, y) == 1

    cpdef bint error(self):
        return self._err

    cpdef int width(self):
        return self._w

    cpdef int height(self):
        return self._h

    def save(self, fn):
        """OMap::save (RangeLib.h:264-291): RGBA PNG, occupied cells black, free cells white, alpha 255.
        Returns False on success like the reference (its return value is lodepng's error code)."""
        from PIL import Image
        np_occ = np.empty((self._w, self._h), np.uint8)
        cdef uint8_t[:, ::1] view = np_occ
        _ck(rl_map_get(self.ptr, &view[0, 0]))
        img = np.full((self._h, self._w, 4), 255, np.uint8)
        img[np_occ.T != 0, :3] = 0
        if isinstance(fn, bytes):
            fn = fn.decode()
        try:
            Image.fromarray(img, "RGBA").save(fn, format="PNG")
        except Exception as ex:
            print("encoder error: %s" % ex)
            return True
        return False


cdef class _Method:
    cdef rl_method* ptr
    cdef float max_range

    def __dealloc__(self):
        if self.ptr != NULL:
            rl_method_destroy(self.ptr)

    cdef _create(self, int kind, PyOMap Map, float max_range, unsigned int theta_disc):
        self.ptr = NULL
        self.max_range = max_range
        _ck(rl_method_create(kind, Map.ptr, max_range, theta_disc, -1, &self.ptr))

    cpdef float calc_range(self, float x, float y, float heading) except? -12345.0:
        cdef float out = 0
        _ck(rl_calc_range(self.ptr, x, y, heading, &out))
        return out

    cpdef calc_range_many(self, float[:, ::1] ins, float[::1] outs):
        if outs.shape[0] == 0:
            return
        if ins.shape[1] != 3 or ins.shape[0] < outs.shape[0]:
            raise ValueError("ins must be [N,3] with N >= len(outs)")
        _ck(rl_numpy_calc_range(self.ptr, &ins[0, 0], &outs[0], <int>outs.shape[0]))

    cpdef calc_range_repeat_angles(self, float[:, ::1] ins, float[::1] angles, float[::1] outs):
        if ins.shape[0] == 0 or angles.shape[0] == 0:
            return
        if ins.shape[1] != 3 or outs.shape[0] < ins.shape[0] * angles.shape[0]:
            raise ValueError("outs must hold N*M floats")
        _ck(rl_numpy_calc_range_angles(self.ptr, &ins[0, 0], &angles[0], &outs[0], <int>ins.shape[0], <int>angles.shape[0]))

    # particle-filter steps either side of the sensor update (extensions: not in the reference; rl_pf.cu)
    def normalize_weights(self, double[::1] weights, double inv_squash=1.0):
        """weights in place: w <- pow(w, inv_squash) / sum; returns the sum of the squashed weights."""
        cdef double total = 0.0
        if weights.shape[0] == 0:
            return 0.0
        _ck(rl_pf_normalize_weights(self.ptr, &weights[0], <int>weights.shape[0], inv_squash, &total))
        return total

    def resample(self, float[:, ::1] particles, double[::1] weights, float[:, ::1] out_particles, double u0):
        """systematic resampling by (normalised) weights with one uniform draw u0 in [0, 1)"""
        if particles.shape[0] == 0:
            return
        if particles.shape[1] != 3 or out_particles.shape[1] != 3 or weights.shape[0] < particles.shape[0] \
                or out_particles.shape[0] < particles.shape[0]:
            raise ValueError("shape mismatch")
        _ck(rl_pf_resample(self.ptr, &particles[0, 0], &weights[0], &out_particles[0, 0], <int>particles.shape[0], u0))

    def motion_update(self, float[:, ::1] particles, float dx, float dy, float dtheta, noise=None):
        """odometry step in place; noise f32[N,3] or None"""
        cdef float[:, ::1] nz
        if particles.shape[0] == 0:
            return
        if particles.shape[1] != 3:
            raise ValueError("particles must be [N,3]")
        if noise is None:
            _ck(rl_pf_motion_update(self.ptr, &particles[0, 0], <int>particles.shape[0], dx, dy, dtheta, NULL))
        else:
            nz = noise
            if nz.shape[0] < particles.shape[0] or nz.shape[1] != 3:
                raise ValueError("noise must be [N,3]")
            _ck(rl_pf_motion_update(self.ptr, &particles[0, 0], <int>particles.shape[0], dx, dy, dtheta, &nz[0, 0]))

    cpdef calc_range_repeat_angles_eval_sensor_model(self, float[:, ::1] ins, float[::1] angles, float[::1] obs,
                                                     double[::1] weights):
        if ins.shape[0] == 0 or angles.shape[0] == 0:
            return
        if ins.shape[1] != 3 or obs.shape[0] < angles.shape[0] or weights.shape[0] < ins.shape[0]:
            raise ValueError("shape mismatch")
        _ck(rl_calc_range_repeat_angles_eval_sensor_model(self.ptr, &ins[0, 0], &angles[0], &obs[0], &weights[0],
                                                          <int>ins.shape[0], <int>angles.shape[0]))

    def peers_init(self, weights0_ptrs, weights1_ptrs, flags_ptrs, int rank):
        """Multi-process extension (no counterpart in the single-GPU reference): peer-mapped device pointers of every
        rank's two gathered-weight buffers and flag array (e.g. torch symmetric memory; see range_libc_b200.parallel)."""
        cdef double* w0[16]
        cdef double* w1[16]
        cdef int64_t* fl[16]
        cdef int n = len(weights0_ptrs)
        if n < 1 or n > 16 or len(weights1_ptrs) != n or len(flags_ptrs) != n:
            raise ValueError("1..16 peers, three pointer lists of equal length")
        for i in range(n):
            w0[i] = <double*><size_t>int(weights0_ptrs[i])
            w1[i] = <double*><size_t>int(weights1_ptrs[i])
            fl[i] = <int64_t*><size_t>int(flags_ptrs[i])
        _ck(rl_method_peers_init(self.ptr, w0, w1, fl, n, rank))

    cpdef calc_range_repeat_angles_eval_sensor_model_sharded(self, float[:, ::1] ins, float[::1] angles, float[::1] obs,
                                                             double[::1] weights_all, long long offset):
        """calc_range_repeat_angles_eval_sensor_model for a cloud sharded over several processes / GPUs: `ins` are
        THIS rank's particles (starting at particle `offset` of the cloud), `weights_all` receives the weights of ALL
        particles (gathered over NVLink by the kernel's epilogue).  Blocking; needs peers_init."""
        if ins.shape[0] == 0 or angles.shape[0] == 0:
            raise ValueError("every rank must pass at least one particle and one angle")
        if ins.shape[1] != 3 or obs.shape[0] < angles.shape[0] or weights_all.shape[0] < offset + ins.shape[0]:
            raise ValueError("shape mismatch")
        _ck(rl_calc_range_repeat_angles_eval_sensor_model_sharded(self.ptr, &ins[0, 0], &angles[0], &obs[0],
                                                                  &weights_all[0], offset, <int>ins.shape[0],
                                                                  <int>angles.shape[0], <int64_t>weights_all.shape[0]))

    cpdef calc_range_many_radial_optimized(self, int num_rays, float min_angle, float max_angle, float[:, ::1] ins,
                                           float[::1] outs):
        if ins.shape[0] == 0:
            return
        if ins.shape[1] != 3 or outs.shape[0] < ins.shape[0] * num_rays:
            raise ValueError("shape mismatch")
        _ck(rl_calc_range_many_radial_optimized(self.ptr, &ins[0, 0], &outs[0], <int>ins.shape[0], num_rays, min_angle,
                                                max_angle))

    cpdef eval_sensor_model(self, float[::1] observation, float[::1] ranges, double[::1] outs, int num_rays,
                            int num_particles):
        if num_particles == 0:
            return
        if observation.shape[0] < num_rays or ranges.shape[0] < num_rays * num_particles or outs.shape[0] < num_particles:
            raise ValueError("shape mismatch")
        _ck(rl_eval_sensor_model(self.ptr, &observation[0], &ranges[0], &outs[0], num_rays, num_particles))

    cpdef set_sensor_model(self, double[:, ::1] table):
        if table.shape[0] != table.shape[1]:
            print("Sensor model must have equal matrix dimensions, failing!")
            return
        _ck(rl_set_sensor_model(self.ptr, &table[0, 0], <int>table.shape[0]))

    def saveTrace(self, path):
        print("WARNING: trace map not generated, must compile with trace support enabled.")


cdef class PyBresenhamsLine(_Method):
    def __cinit__(self, PyOMap Map, float max_range):
        self._create(RL_BL, Map, max_range, 0)


cdef class PyRayMarching(_Method):
    def __cinit__(self, PyOMap Map, float max_range):
        self._create(RL_RM, Map, max_range, 0)


cdef class PyRayMarchingGPU(_Method):
    def __cinit__(self, PyOMap Map, float max_range):
        self._create(RL_RM, Map, max_range, 0)


_pending_checkpoint = None  # set by PyCDDTCast.load around the construction of the object it returns


cdef class PyCDDTCast(_Method):
    def __cinit__(self, PyOMap Map, float max_range, unsigned int theta_disc):
        cdef float mr = 0
        if _pending_checkpoint is not None:
            self.ptr = NULL
            _ck(rl_method_create_from_cddt(Map.ptr, _pending_checkpoint, -1, &self.ptr))
            _ck(rl_method_get_params(self.ptr, &mr, NULL, NULL))
            self.max_range = mr
        else:
            self._create(RL_CDDT, Map, max_range, theta_disc)

    cpdef prune(self, float max_range=-1.0):
        _ck(rl_method_prune(self.ptr, self.max_range if max_range < 0.0 else max_range))

    def save(self, path):
        """Binary checkpoint of the (possibly pruned) table (extension: the reference only dumps YAML / JSON text for
        its viewer, RangeLib.h:1652-1735, and has no loader).  Load with PyCDDTCast.load(omap, path)."""
        p = path if isinstance(path, bytes) else str(path).encode()
        _ck(rl_method_save_cddt(self.ptr, p))

    @staticmethod
    def load(PyOMap Map, path):
        """A CDDT / PCDDT method for `Map` from a checkpoint written by save(): no rebuild, no prune."""
        global _pending_checkpoint
        _pending_checkpoint = path if isinstance(path, bytes) else str(path).encode()
        try:
            return PyCDDTCast(Map, 0.0, 0)
        finally:
            _pending_checkpoint = None


cdef class PyGiantLUTCast(_Method):
    def __cinit__(self, PyOMap Map, float max_range, unsigned int theta_disc):
        self._create(RL_GLT, Map, max_range, theta_disc)


cdef class PyNull:
    """The reference's call-overhead dummy (RangeLibc.pyx:342-352): touches its arguments, computes nothing."""
    def __cinit__(self, PyOMap Map, float max_range, unsigned int theta_disc):
        pass

    cpdef float calc_range(self, float x, float y, float heading):
        return x + y + heading

    cpdef calc_range_many(self, float[:, ::1] ins, float[::1] outs):
        a = ins[0, 0]
        b = outs[0]
        c = outs.shape[0]
