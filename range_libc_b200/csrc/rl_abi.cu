// extern "C" boundary of librangelib_b200.so (include/rangelib_b200.h).
// Handles, pointer classification (host vs device), staging for host-pointer calls, dispatch.
// There is no CPU fallback anywhere in this library: without a usable device every entry point
// that computes fails with RL_E_NO_DEVICE.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rl_internal.cuh"

namespace rl {

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { g_err = msg; }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  g_err = buf;
  cudaGetLastError();  // clear sticky-less errors
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? RL_E_NO_DEVICE : RL_E_CUDA;
}

// where a caller pointer lives
enum Side { SIDE_PAGEABLE = 0, SIDE_DEVICE = 1, SIDE_PINNED = 2 };

static Side pointer_side(const void* p, void** dev_alias) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return SIDE_PAGEABLE;
  }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) {
    if (dev_alias) *dev_alias = const_cast<void*>(p);
    return SIDE_DEVICE;
  }
  if (at.type == cudaMemoryTypeHost && at.devicePointer) {
    if (dev_alias) *dev_alias = at.devicePointer;
    return SIDE_PINNED;
  }
  return SIDE_PAGEABLE;
}

static int is_device_ptr(const void* p) { return pointer_side(p, nullptr) == SIDE_DEVICE; }

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static int ensure_stage(rl_method* m, size_t bytes) {
  if (bytes <= m->d_stage_bytes) return RL_OK;
  if (m->d_stage) cudaFree(m->d_stage);
  m->d_stage = nullptr;
  m->d_stage_bytes = 0;
  size_t want = bytes + bytes / 4 + 4096;
  RL_CUDA(cudaMalloc(&m->d_stage, want));
  m->d_stage_bytes = want;
  return RL_OK;
}

static int ensure_host_stage(rl_method* m, size_t bytes) {
  if (bytes <= m->h_stage_bytes) return RL_OK;
  if (m->h_stage) cudaFreeHost(m->h_stage);
  m->h_stage = nullptr;
  m->h_stage_dev = nullptr;
  m->h_stage_bytes = 0;
  size_t want = bytes + bytes / 2 + 65536;
  RL_CUDA(cudaHostAlloc(&m->h_stage, want, cudaHostAllocMapped | cudaHostAllocPortable));
  RL_CUDA(cudaHostGetDevicePointer(&m->h_stage_dev, m->h_stage, 0));
  m->h_stage_bytes = want;
  return RL_OK;
}

// Makes the handle's device current for the duration of one ABI call and restores the caller's device on exit,
// so a process that drives several GPUs from one thread does not find its current device changed by this library.
class DeviceGuard {
 public:
  ~DeviceGuard() {
    if (switched_) cudaSetDevice(prev_);
  }
  int bind(int device) {
    if (cudaGetDevice(&prev_) != cudaSuccess) {
      cudaGetLastError();
      prev_ = device;
    }
    if (prev_ != device) {
      RL_CUDA(cudaSetDevice(device));
      switched_ = true;
    }
    return RL_OK;
  }
  int bind(rl_method* m) {
    if (!m) {
      set_error("null method handle");
      return RL_E_INVALID;
    }
    return bind(m->device);
  }

 private:
  int prev_ = 0;
  bool switched_ = false;
};

// Marshals the data pointers of one call.
//   all device             -> run in place, asynchronous
//   host, small (<= 4 MB)  -> the inputs are packed into the handle's pinned staging buffer (one
//                             memcpy each) and cross PCIe in ONE async copy; the results are written
//                             by the kernel straight into pinned (mapped) host memory, so no
//                             device-to-host copy is enqueued.  A particle-filter update is
//                             ~50 KB in / 32 KB out: per-copy latency, not bandwidth, is what costs.
//   host, large            -> device staging, one async copy per array on the handle's stream
// and blocks until the results are in the caller's buffer (the reference's semantics).
class Marshal {
 public:
  static constexpr size_t kSmallLimit = 4u << 20;
  static constexpr int kMax = 6;
  explicit Marshal(rl_method* m) : m_(m) {}
  int add(const void* p, size_t bytes, bool output) {
    a_[n_] = Arg{const_cast<void*>(p), bytes, output, nullptr, 0, SIDE_PAGEABLE, false};
    return n_++;
  }
  // a buffer the kernel updates in part: travels to the device with the inputs and comes back whole
  int add_inout(void* p, size_t bytes) {
    a_[n_] = Arg{p, bytes, false, nullptr, 0, SIDE_PAGEABLE, true};
    return n_++;
  }
  void* dev(int i) const { return a_[i].dev; }
  bool on_device() const { return all_device_; }

  int prepare() {
    // results of small calls are written by the kernel straight into the pinned (mapped) staging buffer:
    // no device-to-host copy is enqueued (measured 3-4 us per particle-filter update); RL_ZEROCOPY_OUT=0 disables
    static const bool zc_out = !(getenv("RL_ZEROCOPY_OUT") && atoi(getenv("RL_ZEROCOPY_OUT")) == 0);
    int ndev = 0, nhost = 0;
    size_t in_bytes = 0, out_bytes = 0;
    for (int i = 0; i < n_; ++i) {
      Arg& a = a_[i];
      if (!a.host || a.bytes == 0) continue;
      a.side = pointer_side(a.host, &a.dev);
      if (a.side == SIDE_DEVICE) ++ndev; else ++nhost;
      // inputs first, then outputs, each 256-byte aligned inside the staging buffers
      if (!a.output) { a.off = in_bytes; in_bytes += align256(a.bytes); }
    }
    for (int i = 0; i < n_; ++i) {
      Arg& a = a_[i];
      if (!a.host || a.bytes == 0 || !a.output) continue;
      a.off = in_bytes + out_bytes;
      out_bytes += align256(a.bytes);
    }
    if (ndev && nhost) {
      set_error("host and device pointers mixed in one call");
      return RL_E_MIXED;
    }
    all_device_ = nhost == 0;
    if (all_device_) return RL_OK;
    in_bytes_ = in_bytes;
    out_bytes_ = out_bytes;
    int rc = ensure_stage(m_, in_bytes + out_bytes);
    if (rc) return rc;
    small_ = in_bytes + out_bytes <= kSmallLimit;
    if (small_) {
      rc = ensure_host_stage(m_, in_bytes + out_bytes);
      if (rc) return rc;
      zc_out_ = zc_out;
      for (int i = 0; i < n_; ++i) {
        Arg& a = a_[i];
        if (!a.host || a.bytes == 0) continue;
        if (!a.output) {
          memcpy((char*)m_->h_stage + a.off, a.host, a.bytes);
          a.dev = (char*)m_->d_stage + a.off;
        } else {
          a.dev = zc_out_ ? (char*)m_->h_stage_dev + a.off : (char*)m_->d_stage + a.off;
        }
      }
      if (in_bytes) RL_CUDA(cudaMemcpyAsync(m_->d_stage, m_->h_stage, in_bytes, cudaMemcpyHostToDevice, m_->stream));
      return RL_OK;
    }
    for (int i = 0; i < n_; ++i) {
      Arg& a = a_[i];
      if (!a.host || a.bytes == 0) continue;
      a.dev = (char*)m_->d_stage + a.off;
      if (!a.output) RL_CUDA(cudaMemcpyAsync(a.dev, a.host, a.bytes, cudaMemcpyHostToDevice, m_->stream));
    }
    return RL_OK;
  }

  int finish() {
    if (all_device_) return RL_OK;
    if (small_) {
      if (!zc_out_ && out_bytes_)
        RL_CUDA(cudaMemcpyAsync((char*)m_->h_stage + in_bytes_, (char*)m_->d_stage + in_bytes_, out_bytes_,
                                cudaMemcpyDeviceToHost, m_->stream));
      for (int i = 0; i < n_; ++i) {
        Arg& a = a_[i];
        if (a.host && a.bytes && a.inout)
          RL_CUDA(cudaMemcpyAsync((char*)m_->h_stage + a.off, (char*)m_->d_stage + a.off, a.bytes,
                                  cudaMemcpyDeviceToHost, m_->stream));
      }
      RL_CUDA(cudaStreamSynchronize(m_->stream));
      for (int i = 0; i < n_; ++i) {
        Arg& a = a_[i];
        if (a.host && a.bytes && (a.output || a.inout)) memcpy(a.host, (char*)m_->h_stage + a.off, a.bytes);
      }
      return RL_OK;
    }
    for (int i = 0; i < n_; ++i) {
      Arg& a = a_[i];
      if (a.host && a.bytes && (a.output || a.inout))
        RL_CUDA(cudaMemcpyAsync(a.host, a.dev, a.bytes, cudaMemcpyDeviceToHost, m_->stream));
    }
    RL_CUDA(cudaStreamSynchronize(m_->stream));
    return RL_OK;
  }

 private:
  struct Arg {
    void* host;
    size_t bytes;
    bool output;
    void* dev;
    size_t off;
    Side side;
    bool inout;
  };
  rl_method* m_;
  Arg a_[kMax];
  int n_ = 0;
  size_t in_bytes_ = 0, out_bytes_ = 0;
  bool all_device_ = true, small_ = false, zc_out_ = false;
};

// The fused call with small HOST buffers (a particle-filter update: ~48 KB of poses in, 32 KB of weights out).
// Per-transfer latency is what costs here, so: the poses cross PCIe in one async copy straight from the caller's
// buffer when it is pinned (through the pinned staging buffer otherwise); angles and observation travel inside the
// kernel launch (rl::BeamParams); the kernel stores the weights directly into pinned host memory (the caller's
// buffer if pinned).  Measured on B200, 4000 x 60 RM: 47.1 -> 42.7 us per blocking call (tests/probes/e2e_probe.py).
// Also measured and dropped: CTAs fetching their poses from mapped host memory instead of the copy (56 us: a
// thousand 48-byte PCIe reads), and a completion flag in host memory raised by the last CTA after a system-scope
// fence instead of the stream synchronisation (+2 us).
#ifdef RL_HOST_TIMING
#include <time.h>
static double g_ht[8];
static long g_ht_n;
static inline double ht_now() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
#define RL_HT(i) do { const double t_ = ht_now(); g_ht[i] += t_ - ht_prev; ht_prev = t_; } while (0)
extern "C" void rl_debug_host_timing(double* out, long* n) {
  for (int i = 0; i < 8; ++i) { out[i] = g_ht[i]; g_ht[i] = 0; }
  *n = g_ht_n;
  g_ht_n = 0;
}
#else
#define RL_HT(i)
#endif
static int run_fused_host_small(rl_method* m, const float* ins, const float* angles, const float* obs, double* weights,
                                int n, int M, bool* handled) {
  static const bool enabled = !(getenv("RL_HOST_DIRECT") && atoi(getenv("RL_HOST_DIRECT")) == 0);
#ifdef RL_HOST_TIMING
  double ht_prev = ht_now();
  ++g_ht_n;
#endif
  *handled = false;
  const size_t in_bytes = sizeof(float) * 3 * (size_t)n, out_bytes = sizeof(double) * (size_t)n;
  if (!enabled || M > RL_PARAM_BEAMS || in_bytes + out_bytes > Marshal::kSmallLimit) return RL_OK;
  void *alias = nullptr, *d_w = nullptr;
  const Side s_ins = pointer_side(ins, &alias), s_w = pointer_side(weights, &d_w);
  if (s_ins == SIDE_DEVICE || s_w == SIDE_DEVICE || pointer_side(angles, &alias) == SIDE_DEVICE ||
      pointer_side(obs, &alias) == SIDE_DEVICE)
    return RL_OK;  // device or mixed pointers: the general path decides
  RL_HT(0);  // pointer queries
  *handled = true;
  int rc = ensure_stage(m, in_bytes);
  if (!rc) rc = ensure_host_stage(m, align256(in_bytes) + out_bytes);
  if (rc) return rc;
  const void* src = ins;
  if (s_ins != SIDE_PINNED) {
    memcpy(m->h_stage, ins, in_bytes);
    src = m->h_stage;
  }
  RL_HT(1);  // staging
  RL_CUDA(cudaMemcpyAsync(m->d_stage, src, in_bytes, cudaMemcpyHostToDevice, m->stream));
  RL_HT(2);  // H2D enqueue
  BeamParams beams;
  memcpy(beams.angles, angles, sizeof(float) * (size_t)M);
  memcpy(beams.obs, obs, sizeof(float) * (size_t)M);
  double* w_alias = (s_w == SIDE_PINNED) ? (double*)d_w : (double*)((char*)m->h_stage_dev + align256(in_bytes));
  rc = launch_fused_beam_params(m, (const float*)m->d_stage, beams, w_alias, n, M);
  if (rc) return rc;
  RL_HT(3);  // kernel launch
  RL_CUDA(cudaStreamSynchronize(m->stream));
  RL_HT(4);  // synchronise
  if (s_w != SIDE_PINNED) memcpy(weights, (char*)m->h_stage + align256(in_bytes), out_bytes);
  RL_HT(5);  // copy out
  return RL_OK;
}

// Common driver for the four batched cast entry points.
static int run_cast(rl_method* m, int mode, const float* ins, const float* angles, const float* obs, float* outs,
                    double* weights, int n, int M) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (n < 0 || M < 0) {
    set_error("negative count");
    return RL_E_INVALID;
  }
  if (n == 0 || (mode >= MODE_ANGLES && M == 0)) return RL_OK;
  if (!ins || (mode >= MODE_ANGLES && !angles) || (mode == MODE_FUSED && (!obs || !weights)) ||
      (mode != MODE_FUSED && !outs)) {
    set_error("null data pointer");
    return RL_E_INVALID;
  }
  if (mode == MODE_FUSED) {
    bool handled = false;
    rc = run_fused_host_small(m, ins, angles, obs, weights, n, M, &handled);
    if (rc || handled) return rc;
  }
  const size_t n_out = mode == MODE_ANGLES ? (size_t)n * M : (size_t)n;
  Marshal ms(m);
  const int i_ins = ms.add(ins, sizeof(float) * 3 * (size_t)n, false);
  const int i_ang = ms.add(mode >= MODE_ANGLES ? angles : nullptr, sizeof(float) * (size_t)M, false);
  const int i_obs = ms.add(mode == MODE_FUSED ? obs : nullptr, sizeof(float) * (size_t)M, false);
  const int i_out = mode == MODE_FUSED ? ms.add(weights, sizeof(double) * (size_t)n, true)
                                       : ms.add(outs, sizeof(float) * n_out, true);
  rc = ms.prepare();
  if (rc) return rc;
  rc = launch_cast(m, mode, (const float*)ms.dev(i_ins), (const float*)ms.dev(i_ang), (const float*)ms.dev(i_obs),
                   mode == MODE_FUSED ? nullptr : (float*)ms.dev(i_out),
                   mode == MODE_FUSED ? (double*)ms.dev(i_out) : nullptr, n, M);
  if (rc) return rc;
  return ms.finish();
}

// after the occupancy changed on the device: rebuild what the kind derives from it
static int refresh_structures(rl_method* m) {
  int rc = RL_OK;
  if (m->kind == RL_RM || m->kind == RL_GLT) rc = build_distance_transform(m);
  if (!rc && m->kind == RL_GLT) rc = glt_build(m);
  if (m->kind == RL_CDDT || m->kind == RL_PCDDT) {
    const bool was_pruned = m->pruned;
    rc = cddt_build(m);
    if (!rc && was_pruned) rc = cddt_prune(m, m->max_range);
  }
  return rc;
}

// whole-map replacement from a device- or host-resident source image; `bytes` of it are staged when on the host
template <class F>
static int replace_map(rl_method* m, const void* src, size_t bytes, F ingest) {
  const void* d_src = src;
  if (bytes && !is_device_ptr(src)) {
    int rc = ensure_stage(m, bytes);
    if (rc) return rc;
    RL_CUDA(cudaMemcpyAsync(m->d_stage, src, bytes, cudaMemcpyHostToDevice, m->stream));
    d_src = m->d_stage;
  }
  int rc = ingest(d_src);
  if (rc) return rc;
  return refresh_structures(m);
}

static void free_method(rl_method* m) {
  if (!m) return;
  DeviceGuard dg;
  dg.bind(m->device);
  cddt_free(m);
  sort_free(m);
  pf_free(m);
  cudaFree(m->d_occ);
  cudaFree(m->d_bits_t);
  cudaFree(m->d_dt);
  cudaFree(m->d_glt);
  cudaFree(m->d_table);
  cudaFree(m->d_stage);
  cudaFree(m->d_epoch);
  cudaFree(m->d_counter);
  cudaFree(m->d_work);
  cudaFree(m->d_ts_ranges);
  cudaFree(m->d_ts_poses);
  cudaFree(m->d_radial);
  if (m->h_stage) cudaFreeHost(m->h_stage);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  delete m;
}

}  // namespace rl

using namespace rl;

extern "C" {

const char* rl_last_error(void) { return g_err.c_str(); }
uint64_t rl_stat_kernel_launches(void) { return g_launches.load(); }

int rl_map_create(const uint8_t* occ, int W, int H, rl_map** out) {
  if (!out || W < 0 || H < 0 || (!occ && (size_t)W * H > 0)) {
    set_error("rl_map_create: bad arguments");
    return RL_E_INVALID;
  }
  rl_map* m = new rl_map();
  m->W = W;
  m->H = H;
  m->occ.resize((size_t)W * H);
  for (size_t i = 0; i < m->occ.size(); ++i) m->occ[i] = occ[i] ? 1 : 0;
  *out = m;
  return RL_OK;
}

int rl_map_set_world(rl_map* map, float scale, float angle, float ox, float oy, float sin_a, float cos_a) {
  if (!map) {
    set_error("null map");
    return RL_E_INVALID;
  }
  map->scale = scale;
  map->angle = angle;
  map->ox = ox;
  map->oy = oy;
  map->sin_a = sin_a;
  map->cos_a = cos_a;
  return RL_OK;
}

int rl_map_width(const rl_map* map) { return map ? map->W : RL_E_INVALID; }
int rl_map_height(const rl_map* map) { return map ? map->H : RL_E_INVALID; }

int rl_map_is_occupied(const rl_map* map, int x, int y) {
  if (!map) return RL_E_INVALID;
  if (x < 0 || x >= map->W || y < 0 || y >= map->H) return 0;
  return map->occ[(size_t)x * map->H + y];
}

int rl_map_get(const rl_map* map, uint8_t* out) {
  if (!map || !out) return RL_E_INVALID;
  memcpy(out, map->occ.data(), map->occ.size());
  return RL_OK;
}

int rl_map_update(rl_map* map, const uint8_t* patch, int x0, int y0, int w, int h) {
  if (!map || !patch || x0 < 0 || y0 < 0 || w < 0 || h < 0 || x0 + w > map->W || y0 + h > map->H) {
    set_error("rl_map_update: patch outside the map");
    return RL_E_INVALID;
  }
  for (int x = 0; x < w; ++x)
    for (int y = 0; y < h; ++y) map->occ[(size_t)(x0 + x) * map->H + (y0 + y)] = patch[(size_t)x * h + y] ? 1 : 0;
  return RL_OK;
}

void rl_map_destroy(rl_map* map) { delete map; }

int rl_method_create(int kind, const rl_map* map, float max_range, unsigned td, int device, rl_method** out) {
  if (!map || !out || kind < RL_BL || kind > RL_GLT) {
    set_error("rl_method_create: bad arguments");
    return RL_E_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("rangelib_b200 needs a CUDA device (sm_100); there is no CPU fallback");
    return RL_E_NO_DEVICE;
  }
  if (device < 0) RL_CUDA(cudaGetDevice(&device));
  if (device >= ndev) {
    set_error("rl_method_create: no such device");
    return RL_E_INVALID;
  }
  int major = 0;
  RL_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) {
    set_error("rangelib_b200 is built for sm_100a only; this device is not a B200-class GPU");
    return RL_E_NO_DEVICE;
  }
  DeviceGuard dg;
  {
    const int rc0 = dg.bind(device);
    if (rc0) return rc0;
  }
  rl_method* m = new rl_method();
  m->kind = kind;
  m->device = device;
  m->W = map->W;
  m->H = map->H;
  m->max_range = max_range;
  m->td = td;
  // RangeLib.h:442-450
  m->xf.inv_scale = (float)(1.0 / (double)map->scale);
  m->xf.scale = map->scale;
  m->xf.ox = map->ox;
  m->xf.oy = map->oy;
  m->xf.sin_a = map->sin_a;
  m->xf.cos_a = map->cos_a;
  m->xf.rot = (float)(-1.0 * (double)map->angle - 3.0 * RL_PI / 2.0);
  int rc = RL_OK;
  e = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) rc = cuda_fail(e, "stream create", __FILE__, __LINE__);
  m->stream = m->own_stream;
  if (!rc) rc = upload_occupancy(m, map);
  if (!rc && kind == RL_GLT && td == 0) {
    set_error("GiantLUTCast: theta_discretization must be > 0");
    rc = RL_E_INVALID;
  }
  if (!rc && (kind == RL_RM || kind == RL_GLT)) rc = build_distance_transform(m);
  if (!rc && kind == RL_GLT) rc = glt_build(m);
  if (!rc && (kind == RL_CDDT || kind == RL_PCDDT)) rc = cddt_build(m);
  if (!rc && kind == RL_PCDDT) rc = cddt_prune(m, max_range);
  if (!rc) {
    e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "method build", __FILE__, __LINE__);
  }
  if (rc) {
    free_method(m);
    return rc;
  }
  *out = m;
  return RL_OK;
}

void rl_method_destroy(rl_method* m) { free_method(m); }

int rl_method_get_params(const rl_method* m, float* max_range, unsigned* theta_discretization, int* pruned) {
  if (!m) return RL_E_INVALID;
  if (max_range) *max_range = m->max_range;
  if (theta_discretization) *theta_discretization = m->td;
  if (pruned) *pruned = m->pruned ? 1 : 0;
  return RL_OK;
}

int rl_method_save_cddt(rl_method* m, const char* path) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if ((m->kind != RL_CDDT && m->kind != RL_PCDDT) || !path) {
    set_error("rl_method_save_cddt: not a CDDT method / null path");
    return RL_E_STATE;
  }
  return cddt_save(m, path);
}

int rl_method_create_from_cddt(const rl_map* map, const char* path, int device, rl_method** out) {
  if (!map || !path || !out) {
    set_error("rl_method_create_from_cddt: bad arguments");
    return RL_E_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("rangelib_b200 needs a CUDA device (sm_100); there is no CPU fallback");
    return RL_E_NO_DEVICE;
  }
  if (device < 0) RL_CUDA(cudaGetDevice(&device));
  if (device >= ndev) {
    set_error("rl_method_create_from_cddt: no such device");
    return RL_E_INVALID;
  }
  int major = 0;
  RL_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) {
    set_error("rangelib_b200 is built for sm_100a only; this device is not a B200-class GPU");
    return RL_E_NO_DEVICE;
  }
  DeviceGuard dg;
  int rc = dg.bind(device);
  if (rc) return rc;
  rl_method* m = new rl_method();
  m->kind = RL_CDDT;
  m->device = device;
  m->W = map->W;
  m->H = map->H;
  m->xf.inv_scale = (float)(1.0 / (double)map->scale);  // RangeLib.h:442-450, as in rl_method_create
  m->xf.scale = map->scale;
  m->xf.ox = map->ox;
  m->xf.oy = map->oy;
  m->xf.sin_a = map->sin_a;
  m->xf.cos_a = map->cos_a;
  m->xf.rot = (float)(-1.0 * (double)map->angle - 3.0 * RL_PI / 2.0);
  e = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) rc = cuda_fail(e, "stream create", __FILE__, __LINE__);
  m->stream = m->own_stream;
  if (!rc) rc = upload_occupancy(m, map);
  if (!rc) rc = cddt_load(m, path);
  if (rc) {
    free_method(m);
    return rc;
  }
  *out = m;
  return RL_OK;
}

int rl_method_prune(rl_method* m, float max_range) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (m->kind != RL_CDDT && m->kind != RL_PCDDT) {
    set_error("prune is only defined for CDDT");
    return RL_E_STATE;
  }
  return cddt_prune(m, max_range);
}

int rl_method_set_stream(rl_method* m, void* stream) {
  if (!m) return RL_E_INVALID;
  m->stream = (cudaStream_t)stream;  // 0 is CUDA's default stream, as everywhere in the runtime API
  return RL_OK;
}

int rl_method_use_own_stream(rl_method* m) {
  if (!m) return RL_E_INVALID;
  m->stream = m->own_stream;
  return RL_OK;
}

int rl_method_synchronize(rl_method* m) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  RL_CUDA(cudaStreamSynchronize(m->stream));
  return RL_OK;
}

int rl_method_update_map(rl_method* m, const uint8_t* patch, int x0, int y0, int w, int h) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (!patch || x0 < 0 || y0 < 0 || w <= 0 || h <= 0 || x0 + w > m->W || y0 + h > m->H) {
    set_error("rl_method_update_map: patch outside the map");
    return RL_E_INVALID;
  }
  const uint8_t* d_patch = patch;
  if (!is_device_ptr(patch)) {
    rc = ensure_stage(m, (size_t)w * h);
    if (rc) return rc;
    RL_CUDA(cudaMemcpyAsync(m->d_stage, patch, (size_t)w * h, cudaMemcpyHostToDevice, m->stream));
    d_patch = (const uint8_t*)m->d_stage;
  }
  rc = apply_patch(m, d_patch, x0, y0, w, h);
  if (rc) return rc;
  return refresh_structures(m);
}

int rl_method_set_map_occupancy_grid(rl_method* m, const int8_t* data, int rows, int cols) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (rows != m->W || cols != m->H || (!data && (size_t)rows * cols > 0)) {
    set_error("rl_method_set_map_occupancy_grid: need data[rows][cols] with rows == map width and cols == map height");
    return RL_E_INVALID;
  }
  return replace_map(m, data, (size_t)rows * cols, [&](const void* d) { return ingest_occupancy_grid(m, (const int8_t*)d); });
}

int rl_method_set_map_rgba(rl_method* m, const uint8_t* rgba, int img_w, int img_h, float threshold) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (img_w != m->W || img_h != m->H || (!rgba && (size_t)img_w * img_h > 0)) {
    set_error("rl_method_set_map_rgba: image size differs from the map's");
    return RL_E_INVALID;
  }
  return replace_map(m, rgba, (size_t)img_w * img_h * 4, [&](const void* d) { return ingest_rgba(m, (const uint8_t*)d, threshold); });
}

int rl_debug_set_coop_threshold(rl_method* m, int lanes) {
  if (!m) return RL_E_INVALID;
  m->coop_threshold = lanes < 0 ? 0 : lanes;
  return RL_OK;
}

int rl_debug_set_persistent(rl_method* m, int on) {
  if (!m) return RL_E_INVALID;
  m->persist = on;
  return RL_OK;
}

int rl_debug_set_spatial_sort(rl_method* m, int on) {
  if (!m) return RL_E_INVALID;
  m->spatial_sort = on;
  if (on == 2 && (m->kind == RL_CDDT || m->kind == RL_PCDDT) && !m->use_index) {  // tests: index on a small table
    DeviceGuard dg;
    int rc = dg.bind(m);
    if (rc) return rc;
    return cddt_index_build(m, true);
  }
  return RL_OK;
}

int rl_method_update_map_batch(rl_method* m, const uint8_t* patches, const int* rects, int n) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!patches || !rects))) {
    set_error("rl_method_update_map_batch: bad arguments");
    return RL_E_INVALID;
  }
  if (n == 0) return RL_OK;
  std::vector<long long> offsets(n);
  long long total = 0;
  for (int p = 0; p < n; ++p) {
    const int x0 = rects[4 * p], y0 = rects[4 * p + 1], w = rects[4 * p + 2], h = rects[4 * p + 3];
    if (x0 < 0 || y0 < 0 || w <= 0 || h <= 0 || x0 + w > m->W || y0 + h > m->H) {
      set_error("rl_method_update_map_batch: patch outside the map");
      return RL_E_INVALID;
    }
    offsets[p] = total;
    total += (long long)w * h;
  }
  // cell-wise overlap makes the result depend on the CTA schedule: reject it (sweep over x-sorted rectangles;
  // a frame of BASELINE config 4 has 64 patches -- beyond 4096 the caller's word is taken)
  if (n <= 4096) {
    std::vector<int> order(n);
    for (int p = 0; p < n; ++p) order[p] = p;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return rects[4 * a] < rects[4 * b]; });
    for (int i = 0; i < n; ++i) {
      const int* a = rects + 4 * order[i];
      for (int j = i + 1; j < n && rects[4 * order[j]] < a[0] + a[2]; ++j) {
        const int* b = rects + 4 * order[j];
        if (b[1] < a[1] + a[3] && a[1] < b[1] + b[3]) {
          set_error("rl_method_update_map_batch: patches overlap");
          return RL_E_INVALID;
        }
      }
    }
  }
  // staging: [rects | offsets | patches (if on the host)]
  const size_t b_rects = align256(sizeof(int) * 4 * (size_t)n), b_off = align256(sizeof(long long) * (size_t)n);
  const bool dev_patches = is_device_ptr(patches);
  rc = ensure_stage(m, b_rects + b_off + (dev_patches ? 0 : align256((size_t)total)));
  if (rc) return rc;
  char* base = (char*)m->d_stage;
  RL_CUDA(cudaMemcpyAsync(base, rects, sizeof(int) * 4 * (size_t)n, cudaMemcpyHostToDevice, m->stream));
  RL_CUDA(cudaMemcpyAsync(base + b_rects, offsets.data(), sizeof(long long) * (size_t)n, cudaMemcpyHostToDevice, m->stream));
  const uint8_t* d_patches = patches;
  if (!dev_patches) {
    RL_CUDA(cudaMemcpyAsync(base + b_rects + b_off, patches, (size_t)total, cudaMemcpyHostToDevice, m->stream));
    d_patches = (const uint8_t*)(base + b_rects + b_off);
  }
  rc = apply_patch_batch(m, d_patches, (const int*)base, (const long long*)(base + b_rects), n);
  if (rc) return rc;
  RL_CUDA(cudaStreamSynchronize(m->stream));  // `offsets` and the caller's host arrays are released on return
  return refresh_structures(m);
}

int64_t rl_method_memory(const rl_method* m) {
  if (!m) return RL_E_INVALID;
  int64_t bytes = (int64_t)m->W * m->H + (int64_t)m->tiles8_x() * m->tiles8_y() * 8;
  if (m->kind == RL_RM || m->kind == RL_GLT) bytes += (int64_t)m->dt_elems() * 4;
  if (m->kind == RL_GLT) bytes += (int64_t)m->W * m->H * m->td * 2;
  if (m->kind == RL_CDDT || m->kind == RL_PCDDT) bytes += m->nvalues * 4 + (m->nbins + 1) * 8 + (int64_t)m->td * 16 + ((m->use_index && m->spatial_sort) ? m->nbins * 16 + m->nskip * 2 : 0);
  return bytes;
}

int rl_calc_range(rl_method* m, float x, float y, float heading, float* out) {
  if (!out) return RL_E_INVALID;
  float in[3] = {x, y, heading};
  return run_cast(m, MODE_GRID, in, nullptr, nullptr, out, nullptr, 1, 0);
}

int rl_calc_range_many(rl_method* m, const float* ins, float* outs, int n) {
  return run_cast(m, MODE_GRID, ins, nullptr, nullptr, outs, nullptr, n, 0);
}

int rl_numpy_calc_range(rl_method* m, const float* ins, float* outs, int n) {
  return run_cast(m, MODE_WORLD, ins, nullptr, nullptr, outs, nullptr, n, 0);
}

int rl_numpy_calc_range_angles(rl_method* m, const float* ins, const float* angles, float* outs, int n, int M) {
  return run_cast(m, MODE_ANGLES, ins, angles, nullptr, outs, nullptr, n, M);
}

int rl_set_sensor_model(rl_method* m, const double* table, int k) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (!table || k <= 0) {
    set_error("set_sensor_model: bad table");
    return RL_E_INVALID;
  }
  RL_CUDA(cudaStreamSynchronize(m->stream));
  if (m->d_table) cudaFree(m->d_table);
  m->d_table = nullptr;
  m->K = 0;
  RL_CUDA(cudaMalloc(&m->d_table, sizeof(double) * (size_t)k * k));
  RL_CUDA(cudaMemcpyAsync(m->d_table, table, sizeof(double) * (size_t)k * k, cudaMemcpyDefault, m->stream));
  RL_CUDA(cudaStreamSynchronize(m->stream));
  m->K = k;
  return RL_OK;
}

int rl_eval_sensor_model(rl_method* m, const float* obs, const float* ranges, double* outs, int M, int n) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (n < 0 || M < 0) return RL_E_INVALID;
  if (n == 0) return RL_OK;
  if (!obs || !ranges || !outs) {
    if (M == 0 && outs) {
      // product over zero beams is 1.0 (RangeLib.h:544)
    } else {
      set_error("null data pointer");
      return RL_E_INVALID;
    }
  }
  Marshal ms(m);
  const int i_obs = ms.add(obs, sizeof(float) * (size_t)M, false);
  const int i_rng = ms.add(ranges, sizeof(float) * (size_t)n * M, false);
  const int i_out = ms.add(outs, sizeof(double) * (size_t)n, true);
  rc = ms.prepare();
  if (rc) return rc;
  rc = launch_eval_sensor(m, (const float*)ms.dev(i_obs), (const float*)ms.dev(i_rng), (double*)ms.dev(i_out), M, n);
  if (rc) return rc;
  return ms.finish();
}

int rl_calc_range_repeat_angles_eval_sensor_model(rl_method* m, const float* ins, const float* angles,
                                                  const float* obs, double* weights, int n, int M) {
  return run_cast(m, MODE_FUSED, ins, angles, obs, nullptr, weights, n, M);
}

// ---- particle-filter steps either side of the sensor update (rl_pf.cu; SURVEY 8 f4, not in the reference) ----
int rl_pf_normalize_weights(rl_method* m, double* weights, int n, double inv_squash, double* sum_out) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (n < 0 || (n > 0 && !weights)) {
    set_error("rl_pf_normalize_weights: bad arguments");
    return RL_E_INVALID;
  }
  if (n == 0) {
    if (sum_out) *sum_out = 0.0;
    return RL_OK;
  }
  Marshal ms(m);
  const int iw = ms.add_inout(weights, sizeof(double) * (size_t)n);
  rc = ms.prepare();
  if (rc) return rc;
  rc = pf_normalize(m, (double*)ms.dev(iw), n, inv_squash, sum_out);
  if (rc) return rc;
  return ms.finish();
}

int rl_pf_resample(rl_method* m, const float* particles, const double* weights, float* out_particles, int n, double u0) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!particles || !weights || !out_particles)) || !(u0 >= 0.0 && u0 < 1.0)) {
    set_error("rl_pf_resample: bad arguments (u0 must be in [0, 1))");
    return RL_E_INVALID;
  }
  if (n == 0) return RL_OK;
  Marshal ms(m);
  const int ip = ms.add(particles, sizeof(float) * 3 * (size_t)n, false);
  const int iw = ms.add(weights, sizeof(double) * (size_t)n, false);
  const int io = ms.add(out_particles, sizeof(float) * 3 * (size_t)n, true);
  rc = ms.prepare();
  if (rc) return rc;
  if (ms.dev(ip) == ms.dev(io)) {
    set_error("rl_pf_resample: in-place resampling is not supported");
    return RL_E_INVALID;
  }
  rc = pf_resample(m, (const float*)ms.dev(ip), (const double*)ms.dev(iw), (float*)ms.dev(io), n, u0);
  if (rc) return rc;
  return ms.finish();
}

int rl_pf_motion_update(rl_method* m, float* particles, int n, float dx, float dy, float dtheta, const float* noise) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (n < 0 || (n > 0 && !particles)) {
    set_error("rl_pf_motion_update: bad arguments");
    return RL_E_INVALID;
  }
  if (n == 0) return RL_OK;
  Marshal ms(m);
  const int ip = ms.add_inout(particles, sizeof(float) * 3 * (size_t)n);
  const int in_ = noise ? ms.add(noise, sizeof(float) * 3 * (size_t)n, false) : -1;
  rc = ms.prepare();
  if (rc) return rc;
  rc = pf_motion(m, (float*)ms.dev(ip), n, dx, dy, dtheta, noise ? (const float*)ms.dev(in_) : nullptr);
  if (rc) return rc;
  return ms.finish();
}

// RangeMethod::calc_range_many_radial_optimized RangeLib.h:616-676 (loop constants :635-645 restated with the
// reference's types: float step, double index_offset narrowed to float, roundf).
int rl_calc_range_many_radial_optimized(rl_method* m, const float* ins, float* outs, int n, int num_rays,
                                        float min_angle, float max_angle) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (n < 0 || num_rays < 2 || !std::isfinite(min_angle) || !std::isfinite(max_angle) || !(max_angle != min_angle)) {
    set_error("calc_range_many_radial_optimized: need n >= 0, num_rays >= 2 and finite min_angle != max_angle");
    return RL_E_INVALID;
  }
  if (n == 0) return RL_OK;
  if (!ins || !outs) {
    set_error("null data pointer");
    return RL_E_INVALID;
  }
  if (m->radial_rays != num_rays || m->radial_min != min_angle || m->radial_max != max_angle) {
    const float step = (max_angle - min_angle) / (num_rays - 1);
    const int max_pair = (float)num_rays / 3.0;
    const float index_offset_float = (num_rays - 1.0) * RL_PI / (max_angle - min_angle);
    if (!(fabsf(index_offset_float) < 1e9f)) {
      set_error("calc_range_many_radial_optimized: angular span too small for this beam count");
      return RL_E_INVALID;
    }
    const int index_offset = roundf(index_offset_float);
    const int count = std::min(std::max(max_pair + 1, index_offset), num_rays);
    m->h_radial.resize(count);
    float angle = min_angle;
    for (int a = 0; a < count; ++a) {
      m->h_radial[a] = angle;
      angle += step;
    }
    if (m->radial_cap < count) {
      RL_CUDA(cudaStreamSynchronize(m->stream));
      cudaFree(m->d_radial);
      m->d_radial = nullptr;
      m->radial_cap = 0;
      RL_CUDA(cudaMalloc(&m->d_radial, sizeof(float) * count));
      m->radial_cap = count;
    }
    RL_CUDA(cudaMemcpyAsync(m->d_radial, m->h_radial.data(), sizeof(float) * count, cudaMemcpyHostToDevice, m->stream));
    m->radial_rays = num_rays;
    m->radial_min = min_angle;
    m->radial_max = max_angle;
    m->radial_count = count;
    m->radial_pair = max_pair;
    m->radial_offset = index_offset;
  }
  Marshal ms(m);
  const int i_ins = ms.add(ins, sizeof(float) * 3 * (size_t)n, false);
  const int i_out = ms.add_inout(outs, sizeof(float) * (size_t)n * num_rays);
  rc = ms.prepare();
  if (rc) return rc;
  rc = launch_radial(m, (const float*)ms.dev(i_ins), m->d_radial, (float*)ms.dev(i_out), n, num_rays, m->radial_count,
                     m->radial_pair, m->radial_offset);
  if (rc) return rc;
  return ms.finish();
}

int rl_calc_range_repeat_angles_eval_sensor_model_peers(rl_method* m, const float* ins, const float* angles,
                                                        const float* obs, double* const* peer_weights, int n_peers,
                                                        int64_t offset, int n, int M) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (n < 0 || M < 0 || n_peers < 1 || n_peers > RL_MAX_PEERS || !peer_weights || offset < 0) {
    set_error("rl_calc_range_repeat_angles_eval_sensor_model_peers: bad arguments");
    return RL_E_INVALID;
  }
  if (n == 0) return RL_OK;
  if (!ins || !angles || !obs) {
    set_error("null data pointer");
    return RL_E_INVALID;
  }
  if (!is_device_ptr(ins) || !is_device_ptr(angles) || !is_device_ptr(obs)) {
    set_error("the peer-store variant takes device pointers only");
    return RL_E_MIXED;
  }
  PeerOut po{};
  po.n = n_peers;
  po.offset = offset;
  for (int r = 0; r < n_peers; ++r) {
    if (!peer_weights[r]) {
      set_error("null peer buffer");
      return RL_E_INVALID;
    }
    po.ptr[r] = peer_weights[r];
  }
  if (M == 0) {
    set_error("num_angles must be > 0 for the peer-store variant");
    return RL_E_INVALID;
  }
  return launch_cast(m, MODE_FUSED, ins, angles, obs, nullptr, nullptr, n, M, &po);
}

int rl_method_peers_init(rl_method* m, double* const* weights0, double* const* weights1, int64_t* const* flags,
                         int n_peers, int rank) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (!weights0 || !weights1 || !flags || n_peers < 1 || n_peers > RL_MAX_PEERS || rank < 0 || rank >= n_peers) {
    set_error("rl_method_peers_init: bad arguments");
    return RL_E_INVALID;
  }
  if (!m->d_epoch) {
    RL_CUDA(cudaMalloc(&m->d_epoch, sizeof(long long)));
    RL_CUDA(cudaMalloc(&m->d_counter, sizeof(unsigned)));
  }
  RL_CUDA(cudaMemsetAsync(m->d_epoch, 0, sizeof(long long), m->stream));
  RL_CUDA(cudaMemsetAsync(m->d_counter, 0, sizeof(unsigned), m->stream));
  RL_CUDA(cudaStreamSynchronize(m->stream));
  PeerOut& po = m->peer_cfg;
  po = PeerOut{};
  po.n = n_peers;
  po.sig = 1;
  po.rank = rank;
  for (int r = 0; r < n_peers; ++r) {
    if (!weights0[r] || !weights1[r] || !flags[r]) {
      set_error("rl_method_peers_init: null peer pointer");
      return RL_E_INVALID;
    }
    po.ptr[r] = weights0[r];
    po.ptr1[r] = weights1[r];
    po.flags[r] = (long long*)flags[r];
  }
  po.epoch = m->d_epoch;
  po.counter = m->d_counter;
  m->host_epoch = 0;
  m->peers_ready = true;
  return RL_OK;
}

int rl_calc_range_repeat_angles_eval_sensor_model_signalled(rl_method* m, const float* ins, const float* angles,
                                                            const float* obs, int64_t offset, int n, int M,
                                                            int* buffer_index) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (!m->peers_ready) {
    set_error("rl_method_peers_init has not been called");
    return RL_E_STATE;
  }
  if (n <= 0 || M <= 0 || offset < 0 || !ins || !angles || !obs) {
    set_error("signalled fused call: bad arguments (every rank must launch with n > 0, num_angles > 0)");
    return RL_E_INVALID;
  }
  if (!is_device_ptr(ins) || !is_device_ptr(angles) || !is_device_ptr(obs)) {
    set_error("the signalled variant takes device pointers only");
    return RL_E_MIXED;
  }
  PeerOut po = m->peer_cfg;
  po.offset = offset;
  rc = launch_cast(m, MODE_FUSED, ins, angles, obs, nullptr, nullptr, n, M, &po);
  if (rc) return rc;
  m->host_epoch += 1;
  if (buffer_index) *buffer_index = (int)(m->host_epoch & 1);
  return RL_OK;
}

// The sharded particle-filter update through HOST pointers (one process per GPU): this rank's particles go to the
// device, the fused kernel computes their weights and stores them into every rank's gathered array over NVLink
// (signalled epilogue: the synchronisation is part of the same kernel), a wait kernel holds the stream until every
// rank's slice has arrived, and the whole gathered array comes back to the caller's buffer.  Blocking, like every
// host-pointer call.  With DEVICE pointers the same sequence is enqueued on the handle's stream and the gathered
// array is copied device-to-device without synchronising.
int rl_calc_range_repeat_angles_eval_sensor_model_sharded(rl_method* m, const float* ins, const float* angles,
                                                          const float* obs, double* weights_all, int64_t offset, int n,
                                                          int M, int64_t n_total) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (!m->peers_ready) {
    set_error("rl_method_peers_init has not been called");
    return RL_E_STATE;
  }
  if (n <= 0 || M <= 0 || offset < 0 || n_total < offset + n || !ins || !angles || !obs || !weights_all) {
    set_error("sharded fused call: bad arguments (every rank must call with n > 0, num_angles > 0)");
    return RL_E_INVALID;
  }
  const bool dev = is_device_ptr(ins);
  if (dev != (bool)is_device_ptr(angles) || dev != (bool)is_device_ptr(obs) || dev != (bool)is_device_ptr(weights_all)) {
    set_error("host and device pointers mixed in one call");
    return RL_E_MIXED;
  }
  const float *d_ins = ins, *d_ang = angles, *d_obs = obs;
  if (!dev) {
    const size_t b_ins = align256(sizeof(float) * 3 * (size_t)n), b_m = align256(sizeof(float) * (size_t)M);
    rc = ensure_stage(m, b_ins + 2 * b_m);
    if (rc) return rc;
    char* base = (char*)m->d_stage;
    RL_CUDA(cudaMemcpyAsync(base, ins, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, m->stream));
    RL_CUDA(cudaMemcpyAsync(base + b_ins, angles, sizeof(float) * (size_t)M, cudaMemcpyHostToDevice, m->stream));
    RL_CUDA(cudaMemcpyAsync(base + b_ins + b_m, obs, sizeof(float) * (size_t)M, cudaMemcpyHostToDevice, m->stream));
    d_ins = (const float*)base;
    d_ang = (const float*)(base + b_ins);
    d_obs = (const float*)(base + b_ins + b_m);
  }
  PeerOut po = m->peer_cfg;
  po.offset = offset;
  rc = launch_cast(m, MODE_FUSED, d_ins, d_ang, d_obs, nullptr, nullptr, n, M, &po);
  if (rc) return rc;
  m->host_epoch += 1;
  rc = launch_peers_wait(m);
  if (rc) return rc;
  const double* gathered = (m->host_epoch & 1) ? po.ptr1[po.rank] : po.ptr[po.rank];
  RL_CUDA(cudaMemcpyAsync(weights_all, gathered, sizeof(double) * (size_t)n_total,
                          dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, m->stream));
  if (!dev) RL_CUDA(cudaStreamSynchronize(m->stream));
  return RL_OK;
}

int rl_method_peers_wait(rl_method* m) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (!m->peers_ready) {
    set_error("rl_method_peers_init has not been called");
    return RL_E_STATE;
  }
  return launch_peers_wait(m);
}

int rl_debug_get_occ(rl_method* m, uint8_t* out) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (!out) return RL_E_INVALID;
  const size_t n = (size_t)m->W * m->H;
  if (n) RL_CUDA(cudaMemcpyAsync(out, m->d_occ, n, cudaMemcpyDeviceToHost, m->stream));
  RL_CUDA(cudaStreamSynchronize(m->stream));
  return RL_OK;
}

int rl_debug_get_dt(rl_method* m, float* out) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (m->kind != RL_RM || !m->d_dt || !out) {
    set_error("rl_debug_get_dt: not an RM method");
    return RL_E_STATE;
  }
  RL_CUDA(cudaMemcpyAsync(out, m->d_dt, sizeof(float) * m->dt_elems(), cudaMemcpyDeviceToHost, m->stream));
  RL_CUDA(cudaStreamSynchronize(m->stream));
  return RL_OK;
}

int rl_debug_cddt_dims(rl_method* m, int64_t* n_bins, int64_t* n_values, int* widths, float* translations) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (m->kind != RL_CDDT && m->kind != RL_PCDDT) {
    set_error("not a CDDT method");
    return RL_E_STATE;
  }
  if (n_bins) *n_bins = m->nbins;
  if (n_values) *n_values = m->nvalues;
  if (widths) memcpy(widths, m->h_widths.data(), sizeof(int) * m->td);
  if (translations) memcpy(translations, m->h_trans.data(), sizeof(float) * m->td);
  return RL_OK;
}

int rl_debug_cddt_dump(rl_method* m, int64_t* offsets, float* values) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if ((m->kind != RL_CDDT && m->kind != RL_PCDDT) || !offsets || !values) {
    set_error("not a CDDT method");
    return RL_E_STATE;
  }
  RL_CUDA(cudaMemcpyAsync(offsets, m->d_offsets, sizeof(int64_t) * ((size_t)m->nbins + 1), cudaMemcpyDeviceToHost, m->stream));
  if (m->nvalues)
    RL_CUDA(cudaMemcpyAsync(values, m->d_values, sizeof(float) * (size_t)m->nvalues, cudaMemcpyDeviceToHost, m->stream));
  RL_CUDA(cudaStreamSynchronize(m->stream));
  return RL_OK;
}

int rl_debug_glt_dump(rl_method* m, uint16_t* out) {
  DeviceGuard dg;
  int rc = dg.bind(m);
  if (rc) return rc;
  if (m->kind != RL_GLT || !m->d_glt || !out) {
    set_error("rl_debug_glt_dump: not a GiantLUT method");
    return RL_E_STATE;
  }
  RL_CUDA(cudaMemcpyAsync(out, m->d_glt, sizeof(uint16_t) * (size_t)m->W * m->H * m->td, cudaMemcpyDeviceToHost, m->stream));
  RL_CUDA(cudaStreamSynchronize(m->stream));
  return RL_OK;
}

int rl_debug_sincosf(const float* x, float* s, float* c, int n) {
  if (n <= 0) return RL_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("rangelib_b200 needs a CUDA device");
    return RL_E_NO_DEVICE;
  }
  float *dx = nullptr, *ds = nullptr, *dc = nullptr;
  RL_CUDA(cudaMalloc(&dx, sizeof(float) * n));
  RL_CUDA(cudaMalloc(&ds, sizeof(float) * n));
  RL_CUDA(cudaMalloc(&dc, sizeof(float) * n));
  RL_CUDA(cudaMemcpy(dx, x, sizeof(float) * n, cudaMemcpyHostToDevice));
  int rc = launch_sincosf(dx, ds, dc, n, 0);
  if (!rc) {
    RL_CUDA(cudaMemcpy(s, ds, sizeof(float) * n, cudaMemcpyDeviceToHost));
    RL_CUDA(cudaMemcpy(c, dc, sizeof(float) * n, cudaMemcpyDeviceToHost));
  }
  cudaFree(dx);
  cudaFree(ds);
  cudaFree(dc);
  return rc;
}

}  // extern "C"
