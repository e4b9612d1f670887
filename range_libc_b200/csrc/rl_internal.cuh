// Internal declarations shared by the translation units of librangelib_b200.so.
// The whole library is compiled with --fmad=false: every float operation in the parity-critical
// paths rounds exactly like the reference built without FP contraction (the pinned STRICT
// oracle, see DESIGN.md "Numerics").  Where a fused multiply-add is wanted it is written fma().
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "rangelib_b200.h"

#define RL_EPSILON 0.00001             // RangeLib.h:60
#define RL_M_2PI 6.28318530718         // RangeLib.h:61
#define RL_BINARY_SEARCH_THRESHOLD 64  // RangeLib.h:62
#define RL_PI 3.14159265358979323846   // M_PI

namespace rl {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);

#define RL_CUDA(expr)                                                    \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) return rl::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define RL_CHECK_LAUNCH()                                                \
  do {                                                                   \
    cudaError_t _e = cudaGetLastError();                                 \
    if (_e != cudaSuccess) return rl::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

// world -> grid conversion constants, RangeLib.h:442-450
struct WorldXform {
  float inv_scale, scale, ox, oy, sin_a, cos_a, rot;
};

// Distance-transform residency layout: plain x-major floats, dt[x*H + y], the order of the reference's
// DistanceTransform::grid[x][y] (RangeLib.h:328).  A 128-byte-tile layout with bit-interleaved 4x2
// sectors was built and measured in round 1: it costs ~8 more integer instructions per
// sphere-tracing step and was 4-10 % SLOWER on every RM workload (the kernels are
// instruction-issue / latency bound, not L1-hit bound; profiles/ab_dt_layout_r01.log), so it was removed.
__host__ __device__ __forceinline__ unsigned dt_index(int px, int py, int H) {
  return (unsigned)px * (unsigned)H + (unsigned)py;
}

// read-only view of the resident structures handed to kernels
struct MapView {
  int W, H;
  const uint8_t* occ;      // x-major bytes occ[x*H+y]
  // occupancy bits in 8x8-cell tiles, one 64-bit word per tile: tile (x>>3, y>>3) at index
  // (x>>3)*tiles8_y + (y>>3), bit (x&7)*8 + (y&7).  A Bresenham walk crosses a tile in ~8-11 steps
  // whatever its direction, so one load serves that many cell tests.
  const unsigned long long* bits_t;
  int tiles8_y;
  const float* dt;         // float distance transform (RM), x-major, see dt_index
  int coop_threshold;      // RM: a CTA with at most this many live rays finishes them cooperatively, one per warp (0 = off)
  // GiantLUTCast (RangeLib.h:1772-1904): uint16 range per (x, y, theta bin), glt[(x*H + y)*td + i]
  const uint16_t* glt;
  unsigned glt_td;
  float glt_td_div_2pi, glt_twopi_div_td;  // :1783-1784
  float glt_max_div_limits, glt_limits_div_max;  // :1785-1786
  int block_burst_pairs = 6;  // rm_march_block: own-ray steps between two CTA-wide counts = 2 * this
};

// Per-bin record of the L2-resident query index (rl_cddt.cu: cddt_index_build): everything a query needs to know
// about its bin before it touches the zero points themselves, in one 16-byte load.
struct __align__(16) CddtBinMeta {
  unsigned off;   // first zero point of the bin in values[]
  unsigned size;  // number of zero points
  float first, last;
};
#define RL_CDDT_BLOCK 16  // zero points per skip entry: one aligned 64-byte block of values[]
// values[] is allocated with this many floats so that the last aligned block and the pad element exist
inline size_t cddt_values_alloc(int64_t nvalues) {
  return (((size_t)nvalues + RL_CDDT_BLOCK) / RL_CDDT_BLOCK + 1) * RL_CDDT_BLOCK;
}
#ifdef __CUDACC__
// 16-bit position code of a zero point inside its bin [first, last]: MONOTONE non-decreasing in v (every operation
// is: subtraction of a constant, multiplication by a non-negative constant, clamp, truncation), so for two values of
// one bin  code(a) < code(b)  implies  a < b;  equal codes decide nothing and the caller compares the values.
__device__ __forceinline__ float cddt_code_scale(float first, float last) {
  return last > first ? __fdiv_rn(65535.0f, __fsub_rn(last, first)) : 0.0f;
}
__device__ __forceinline__ unsigned cddt_code(float v, float first, float scale) {
  const float t = fminf(fmaxf(__fmul_rn(__fsub_rn(v, first), scale), 0.0f), 65535.0f);
  return (unsigned)__float2int_rz(t);
}
#endif

struct CddtView {
  unsigned td;
  const int* widths;        // [td]
  const float* trans;       // [td]
  const float* cosv;        // [td] host-tabulated libm cosf(discrete angle)
  const float* sinv;        // [td]
  const int64_t* slice0;    // [td+1]
  const int64_t* offsets;   // [nbins+1]
  const float* values;      // [nvalues + pad]
  float td_div_2pi;         // (float)(td / M_2PI)           RangeLib.h:976
  float twopi_div_td;       // (float)(M_2PI / (float)td)    RangeLib.h:977
  const CddtBinMeta* meta;  // [nbins] query index, or nullptr: search values[] directly
  const uint16_t* skip;     // skip[k] = cddt_code of values[16 k] relative to the bin that holds it
};

struct SensorView {
  const double* table;  // K*K row-major table[r*K + d]
  int K;
};

enum Mode { MODE_GRID = 0, MODE_WORLD = 1, MODE_ANGLES = 2, MODE_FUSED = 3, MODE_GLT_BUILD = 4 };

// Multi-GPU epilogue of the fused kernel: instead of one local array, the per-particle weights are
// stored straight into the weight buffer of every peer GPU (peer-mapped device pointers over
// NVLink), at this rank's offset -- the all-gather is the kernel's own store phase.
#define RL_MAX_PEERS 16
struct PeerOut {
  double* ptr[RL_MAX_PEERS];
  int n;             // 0: write only the local `weights` array
  long long offset;  // first particle of this rank inside the gathered array
  // Signalled mode (sig != 0): no separate barrier.  Every launch is one "epoch" e = *epoch + 1.  The
  // kernel first waits until every rank has finished epoch e-1 (flags written by the peers into OUR
  // flag array), stores its weights into buffer e & 1 of every rank (ptr = buffer 0, ptr1 = buffer 1),
  // and the last CTA to finish publishes flags[r][rank] = e on every rank r.  Double buffering makes
  // the start-of-kernel wait sufficient for reuse safety.
  int sig;
  int rank;
  double* ptr1[RL_MAX_PEERS];
  long long* flags[RL_MAX_PEERS];  // flags[r] = rank r's flag array (n entries), peer mapped
  long long* epoch;                // local: epochs completed by this rank
  unsigned* counter;               // local: CTAs finished in the current launch
};

// Small fused calls through HOST pointers carry the beam angles and the observation inside the launch (kernel
// parameters) instead of a host-to-device copy.
#define RL_PARAM_BEAMS 256
struct BeamParams {
  float angles[RL_PARAM_BEAMS];
  float obs[RL_PARAM_BEAMS];
};
struct NoBeamParams {};

}  // namespace rl

struct rl_map {
  int W = 0, H = 0;
  std::vector<uint8_t> occ;  // x-major
  float scale = 1.f, angle = 0.f, ox = 0.f, oy = 0.f, sin_a = 0.f, cos_a = 1.f;
};

struct rl_method {
  int kind = 0;
  int device = 0;
  int W = 0, H = 0;
  float max_range = 0.f;
  unsigned td = 0;
  rl::WorldXform xf{};
  cudaStream_t own_stream = nullptr, stream = nullptr;

  // occupancy on device
  uint8_t* d_occ = nullptr;
  unsigned long long* d_bits_t = nullptr;
  int tiles8_x() const { return (W + 7) >> 3; }
  int tiles8_y() const { return (H + 7) >> 3; }
  // RM
  float* d_dt = nullptr;
  // GiantLUT
  uint16_t* d_glt = nullptr;
  // CDDT
  int* d_widths = nullptr;
  float *d_trans = nullptr, *d_cosv = nullptr, *d_sinv = nullptr;
  int64_t *d_slice0 = nullptr, *d_offsets = nullptr;
  float* d_values = nullptr;
  int64_t nbins = 0, nvalues = 0;
  rl::CddtBinMeta* d_meta = nullptr;  // query index over the table (rebuilt whenever the table changes)
  uint16_t* d_skip = nullptr;
  int64_t nskip = 0;
  bool use_index = false;  // queries go through the index (tables larger than L2; RL_CDDT_INDEX=0|1 overrides)
  std::vector<int> h_widths;
  std::vector<float> h_trans, h_cosv, h_sinv;
  std::vector<int64_t> h_slice0;
  float td_div_2pi = 0.f, twopi_div_td = 0.f;
  bool pruned = false;
  // sensor model
  double* d_table = nullptr;
  int K = 0;
  // multi-GPU signalled all-gather state (rl_method_peers_init)
  rl::PeerOut peer_cfg{};
  bool peers_ready = false;
  long long* d_epoch = nullptr;
  unsigned* d_counter = nullptr;
  unsigned long long* d_work = nullptr;  // ray cursor of the persistent kernels (zeroed before each launch)
  long long host_epoch = 0;
  // staging for host-pointer calls
  void* d_stage = nullptr;
  size_t d_stage_bytes = 0;
  void* h_stage = nullptr;      // pinned + mapped
  void* h_stage_dev = nullptr;  // its device alias (zero-copy)
  size_t h_stage_bytes = 0;
  // spatial ordering of large particle sets (rl_sort.cu)
  unsigned* d_sort_keys = nullptr;
  int* d_sort_idx = nullptr;
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  int sort_cap = 0;
  // deep fused updates as cast + eval (rl_cast.cu, launch_fused_twostep): ranges of one chunk of particles, and the
  // chunk's poses in processing order when the cloud is tile-ordered
  float* d_ts_ranges = nullptr;
  size_t ts_ranges_bytes = 0;
  float* d_ts_poses = nullptr;
  size_t ts_poses_bytes = 0;
  // particle-filter steps (rl_pf.cu): reduction result, fixed-point weights / prefix sums, cub scratch
  void* d_pf = nullptr;
  size_t pf_bytes = 0;
  // calc_range_many_radial_optimized: beam-angle table of the last call
  float* d_radial = nullptr;
  int radial_cap = 0, radial_rays = -1, radial_count = 0;
  float radial_min = 0.f, radial_max = 0.f;
  int radial_pair = 0, radial_offset = 0;
  std::vector<float> h_radial;

  size_t dt_elems() const { return (size_t)W * H; }
  int coop_threshold = 8;
  int spatial_sort = 1;  // big clouds on > L2 structures are processed in tile order (rl_sort.cu); 0 also switches the CDDT query index off
  int persist = 1;  // RM large batches: 0 one ray per thread, 1 persistent warps with lane re-queuing
  rl::MapView map_view() const {
    rl::MapView v{W, H, d_occ, d_bits_t, tiles8_y(), d_dt, coop_threshold, d_glt, td, 0.f, 0.f, 0.f, 0.f};
    if (kind == RL_GLT && td) {
      v.glt_td_div_2pi = (float)((double)td / RL_M_2PI);
      v.glt_twopi_div_td = (float)(RL_M_2PI / (double)((float)td));
      v.glt_max_div_limits = max_range / (float)65535;
      v.glt_limits_div_max = (float)65535 / max_range;
    }
    return v;
  }
  rl::CddtView cddt_view() const {
    return rl::CddtView{td, d_widths, d_trans, d_cosv, d_sinv, d_slice0, d_offsets, d_values, td_div_2pi, twopi_div_td,
                        (use_index && spatial_sort) ? d_meta : nullptr, d_skip};
  }
  rl::SensorView sensor_view() const { return rl::SensorView{d_table, K}; }
};

namespace rl {
// rl_edt.cu
int build_distance_transform(rl_method* m);
// rl_occ.cu
int upload_occupancy(rl_method* m, const rl_map* map);
int apply_patch(rl_method* m, const uint8_t* d_patch, int x0, int y0, int w, int h);
int ingest_occupancy_grid(rl_method* m, const int8_t* d_data);
int ingest_rgba(rl_method* m, const uint8_t* d_rgba, float threshold);
int apply_patch_batch(rl_method* m, const uint8_t* d_patches, const int* d_rects, const long long* d_offsets, int n);
// rl_cddt.cu
int cddt_build(rl_method* m);
int cddt_prune(rl_method* m, float max_range);
void cddt_free(rl_method* m);
int cddt_save(rl_method* m, const char* path);
int cddt_load(rl_method* m, const char* path);
int cddt_index_build(rl_method* m, bool force_on = false);
// rl_sort.cu
int spatial_order(rl_method* m, const float* d_ins, int n, const int** d_perm);
void sort_free(rl_method* m);
// rl_pf.cu (device pointers)
int pf_normalize(rl_method* m, double* d_w, int n, double inv_squash, double* h_sum);
int pf_resample(rl_method* m, const float* d_particles, const double* d_w, float* d_out, int n, double u0);
int pf_motion(rl_method* m, float* d_particles, int n, float dx, float dy, float dth, const float* d_noise);
void pf_free(rl_method* m);
// rl_cast.cu -- the batched query kernels (all kinds, all modes)
int launch_cast(rl_method* m, int mode, const float* d_ins, const float* d_angles, const float* d_obs, float* d_outs,
                double* d_weights, int n, int num_angles, const PeerOut* peers = nullptr);
int launch_fused_beam_params(rl_method* m, const float* d_ins, const BeamParams& beams, double* d_weights, int n,
                             int num_angles);
int launch_radial(rl_method* m, const float* d_ins, const float* d_beam_angles, float* d_outs, int n, int num_rays,
                  int count, int max_pair, int index_offset);
int launch_eval_sensor(rl_method* m, const float* d_obs, const float* d_ranges, double* d_outs, int m_rays, int n);
int launch_sincosf(const float* d_x, float* d_s, float* d_c, int n, cudaStream_t st);
int launch_peers_wait(rl_method* m);
int glt_build(rl_method* m);
}  // namespace rl
