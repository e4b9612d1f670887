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

// Distance-transform residency layout.  The grid is cut into tiles of 8 (x) by 4 (y) cells = 32
// floats = one 128-byte L1/L2 line; inside a tile the cells are bit-interleaved so that each
// 32-byte sector (the unit the L1 actually fills on a miss) covers a 4x2 block.  A ray that
// creeps along a wall one pixel at a time therefore stays inside one sector / line for
// several consecutive sphere-tracing steps whatever its direction, instead of touching a new
// line on every step as it does in the reference's x-major vector<vector<float>>.
#define RL_DT_TILE_X_SHIFT 3
#define RL_DT_TILE_Y_SHIFT 2
__host__ __device__ __forceinline__ unsigned dt_tiled_index(int px, int py, int tiles_y) {
  const unsigned ux = (unsigned)px, uy = (unsigned)py;
  const unsigned tile = (ux >> RL_DT_TILE_X_SHIFT) * (unsigned)tiles_y + (uy >> RL_DT_TILE_Y_SHIFT);
  const unsigned low = (ux & 3u) | ((uy & 1u) << 2) | ((ux & 4u) << 1) | ((uy & 2u) << 3);
  return (tile << 5) | low;
}

// read-only view of the resident structures handed to kernels
struct MapView {
  int W, H;
  const uint8_t* occ;      // x-major bytes occ[x*H+y]
  const uint32_t* bits_y;  // bit grid packed along y: word (x, y>>5), bit y&31; row stride wpy words
  int wpy;
  const uint32_t* bits_x;  // the same grid packed along x: word (y, x>>5), bit x&31; row stride wpx words
  int wpx;
  const float* dt;         // tiled float distance transform (RM), see dt_tiled_index
  int dt_tiles_y;
  int coop_threshold;      // RM: warps with at most this many live rays finish them cooperatively (0 = off)
};

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
};

struct SensorView {
  const double* table;  // K*K row-major table[r*K + d]
  int K;
};

enum Mode { MODE_GRID = 0, MODE_WORLD = 1, MODE_ANGLES = 2, MODE_FUSED = 3 };

// Multi-GPU epilogue of the fused kernel: instead of one local array, the per-particle weights are
// stored straight into the weight buffer of every peer GPU (peer-mapped device pointers over
// NVLink), at this rank's offset -- the all-gather is the kernel's own store phase.
#define RL_MAX_PEERS 16
struct PeerOut {
  double* ptr[RL_MAX_PEERS];
  int n;             // 0: write only the local `weights` array
  long long offset;  // first particle of this rank inside the gathered array
};

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
  cudaEvent_t ev = nullptr;

  // occupancy on device
  uint8_t* d_occ = nullptr;
  uint32_t* d_bits_y = nullptr;
  int wpy = 0;
  uint32_t* d_bits_x = nullptr;
  int wpx = 0;
  // RM
  float* d_dt = nullptr;
  // CDDT
  int* d_widths = nullptr;
  float *d_trans = nullptr, *d_cosv = nullptr, *d_sinv = nullptr;
  int64_t *d_slice0 = nullptr, *d_offsets = nullptr;
  float* d_values = nullptr;
  int64_t nbins = 0, nvalues = 0;
  std::vector<int> h_widths;
  std::vector<float> h_trans, h_cosv, h_sinv;
  std::vector<int64_t> h_slice0;
  float td_div_2pi = 0.f, twopi_div_td = 0.f;
  bool pruned = false;
  // sensor model
  double* d_table = nullptr;
  int K = 0;
  // staging for host-pointer calls
  void* d_stage = nullptr;
  size_t d_stage_bytes = 0;
  void* h_stage = nullptr;      // pinned + mapped
  void* h_stage_dev = nullptr;  // its device alias (zero-copy)
  size_t h_stage_bytes = 0;

  int dt_tiles_x() const { return (W + 7) >> 3; }
  int dt_tiles_y() const { return (H + 3) >> 2; }
  size_t dt_elems() const { return (size_t)dt_tiles_x() * dt_tiles_y() * 32; }
  int coop_threshold = 3;
  int persist = 1;  // RM large batches: 0 one ray per thread, 1 persistent warps with lane re-queuing
  rl::MapView map_view() const { return rl::MapView{W, H, d_occ, d_bits_y, wpy, d_bits_x, wpx, d_dt, dt_tiles_y(), coop_threshold}; }
  rl::CddtView cddt_view() const {
    return rl::CddtView{td, d_widths, d_trans, d_cosv, d_sinv, d_slice0, d_offsets, d_values, td_div_2pi, twopi_div_td};
  }
  rl::SensorView sensor_view() const { return rl::SensorView{d_table, K}; }
};

namespace rl {
// rl_edt.cu
int build_distance_transform(rl_method* m);
// rl_occ.cu
int upload_occupancy(rl_method* m, const rl_map* map);
int apply_patch(rl_method* m, const uint8_t* d_patch, int x0, int y0, int w, int h);
// rl_cddt.cu
int cddt_build(rl_method* m);
int cddt_prune(rl_method* m, float max_range);
void cddt_free(rl_method* m);
// rl_cast.cu -- the batched query kernels (all kinds, all modes)
int launch_cast(rl_method* m, int mode, const float* d_ins, const float* d_angles, const float* d_obs, float* d_outs,
                double* d_weights, int n, int num_angles, const PeerOut* peers = nullptr);
int launch_eval_sensor(rl_method* m, const float* d_obs, const float* d_ranges, double* d_outs, int m_rays, int n);
int launch_sincosf(const float* d_x, float* d_s, float* d_c, int n, cudaStream_t st);
}  // namespace rl
