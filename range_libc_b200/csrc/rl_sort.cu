// Spatial ordering of large particle sets for the fused sensor-model call.
//
// When the structure a ray reads is larger than L2 (RM on an 8192^2 map: a 268 MB distance transform; the
// GiantLUT table), what bounds a big update is HBM sector traffic: the particles of a global-localisation
// cloud arrive in random order, so the CTAs resident at any moment read all over the map.  Processing the
// particles in the order of the 64 x 64-cell tile they stand in (Morton order of the tiles) makes the resident
// CTAs work on one neighbourhood at a time, which then stays in L2.  Only the ORDER of processing changes: every
// particle's weight is computed by the same arithmetic and stored at its own index.
#include <cub/device/device_radix_sort.cuh>

#include "rl_internal.cuh"
#include "rl_math.cuh"

#ifndef RL_SORT_TILE_SHIFT
#define RL_SORT_TILE_SHIFT 6  // tiles of 64 x 64 cells
#endif

namespace rl {

__device__ __forceinline__ unsigned spread_bits(unsigned v) {  // 0000abcd -> 0a0b0c0d (up to 16 bits)
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}

// key of a particle: Morton code of the tile holding the cell calc_range starts from -- the pose goes through
// the same world -> grid transform as the cast (RangeLib.h:464-475: calc_range(y, x, theta))
__global__ void tile_key_kernel(WorldXform xf, const float* __restrict__ ins, int n, int W, int H, unsigned* keys,
                                int* idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xw = __ldg(ins + 3 * (size_t)i), yw = __ldg(ins + 3 * (size_t)i + 1);
  float xx = fmul(fsub(xw, xf.ox), xf.inv_scale);
  float yy = fmul(fsub(yw, xf.oy), xf.inv_scale);
  const float tmp = xx;
  xx = fsub(fmul(xf.cos_a, xx), fmul(xf.sin_a, yy));
  yy = fadd(fmul(xf.sin_a, tmp), fmul(xf.cos_a, yy));
  // first argument of calc_range (yy) runs along the map's x, the second (xx) along its y
  int cx = (yy >= 0.0f && yy < (float)W) ? __float2int_rz(yy) : 0;
  int cy = (xx >= 0.0f && xx < (float)H) ? __float2int_rz(xx) : 0;
  keys[i] = spread_bits((unsigned)cx >> RL_SORT_TILE_SHIFT) | (spread_bits((unsigned)cy >> RL_SORT_TILE_SHIFT) << 1);
  idx[i] = i;
}

static int ensure_sort_buffers(rl_method* m, int n) {
  if (n <= m->sort_cap) return RL_OK;
  cudaFree(m->d_sort_keys);
  cudaFree(m->d_sort_idx);
  cudaFree(m->d_sort_tmp);
  m->d_sort_keys = nullptr;
  m->d_sort_idx = nullptr;
  m->d_sort_tmp = nullptr;
  m->sort_cap = 0;
  const int cap = n + n / 4;
  RL_CUDA(cudaMalloc(&m->d_sort_keys, sizeof(unsigned) * 2 * (size_t)cap));
  RL_CUDA(cudaMalloc(&m->d_sort_idx, sizeof(int) * 2 * (size_t)cap));
  cub::DoubleBuffer<unsigned> k(m->d_sort_keys, m->d_sort_keys + cap);
  cub::DoubleBuffer<int> v(m->d_sort_idx, m->d_sort_idx + cap);
  size_t tb = 0;
  RL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k, v, cap, 0, 32, m->stream));
  RL_CUDA(cudaMalloc(&m->d_sort_tmp, tb ? tb : 1));
  m->sort_tmp_bytes = tb;
  m->sort_cap = cap;
  return RL_OK;
}

// Leaves in *perm a device array of n particle indices in tile order (owned by the handle, valid until the next call).
int spatial_order(rl_method* m, const float* d_ins, int n, const int** perm) {
  int rc = ensure_sort_buffers(m, n);
  if (rc) return rc;
  const int cap = m->sort_cap;
  tile_key_kernel<<<(n + 255) / 256, 256, 0, m->stream>>>(m->xf, d_ins, n, m->W, m->H, m->d_sort_keys, m->d_sort_idx);
  count_launch();
  RL_CHECK_LAUNCH();
  int bits = 2;
  while (bits < 32 && (1u << (bits / 2)) <
                          (unsigned)((max(m->W, m->H) + (1 << RL_SORT_TILE_SHIFT) - 1) >> RL_SORT_TILE_SHIFT))
    bits += 2;
  cub::DoubleBuffer<unsigned> k(m->d_sort_keys, m->d_sort_keys + cap);
  cub::DoubleBuffer<int> v(m->d_sort_idx, m->d_sort_idx + cap);
  size_t tb = m->sort_tmp_bytes;
  RL_CUDA(cub::DeviceRadixSort::SortPairs(m->d_sort_tmp, tb, k, v, n, 0, bits, m->stream));
  count_launch(2);
  *perm = v.Current();
  return RL_OK;
}

void sort_free(rl_method* m) {
  cudaFree(m->d_sort_keys);
  cudaFree(m->d_sort_idx);
  cudaFree(m->d_sort_tmp);
}

}  // namespace rl
