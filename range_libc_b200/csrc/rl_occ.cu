// Occupancy grid residency: byte grid (x-major, the reference's OMap::grid[x][y] order,
// RangeLib.h:126) plus a bit-packed copy (32 cells of one x-column per word) that the BL walk
// and the CDDT/BL "standing on an obstacle" test read.  Dynamic maps (BASELINE config 4)
// patch both on the device.
#include "rl_internal.cuh"

namespace rl {

// one thread per output word
__global__ void pack_bits_kernel(const uint8_t* __restrict__ occ, uint32_t* __restrict__ bits, int W, int H, int wpy,
                                 int x_begin, int x_end, int word_begin, int word_end) {
  const int nw = word_end - word_begin;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(x_end - x_begin) * nw;
  if (idx >= total) return;
  const int x = x_begin + (int)(idx / nw);
  const int wd = word_begin + (int)(idx % nw);
  const int y0 = wd << 5;
  uint32_t v = 0;
  const uint8_t* col = occ + (size_t)x * H;
#pragma unroll 4
  for (int b = 0; b < 32; ++b) {
    int y = y0 + b;
    if (y < H && col[y]) v |= (1u << b);
  }
  bits[(size_t)x * wpy + wd] = v;
}

// x-packed copy: one thread per output word (y, x>>5)
__global__ void pack_bits_x_kernel(const uint8_t* __restrict__ occ, uint32_t* __restrict__ bits, int W, int H, int wpx,
                                   int y_begin, int y_end, int word_begin, int word_end) {
  const int nw = word_end - word_begin;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(y_end - y_begin) * nw;
  if (idx >= total) return;
  // consecutive threads take consecutive y so the byte reads of a warp are contiguous
  const int ny = y_end - y_begin;
  const int y = y_begin + (int)(idx % ny);
  const int wd = word_begin + (int)(idx / ny);
  const int x0 = wd << 5;
  uint32_t v = 0;
#pragma unroll 4
  for (int b = 0; b < 32; ++b) {
    int x = x0 + b;
    if (x < W && occ[(size_t)x * H + y]) v |= (1u << b);
  }
  bits[(size_t)y * wpx + wd] = v;
}

__global__ void patch_kernel(uint8_t* __restrict__ occ, const uint8_t* __restrict__ patch, int H, int x0, int y0, int w,
                             int h) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w * h) return;
  int px = idx / h, py = idx - px * h;
  occ[(size_t)(x0 + px) * H + (y0 + py)] = patch[idx] ? 1 : 0;
}

// ---- batched patches (dynamic maps, BASELINE config 4): one CTA per patch, three phases in one launch ----
// rects[4*p .. 4*p+3] = x0, y0, w, h; patch p's bytes start at offsets[p] in `patches` (x-major inside the patch)
__global__ void patch_batch_kernel(uint8_t* __restrict__ occ, uint32_t* __restrict__ bits_y, uint32_t* __restrict__ bits_x,
                                   const uint8_t* __restrict__ patches, const int* __restrict__ rects,
                                   const long long* __restrict__ offsets, int W, int H, int wpy, int wpx) {
  const int p = blockIdx.x;
  const int x0 = rects[4 * p], y0 = rects[4 * p + 1], w = rects[4 * p + 2], h = rects[4 * p + 3];
  const uint8_t* src = patches + offsets[p];
  for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
    const int px = i / h, py = i - px * h;
    occ[(size_t)(x0 + px) * H + (y0 + py)] = src[i] ? 1 : 0;
  }
  __syncthreads();  // the words below are rebuilt from occ; patches of one batch must not overlap
  const int wb = y0 >> 5, we = ((y0 + h - 1) >> 5) + 1;
  for (int i = threadIdx.x; i < w * (we - wb); i += blockDim.x) {
    const int x = x0 + i / (we - wb), wd = wb + i % (we - wb);
    uint32_t v = 0;
    for (int b = 0; b < 32; ++b) {
      const int y = (wd << 5) + b;
      if (y < H && occ[(size_t)x * H + y]) v |= (1u << b);
    }
    bits_y[(size_t)x * wpy + wd] = v;
  }
  const int xb = x0 >> 5, xe = ((x0 + w - 1) >> 5) + 1;
  for (int i = threadIdx.x; i < h * (xe - xb); i += blockDim.x) {
    const int y = y0 + i % h, wd = xb + i / h;
    uint32_t v = 0;
    for (int b = 0; b < 32; ++b) {
      const int x = (wd << 5) + b;
      if (x < W && occ[(size_t)x * H + y]) v |= (1u << b);
    }
    bits_x[(size_t)y * wpx + wd] = v;
  }
}

int apply_patch_batch(rl_method* m, const uint8_t* d_patches, const int* d_rects, const long long* d_offsets, int n) {
  if (n <= 0) return RL_OK;
  patch_batch_kernel<<<n, 256, 0, m->stream>>>(m->d_occ, m->d_bits_y, m->d_bits_x, d_patches, d_rects, d_offsets, m->W,
                                               m->H, m->wpy, m->wpx);
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

int upload_occupancy(rl_method* m, const rl_map* map) {
  const size_t n = (size_t)m->W * m->H;
  m->wpy = (m->H + 31) / 32;
  RL_CUDA(cudaMalloc(&m->d_occ, n ? n : 1));
  m->wpx = (m->W + 31) / 32;
  RL_CUDA(cudaMalloc(&m->d_bits_y, sizeof(uint32_t) * (size_t)m->W * m->wpy + 4));
  RL_CUDA(cudaMalloc(&m->d_bits_x, sizeof(uint32_t) * (size_t)m->H * m->wpx + 4));
  if (n) RL_CUDA(cudaMemcpyAsync(m->d_occ, map->occ.data(), n, cudaMemcpyHostToDevice, m->stream));
  const long long total = (long long)m->W * m->wpy;
  if (total > 0) {
    pack_bits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, m->stream>>>(m->d_occ, m->d_bits_y, m->W, m->H, m->wpy,
                                                                            0, m->W, 0, m->wpy);
    count_launch();
    RL_CHECK_LAUNCH();
  }
  const long long total_x = (long long)m->H * m->wpx;
  if (total_x > 0) {
    pack_bits_x_kernel<<<(unsigned)((total_x + 255) / 256), 256, 0, m->stream>>>(m->d_occ, m->d_bits_x, m->W, m->H,
                                                                               m->wpx, 0, m->H, 0, m->wpx);
    count_launch();
    RL_CHECK_LAUNCH();
  }
  return RL_OK;
}

int apply_patch(rl_method* m, const uint8_t* d_patch, int x0, int y0, int w, int h) {
  patch_kernel<<<(w * h + 255) / 256, 256, 0, m->stream>>>(m->d_occ, d_patch, m->H, x0, y0, w, h);
  count_launch();
  RL_CHECK_LAUNCH();
  const int wb = y0 >> 5, we = ((y0 + h - 1) >> 5) + 1;
  const long long total = (long long)w * (we - wb);
  pack_bits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, m->stream>>>(m->d_occ, m->d_bits_y, m->W, m->H, m->wpy,
                                                                          x0, x0 + w, wb, we);
  count_launch();
  RL_CHECK_LAUNCH();
  const int xb = x0 >> 5, xe = ((x0 + w - 1) >> 5) + 1;
  const long long total_x = (long long)h * (xe - xb);
  pack_bits_x_kernel<<<(unsigned)((total_x + 255) / 256), 256, 0, m->stream>>>(m->d_occ, m->d_bits_x, m->W, m->H, m->wpx,
                                                                             y0, y0 + h, xb, xe);
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

}  // namespace rl
