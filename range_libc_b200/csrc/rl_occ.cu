// Occupancy grid residency: byte grid (x-major, the reference's OMap::grid[x][y] order,
// RangeLib.h:126) plus a bit-packed copy in 8x8-cell tiles (one 64-bit word per tile) that the BL
// walk and the CDDT/BL "standing on an obstacle" test read.  Dynamic maps (BASELINE config 4)
// patch both on the device.
#include <algorithm>

#include "rl_internal.cuh"

namespace rl {

// one thread per 8x8 tile in the tile range [tx0, tx1) x [ty0, ty1)
__global__ void pack_tiles_kernel(const uint8_t* __restrict__ occ, unsigned long long* __restrict__ bits, int W, int H,
                                  int tiles_y, int tx0, int tx1, int ty0, int ty1) {
  const int nty = ty1 - ty0;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)(tx1 - tx0) * nty) return;
  const int tx = tx0 + (int)(idx / nty), ty = ty0 + (int)(idx % nty);
  unsigned long long v = 0;
  for (int i = 0; i < 8; ++i) {
    const int x = (tx << 3) + i;
    if (x >= W) break;
    const uint8_t* col = occ + (size_t)x * H;
    for (int j = 0; j < 8; ++j) {
      const int y = (ty << 3) + j;
      if (y < H && col[y]) v |= 1ULL << (i * 8 + j);
    }
  }
  bits[(size_t)tx * tiles_y + ty] = v;
}

__global__ void patch_kernel(uint8_t* __restrict__ occ, const uint8_t* __restrict__ patch, int H, int x0, int y0, int w,
                             int h) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w * h) return;
  int px = idx / h, py = idx - px * h;
  occ[(size_t)(x0 + px) * H + (y0 + py)] = patch[idx] ? 1 : 0;
}

// ---- batched patches (dynamic maps, BASELINE config 4): one CTA per patch, two launches ----
// rects[4*p .. 4*p+3] = x0, y0, w, h; patch p's bytes start at offsets[p] in `patches` (x-major inside the patch).
// Phase 1 writes the cells of every patch; phase 2 (a second launch, so that ALL cells of the batch are in place)
// rebuilds the 8x8-tile words each patch touches.  Patches that are not 8-aligned may share a tile: both CTAs then
// rebuild the same word from the same, final cells -- identical stores, no ordering requirement.  (Round 1 did both
// phases in one launch, which raced on shared tiles.)  Patches must still not overlap CELL-wise: which value an
// overlapped cell ends up with would depend on the CTA schedule.
__global__ void patch_batch_write_kernel(uint8_t* __restrict__ occ, const uint8_t* __restrict__ patches,
                                         const int* __restrict__ rects, const long long* __restrict__ offsets, int H) {
  const int p = blockIdx.x;
  const int x0 = rects[4 * p], y0 = rects[4 * p + 1], w = rects[4 * p + 2], h = rects[4 * p + 3];
  const uint8_t* src = patches + offsets[p];
  for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
    const int px = i / h, py = i - px * h;
    occ[(size_t)(x0 + px) * H + (y0 + py)] = src[i] ? 1 : 0;
  }
}

__global__ void patch_batch_pack_kernel(const uint8_t* __restrict__ occ, unsigned long long* __restrict__ bits,
                                        const int* __restrict__ rects, int W, int H, int tiles_y) {
  const int p = blockIdx.x;
  const int x0 = rects[4 * p], y0 = rects[4 * p + 1], w = rects[4 * p + 2], h = rects[4 * p + 3];
  const int tx0 = x0 >> 3, tx1 = ((x0 + w - 1) >> 3) + 1, ty0 = y0 >> 3, ty1 = ((y0 + h - 1) >> 3) + 1;
  const int nty = ty1 - ty0;
  for (int t = threadIdx.x; t < (tx1 - tx0) * nty; t += blockDim.x) {
    const int tx = tx0 + t / nty, ty = ty0 + t % nty;
    unsigned long long v = 0;
    for (int i = 0; i < 8; ++i) {
      const int x = (tx << 3) + i;
      if (x >= W) break;
      for (int j = 0; j < 8; ++j) {
        const int y = (ty << 3) + j;
        if (y < H && occ[(size_t)x * H + y]) v |= 1ULL << (i * 8 + j);
      }
    }
    bits[(size_t)tx * tiles_y + ty] = v;
  }
}

int apply_patch_batch(rl_method* m, const uint8_t* d_patches, const int* d_rects, const long long* d_offsets, int n) {
  if (n <= 0) return RL_OK;
  patch_batch_write_kernel<<<n, 256, 0, m->stream>>>(m->d_occ, d_patches, d_rects, d_offsets, m->H);
  count_launch();
  RL_CHECK_LAUNCH();
  patch_batch_pack_kernel<<<n, 64, 0, m->stream>>>(m->d_occ, m->d_bits_t, d_rects, m->W, m->H, m->tiles8_y());
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

int upload_occupancy(rl_method* m, const rl_map* map) {
  const size_t n = (size_t)m->W * m->H;
  const long long tiles = (long long)m->tiles8_x() * m->tiles8_y();
  RL_CUDA(cudaMalloc(&m->d_occ, n ? n : 1));
  RL_CUDA(cudaMalloc(&m->d_bits_t, sizeof(unsigned long long) * (size_t)(tiles > 0 ? tiles : 1)));
  if (n) RL_CUDA(cudaMemcpyAsync(m->d_occ, map->occ.data(), n, cudaMemcpyHostToDevice, m->stream));
  if (tiles > 0) {
    pack_tiles_kernel<<<(unsigned)((tiles + 255) / 256), 256, 0, m->stream>>>(m->d_occ, m->d_bits_t, m->W, m->H,
                                                                            m->tiles8_y(), 0, m->tiles8_x(), 0,
                                                                            m->tiles8_y());
    count_launch();
    RL_CHECK_LAUNCH();
  }
  return RL_OK;
}

int apply_patch(rl_method* m, const uint8_t* d_patch, int x0, int y0, int w, int h) {
  patch_kernel<<<(w * h + 255) / 256, 256, 0, m->stream>>>(m->d_occ, d_patch, m->H, x0, y0, w, h);
  count_launch();
  RL_CHECK_LAUNCH();
  const int tx0 = x0 >> 3, tx1 = ((x0 + w - 1) >> 3) + 1, ty0 = y0 >> 3, ty1 = ((y0 + h - 1) >> 3) + 1;
  const long long tiles = (long long)(tx1 - tx0) * (ty1 - ty0);
  pack_tiles_kernel<<<(unsigned)((tiles + 255) / 256), 256, 0, m->stream>>>(m->d_occ, m->d_bits_t, m->W, m->H,
                                                                          m->tiles8_y(), tx0, tx1, ty0, ty1);
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

// ---- whole-map ingest on the device: the map never visits the host ----

// ROS nav_msgs/OccupancyGrid data (int8, row-major [rows][cols]; 0 free, -1 unknown, 100 blocked) as the
// reference's PyOMap(OccupancyGrid) reads it (RangeLibc.pyx:146-157): OMap(rows, cols) with
// grid[x][y] = data[x*cols + y] > 10 -- the message's row index is the map's x, so the layout already is x-major.
__global__ void ingest_grid_kernel(const int8_t* __restrict__ data, uint8_t* __restrict__ occ, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) occ[i] = data[i] > 10 ? 1 : 0;
}

// RGBA8 image rows (as lodepng_decode32 returns them) -> occupancy, the reference's OMap(filename, threshold)
// loop (RangeLib.h:189-199): r = byte 2, g = byte 1, b = byte 0, gray = (int)rgb2gray(r,g,b) with rgb2gray's
// double sum narrowed to its float return type (RangeUtils.h:30-32), occupied iff gray < threshold.
// The image is row-major [H][W] and the grid x-major [W][H]: 32x32 tiles are transposed through shared memory
// so both the pixel reads and the cell writes are coalesced.
__global__ void ingest_rgba_kernel(const uchar4* __restrict__ img, uint8_t* __restrict__ occ, int W, int H,
                                   float threshold) {
  __shared__ uint8_t tile[32][33];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int x = x0 + threadIdx.x, y = y0 + j;
    uint8_t o = 0;
    if (x < W && y < H) {
      const uchar4 p = __ldg(img + (size_t)y * W + x);
      const double r = (double)(float)(int)p.z, g = (double)(float)(int)p.y, b = (double)(float)(int)p.x;
      const float grayf = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(0.229, r), __dmul_rn(0.587, g)), __dmul_rn(0.114, b)));
      const int gray = __float2int_rz(grayf);
      o = ((float)gray < threshold) ? 1 : 0;
    }
    tile[j][threadIdx.x] = o;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int x = x0 + i, y = y0 + threadIdx.x;
    if (x < W && y < H) occ[(size_t)x * H + y] = tile[threadIdx.x][i];
  }
}

static int repack_all(rl_method* m) {
  const long long tiles = (long long)m->tiles8_x() * m->tiles8_y();
  if (tiles <= 0) return RL_OK;
  pack_tiles_kernel<<<(unsigned)((tiles + 255) / 256), 256, 0, m->stream>>>(m->d_occ, m->d_bits_t, m->W, m->H,
                                                                          m->tiles8_y(), 0, m->tiles8_x(), 0,
                                                                          m->tiles8_y());
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

int ingest_occupancy_grid(rl_method* m, const int8_t* d_data) {
  const size_t n = (size_t)m->W * m->H;
  if (n == 0) return RL_OK;
  const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)148 * 16);
  ingest_grid_kernel<<<grid, 256, 0, m->stream>>>(d_data, m->d_occ, n);
  count_launch();
  RL_CHECK_LAUNCH();
  return repack_all(m);
}

int ingest_rgba(rl_method* m, const uint8_t* d_rgba, float threshold) {
  if ((size_t)m->W * m->H == 0) return RL_OK;
  const dim3 grid((m->W + 31) / 32, (m->H + 31) / 32), block(32, 8);
  ingest_rgba_kernel<<<grid, block, 0, m->stream>>>((const uchar4*)d_rgba, m->d_occ, m->W, m->H, threshold);
  count_launch();
  RL_CHECK_LAUNCH();
  return repack_all(m);
}

}  // namespace rl
