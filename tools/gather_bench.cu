// Microbenchmark: what does one divergent 4-byte gather cost on B200 through each path?
//   ldg     global load, linear float array (random index per lane)
//   tex     tex2D point fetch from a cudaArray (block-linear layout)
//   smem    shared-memory gather (random bank per lane)
// dependent = 1: the next index depends on the loaded value (ray-marching-like chain)
// Usage: gather_bench <W> <H>   (float grid of W*H cells)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned lcg(unsigned s) { return s * 1664525u + 1013904223u; }

template <int MODE, bool DEP, bool COHERENT>
__global__ void __launch_bounds__(256) k_gather(const float* __restrict__ a, cudaTextureObject_t tex, int W, int H,
                                                int steps, float* out) {
  extern __shared__ float sm[];
  const unsigned n = (unsigned)W * H;
  unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (MODE == 2) {
    for (int i = threadIdx.x; i < 12288; i += blockDim.x) sm[i] = a[i];
    __syncthreads();
  }
  unsigned s = COHERENT ? (tid >> 5) * 2654435761u : tid * 2654435761u + 12345u;
  float acc = 0.f;
  for (int it = 0; it < steps; ++it) {
    s = lcg(s);
    unsigned idx = (s >> 4) % n;
    if (COHERENT) idx = (idx & ~31u) + (threadIdx.x & 31);
    if (idx >= n) idx = n - 1;
    float v;
    if (MODE == 0) v = __ldg(a + idx);
    else if (MODE == 1) v = tex2D<float>(tex, (float)(idx % W) + 0.5f, (float)(idx / W) + 0.5f);
    else v = sm[idx % 12288];
    acc += v;
    if (DEP) s += __float_as_uint(v) & 0xff;
  }
  out[tid] = acc;
}

template <int MODE, bool DEP, bool COHERENT>
void run(const char* name, const float* d, cudaTextureObject_t tex, int W, int H, float* out, int blocks) {
  const int steps = 256;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  size_t smem = MODE == 2 ? 12288 * 4 : 0;
  k_gather<MODE, DEP, COHERENT><<<blocks, 256, smem>>>(d, tex, W, H, steps, out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < 5; ++r) k_gather<MODE, DEP, COHERENT><<<blocks, 256, smem>>>(d, tex, W, H, steps, out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double g = 5.0 * blocks * 256.0 * steps / (ms * 1e-3) / 1e9;
  printf("%-34s grid %5d: %8.2f G gathers/s  (%.3f ms/launch)\n", name, blocks, g, ms / 5);
}

int main(int argc, char** argv) {
  int W = argc > 1 ? atoi(argv[1]) : 1200, H = argc > 2 ? atoi(argv[2]) : 1200;
  size_t n = (size_t)W * H;
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 97);
  float* d;
  CK(cudaMalloc(&d, n * 4));
  CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
  cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
  cudaArray_t arr;
  CK(cudaMallocArray(&arr, &desc, W, H));
  CK(cudaMemcpy2DToArray(arr, 0, 0, h.data(), (size_t)W * 4, (size_t)W * 4, H, cudaMemcpyHostToDevice));
  cudaResourceDesc rd{};
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = arr;
  cudaTextureDesc td{};
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  cudaTextureObject_t tex;
  CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  float* out;
  CK(cudaMalloc(&out, 148 * 64 * 256 * 4));
  printf("grid %dx%d floats = %.1f MB\n", W, H, n * 4 / 1e6);
  for (int blocks : {148 * 8, 148 * 32}) {
    run<0, false, false>("ldg random independent", d, tex, W, H, out, blocks);
    run<0, true, false>("ldg random dependent", d, tex, W, H, out, blocks);
    run<0, false, true>("ldg warp-coherent independent", d, tex, W, H, out, blocks);
    run<0, true, true>("ldg warp-coherent dependent", d, tex, W, H, out, blocks);
    run<1, false, false>("tex2D random independent", d, tex, W, H, out, blocks);
    run<1, true, false>("tex2D random dependent", d, tex, W, H, out, blocks);
    run<1, true, true>("tex2D warp-coherent dependent", d, tex, W, H, out, blocks);
    run<2, false, false>("smem random independent", d, tex, W, H, out, blocks);
    run<2, true, false>("smem random dependent", d, tex, W, H, out, blocks);
  }
  return 0;
}
