// Dependent-load latency on B200 at natural clocks: one thread chases a random cyclic permutation.
//   8 KB buffer  -> L1 hits     4 MB -> L2 hits     1 GB -> HBM
// ld.global.nc (what __ldg emits) and plain ld.global; 4-byte elements (index of next element).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <numeric>
#include <algorithm>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <bool NC>
__global__ void chase(const unsigned* __restrict__ a, int steps, unsigned* out, long long* cyc) {
  unsigned i = 0;
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) i = NC ? __ldg(a + i) : ((volatile const unsigned*)a)[i];
  long long t1 = clock64();
  *out = i;
  *cyc = t1 - t0;
}

int main() {
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("SM clock (attr) %d MHz\n", clk_khz / 1000);
  for (size_t bytes : {(size_t)8 << 10, (size_t)64 << 10, (size_t)4 << 20, (size_t)32 << 20, (size_t)1 << 30}) {
    size_t n = bytes / 4;
    // stride through 128-byte lines in random order so every hop is a new line
    size_t lines = n / 32;
    std::vector<unsigned> perm(lines);
    std::iota(perm.begin(), perm.end(), 0u);
    std::mt19937 rng(1);
    std::shuffle(perm.begin(), perm.end(), rng);
    std::vector<unsigned> h(n, 0);
    for (size_t k = 0; k < lines; ++k) h[(size_t)perm[k] * 32] = perm[(k + 1) % lines] * 32;
    unsigned* d;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice));
    unsigned* out; long long* cyc;
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, 8));
    int steps = 20000;
    for (int nc = 0; nc < 2; ++nc) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        if (nc) chase<true><<<1, 1>>>(d, steps, out, cyc); else chase<false><<<1, 1>>>(d, steps, out, cyc);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        if (rep == 1)
          printf("%8zu KB %-12s: %7.1f ns/hop  %7.1f cycles/hop\n", bytes >> 10, nc ? "ld.global.nc" : "ld.global", ms * 1e6 / steps, (double)c / steps);
      }
    }
    cudaFree(d);
  }
  return 0;
}
