// Cycles per replay step of the cooperative RM tail for different lane-matching schemes (1 warp).
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

template <int V>
__global__ void k(float x0, float y0, float dx, float dy, int iters, float* out, long long* cyc) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  // probes: lane j holds cell at t = 0.75 j (ray along +x): key, step bits
  float s = lane * 0.75f;
  int cx = __float2int_rz(x0 + dx * s), cy = __float2int_rz(y0 + dy * s);
  int key = (cx << 16) | cy;
  unsigned stepbits = __float_as_uint(1.0f + 0.001f * lane);
  float t = 0.f;
  float acc = 0.f;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    int k0 = 0;
    if (V < 4) {
      int px = __float2int_rz(__fadd_rn(x0, __fmul_rn(dx, t)));
      int py = __float2int_rz(__fadd_rn(y0, __fmul_rn(dy, t)));
      k0 = ((unsigned)px < 1200u && (unsigned)py < 1200u) ? ((px << 16) | py) : -2;
    }
    float st;
    if (V == 4) {
      // interval test instead of the cell computation: lane j claims t in [0.75 j, 0.75 (j + 1)) (round 2)
      const float lo = lane * 0.75f, hi = lo + 0.75f;
      const bool in = t >= lo && t < hi;
      st = __uint_as_float(__reduce_min_sync(FULL, in ? stepbits : 0x7f800000u));
    } else if (V == 5) {
      // sorted breakpoints: index = popc(ballot(lo <= t)) - 1, then one shuffle
      const float lo = lane * 0.75f;
      const unsigned m = __ballot_sync(FULL, lo <= t);
      st = __shfl_sync(FULL, __uint_as_float(stepbits), __popc(m) - 1);
    } else if (V == 6) {
      // interval test, ballot + ffs + shfl
      const float lo = lane * 0.75f, hi = lo + 0.75f;
      const unsigned m = __ballot_sync(FULL, t >= lo && t < hi);
      st = m ? __shfl_sync(FULL, __uint_as_float(stepbits), __ffs(m) - 1) : __uint_as_float(0x7f800000u);
    } else if (V == 0) {
      unsigned r = __reduce_max_sync(FULL, key == k0 ? stepbits : 0u);
      st = __uint_as_float(r ? r : 0x7f800000u);
    } else if (V == 1) {
      unsigned found = __ballot_sync(FULL, key == k0);
      st = found ? __shfl_sync(FULL, __uint_as_float(stepbits), __ffs(found) - 1) : __uint_as_float(0x7f800000u);
    } else if (V == 2) {
      // predicted lane: j = (int)(t / 0.75), check j and j+1
      int j = __float2int_rz(t * (1.0f / 0.75f));
      int ka = __shfl_sync(FULL, key, j), kb = __shfl_sync(FULL, key, j + 1);
      unsigned sa = __shfl_sync(FULL, stepbits, j), sb = __shfl_sync(FULL, stepbits, j + 1);
      unsigned r = ka == k0 ? sa : (kb == k0 ? sb : 0u);
      st = __uint_as_float(r ? r : 0x7f800000u);
    } else {
      // shared-memory table indexed by predicted lane
      __shared__ int skey[32];
      __shared__ unsigned sstep[32];
      if (i == 0) { skey[lane] = key; sstep[lane] = stepbits; __syncwarp(); }
      int j = __float2int_rz(t * (1.0f / 0.75f)) & 31;
      int ka = skey[j], kb = skey[(j + 1) & 31];
      unsigned sa = sstep[j], sb = sstep[(j + 1) & 31];
      unsigned r = ka == k0 ? sa : (kb == k0 ? sb : 0u);
      st = __uint_as_float(r ? r : 0x7f800000u);
    }
    t = __fadd_rn(t, st);
    if (!(t < 20.0f)) { acc += t; t = 0.0f; }   // wrap inside the window
  }
  long long t1 = clock64();
  if (lane == 0) { *out = acc + t; *cyc = t1 - t0; }
}

int main() {
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, 8));
  const int iters = 100000;
  const char* names[7] = {"reduce_max (CREDUX)", "ballot+ffs+shfl", "predicted lane 4x shfl", "predicted lane smem",
                          "interval + reduce_min", "breakpoints ballot+popc+shfl", "interval ballot+ffs+shfl"};
  for (int v = 0; v < 7; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      if (v == 0) k<0><<<1, 32>>>(100.3f, 200.7f, 0.9998f, 0.02f, iters, out, cyc);
      if (v == 1) k<1><<<1, 32>>>(100.3f, 200.7f, 0.9998f, 0.02f, iters, out, cyc);
      if (v == 2) k<2><<<1, 32>>>(100.3f, 200.7f, 0.9998f, 0.02f, iters, out, cyc);
      if (v == 3) k<3><<<1, 32>>>(100.3f, 200.7f, 0.9998f, 0.02f, iters, out, cyc);
      if (v == 4) k<4><<<1, 32>>>(100.3f, 200.7f, 0.9998f, 0.02f, iters, out, cyc);
      if (v == 5) k<5><<<1, 32>>>(100.3f, 200.7f, 0.9998f, 0.02f, iters, out, cyc);
      if (v == 6) k<6><<<1, 32>>>(100.3f, 200.7f, 0.9998f, 0.02f, iters, out, cyc);
      CK(cudaDeviceSynchronize());
    }
    long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    printf("%-26s %7.1f cycles/step\n", names[v], (double)c / iters);
  }
  return 0;
}
