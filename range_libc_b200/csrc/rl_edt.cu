// Distance transform built on the device, restating the REFERENCE's arithmetic -- not a "better"
// exact EDT: DistanceTransform(OMap*) RangeLib.h:345-373 feeding the Felzenszwalb-Huttenlocher
// lower-envelope code of vendor/distance_transform.h (:873-910 pass order, :1054-1095 1-D kernel,
// :1117-1121 sqrt).  The reference's float rounding of q*q makes its result differ from the
// exact EDT on maps with a side > 4096 px; the per-scanline recurrence below reproduces that
// bit for bit (f32 sums, f64 intersection, f32 fill), so it stays sequential along a scanline
// and parallel across scanlines: one thread per scanline, envelope stack in global scratch,
// one contiguous run of 16-byte entries per scanline (its hot top stays in L1).
#include <float.h>

#include "rl_internal.cuh"
#include "rl_math.cuh"

namespace rl {

// element q of scanline `line`: source is either the occupancy bytes (f = 0 occupied /
// FLT_MAX free, RangeLib.h:353-356) or the float output of the previous pass.
template <bool FROM_OCC>
__device__ __forceinline__ float edt_load(const uint8_t* __restrict__ occ, const float* __restrict__ fin, long long base,
                                          long long es, int q) {
  if (FROM_OCC) return occ[base + q * es] ? 0.0f : FLT_MAX;
  return fin[base + q * es];
}

// One envelope entry below the top of a scanline's stack: parabola position, its raw input value and the left end of
// its interval.  16 bytes, stored contiguously per scanline (stk[line * (n + 1) + k]): a push is one 16-byte store, a
// pop one 16-byte load that hits L1 (the top of a scanline's stack stays within one or two 128-byte lines).
struct __align__(16) EdtEntry {
  int v;
  float fraw;  // f[v]; the sum f[v] + v*v is re-formed on a pop (same float operation as when it was pushed)
  double z;
};

// FINAL_SQRT: second pass -- square-rooted (line = y, q = x, written x-major)
//
// Round 2: the recurrence itself is unchanged (it is the reference's, operation by operation); what changed is
// how its operands reach the thread.  Round 1 spent ~2 us per element (27 ms for an 8192^2 map) on dependent
// DRAM / L2 round trips: the input element f[q] (a new cache line every step, requested only when the previous
// step had finished) and three separate scratch arrays laid out [k][line].  Now the next EDT_PF input elements are
// requested while the current EDT_PF are processed (they do not depend on the recurrence), and the stack is an
// array of 16-byte entries per scanline.
#define EDT_PF 8
template <bool FROM_OCC, bool FINAL_SQRT>
__global__ void __launch_bounds__(32)
edt_pass_kernel(const uint8_t* __restrict__ occ, const float* __restrict__ fin, float* __restrict__ out, int nlines,
                int n, long long in_ls, long long in_es, long long out_ls, long long out_es,
                EdtEntry* __restrict__ stack) {
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= nlines) return;
  const long long ib = (long long)line * in_ls, ob = (long long)line * out_ls;
  if (n == 1) {  // distance_transform.h:1058-1062
    float v = edt_load<FROM_OCC>(occ, fin, ib, in_es, 0);
    out[ob] = FINAL_SQRT ? __fsqrt_rn(v) : v;
    return;
  }
  EdtEntry* __restrict__ stk = stack + (size_t)line * ((size_t)n + 1);
  // lower envelope (:1072-1083).  The top entry lives in registers; entries below it in scratch.
  int k = 0;
  int v_top = 0;
  float fraw_top = edt_load<FROM_OCC>(occ, fin, ib, in_es, 0);
  float fv_top = fadd(fraw_top, 0.0f);  // f[0] + (float)(0*0)
  double z_top = -DBL_MAX;
  float nxt[EDT_PF];
#pragma unroll
  for (int j = 0; j < EDT_PF; ++j) nxt[j] = (1 + j < n) ? edt_load<FROM_OCC>(occ, fin, ib, in_es, 1 + j) : 0.0f;
  for (int q0 = 1; q0 < n; q0 += EDT_PF) {
    float cur[EDT_PF];
#pragma unroll
    for (int j = 0; j < EDT_PF; ++j) cur[j] = nxt[j];
#pragma unroll
    for (int j = 0; j < EDT_PF; ++j) {
      const int qn = q0 + EDT_PF + j;
      nxt[j] = (qn < n) ? edt_load<FROM_OCC>(occ, fin, ib, in_es, qn) : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < EDT_PF; ++j) {
      const int q = q0 + j;
      if (q >= n) break;
      const float fq = fadd(cur[j], (float)((unsigned long long)q * q));
      double s;
      while (true) {
        s = __ddiv_rn(__dsub_rn((double)fq, (double)fv_top),
                      __dsub_rn((double)(2 * (long long)q), (double)(2 * (long long)v_top)));
        if (s <= z_top && k > 0) {  // pop
          --k;
          const EdtEntry e = stk[k];
          v_top = e.v;
          fraw_top = e.fraw;
          fv_top = fadd(e.fraw, (float)((unsigned long long)e.v * e.v));
          z_top = e.z;
          continue;
        }
        break;
      }
      stk[k] = EdtEntry{v_top, fraw_top, z_top};  // push the old top down
      ++k;
      v_top = q;
      fraw_top = cur[j];
      fv_top = fq;
      z_top = s;
    }
  }
  stk[k] = EdtEntry{v_top, fraw_top, z_top};
  const int ktop = k;
  // fill (:1086-1091)
  int kk = 0;
  EdtEntry e = stk[0];
  int cur_v = e.v;
  float cur_f = e.fraw;
  EdtEntry en = (kk < ktop) ? stk[1] : EdtEntry{0, 0.0f, DBL_MAX};
  double next_z = (kk < ktop) ? en.z : DBL_MAX;
  for (int q = 0; q < n; ++q) {
    while (next_z < (double)q) {
      ++kk;
      cur_v = en.v;
      cur_f = en.fraw;
      if (kk < ktop) {
        en = stk[kk + 1];
        next_z = en.z;
      } else {
        next_z = DBL_MAX;
      }
    }
    const float dq = fsub((float)q, (float)cur_v);
    const float D = fadd(cur_f, fmul(dq, dq));
    out[ob + q * out_es] = FINAL_SQRT ? __fsqrt_rn(D) : D;
  }
}

int build_distance_transform(rl_method* m) {
  const int W = m->W, H = m->H;
  const size_t cells = (size_t)W * H;
  if (!m->d_dt) {
    RL_CUDA(cudaMalloc(&m->d_dt, sizeof(float) * (m->dt_elems() ? m->dt_elems() : 1)));
    RL_CUDA(cudaMemsetAsync(m->d_dt, 0, sizeof(float) * m->dt_elems(), m->stream));
  }
  if (cells == 0) return RL_OK;
  float* d_tmp = nullptr;
  EdtEntry* stack = nullptr;
  const size_t nmax = (size_t)(W > H ? W : H);
  const size_t stack_elems = (nmax + 1) * nmax;  // (n+1) entries for max(W,H) lines
  RL_CUDA(cudaMalloc(&d_tmp, sizeof(float) * cells));
  cudaError_t ea = cudaMalloc(&stack, sizeof(EdtEntry) * stack_elems);
  if (ea != cudaSuccess) {
    cudaFree(d_tmp);
    return cuda_fail(ea, "edt scratch", __FILE__, __LINE__);
  }
  const int threads = 32;
  // pass 1: for each x a scanline along y (slices of dimension 0 first, :893-900)
  edt_pass_kernel<true, false><<<(W + threads - 1) / threads, threads, 0, m->stream>>>(
      m->d_occ, nullptr, d_tmp, W, H, (long long)H, 1LL, (long long)H, 1LL, stack);
  count_launch();
  // pass 2: for each y a scanline along x; result square-rooted (:1117-1121)
  edt_pass_kernel<false, true><<<(H + threads - 1) / threads, threads, 0, m->stream>>>(
      nullptr, d_tmp, m->d_dt, H, W, 1LL, (long long)H, 1LL, (long long)H, stack);
  count_launch();
  cudaError_t e = cudaGetLastError();
  cudaError_t e2 = cudaStreamSynchronize(m->stream);
  cudaFree(d_tmp);
  cudaFree(stack);
  if (e != cudaSuccess) return cuda_fail(e, "edt launch", __FILE__, __LINE__);
  if (e2 != cudaSuccess) return cuda_fail(e2, "edt sync", __FILE__, __LINE__);
  return RL_OK;
}

}  // namespace rl
