// Distance transform built on the device, restating the REFERENCE's arithmetic -- not a "better"
// exact EDT: DistanceTransform(OMap*) RangeLib.h:345-373 feeding the Felzenszwalb-Huttenlocher
// lower-envelope code of vendor/distance_transform.h (:873-910 pass order, :1054-1095 1-D kernel,
// :1117-1121 sqrt).  The reference's float rounding of q*q makes its result differ from the
// exact EDT on maps with a side > 4096 px; the per-scanline recurrence below reproduces that
// bit for bit (f32 sums, f64 intersection, f32 fill), so it stays sequential along a scanline
// and parallel across scanlines: one thread per scanline, envelope stack in global scratch
// laid out [k][line] so the threads of a warp touch neighbouring addresses.
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

// FINAL_SQRT: second pass -- square-rooted (line = y, q = x, written x-major)
template <bool FROM_OCC, bool FINAL_SQRT>
__global__ void __launch_bounds__(64)
edt_pass_kernel(const uint8_t* __restrict__ occ, const float* __restrict__ fin, float* __restrict__ out, int nlines,
                int n, long long in_ls, long long in_es, long long out_ls, long long out_es, int* __restrict__ vstk,
                float* __restrict__ fstk, double* __restrict__ zstk) {
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= nlines) return;
  const long long ib = (long long)line * in_ls, ob = (long long)line * out_ls;
  if (n == 1) {  // distance_transform.h:1058-1062
    float v = edt_load<FROM_OCC>(occ, fin, ib, in_es, 0);
    out[ob] = FINAL_SQRT ? __fsqrt_rn(v) : v;
    return;
  }
  // lower envelope (:1072-1083).  The top entry lives in registers; entries below it in scratch.
  int k = 0;
  int v_top = 0;
  float fv_top = fadd(edt_load<FROM_OCC>(occ, fin, ib, in_es, 0), 0.0f);  // f[0] + (float)(0*0)
  double z_top = -DBL_MAX;
  for (int q = 1; q < n; ++q) {
    const float fq = fadd(edt_load<FROM_OCC>(occ, fin, ib, in_es, q), (float)((unsigned long long)q * q));
    double s;
    while (true) {
      s = __ddiv_rn(__dsub_rn((double)fq, (double)fv_top), __dsub_rn((double)(2 * (long long)q), (double)(2 * (long long)v_top)));
      if (s <= z_top && k > 0) {  // pop
        --k;
        const size_t o = (size_t)k * nlines + line;
        v_top = vstk[o];
        fv_top = fstk[o];
        z_top = zstk[o];
        continue;
      }
      break;
    }
    const size_t o = (size_t)k * nlines + line;  // push the old top down
    vstk[o] = v_top;
    fstk[o] = fv_top;
    zstk[o] = z_top;
    ++k;
    v_top = q;
    fv_top = fq;
    z_top = s;
  }
  {
    const size_t o = (size_t)k * nlines + line;
    vstk[o] = v_top;
    zstk[o] = z_top;
  }
  const int ktop = k;
  // fill (:1086-1091)
  int kk = 0;
  int cur_v = vstk[line];
  float cur_f = edt_load<FROM_OCC>(occ, fin, ib, in_es, cur_v);
  double next_z = (kk < ktop) ? zstk[(size_t)(kk + 1) * nlines + line] : DBL_MAX;
  for (int q = 0; q < n; ++q) {
    while (next_z < (double)q) {
      ++kk;
      cur_v = vstk[(size_t)kk * nlines + line];
      cur_f = edt_load<FROM_OCC>(occ, fin, ib, in_es, cur_v);
      next_z = (kk < ktop) ? zstk[(size_t)(kk + 1) * nlines + line] : DBL_MAX;
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
  int* vstk = nullptr;
  float* fstk = nullptr;
  double* zstk = nullptr;
  const size_t nmax = (size_t)(W > H ? W : H);
  const size_t stack_elems = (nmax + 1) * nmax;  // (n+1) entries for max(W,H) lines
  RL_CUDA(cudaMalloc(&d_tmp, sizeof(float) * cells));
  RL_CUDA(cudaMalloc(&vstk, sizeof(int) * stack_elems));
  RL_CUDA(cudaMalloc(&fstk, sizeof(float) * stack_elems));
  RL_CUDA(cudaMalloc(&zstk, sizeof(double) * stack_elems));
  const int threads = 64;
  // pass 1: for each x a scanline along y (slices of dimension 0 first, :893-900)
  edt_pass_kernel<true, false><<<(W + threads - 1) / threads, threads, 0, m->stream>>>(
      m->d_occ, nullptr, d_tmp, W, H, (long long)H, 1LL, (long long)H, 1LL, vstk, fstk, zstk);
  count_launch();
  // pass 2: for each y a scanline along x; result square-rooted (:1117-1121)
  edt_pass_kernel<false, true><<<(H + threads - 1) / threads, threads, 0, m->stream>>>(
      nullptr, d_tmp, m->d_dt, H, W, 1LL, (long long)H, 1LL, (long long)H, vstk, fstk, zstk);
  count_launch();
  cudaError_t e = cudaGetLastError();
  cudaError_t e2 = cudaStreamSynchronize(m->stream);
  cudaFree(d_tmp);
  cudaFree(vstk);
  cudaFree(fstk);
  cudaFree(zstk);
  if (e != cudaSuccess) return cuda_fail(e, "edt launch", __FILE__, __LINE__);
  if (e2 != cudaSuccess) return cuda_fail(e2, "edt sync", __FILE__, __LINE__);
  return RL_OK;
}

}  // namespace rl
