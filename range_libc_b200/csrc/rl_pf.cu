// Particle-filter steps either side of the sensor update.  NOT in the reference: range_libc
// ends at the per-particle weights (RangeLib.h:558-612); these are the three steps its downstream user (the
// mit-racecar particle_filter MCL loop, README.md:57 of the reference) runs on the host between two sensor updates --
// weight squash + normalisation, resampling, motion update -- so that particles and weights can stay in HBM from one
// update to the next.  Their checker is oracle/pf_oracle.py (numpy; "parity unpinned": there is no reference source).
//
// Each step is specified so that it does not depend on the order of a parallel reduction:
//  * normalise:  w_i <- pow(w_i, inv_squash) (skipped for inv_squash == 1), S = sum w_i, w_i <- w_i / S.  S is a
//                floating-point sum (cub::DeviceReduce), compared with the oracle's within 1e-12 relative.
//  * resample:   systematic ("low variance") resampling in FIXED POINT: f_i = trunc(w_i * 2^40) as uint64, C = inclusive
//                prefix sums of f (integer, associative: any scan order gives the same C), T = C_{n-1};
//                threshold_j = ((u0 + j) / n) * T in IEEE double (add, divide, multiply, each correctly rounded; the
//                library is compiled without FMA contraction); out[j] = particles[first i with (double)C_i > threshold_j]
//                (clamped to n - 1).  Bit-exact against the oracle.
//  * motion:     the odometry step of a planar pose, float32 in this order:
//                x' = (x + (cos(th) dx - sin(th) dy)) + nx,  y' = (y + (sin(th) dx + cos(th) dy)) + ny,
//                th' = (th + dth) + nth, with the device sinf / cosf that equal libm's bit for bit (rl_math.cuh);
//                the noise (n x 3) is the caller's (NULL = none).  Bit-exact against the oracle.
#include <cub/cub.cuh>

#include "rl_internal.cuh"
#include "rl_math.cuh"

namespace rl {

namespace {

__global__ void __launch_bounds__(256)
pf_pow_kernel(double* __restrict__ w, int n, double inv_squash) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] = pow(w[i], inv_squash);
}

__global__ void __launch_bounds__(256)
pf_scale_kernel(double* __restrict__ w, int n, const double* __restrict__ sum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] = __ddiv_rn(w[i], *sum);
}

__global__ void __launch_bounds__(256)
pf_fixed_kernel(const double* __restrict__ w, int n, unsigned long long* __restrict__ f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = __dmul_rn(w[i], 1099511627776.0);  // 2^40: exact scaling
  f[i] = (v > 0.0) ? __double2ull_rz(v) : 0ULL;        // NaN and negative weights count as zero
}

// one thread per output particle: bisection over the prefix sums (L2-resident: 8 B per particle)
__global__ void __launch_bounds__(256)
pf_resample_kernel(const float* __restrict__ particles, const unsigned long long* __restrict__ C, int n, double u0,
                   float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double T = (double)C[n - 1];
  const double thr = __dmul_rn(__ddiv_rn(__dadd_rn(u0, (double)j), (double)n), T);
  int lo = 0, cnt = n;  // first i with (double)C[i] > thr
  while (cnt > 0) {
    const int half = cnt >> 1;
    const bool go_right = !((double)__ldg(C + lo + half) > thr);
    lo = go_right ? lo + half + 1 : lo;
    cnt = go_right ? cnt - half - 1 : half;
  }
  const int i = min(lo, n - 1);
  out[3 * j] = __ldg(particles + 3 * i);
  out[3 * j + 1] = __ldg(particles + 3 * i + 1);
  out[3 * j + 2] = __ldg(particles + 3 * i + 2);
}

__global__ void __launch_bounds__(256)
pf_motion_kernel(float* __restrict__ p, int n, float dx, float dy, float dth, const float* __restrict__ noise) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = p[3 * i], y = p[3 * i + 1], th = p[3 * i + 2];
  float sn, cs;
  rl_sincosf(th, &sn, &cs);
  float nx = fadd(x, fsub(fmul(cs, dx), fmul(sn, dy)));
  float ny = fadd(y, fadd(fmul(sn, dx), fmul(cs, dy)));
  float nth = fadd(th, dth);
  if (noise) {
    nx = fadd(nx, noise[3 * i]);
    ny = fadd(ny, noise[3 * i + 1]);
    nth = fadd(nth, noise[3 * i + 2]);
  }
  p[3 * i] = nx;
  p[3 * i + 1] = ny;
  p[3 * i + 2] = nth;
}

// scratch owned by the handle: [0, 8) the sum, then n uint64 (fixed-point weights / prefix sums), then cub's
int ensure_pf_scratch(rl_method* m, int n, size_t cub_bytes, unsigned long long** fixed, void** cub_tmp) {
  const size_t need = 256 + sizeof(unsigned long long) * (size_t)n + 256 + cub_bytes;
  if (need > m->pf_bytes) {
    RL_CUDA(cudaStreamSynchronize(m->stream));
    cudaFree(m->d_pf);
    m->d_pf = nullptr;
    m->pf_bytes = 0;
    RL_CUDA(cudaMalloc(&m->d_pf, need + need / 4));
    m->pf_bytes = need + need / 4;
  }
  *fixed = reinterpret_cast<unsigned long long*>((char*)m->d_pf + 256);
  *cub_tmp = (char*)m->d_pf + 256 + ((sizeof(unsigned long long) * (size_t)n + 255) & ~(size_t)255);
  return RL_OK;
}

}  // namespace

int pf_normalize(rl_method* m, double* d_w, int n, double inv_squash, double* h_sum) {
  cudaStream_t st = m->stream;
  size_t tb = 0;
  RL_CUDA(cub::DeviceReduce::Sum(nullptr, tb, d_w, (double*)nullptr, n, st));
  unsigned long long* fixed = nullptr;
  void* tmp = nullptr;
  int rc = ensure_pf_scratch(m, 0, tb, &fixed, &tmp);
  if (rc) return rc;
  double* d_sum = reinterpret_cast<double*>(m->d_pf);
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (inv_squash != 1.0) {
    pf_pow_kernel<<<grid, 256, 0, st>>>(d_w, n, inv_squash);
    count_launch();
  }
  RL_CUDA(cub::DeviceReduce::Sum(tmp, tb, d_w, d_sum, n, st));
  pf_scale_kernel<<<grid, 256, 0, st>>>(d_w, n, d_sum);
  count_launch(2);
  RL_CHECK_LAUNCH();
  if (h_sum) {
    RL_CUDA(cudaMemcpyAsync(h_sum, d_sum, sizeof(double), cudaMemcpyDeviceToHost, st));
    RL_CUDA(cudaStreamSynchronize(st));
  }
  return RL_OK;
}

int pf_resample(rl_method* m, const float* d_particles, const double* d_w, float* d_out, int n, double u0) {
  cudaStream_t st = m->stream;
  size_t tb = 0;
  RL_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, (unsigned long long*)nullptr, (unsigned long long*)nullptr, n, st));
  unsigned long long* fixed = nullptr;
  void* tmp = nullptr;
  int rc = ensure_pf_scratch(m, n, tb, &fixed, &tmp);
  if (rc) return rc;
  const unsigned grid = (unsigned)((n + 255) / 256);
  pf_fixed_kernel<<<grid, 256, 0, st>>>(d_w, n, fixed);
  RL_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb, fixed, fixed, n, st));
  pf_resample_kernel<<<grid, 256, 0, st>>>(d_particles, fixed, n, u0, d_out);
  count_launch(3);
  RL_CHECK_LAUNCH();
  return RL_OK;
}

int pf_motion(rl_method* m, float* d_particles, int n, float dx, float dy, float dth, const float* d_noise) {
  pf_motion_kernel<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(d_particles, n, dx, dy, dth, d_noise);
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

void pf_free(rl_method* m) {
  cudaFree(m->d_pf);
  m->d_pf = nullptr;
  m->pf_bytes = 0;
}

}  // namespace rl
