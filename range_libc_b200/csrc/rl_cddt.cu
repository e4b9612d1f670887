// CDDT / PCDDT compressed zero-point tables, built and pruned on the device and kept resident
// in HBM in CSR form (slice-major bins, int64 offsets, float zero points).
//
//   build  CDDTCast::CDDTCast  RangeLib.h:974-1159  (per-slice constants :991-1061 on the host with
//          the host libm, because the reference evaluates cosf/sinf of the td discrete angles with
//          glibc; edge map :293-312; projection :1083-1129; per-bin sort + unique :1132-1142)
//   prune  CDDTCast::prune     RangeLib.h:1176-1283
//
// Pipeline: edge pixels -> (count, scan, fill) of projected zero points per bin -> segmented sort
// -> per-bin unique -> CSR.  Sorting/scanning/selection use CUB device primitives; projection,
// marking and compaction kernels are ours.  The result is order independent, so it equals the
// reference's sequentially built table bit for bit.
#include <math.h>

#include <cub/cub.cuh>

#include "rl_internal.cuh"
#include "rl_math.cuh"

namespace rl {

namespace {

struct SliceConsts {
  const int* widths;
  const float* trans;
  const float* cosv;
  const float* sinv;
  const int64_t* slice0;
};

// OMap::make_edge_map(true) RangeLib.h:293-312 with the 8-neighbourhood of RangeUtils.h:36-52:
// occupied and at least one in-bounds free neighbour.
__global__ void edge_flags_kernel(const uint8_t* __restrict__ occ, int W, int H, uint8_t* __restrict__ flags) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)W * H) return;
  const int x = (int)(idx / H), y = (int)(idx % H);
  uint8_t e = 0;
  if (occ[idx]) {
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        if (dx == 0 && dy == 0) continue;
        const int cx = x + dx, cy = y + dy;
        if (cx >= 0 && cy >= 0 && cx < W && cy < H && !occ[(size_t)cx * H + cy]) e = 1;
      }
  }
  flags[idx] = e;
}

// projection of a pixel centre into slice a (:1099-1105); returns lut_space_x
__device__ __forceinline__ float project(float pcx, float pcy, float ca, float sa, float tr, int* lower, int* upper) {
  const float half = __double2float_rn((double)fadd(fabsf(sa), fabsf(ca)) / 2.0);
  const float lx = fsub(fmul(pcx, ca), fmul(pcy, sa));
  const float ly = fadd(fadd(fmul(pcx, sa), fmul(pcy, ca)), tr);
  *upper = __double2int_rz((double)fadd(ly, half) - RL_EPSILON);
  *lower = __double2int_rz((double)fsub(ly, half) + RL_EPSILON);
  return lx;
}

// one thread per (edge pixel, slice).  PASS 0 counts, PASS 1 writes through per-bin cursors.
template <int PASS>
__global__ void project_kernel(const long long* __restrict__ edge_idx, long long n_edge, int H, int na, SliceConsts sc,
                               unsigned long long* __restrict__ counts, const int64_t* __restrict__ raw_off,
                               unsigned long long* __restrict__ cursors, float* __restrict__ raw_vals) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_edge * na) return;
  const long long e = t / na;
  const int a = (int)(t - e * na);
  const long long cell = edge_idx[e];
  const int x = (int)(cell / H), y = (int)(cell % H);
  const float pcx = __double2float_rn((double)x + 0.5), pcy = __double2float_rn((double)y + 0.5);  // :1088
  int lower, upper;
  const float lx = project(pcx, pcy, sc.cosv[a], sc.sinv[a], sc.trans[a], &lower, &upper);
  const int width = sc.widths[a];
  const int64_t s0 = sc.slice0[a];
  for (int i = lower; i <= upper; ++i) {
    if (i < 0 || i >= width) continue;  // the reference would index out of bounds here
    const int64_t b = s0 + i;
    if (PASS == 0) {
      atomicAdd(counts + b, 1ULL);
    } else {
      const unsigned long long slot = atomicAdd(cursors + b, 1ULL);
      raw_vals[raw_off[b] + (int64_t)slot] = lx;
    }
  }
}

// one warp per bin: number of distinct values in a sorted segment
__global__ void unique_count_kernel(const float* __restrict__ vals, const int64_t* __restrict__ off, int64_t nbins,
                                    int64_t* __restrict__ out_count) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nbins) return;
  const int64_t lo = off[b], hi = off[b + 1];
  int cnt = 0;
  for (int64_t i = lo + lane; i < hi; i += 32) cnt += (i == lo || vals[i] != vals[i - 1]) ? 1 : 0;
  for (int s = 16; s > 0; s >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, s);
  if (lane == 0) out_count[b] = cnt;
}

// one warp per bin: copy the kept entries (distinct values, or entries whose `used` flag is set)
template <bool BY_FLAG>
__global__ void compact_kernel(const float* __restrict__ vals, const int64_t* __restrict__ off, int64_t nbins,
                               const uint8_t* __restrict__ used, const int64_t* __restrict__ new_off,
                               float* __restrict__ out_vals) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nbins) return;
  const int64_t lo = off[b], hi = off[b + 1];
  int64_t w = new_off[b];
  for (int64_t base = lo; base < hi; base += 32) {
    const int64_t i = base + lane;
    bool keep = false;
    float v = 0.f;
    if (i < hi) {
      v = vals[i];
      keep = BY_FLAG ? (used[i] != 0) : (i == lo || v != vals[i - 1]);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) out_vals[w + __popc(mask & ((1u << lane) - 1))] = v;
    w += __popc(mask);
  }
}

__global__ void used_count_kernel(const uint8_t* __restrict__ used, const int64_t* __restrict__ off, int64_t nbins,
                                  int64_t* __restrict__ out_count) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nbins) return;
  const int64_t lo = off[b], hi = off[b + 1];
  int cnt = 0;
  for (int64_t i = lo + lane; i < hi; i += 32) cnt += used[i] ? 1 : 0;
  for (int s = 16; s > 0; s >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, s);
  if (lane == 0) out_count[b] = cnt;
}

// ---- prune (RangeLib.h:1188-1262), one slice at a time --------------------------------------
// res[1 + cell] = index assigned at this pixel, or -1 when the pixel assigns nothing
// (bin empty, occupied, handled by the "last entry behind the query" rule, or the reference's
// unassigned-index path).  ub[cell] = 1 for the unassigned-index pixels.  res[0] holds the
// carry-in: the last index assigned in earlier slices.
#define RL_PRUNE_NONE (-1)
__global__ void prune_mark_kernel(const uint8_t* __restrict__ occ, int W, int H, int a, SliceConsts sc,
                                  const int64_t* __restrict__ off, const float* __restrict__ vals, float max_range,
                                  uint8_t* __restrict__ used, int* __restrict__ res, uint8_t* __restrict__ ub) {
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= (long long)W * H) return;
  const int x = (int)(cell / H), y = (int)(cell % H);
  int r = RL_PRUNE_NONE;
  uint8_t is_ub = 0;
  const float _x = __double2float_rn(0.5 + (double)x), _y = __double2float_rn(0.5 + (double)y);
  const float ca = sc.cosv[a], sa = sc.sinv[a];
  const float lx = fsub(fmul(_x, ca), fmul(_y, sa));
  const float ly = fadd(fadd(fmul(_x, sa), fmul(_y, ca)), sc.trans[a]);
  const unsigned li = (unsigned)f2i(ly);
  if (li < (unsigned)sc.widths[a]) {
    const int64_t b = sc.slice0[a] + li;
    const int64_t o0 = off[b];
    const int size = (int)(off[b + 1] - o0);
    const int high = size - 1;
    if (high != -1 && !occ[cell]) {
      const float* B = vals + o0;
      uint8_t* U = used + o0;
      const float last = B[high];
      if (last < lx && fsub(lx, last) < max_range) {
        U[high] = 1;
      } else {
        int index = -1;
        if (high > RL_BINARY_SEARCH_THRESHOLD) {  // std::lower_bound: first element >= lx
          int lo = 0, hi = size;
          while (lo < hi) {
            int mid = lo + ((hi - lo) >> 1);
            if (B[mid] < lx) lo = mid + 1; else hi = mid;
          }
          index = lo;
        } else {
          for (int i = 0; i < size; ++i)
            if (B[i] >= lx) { index = i; break; }
          if (index < 0) is_ub = 1;
        }
        if (!is_ub) {
          r = index;
          if (index < size) U[index] = 1;
          if (index - 1 >= 0) U[index - 1] = 1;
        }
      }
    }
  }
  res[1 + cell] = r;
  ub[cell] = is_ub;
}

struct LastValid {
  __device__ __forceinline__ int operator()(const int& a, const int& b) const { return (b != RL_PRUNE_NONE) ? b : a; }
};

// unassigned-index pixels mark {stale, stale-1} in their own bin, where stale = filled[cell]
// (inclusive last-valid scan up to the previous pixel in (x, y) order, seeded with the carry)
__global__ void prune_stale_kernel(int W, int H, int a, SliceConsts sc, const int64_t* __restrict__ off,
                                   const int* __restrict__ filled, const uint8_t* __restrict__ ub,
                                   uint8_t* __restrict__ used) {
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= (long long)W * H) return;
  if (!ub[cell]) return;
  const int x = (int)(cell / H), y = (int)(cell % H);
  const float _x = __double2float_rn(0.5 + (double)x), _y = __double2float_rn(0.5 + (double)y);
  const float ly = fadd(fadd(fmul(_x, sc.sinv[a]), fmul(_y, sc.cosv[a])), sc.trans[a]);
  const unsigned li = (unsigned)f2i(ly);
  const int64_t b = sc.slice0[a] + li;
  const int64_t o0 = off[b];
  const int size = (int)(off[b + 1] - o0);
  const int index = filled[cell];  // filled[] is the scan of res[]; entry `cell` covers res[0..cell] = pixels < cell
  if (index >= 0 && index < size) used[o0 + index] = 1;
  if (index - 1 >= 0 && index - 1 < size) used[o0 + index - 1] = 1;
}

__global__ void carry_kernel(const int* __restrict__ filled, long long cells, int* __restrict__ res) {
  res[0] = filled[cells];  // last valid index after this slice -> carry-in of the next
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

int upload_consts(rl_method* m) {
  const unsigned td = m->td;
  // RangeLib.h:976-977
  m->td_div_2pi = (float)((double)td / RL_M_2PI);
  m->twopi_div_td = (float)(RL_M_2PI / (double)((float)td));
  m->h_widths.assign(td, 0);
  m->h_trans.assign(td, 0.f);
  m->h_cosv.assign(td, 0.f);
  m->h_sinv.assign(td, 0.f);
  m->h_slice0.assign(td + 1, 0);
  const float Wf = (float)(unsigned)m->W, Hf = (float)(unsigned)m->H;
  for (unsigned i = 0; i < td; ++i) {  // :991-1061, host libm like the reference
    const float angle = (float)(int)i * m->twopi_div_td;
    const float ca = cosf(angle), sa = sinf(angle);
    m->h_cosv[i] = ca;
    m->h_sinv[i] = sa;
    volatile float ws = Wf * sa, hc = Hf * ca;  // volatile: keep the two products as rounded floats
    const float rotated_height = fabsf(ws) + fabsf(hc);
    m->h_widths[i] = (int)(unsigned)ceil((double)rotated_height - RL_EPSILON);
    const float ltc = hc;
    volatile float rtc_v = ws + hc;
    const float rtc = rtc_v, rbc = ws;
    const float inner = (rbc < rtc) ? rbc : rtc;
    const float mn = (inner < ltc) ? inner : ltc;
    const double tr = -1.0 * (double)mn - RL_EPSILON;
    m->h_trans[i] = (float)(0.0 < tr ? tr : 0.0);
    m->h_slice0[i + 1] = m->h_slice0[i] + m->h_widths[i];
  }
  m->nbins = m->h_slice0[td];
  RL_CUDA(cudaMalloc(&m->d_widths, sizeof(int) * td));
  RL_CUDA(cudaMalloc(&m->d_trans, sizeof(float) * td));
  RL_CUDA(cudaMalloc(&m->d_cosv, sizeof(float) * td));
  RL_CUDA(cudaMalloc(&m->d_sinv, sizeof(float) * td));
  RL_CUDA(cudaMalloc(&m->d_slice0, sizeof(int64_t) * (td + 1)));
  RL_CUDA(cudaMemcpyAsync(m->d_widths, m->h_widths.data(), sizeof(int) * td, cudaMemcpyHostToDevice, m->stream));
  RL_CUDA(cudaMemcpyAsync(m->d_trans, m->h_trans.data(), sizeof(float) * td, cudaMemcpyHostToDevice, m->stream));
  RL_CUDA(cudaMemcpyAsync(m->d_cosv, m->h_cosv.data(), sizeof(float) * td, cudaMemcpyHostToDevice, m->stream));
  RL_CUDA(cudaMemcpyAsync(m->d_sinv, m->h_sinv.data(), sizeof(float) * td, cudaMemcpyHostToDevice, m->stream));
  RL_CUDA(cudaMemcpyAsync(m->d_slice0, m->h_slice0.data(), sizeof(int64_t) * (td + 1), cudaMemcpyHostToDevice, m->stream));
  RL_CUDA(cudaStreamSynchronize(m->stream));
  return RL_OK;
}

struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
  }
  template <class T>
  cudaError_t alloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc((void**)p, sizeof(T) * (n ? n : 1));
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};

inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }
inline int slices_to_fill(unsigned td) {  // the reference iterates a < td / 2.0 (:1089, :1179, :1188)
  int na = 0;
  while ((double)na < (double)td / 2.0) ++na;
  return na;
}

}  // namespace

void cddt_free(rl_method* m) {
  cudaFree(m->d_widths);
  cudaFree(m->d_trans);
  cudaFree(m->d_cosv);
  cudaFree(m->d_sinv);
  cudaFree(m->d_slice0);
  cudaFree(m->d_offsets);
  cudaFree(m->d_values);
  m->d_widths = nullptr;
  m->d_trans = m->d_cosv = m->d_sinv = nullptr;
  m->d_slice0 = m->d_offsets = nullptr;
  m->d_values = nullptr;
  m->nbins = m->nvalues = 0;
}

int cddt_build(rl_method* m) {
  cddt_free(m);
  m->pruned = false;
  if (m->td == 0) {
    set_error("CDDT: theta_discretization must be > 0");
    return RL_E_INVALID;
  }
  int rc = upload_consts(m);
  if (rc) return rc;
  const int W = m->W, H = m->H;
  const long long cells = (long long)W * H;
  const int64_t nbins = m->nbins;
  const int na = slices_to_fill(m->td);
  cudaStream_t st = m->stream;
  Scratch sc;
  SliceConsts consts{m->d_widths, m->d_trans, m->d_cosv, m->d_sinv, m->d_slice0};

  // 1. edge pixels
  uint8_t* d_flags = nullptr;
  long long* d_edge = nullptr;
  long long* d_nedge = nullptr;
  RL_CUDA(sc.alloc(&d_flags, (size_t)cells));
  RL_CUDA(sc.alloc(&d_edge, (size_t)cells));
  RL_CUDA(sc.alloc(&d_nedge, 1));
  if (cells > 0) {
    edge_flags_kernel<<<blocks_for(cells, 256), 256, 0, st>>>(m->d_occ, W, H, d_flags);
    count_launch();
    RL_CHECK_LAUNCH();
  }
  RL_CUDA(cudaMemsetAsync(d_nedge, 0, sizeof(long long), st));
  if (cells > 0) {
    size_t tb = 0;
    cub::CountingInputIterator<long long> it(0);
    RL_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, it, d_flags, d_edge, d_nedge, cells, st));
    void* tmp = nullptr;
    RL_CUDA(sc.alloc((uint8_t**)&tmp, tb));
    RL_CUDA(cub::DeviceSelect::Flagged(tmp, tb, it, d_flags, d_edge, d_nedge, cells, st));
    count_launch(2);
  }
  long long n_edge = 0;
  RL_CUDA(cudaMemcpyAsync(&n_edge, d_nedge, sizeof(long long), cudaMemcpyDeviceToHost, st));
  RL_CUDA(cudaStreamSynchronize(st));

  // 2. count projected zero points per bin, scan, fill
  unsigned long long* d_counts = nullptr;
  unsigned long long* d_cursors = nullptr;
  int64_t* d_raw_off = nullptr;
  RL_CUDA(sc.alloc(&d_counts, (size_t)nbins + 1));
  RL_CUDA(sc.alloc(&d_cursors, (size_t)nbins + 1));
  RL_CUDA(sc.alloc(&d_raw_off, (size_t)nbins + 1));
  RL_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * ((size_t)nbins + 1), st));
  RL_CUDA(cudaMemsetAsync(d_cursors, 0, sizeof(unsigned long long) * ((size_t)nbins + 1), st));
  const long long work = n_edge * na;
  if (work > 0) {
    project_kernel<0><<<blocks_for(work, 256), 256, 0, st>>>(d_edge, n_edge, H, na, consts, d_counts, nullptr, nullptr,
                                                            nullptr);
    count_launch();
    RL_CHECK_LAUNCH();
  }
  {
    size_t tb = 0;
    RL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t*)d_counts, d_raw_off, nbins + 1, st));
    void* tmp = nullptr;
    RL_CUDA(sc.alloc((uint8_t**)&tmp, tb));
    RL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, (int64_t*)d_counts, d_raw_off, nbins + 1, st));
    count_launch();
  }
  int64_t n_raw = 0;
  RL_CUDA(cudaMemcpyAsync(&n_raw, d_raw_off + nbins, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  RL_CUDA(cudaStreamSynchronize(st));
  float* d_raw = nullptr;
  float* d_sorted = nullptr;
  RL_CUDA(sc.alloc(&d_raw, (size_t)n_raw));
  RL_CUDA(sc.alloc(&d_sorted, (size_t)n_raw));
  if (work > 0) {
    project_kernel<1><<<blocks_for(work, 256), 256, 0, st>>>(d_edge, n_edge, H, na, consts, nullptr, d_raw_off,
                                                            d_cursors, d_raw);
    count_launch();
    RL_CHECK_LAUNCH();
  }
  // 3. per-bin sort (std::sort :1137)
  if (n_raw > 0) {
    size_t tb = 0;
    RL_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, tb, d_raw, d_sorted, n_raw, nbins, d_raw_off, d_raw_off + 1, st));
    void* tmp = nullptr;
    RL_CUDA(sc.alloc((uint8_t**)&tmp, tb));
    RL_CUDA(cub::DeviceSegmentedSort::SortKeys(tmp, tb, d_raw, d_sorted, n_raw, nbins, d_raw_off, d_raw_off + 1, st));
    count_launch(3);
  }
  // 4. per-bin unique (:1140) -> final CSR
  int64_t* d_ucount = nullptr;
  RL_CUDA(sc.alloc(&d_ucount, (size_t)nbins + 1));
  RL_CUDA(cudaMemsetAsync(d_ucount, 0, sizeof(int64_t) * ((size_t)nbins + 1), st));
  if (nbins > 0) {
    unique_count_kernel<<<blocks_for(nbins * 32, 256), 256, 0, st>>>(d_sorted, d_raw_off, nbins, d_ucount);
    count_launch();
    RL_CHECK_LAUNCH();
  }
  RL_CUDA(cudaMalloc(&m->d_offsets, sizeof(int64_t) * ((size_t)nbins + 1)));
  {
    size_t tb = 0;
    RL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_ucount, m->d_offsets, nbins + 1, st));
    void* tmp = nullptr;
    RL_CUDA(sc.alloc((uint8_t**)&tmp, tb));
    RL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, d_ucount, m->d_offsets, nbins + 1, st));
    count_launch();
  }
  int64_t nvalues = 0;
  RL_CUDA(cudaMemcpyAsync(&nvalues, m->d_offsets + nbins, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  RL_CUDA(cudaStreamSynchronize(st));
  m->nvalues = nvalues;
  RL_CUDA(cudaMalloc(&m->d_values, sizeof(float) * ((size_t)nvalues + 1)));  // +1 pad, see cddt_cast
  RL_CUDA(cudaMemsetAsync(m->d_values + nvalues, 0, sizeof(float), st));
  if (nbins > 0 && nvalues > 0) {
    compact_kernel<false><<<blocks_for(nbins * 32, 256), 256, 0, st>>>(d_sorted, d_raw_off, nbins, nullptr, m->d_offsets,
                                                                      m->d_values);
    count_launch();
    RL_CHECK_LAUNCH();
  }
  RL_CUDA(cudaStreamSynchronize(st));
  return RL_OK;
}

int cddt_prune(rl_method* m, float max_range) {
  const int W = m->W, H = m->H;
  const long long cells = (long long)W * H;
  const int64_t nbins = m->nbins;
  const int na = slices_to_fill(m->td);
  cudaStream_t st = m->stream;
  if (nbins == 0 || m->nvalues == 0 || cells == 0) {
    m->pruned = true;
    return RL_OK;
  }
  Scratch sc;
  SliceConsts consts{m->d_widths, m->d_trans, m->d_cosv, m->d_sinv, m->d_slice0};
  uint8_t* d_used = nullptr;
  int* d_res = nullptr;
  int* d_filled = nullptr;
  uint8_t* d_ub = nullptr;
  RL_CUDA(sc.alloc(&d_used, (size_t)m->nvalues));
  RL_CUDA(sc.alloc(&d_res, (size_t)cells + 1));
  RL_CUDA(sc.alloc(&d_filled, (size_t)cells + 1));
  RL_CUDA(sc.alloc(&d_ub, (size_t)cells));
  RL_CUDA(cudaMemsetAsync(d_used, 0, (size_t)m->nvalues, st));
  set_int_kernel<<<1, 1, 0, st>>>(d_res, -2);  // nothing assigned yet: marks nothing
  count_launch();
  size_t tb = 0;
  RL_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb, d_res, d_filled, LastValid(), cells + 1, st));
  void* tmp = nullptr;
  RL_CUDA(sc.alloc((uint8_t**)&tmp, tb));
  for (int a = 0; a < na; ++a) {
    prune_mark_kernel<<<blocks_for(cells, 256), 256, 0, st>>>(m->d_occ, W, H, a, consts, m->d_offsets, m->d_values,
                                                             max_range, d_used, d_res, d_ub);
    RL_CHECK_LAUNCH();
    RL_CUDA(cub::DeviceScan::InclusiveScan(tmp, tb, d_res, d_filled, LastValid(), cells + 1, st));
    prune_stale_kernel<<<blocks_for(cells, 256), 256, 0, st>>>(W, H, a, consts, m->d_offsets, d_filled, d_ub, d_used);
    carry_kernel<<<1, 1, 0, st>>>(d_filled, cells, d_res);
    count_launch(4);
    RL_CHECK_LAUNCH();
  }
  // rebuild the bins with the marked entries only (:1268-1278)
  int64_t* d_count = nullptr;
  int64_t* d_new_off = nullptr;
  RL_CUDA(sc.alloc(&d_count, (size_t)nbins + 1));
  RL_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t) * ((size_t)nbins + 1), st));
  used_count_kernel<<<blocks_for(nbins * 32, 256), 256, 0, st>>>(d_used, m->d_offsets, nbins, d_count);
  count_launch();
  RL_CHECK_LAUNCH();
  RL_CUDA(cudaMalloc(&d_new_off, sizeof(int64_t) * ((size_t)nbins + 1)));
  {
    size_t tb2 = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tb2, d_count, d_new_off, nbins + 1, st);
    void* tmp2 = nullptr;
    if (e == cudaSuccess) e = sc.alloc((uint8_t**)&tmp2, tb2);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp2, tb2, d_count, d_new_off, nbins + 1, st);
    if (e != cudaSuccess) {
      cudaFree(d_new_off);
      return cuda_fail(e, "prune scan", __FILE__, __LINE__);
    }
    count_launch();
  }
  int64_t nvalues = 0;
  cudaError_t e = cudaMemcpyAsync(&nvalues, d_new_off + nbins, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  float* d_new_vals = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&d_new_vals, sizeof(float) * ((size_t)nvalues + 1));
  if (e != cudaSuccess) {
    cudaFree(d_new_off);
    return cuda_fail(e, "prune alloc", __FILE__, __LINE__);
  }
  cudaMemsetAsync(d_new_vals + nvalues, 0, sizeof(float), st);
  compact_kernel<true><<<blocks_for(nbins * 32, 256), 256, 0, st>>>(m->d_values, m->d_offsets, nbins, d_used, d_new_off,
                                                                   d_new_vals);
  count_launch();
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    cudaFree(d_new_off);
    cudaFree(d_new_vals);
    return cuda_fail(e, "prune compact", __FILE__, __LINE__);
  }
  cudaFree(m->d_offsets);
  cudaFree(m->d_values);
  m->d_offsets = d_new_off;
  m->d_values = d_new_vals;
  m->nvalues = nvalues;
  m->pruned = true;
  return RL_OK;
}

}  // namespace rl
