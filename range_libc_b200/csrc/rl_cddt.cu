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

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "rl_internal.cuh"
#include "rl_math.cuh"

namespace rl {

void cddt_free(rl_method* m);


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

// ------------------------------------------------------------------------------------------
// Query index (round 2).  On BASELINE config 3 the zero points (636 MB, 235 MB pruned) do not fit L2, and a query used
// to touch ~9 separate DRAM sectors: two offsets, the bin's first and last value, and the probes of a binary search
// over ~200 values (ncu: 534 B of DRAM traffic per query, 8.4 G rays/s, DRAM 54 % busy on 32-byte gathers).
// Partitioning the queries by bin first (key + radix sort, then a hand-written histogram / scatter) made the table
// stream through L2 once, but the partition itself cost as much as it saved (12.3 and 13.6 G rays/s).  The index
// makes the part of the search that decides WHERE to look small enough to live in L2:
//   meta[bin]  one 16-byte record: offset, size, first and last value of the bin            (13 MB on C3)
//   skip[k]    a 16-bit monotone position code (cddt_code) of values[16 k], the first value of every 64-byte aligned
//              block of values[], relative to the bin that holds it                         (20 MB on C3, 7 MB pruned)
// A query reads its record (the early exits need nothing else), bisects the skip codes of the blocks that start
// inside its bin (L2; a code equal to the query's own code decides nothing and that probe reads the value itself --
// 0.5-pixel resolution on gigantic_map, so it is rare), and then reads the ONE aligned 64-byte block of values[] that holds its answer with four
// 16-byte loads -- one DRAM access instead of nine.  The search returns the same element as before: its predicate is
// monotone along a bin, so "count the blocks whose first element satisfies it, then the elements inside the last such
// block" is the same bisection split in two (cddt_search_indexed, rl_cast.cu).
// ------------------------------------------------------------------------------------------
namespace {
__global__ void index_meta_kernel(const int64_t* __restrict__ off, const float* __restrict__ vals, int64_t nbins,
                                  CddtBinMeta* __restrict__ meta) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  const int64_t o0 = off[b];
  const unsigned size = (unsigned)(off[b + 1] - o0);
  CddtBinMeta mt;
  mt.off = (unsigned)o0;
  mt.size = size;
  mt.first = size ? vals[o0] : 0.0f;
  mt.last = size ? vals[o0 + size - 1] : 0.0f;
  meta[b] = mt;
}
// one warp per bin: codes of the blocks that start inside the bin (after its first element; the block that holds the
// bin's first element is never probed, see cddt_search_indexed)
__global__ void index_skip_kernel(const CddtBinMeta* __restrict__ meta, const float* __restrict__ vals, int64_t nbins,
                                  uint16_t* __restrict__ skip) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nbins) return;
  const CddtBinMeta mt = meta[b];
  if (mt.size == 0) return;
  const unsigned k0 = mt.off >> 4, k1 = (mt.off + mt.size - 1) >> 4;
  const float scale = cddt_code_scale(mt.first, mt.last);
  for (unsigned k = k0 + 1 + lane; k <= k1; k += 32)
    skip[k] = (uint16_t)cddt_code(vals[(size_t)k * RL_CDDT_BLOCK], mt.first, scale);
}
}  // namespace

int cddt_index_build(rl_method* m, bool force_on) {
  cudaFree(m->d_meta);
  cudaFree(m->d_skip);
  m->d_meta = nullptr;
  m->d_skip = nullptr;
  m->nskip = 0;
  m->use_index = false;
  const int64_t nbins = m->nbins;
  if (nbins == 0 || m->nvalues >= 0xfffffff0LL) return RL_OK;  // no index: plain search
  // tables that fit L2 are searched directly (one pass less through L2 per query); RL_CDDT_INDEX=1 / 0 forces
  static const int force = getenv("RL_CDDT_INDEX") ? atoi(getenv("RL_CDDT_INDEX")) : -1;
  const bool want = force_on ? true : force >= 0 ? force != 0 : (size_t)m->nvalues * sizeof(float) > ((size_t)48 << 20);
  if (!want) return RL_OK;
  cudaStream_t st = m->stream;
  const int64_t nskip = (int64_t)(cddt_values_alloc(m->nvalues) / RL_CDDT_BLOCK);
  RL_CUDA(cudaMalloc(&m->d_meta, sizeof(CddtBinMeta) * (size_t)nbins));
  RL_CUDA(cudaMalloc(&m->d_skip, sizeof(uint16_t) * (size_t)nskip));
  RL_CUDA(cudaMemsetAsync(m->d_skip, 0, sizeof(uint16_t) * (size_t)nskip, st));
  index_meta_kernel<<<blocks_for(nbins, 256), 256, 0, st>>>(m->d_offsets, m->d_values, nbins, m->d_meta);
  index_skip_kernel<<<blocks_for(nbins * 32, 256), 256, 0, st>>>(m->d_meta, m->d_values, nbins, m->d_skip);
  count_launch(2);
  RL_CHECK_LAUNCH();
  RL_CUDA(cudaStreamSynchronize(st));
  m->nskip = nskip;
  m->use_index = true;
  return RL_OK;
}

// ------------------------------------------------------------------------------------------
// Binary checkpoint of a built (and possibly pruned) table.  The reference can only dump its table as YAML / JSON text
// for a viewer (CDDTCast::serializeYaml / serializeJson, RangeLib.h:1652-1735, helpers RangeUtils.h:210-238) and has no
// loader; the content is the same -- theta_discretization, lut_translations, max_range, map size, and per slice and bin
// the sorted zero points -- stored as the CSR arrays the kernels read, so that a PCDDT table whose prune takes the CPU
// reference 24 minutes on gigantic_map is loaded without rebuilding.
//   header  : magic "RLCDDT\0\1", u32 version, u32 theta_discretization, i32 W, i32 H, f32 max_range, u32 pruned,
//             u64 FNV-1a of the occupancy bytes, i64 nbins, i64 nvalues
//   payload : i32 widths[td], f32 translations[td], i64 offsets[nbins + 1], f32 values[nvalues]   (little endian)
// ------------------------------------------------------------------------------------------
namespace {
struct CkptHeader {
  char magic[8];
  uint32_t version, td;
  int32_t W, H;
  float max_range;
  uint32_t pruned;
  uint64_t occ_hash;
  int64_t nbins, nvalues;
};
const char kMagic[8] = {'R', 'L', 'C', 'D', 'D', 'T', 0, 1};

uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ULL;
  for (size_t i = 0; i < n; ++i) h = (h ^ (uint64_t)(p[i] ? 1 : 0)) * 1099511628211ULL;
  return h;
}
}  // namespace

int cddt_save(rl_method* m, const char* path) {
  const size_t cells = (size_t)m->W * m->H;
  std::vector<uint8_t> occ(cells ? cells : 1);
  std::vector<int64_t> off((size_t)m->nbins + 1);
  std::vector<float> vals((size_t)m->nvalues ? (size_t)m->nvalues : 1);
  if (cells) RL_CUDA(cudaMemcpyAsync(occ.data(), m->d_occ, cells, cudaMemcpyDeviceToHost, m->stream));
  RL_CUDA(cudaMemcpyAsync(off.data(), m->d_offsets, sizeof(int64_t) * off.size(), cudaMemcpyDeviceToHost, m->stream));
  if (m->nvalues)
    RL_CUDA(cudaMemcpyAsync(vals.data(), m->d_values, sizeof(float) * (size_t)m->nvalues, cudaMemcpyDeviceToHost, m->stream));
  RL_CUDA(cudaStreamSynchronize(m->stream));
  CkptHeader h{};
  memcpy(h.magic, kMagic, 8);
  h.version = 1;
  h.td = m->td;
  h.W = m->W;
  h.H = m->H;
  h.max_range = m->max_range;
  h.pruned = m->pruned ? 1u : 0u;
  h.occ_hash = fnv1a(occ.data(), cells);
  h.nbins = m->nbins;
  h.nvalues = m->nvalues;
  FILE* f = fopen(path, "wb");
  if (!f) {
    set_error(std::string("cannot open for writing: ") + path);
    return RL_E_INVALID;
  }
  bool ok = fwrite(&h, sizeof h, 1, f) == 1;
  ok = ok && fwrite(m->h_widths.data(), sizeof(int), m->td, f) == m->td;
  ok = ok && fwrite(m->h_trans.data(), sizeof(float), m->td, f) == m->td;
  ok = ok && fwrite(off.data(), sizeof(int64_t), off.size(), f) == off.size();
  ok = ok && (m->nvalues == 0 || fwrite(vals.data(), sizeof(float), (size_t)m->nvalues, f) == (size_t)m->nvalues);
  ok = (fclose(f) == 0) && ok;
  if (!ok) {
    set_error(std::string("short write: ") + path);
    return RL_E_INVALID;
  }
  return RL_OK;
}

// fills the CDDT part of a handle whose occupancy is already resident from a checkpoint written by cddt_save
int cddt_load(rl_method* m, const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    set_error(std::string("cannot open: ") + path);
    return RL_E_INVALID;
  }
  struct Closer {
    FILE* f;
    ~Closer() { fclose(f); }
  } closer{f};
  CkptHeader h{};
  if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, kMagic, 8) != 0 || h.version != 1) {
    set_error("not a CDDT checkpoint (magic / version)");
    return RL_E_INVALID;
  }
  if (h.W != m->W || h.H != m->H || h.td == 0 || h.nbins < 0 || h.nvalues < 0) {
    set_error("CDDT checkpoint: map size differs from the map passed in");
    return RL_E_INVALID;
  }
  const size_t cells = (size_t)m->W * m->H;
  std::vector<uint8_t> occ(cells ? cells : 1);
  if (cells) RL_CUDA(cudaMemcpyAsync(occ.data(), m->d_occ, cells, cudaMemcpyDeviceToHost, m->stream));
  RL_CUDA(cudaStreamSynchronize(m->stream));
  if (fnv1a(occ.data(), cells) != h.occ_hash) {
    set_error("CDDT checkpoint: it was built from a different occupancy grid");
    return RL_E_INVALID;
  }
  cddt_free(m);
  m->td = h.td;
  m->max_range = h.max_range;
  int rc = upload_consts(m);  // per-slice constants are recomputed (host libm, as at build time) and cross-checked
  if (rc) return rc;
  std::vector<int> widths(h.td);
  std::vector<float> trans(h.td);
  if (fread(widths.data(), sizeof(int), h.td, f) != h.td || fread(trans.data(), sizeof(float), h.td, f) != h.td) {
    set_error("CDDT checkpoint: truncated");
    return RL_E_INVALID;
  }
  if (m->nbins != h.nbins || memcmp(widths.data(), m->h_widths.data(), sizeof(int) * h.td) != 0 ||
      memcmp(trans.data(), m->h_trans.data(), sizeof(float) * h.td) != 0) {
    set_error("CDDT checkpoint: slice geometry differs from what this build derives for the map");
    return RL_E_INVALID;
  }
  std::vector<int64_t> off((size_t)h.nbins + 1);
  std::vector<float> vals(cddt_values_alloc(h.nvalues), 0.0f);  // + the pad element / last aligned block cddt_cast may read
  if (fread(off.data(), sizeof(int64_t), off.size(), f) != off.size() ||
      (h.nvalues && fread(vals.data(), sizeof(float), (size_t)h.nvalues, f) != (size_t)h.nvalues)) {
    set_error("CDDT checkpoint: truncated");
    return RL_E_INVALID;
  }
  if (off[0] != 0 || off[(size_t)h.nbins] != h.nvalues) {
    set_error("CDDT checkpoint: inconsistent offsets");
    return RL_E_INVALID;
  }
  for (size_t i = 0; i < (size_t)h.nbins; ++i) {
    if (off[i] > off[i + 1]) {
      set_error("CDDT checkpoint: inconsistent offsets");
      return RL_E_INVALID;
    }
  }
  RL_CUDA(cudaMalloc(&m->d_offsets, sizeof(int64_t) * off.size()));
  RL_CUDA(cudaMalloc(&m->d_values, sizeof(float) * vals.size()));
  RL_CUDA(cudaMemcpyAsync(m->d_offsets, off.data(), sizeof(int64_t) * off.size(), cudaMemcpyHostToDevice, m->stream));
  RL_CUDA(cudaMemcpyAsync(m->d_values, vals.data(), sizeof(float) * vals.size(), cudaMemcpyHostToDevice, m->stream));
  RL_CUDA(cudaStreamSynchronize(m->stream));
  m->nvalues = h.nvalues;
  m->pruned = h.pruned != 0;
  return cddt_index_build(m);
}

void cddt_free(rl_method* m) {
  cudaFree(m->d_widths);
  cudaFree(m->d_trans);
  cudaFree(m->d_cosv);
  cudaFree(m->d_sinv);
  cudaFree(m->d_slice0);
  cudaFree(m->d_offsets);
  cudaFree(m->d_values);
  cudaFree(m->d_meta);
  cudaFree(m->d_skip);
  m->d_meta = nullptr;
  m->d_skip = nullptr;
  m->nskip = 0;
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
  RL_CUDA(cudaMalloc(&m->d_values, sizeof(float) * cddt_values_alloc(nvalues)));  // pad, see cddt_cast
  RL_CUDA(cudaMemsetAsync(m->d_values + nvalues, 0, sizeof(float) * (cddt_values_alloc(nvalues) - (size_t)nvalues), st));
  if (nbins > 0 && nvalues > 0) {
    compact_kernel<false><<<blocks_for(nbins * 32, 256), 256, 0, st>>>(d_sorted, d_raw_off, nbins, nullptr, m->d_offsets,
                                                                      m->d_values);
    count_launch();
    RL_CHECK_LAUNCH();
  }
  RL_CUDA(cudaStreamSynchronize(st));
  return cddt_index_build(m);
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
  if (e == cudaSuccess) e = cudaMalloc(&d_new_vals, sizeof(float) * cddt_values_alloc(nvalues));
  if (e != cudaSuccess) {
    cudaFree(d_new_off);
    return cuda_fail(e, "prune alloc", __FILE__, __LINE__);
  }
  cudaMemsetAsync(d_new_vals + nvalues, 0, sizeof(float) * (cddt_values_alloc(nvalues) - (size_t)nvalues), st);
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
  return cddt_index_build(m);
}

}  // namespace rl
