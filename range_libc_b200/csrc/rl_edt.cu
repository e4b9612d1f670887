// Distance transform built on the device, restating the REFERENCE's arithmetic -- not a "better"
// exact EDT: DistanceTransform(OMap*) RangeLib.h:345-373 feeding the Felzenszwalb-Huttenlocher
// lower-envelope code of vendor/distance_transform.h (:873-910 pass order, :1054-1095 1-D kernel,
// :1117-1121 sqrt).  The reference's float rounding of q*q makes its result differ from the
// exact EDT on maps with a side > 4096 px; the per-scanline recurrence below reproduces that
// bit for bit (f32 sums, f64 intersection, f32 fill), so it stays sequential along a scanline
// and parallel across scanlines: one thread per scanline, envelope stack in global scratch,
// one contiguous run of 16-byte entries per scanline (its hot top stays in L1).
#include <float.h>

#include <algorithm>
#include <cstdlib>

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
#define EDT_WIN 16  // entries of a scanline's stack kept in shared memory (per thread)
// Round 2, second step (after ncu / timing: ~1 us per element even with prefetched inputs): what a thread waits for is
// the POP -- a load of an entry it stored a few steps earlier, which misses L1 (stores do not allocate) and comes back
// from L2 -- multiplied by the warp's divergence (every lane waits for the lane with the most pops).  So
//  * the most recent EDT_WIN entries of every scanline's stack live in a shared-memory ring (older entries are
//    spilled to the global stack, which the fill phase reads anyway; a pop below the ring reloads one entry), and
//  * pass 1 (FROM_OCC: f is 0 on occupied cells and FLT_MAX elsewhere) skips free cells altogether.  That is exact:
//    a free cell's parabola lies 3.4e38 above every occupied cell's, so in the reference's recurrence a free cell is
//    pushed with an intersection of +1.7e38 / (q - v) (beyond any scanline) or -- after another free cell, s = 0 -- pops
//    that one first, and the next occupied cell pops it again (s = -1.7e38 / .. <= anything on the stack above entry
//    0, which is never popped): free cells never change the entries of occupied cells, and their own intervals
//    ((-inf, -1.7e38 / q) for a free cell 0, (+1.7e38 / .., ..) for a trailing one) hold no integer position.
template <bool FROM_OCC, bool FINAL_SQRT>
__global__ void __launch_bounds__(32)
edt_pass_kernel(const uint8_t* __restrict__ occ, const float* __restrict__ fin, float* __restrict__ out, int nlines,
                int n, long long in_ls, long long in_es, long long out_ls, long long out_es,
                EdtEntry* __restrict__ stack) {
  __shared__ int w_v[EDT_WIN][32];
  __shared__ float w_f[EDT_WIN][32];
  __shared__ double w_z[EDT_WIN][32];
  const int lane = threadIdx.x;
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= nlines) return;
  const long long ib = (long long)line * in_ls, ob = (long long)line * out_ls;
  if (n == 1) {  // distance_transform.h:1058-1062
    float v = edt_load<FROM_OCC>(occ, fin, ib, in_es, 0);
    out[ob] = FINAL_SQRT ? __fsqrt_rn(v) : v;
    return;
  }
  EdtEntry* __restrict__ stk = stack + (size_t)line * ((size_t)n + 1);
  // lower envelope (:1072-1083).  The top entry lives in registers; entries [wlo, k) in the ring, [0, wlo) in scratch.
  int k = 0, wlo = 0;
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
      if (FROM_OCC && cur[j] != 0.0f) continue;  // a free cell (see above)
      const float fq = fadd(cur[j], (float)((unsigned long long)q * q));
      double s;
      while (true) {
        s = __ddiv_rn(__dsub_rn((double)fq, (double)fv_top),
                      __dsub_rn((double)(2 * (long long)q), (double)(2 * (long long)v_top)));
        if (s <= z_top && k > 0) {  // pop
          --k;
          if (k < wlo) {  // below the ring: one entry back from scratch
            const EdtEntry e = stk[k];
            v_top = e.v;
            fraw_top = e.fraw;
            z_top = e.z;
            wlo = k;
          } else {
            const int slot = k & (EDT_WIN - 1);
            v_top = w_v[slot][lane];
            fraw_top = w_f[slot][lane];
            z_top = w_z[slot][lane];
          }
          fv_top = fadd(fraw_top, (float)((unsigned long long)v_top * v_top));
          continue;
        }
        break;
      }
      // push the old top down
      if (k - wlo == EDT_WIN) {  // ring full: its oldest entry goes to scratch
        const int slot = wlo & (EDT_WIN - 1);
        stk[wlo] = EdtEntry{w_v[slot][lane], w_f[slot][lane], w_z[slot][lane]};
        ++wlo;
      }
      {
        const int slot = k & (EDT_WIN - 1);
        w_v[slot][lane] = v_top;
        w_f[slot][lane] = fraw_top;
        w_z[slot][lane] = z_top;
      }
      ++k;
      v_top = q;
      fraw_top = cur[j];
      fv_top = fq;
      z_top = s;
    }
  }
  for (int i = wlo; i < k; ++i) {  // the fill phase reads the whole stack from scratch
    const int slot = i & (EDT_WIN - 1);
    stk[i] = EdtEntry{w_v[slot][lane], w_f[slot][lane], w_z[slot][lane]};
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

// ------------------------------------------------------------------------------------------
// Integer form of the same recurrence (maps with sides <= 16384; round 2, third step).
//
// ncu on the kernel above: ~200 warp-instructions per element at one warp per scheduler (every instruction waits
// for the one before, ~7 cycles each) -- the double-precision division behind every intersection, the 64-bit
// conversions around it and the reloads of popped entries.  None of it is needed to take the reference's decisions:
//  * every finite input is an integer-valued float: 0 on occupied cells, and after pass 1 the float (q - v)^2, which
//    is an integer even where it was rounded (floats >= 2^24 are integers); so g(q) = fl(f[q] + fl(q^2)), formed with
//    the reference's two float operations, is an integer below 2^30;
//  * an intersection is s = fl((g(q) - g(v)) / (2q - 2v)) = fl(n / d) with integers |n| < 2^30, 0 < d <= 2^15, and the
//    reference only ever COMPARES intersections: s <= z[k] when it pops, z[k+1] < q when it fills.  Two different
//    rationals n1/d1 != n2/d2 of this size differ by at least 1 / (d1 d2) >= 2^-30, i.e. by more than 2^-52 of their
//    magnitude (< 2^30 * 2^15 / 2^52 of it would be needed for both to round to one double), so
//        fl(n1/d1) <= fl(n2/d2)  <=>  n1 d2 <= n2 d1        and        fl(n/d) < q  <=>  n < q d
//    in 64-bit integers -- the same decisions, no division;
//  * FLT_MAX inputs (free cells in pass 1; in pass 2 the cells of a column without any obstacle) are skipped.  Such a
//    parabola lies 3.4e38 above every finite one: pushed, its interval starts at +1.7e38 / (q - v), beyond any
//    scanline, and the next finite element pops it again (s = -1.7e38 / .. is below every intersection on the stack,
//    and entry 0 is never popped); after another infinite element (s = 0) it is replaced by it.  So infinite elements
//    never change an entry of a finite one and own no integer position -- except element 0, which stays at the bottom
//    of the stack; the first finite parabola above it starts at -1.7e38 / q, kept as "d = 0: minus infinity".
// The fill (f[v] + (q - v)^2 in float, then sqrt) is unchanged.  Maps with a side above 16384 take the kernel above.
// ------------------------------------------------------------------------------------------
struct __align__(16) EdtEntryI {
  int v;
  float fraw;
  int n, d;  // the entry's interval starts at n / d; d == 0: at minus infinity
};

__device__ __forceinline__ bool edt_le(int n1, int d1, int n2, int d2) {  // n1/d1 <= n2/d2 (d1 > 0, d2 >= 0)
  return d2 != 0 && (long long)n1 * d2 <= (long long)n2 * d1;
}

// scanlines of length 1 (a map one cell wide / high): copied through (distance_transform.h:1058-1062)
template <bool FROM_OCC, bool FINAL_SQRT>
__global__ void edt_pass_int1_kernel(const uint8_t* __restrict__ occ, const float* __restrict__ fin, float* __restrict__ out,
                                     int nlines, long long in_ls, long long out_ls) {
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= nlines) return;
  const float v = edt_load<FROM_OCC>(occ, fin, (long long)line * in_ls, 1LL, 0);
  out[(long long)line * out_ls] = FINAL_SQRT ? __fsqrt_rn(v) : v;
}

// one envelope step of the integer recurrence for a finite element (q, raw value fq_raw); everything by reference
struct EdtTop {
  int v, g, n, d;
  float fraw;
  bool inf;
};

// Envelope phase: one thread per scanline SEGMENT (the recurrence is sequential along it); writes the segment's stack
// to scratch (entries 0 .. ktop) and ktop[line * nseg + seg].  The fill phase is a separate, fully parallel kernel.
//
// Segments (round 2, fifth step).  With exact comparisons the reference's envelope is the exact lower envelope of the
// parabolas P_v(x) = g(v) - 2 v x + x^2, and its fill gives position x to the parabola that is lowest there, the
// LEFTMOST one among equals (`while (z[k+1] < q) k++` stays on the left entry when an intersection falls on x; a
// parabola whose interval has shrunk to that single point was popped, `s <= z[k]`).  That owner does not depend on the
// order in which parabolas were offered, so a scanline of an n-cell map can be cut into nseg pieces whose envelopes are
// built independently -- nseg times more threads for the only sequential part of the transform (a 1200^2 map has just
// 1200 scanlines for 148 SMs) -- and the fill takes, per position, the owner each piece proposes and keeps the one with
// the smallest integer key g(v) - 2 v x (first piece on ties = smallest v).  Positions and squares stay global.
template <bool FROM_OCC>
__global__ void __launch_bounds__(32)
edt_envelope_int_kernel(const uint8_t* __restrict__ occ, const float* __restrict__ fin, int nlines, int n,
                        long long in_ls, long long in_es, EdtEntryI* __restrict__ stack, int* __restrict__ ktop_out,
                        int nseg, int seglen) {
  __shared__ int w_v[EDT_WIN][32];
  __shared__ float w_f[EDT_WIN][32];
  __shared__ int w_n[EDT_WIN][32];
  __shared__ int w_d[EDT_WIN][32];
  __shared__ int w_g[EDT_WIN][32];  // g of the entry (recomputed when an entry comes back from scratch)
  const int lane = threadIdx.x;
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  const int seg = blockIdx.y;
  if (line >= nlines) return;
  const int q_lo = seg * seglen, q_hi = min(n, q_lo + seglen);  // the host makes every segment non-empty
  const long long ib = (long long)line * in_ls;
  EdtEntryI* __restrict__ stk = stack + ((size_t)line * nseg + seg) * ((size_t)seglen + 1);
  // lower envelope (:1072-1083).  The top entry lives in registers; entries [wlo, k) in the ring, [0, wlo) in scratch.
  int k = 0, wlo = 0;
  EdtTop top;
  top.v = q_lo;
  top.fraw = edt_load<FROM_OCC>(occ, fin, ib, in_es, q_lo);
  top.inf = top.fraw == FLT_MAX;
  top.g = top.inf ? 0 : __float2int_rz(fadd(top.fraw, (float)((unsigned)q_lo * (unsigned)q_lo)));
  top.n = 0;
  top.d = 0;

  auto step = [&](int q, float raw) {  // a finite element
    const int gq = __float2int_rz(fadd(raw, (float)((unsigned)q * (unsigned)q)));
    int sn = -1, sd = 0;
    while (!top.inf) {  // only entry 0 can be infinite: above it the new interval starts at minus infinity
      sn = gq - top.g;
      sd = 2 * (q - top.v);
      if (!(k > 0 && edt_le(sn, sd, top.n, top.d))) break;
      --k;  // pop
      if (k < wlo) {  // below the ring: one entry back from scratch
        const EdtEntryI e = stk[k];
        top.v = e.v; top.fraw = e.fraw; top.n = e.n; top.d = e.d;
        top.inf = top.fraw == FLT_MAX;
        top.g = top.inf ? 0 : __float2int_rz(fadd(top.fraw, (float)((unsigned)top.v * (unsigned)top.v)));
        wlo = k;
      } else {
        const int slot = k & (EDT_WIN - 1);
        top.v = w_v[slot][lane]; top.fraw = w_f[slot][lane]; top.n = w_n[slot][lane]; top.d = w_d[slot][lane];
        top.g = w_g[slot][lane];
        top.inf = top.fraw == FLT_MAX;
      }
      sn = -1;
      sd = 0;
    }
    if (k - wlo == EDT_WIN) {  // ring full: its oldest entry goes to scratch
      const int slot = wlo & (EDT_WIN - 1);
      stk[wlo] = EdtEntryI{w_v[slot][lane], w_f[slot][lane], w_n[slot][lane], w_d[slot][lane]};
      ++wlo;
    }
    const int slot = k & (EDT_WIN - 1);  // push the old top down
    w_v[slot][lane] = top.v; w_f[slot][lane] = top.fraw; w_n[slot][lane] = top.n; w_d[slot][lane] = top.d;
    w_g[slot][lane] = top.g;
    ++k;
    top.v = q; top.fraw = raw; top.inf = false; top.g = gq; top.n = sn; top.d = sd;
  };

  if (FROM_OCC) {
    // f is 0 on occupied cells and FLT_MAX elsewhere; free cells are skipped (see above), sixteen at a time where
    // the column's bytes allow 16-byte loads (in_es == 1: a column is contiguous)
    const uint8_t* col = occ + ib;
    int q = q_lo + 1;
    if (in_es == 1) {
      while (q < q_hi && ((uintptr_t)(col + q) & 15)) {
        if (col[q]) step(q, 0.0f);
        ++q;
      }
      for (; q + 16 <= q_hi; q += 16) {
        const uint4 wd = *reinterpret_cast<const uint4*>(col + q);
        if ((wd.x | wd.y | wd.z | wd.w) == 0) continue;
        const unsigned ws[4] = {wd.x, wd.y, wd.z, wd.w};
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if ((ws[i >> 2] >> (8 * (i & 3))) & 0xffu) step(q + i, 0.0f);
      }
    }
    for (; q < q_hi; ++q)
      if (occ[ib + q * in_es]) step(q, 0.0f);
  } else {
    float nxt[EDT_PF];
#pragma unroll
    for (int j = 0; j < EDT_PF; ++j) nxt[j] = (q_lo + 1 + j < q_hi) ? fin[ib + (q_lo + 1 + j) * in_es] : FLT_MAX;
    for (int q0 = q_lo + 1; q0 < q_hi; q0 += EDT_PF) {
      float cur[EDT_PF];
#pragma unroll
      for (int j = 0; j < EDT_PF; ++j) cur[j] = nxt[j];
#pragma unroll
      for (int j = 0; j < EDT_PF; ++j) {
        const int qn = q0 + EDT_PF + j;
        nxt[j] = (qn < q_hi) ? fin[ib + qn * in_es] : FLT_MAX;
      }
#pragma unroll
      for (int j = 0; j < EDT_PF; ++j) {
        const int q = q0 + j;
        if (q < q_hi && cur[j] != FLT_MAX) step(q, cur[j]);  // infinite elements are skipped (see above)
      }
    }
  }
  for (int i = wlo; i < k; ++i) {
    const int slot = i & (EDT_WIN - 1);
    stk[i] = EdtEntryI{w_v[slot][lane], w_f[slot][lane], w_n[slot][lane], w_d[slot][lane]};
  }
  stk[k] = EdtEntryI{top.v, top.fraw, top.n, top.d};
  ktop_out[(size_t)line * nseg + seg] = k;
}

// Fill phase (:1086-1091): one thread per output element.  Within a segment, position q belongs to the last stack
// entry whose interval starts below q -- the reference walks the entries with `while (z[k+1] < q) k++`; interval
// starts increase along the stack, so that entry is found by bisection (integer comparison n < q d, see above).  Among
// the segments' owners the lowest parabola at q wins (see the envelope kernel).
// A CTA covers 32 positions x 32 scanlines.  Each warp searches one scanline at a time with its lanes on 32 consecutive
// positions: their bisections read the same few stack entries (one or two sectors per request instead of 32).
// OUT_ALONG_Q: the output is contiguous along the scanline (pass 1) and is stored directly; otherwise (pass 2: the
// output is contiguous ACROSS scanlines) the tile is transposed through shared memory so that the stores coalesce too.
template <bool OUT_ALONG_Q, bool FINAL_SQRT>
__global__ void __launch_bounds__(256)
edt_fill_int_kernel(float* __restrict__ out, int nlines, int n, long long out_ls, long long out_es,
                    const EdtEntryI* __restrict__ stack, const int* __restrict__ ktop, int nseg, int seglen) {
  __shared__ float tile[32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int q0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
  const int q = q0 + lane;
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const int ll = w + 8 * i, line = l0 + ll;
    float D = 0.0f;
    const bool valid = q < n && line < nlines;
    if (valid) {
      long long best = 0;
      int best_v = 0;
      float best_f = FLT_MAX;
      bool have = false;
      for (int s = 0; s < nseg; ++s) {
        const EdtEntryI* __restrict__ stk = stack + ((size_t)line * nseg + s) * ((size_t)seglen + 1);
        int lo = 0, hi = __ldg(ktop + (size_t)line * nseg + s);  // invariant: entry lo starts below q, entries > hi do not
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          const int2 nd = __ldg(reinterpret_cast<const int2*>(&stk[mid].n));
          if (nd.y == 0 || (long long)nd.x < (long long)q * nd.y) lo = mid; else hi = mid - 1;
        }
        const int2 vf = __ldg(reinterpret_cast<const int2*>(&stk[lo].v));
        const int v = vf.x;
        const float fraw = __int_as_float(vf.y);
        if (fraw == FLT_MAX) {  // an all-infinite stretch: it owns the position only if nothing finite exists (then segment 0's)
          if (s == 0) best_v = v;
          continue;
        }
        const long long key = (long long)__float2int_rz(fadd(fraw, (float)((unsigned)v * (unsigned)v))) - 2LL * v * q;
        if (!have || key < best) {
          best = key;
          best_v = v;
          best_f = fraw;
          have = true;
        }
      }
      const float dq = fsub((float)q, (float)best_v);
      D = fadd(best_f, fmul(dq, dq));
      if (FINAL_SQRT) D = __fsqrt_rn(D);
    }
    if (OUT_ALONG_Q) {
      if (valid) out[(long long)line * out_ls + (long long)q * out_es] = D;
    } else {
      tile[ll][lane] = D;
    }
  }
  if (!OUT_ALONG_Q) {
    __syncthreads();
    const int line = l0 + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qq = w + 8 * i;
      if (q0 + qq < n && line < nlines) out[(long long)line * out_ls + (long long)(q0 + qq) * out_es] = tile[lane][qq];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Pass 1 without an envelope (columns of at most 4096 cells; round 2, fourth step).
// Pass 1 sees f = 0 on occupied cells and FLT_MAX elsewhere.  For q, v <= 4096 the reference's float q*q and v*v are
// exact, so the parabolas of two occupied cells v1 < v2 meet exactly at (v1 + v2) / 2 and the envelope gives every
// position to its NEAREST occupied cell of the column (a position exactly half way between two gets the left one,
// `z[k+1] < q` is false at equality; both give the same value); free cells own no position (see the integer form
// above), and a column without obstacle keeps entry 0 = FLT_MAX everywhere (FLT_MAX + dq^2 rounds to FLT_MAX).  The
// value written is the reference's fill expression f[v] + (q - v)^2 = 0 + fl(dq * dq), exact here.  So the pass is
// two nearest-set-bit queries per cell: one CTA per column, the column as a bit mask in shared memory plus the
// running "last set bit at or before word w" / "first set bit at or after word w".
// ------------------------------------------------------------------------------------------
#define EDT_DIRECT_MAX 4096
__global__ void __launch_bounds__(128)
edt_pass1_direct_kernel(const uint8_t* __restrict__ occ, float* __restrict__ out, int n) {
  __shared__ unsigned mask[EDT_DIRECT_MAX / 32];
  __shared__ int last_le[EDT_DIRECT_MAX / 32], first_ge[EDT_DIRECT_MAX / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint8_t* __restrict__ col = occ + (size_t)blockIdx.x * n;
  float* __restrict__ dst = out + (size_t)blockIdx.x * n;
  const int nw = (n + 31) >> 5;
  const int FAR = 1 << 20;
  for (int w = warp; w < nw; w += 4) {
    const int q = 32 * w + lane;
    const unsigned m = __ballot_sync(0xffffffffu, q < n && col[q] != 0);
    if (lane == 0) mask[w] = m;
  }
  __syncthreads();
  if (warp < 2) {  // warp 0: last set bit at or before the end of word w; warp 1: first set bit at or after its start
    const int per = (nw + 31) >> 5;  // words per lane (<= 4)
    const int w0 = lane * per;
    int run = warp == 0 ? -FAR : FAR;
    for (int i = 0; i < per; ++i) {
      const int w = warp == 0 ? w0 + i : w0 + per - 1 - i;
      if (w < nw && mask[w]) run = warp == 0 ? 32 * w + 31 - __clz(mask[w]) : min(run, 32 * w + __ffs(mask[w]) - 1);
    }
    // inclusive scan over lanes: max from the left (warp 0), min from the right (warp 1)
    for (int d = 1; d < 32; d <<= 1) {
      const int o = warp == 0 ? __shfl_up_sync(0xffffffffu, run, d) : __shfl_down_sync(0xffffffffu, run, d);
      if (warp == 0 ? lane >= d : lane + d < 32) run = warp == 0 ? max(run, o) : min(run, o);
    }
    int carry = warp == 0 ? __shfl_up_sync(0xffffffffu, run, 1) : __shfl_down_sync(0xffffffffu, run, 1);
    if (warp == 0 ? lane == 0 : lane == 31) carry = warp == 0 ? -FAR : FAR;
    for (int i = 0; i < per; ++i) {
      const int w = warp == 0 ? w0 + i : w0 + per - 1 - i;
      if (w < nw) {
        if (mask[w]) carry = warp == 0 ? 32 * w + 31 - __clz(mask[w]) : 32 * w + __ffs(mask[w]) - 1;
        (warp == 0 ? last_le : first_ge)[w] = carry;
      }
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < n; q += 128) {
    const int w = q >> 5, b = q & 31;
    const unsigned below = mask[w] & (0xffffffffu >> (31 - b)), above = mask[w] & (0xffffffffu << b);
    const int prev = below ? 32 * w + 31 - __clz(below) : (w > 0 ? last_le[w - 1] : -FAR);
    const int next = above ? 32 * w + __ffs(above) - 1 : (w + 1 < nw ? first_ge[w + 1] : FAR);
    const int dist = min(q - prev, next - q);
    const float dq = (float)dist;
    dst[q] = dist >= FAR / 2 ? FLT_MAX : fadd(0.0f, fmul(dq, dq));
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
  const size_t stack_elems = (nmax + 8) * nmax;  // (n + 1) entries per scanline, or (seglen + 1) per segment, <= 8 segments
  RL_CUDA(cudaMalloc(&d_tmp, sizeof(float) * cells));
  cudaError_t ea = cudaMalloc(&stack, sizeof(EdtEntry) * stack_elems);
  if (ea != cudaSuccess) {
    cudaFree(d_tmp);
    return cuda_fail(ea, "edt scratch", __FILE__, __LINE__);
  }
  const int threads = 32;
  // RL_EDT_EXACT_DIV=1 (tests): the double-precision form of the recurrence on every map
  const bool integer_form = nmax <= 16384 && !(getenv("RL_EDT_EXACT_DIV") && atoi(getenv("RL_EDT_EXACT_DIV")) != 0);
  // pass 1: for each x a scanline along y (slices of dimension 0 first, :893-900)
  // pass 2: for each y a scanline along x; result square-rooted (:1117-1121)
  if (integer_form) {
    static_assert(sizeof(EdtEntryI) == sizeof(EdtEntry), "the two entry types share the scratch allocation");
    int* ktop = nullptr;
    cudaError_t ek = cudaMalloc(&ktop, sizeof(int) * nmax * 8);
    if (ek != cudaSuccess) {
      cudaFree(d_tmp);
      cudaFree(stack);
      return cuda_fail(ek, "edt scratch", __FILE__, __LINE__);
    }
    EdtEntryI* stk = (EdtEntryI*)stack;
    // segments per scanline: enough threads for the chip (~64 per SM), pieces of at least 64 cells, at most 8
    auto segments = [&](int nlines, int n, int* nseg, int* seglen) {
      const int forced = getenv("RL_EDT_SEGMENTS") ? atoi(getenv("RL_EDT_SEGMENTS")) : 0;
      int want = forced > 0 ? forced : (9472 + nlines - 1) / nlines;
      want = std::max(1, std::min(std::min(want, 8), n / 64));
      if (want < 1) want = 1;
      *seglen = (n + want - 1) / want;
      *nseg = (n + *seglen - 1) / *seglen;
    };
    int ns1 = 1, sl1 = H, ns2 = 1, sl2 = W;
    segments(W, H, &ns1, &sl1);
    segments(H, W, &ns2, &sl2);
    const bool direct1 = H <= EDT_DIRECT_MAX && !(getenv("RL_EDT_DIRECT_PASS1") && atoi(getenv("RL_EDT_DIRECT_PASS1")) == 0);
    if (H == 1) {  // scanlines of length 1 are copied through (distance_transform.h:1058-1062)
      edt_pass_int1_kernel<true, false><<<(W + 255) / 256, 256, 0, m->stream>>>(m->d_occ, nullptr, d_tmp, W, (long long)H, (long long)H);
    } else if (direct1) {
      edt_pass1_direct_kernel<<<W, 128, 0, m->stream>>>(m->d_occ, d_tmp, H);
    } else {
      edt_envelope_int_kernel<true><<<dim3((W + threads - 1) / threads, ns1), threads, 0, m->stream>>>(
          m->d_occ, nullptr, W, H, (long long)H, 1LL, stk, ktop, ns1, sl1);
      edt_fill_int_kernel<true, false><<<dim3((H + 31) / 32, (W + 31) / 32), 256, 0, m->stream>>>(d_tmp, W, H, (long long)H, 1LL, stk, ktop, ns1, sl1);
    }
    if (W == 1) {
      edt_pass_int1_kernel<false, true><<<(H + 255) / 256, 256, 0, m->stream>>>(nullptr, d_tmp, m->d_dt, H, 1LL, 1LL);
    } else {
      edt_envelope_int_kernel<false><<<dim3((H + threads - 1) / threads, ns2), threads, 0, m->stream>>>(
          nullptr, d_tmp, H, W, 1LL, (long long)H, stk, ktop, ns2, sl2);
      edt_fill_int_kernel<false, true><<<dim3((W + 31) / 32, (H + 31) / 32), 256, 0, m->stream>>>(m->d_dt, H, W, 1LL, (long long)H, stk, ktop, ns2, sl2);
    }
    count_launch(4);
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(m->stream);
    cudaFree(ktop);
    cudaFree(d_tmp);
    cudaFree(stack);
    if (e != cudaSuccess) return cuda_fail(e, "edt launch", __FILE__, __LINE__);
    if (e2 != cudaSuccess) return cuda_fail(e2, "edt sync", __FILE__, __LINE__);
    return RL_OK;
  } else {
    edt_pass_kernel<true, false><<<(W + threads - 1) / threads, threads, 0, m->stream>>>(
        m->d_occ, nullptr, d_tmp, W, H, (long long)H, 1LL, (long long)H, 1LL, stack);
    edt_pass_kernel<false, true><<<(H + threads - 1) / threads, threads, 0, m->stream>>>(
        nullptr, d_tmp, m->d_dt, H, W, 1LL, (long long)H, 1LL, (long long)H, stack);
  }
  count_launch(2);
  cudaError_t e = cudaGetLastError();
  cudaError_t e2 = cudaStreamSynchronize(m->stream);
  cudaFree(d_tmp);
  cudaFree(stack);
  if (e != cudaSuccess) return cuda_fail(e, "edt launch", __FILE__, __LINE__);
  if (e2 != cudaSuccess) return cuda_fail(e2, "edt sync", __FILE__, __LINE__);
  return RL_OK;
}

}  // namespace rl
