// Batched range queries: the hot path.
//
//   cast_kernel<KIND, MODE>   one ray per thread; MODE selects the reference entry point
//                             GRID   RayMarchingGPU::calc_range_many      RangeLib.h:819-831
//                             WORLD  RangeMethod::numpy_calc_range        RangeLib.h:439-480
//                             ANGLES RangeMethod::numpy_calc_range_angles RangeLib.h:482-520
//                             RM marches at CTA level (rm_march_block: own-ray bursts, then the last live
//                             rays are finished cooperatively, one per warp)
//   rm_persist_kernel<MODE,..> RM, launches many waves deep: persistent warps with lane re-queuing, predicated
//                             straight-line step bursts, no shared memory; also fills the GiantLUT table
//                             (MODE_GLT_BUILD)
//   fused_kernel<KIND,..>     RangeMethod::calc_range_repeat_angles_eval_sensor_model :558-612
//                             a CTA owns whole particles; ranges go straight to the sensor
//                             table lookup, the per-particle product is formed in the
//                             reference's order (beam 0..M-1) so weights are bit-identical;
//                             multi-GPU epilogues (peer stores, epoch flags) live here too
//   fused_rm_persist_kernel   the same call for RM clouds many waves deep: re-queuing inside particle groups
//                             (short fans on structures within L2)
//   fused_overlap_kernel<KIND> the same call, deep launches below ~1.8 M rays: the products on a ninth warp
//                             behind named barriers, two value buffers
//   launch_fused_twostep      the same call, deep launches from ~1.8 M rays on: the kind's big-batch cast
//                             into a scratch array + eval_overlap_kernel (config 5 runs this way)
//   radial_kernel<KIND>       RangeMethod::calc_range_many_radial_optimized :616-676 (CDDT calc_range_pair)
//   eval_sensor_kernel        RangeMethod::eval_sensor_model              RangeLib.h:533-555 (small calls)
//   eval_overlap_kernel       the same for big batches and as the second half of a deep fused update:
//                             loader warps own particle rows (table row fixed per lane), a ninth warp
//                             multiplies in beam order; multi-GPU epilogues as in fused_kernel
//
// KIND: RL_BL (:696-769), RL_RM (:927-962), RL_CDDT / RL_PCDDT (:1342-1516), RL_GLT (:1869-1880).
#include <cstdlib>

#include <type_traits>

#include "rl_internal.cuh"
#include "rl_math.cuh"

namespace rl {

// Development aid (tools/trace_fused.py builds a separate library with -DRL_TRACE; the product library never
// defines it): per-CTA globaltimer stamps at the phase boundaries of the small fused launch.
#ifdef RL_TRACE
__device__ unsigned long long* g_rl_trace = nullptr;
__device__ __forceinline__ void rl_trace_mark(int slot) {
  if (g_rl_trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_rl_trace[(size_t)blockIdx.x * 16 + slot] = t;
  }
}
__device__ __forceinline__ void rl_trace_value(int slot, unsigned long long v) {
  if (g_rl_trace && threadIdx.x == 0) g_rl_trace[(size_t)blockIdx.x * 16 + slot] = v;
}
__device__ __forceinline__ void rl_trace_add(int slot, unsigned long long v) {  // any thread
  if (g_rl_trace) atomicAdd(&g_rl_trace[(size_t)blockIdx.x * 16 + slot], v);
}
#define RL_TRACE_MARK(slot) rl_trace_mark(slot)
#define RL_TRACE_VALUE(slot, v) rl_trace_value(slot, (unsigned long long)(v))
#define RL_TRACE_ADD(slot, v) rl_trace_add(slot, (unsigned long long)(v))
__device__ __forceinline__ void rl_trace_max(int slot, unsigned long long v) {  // any thread
  if (g_rl_trace) atomicMax(&g_rl_trace[(size_t)blockIdx.x * 16 + slot], v);
}
#else
#define RL_TRACE_MARK(slot)
#define RL_TRACE_VALUE(slot, v)
#define RL_TRACE_ADD(slot, v)
#endif

// ------------------------------------------------------------------------------------------
// single-ray device functions
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ bool finite3(float a, float b, float c) {
  return (fabsf(a) <= 3.402823466e38f) && (fabsf(b) <= 3.402823466e38f) && (fabsf(c) <= 3.402823466e38f);
}

// ---- RayMarching::calc_range, RangeLib.h:927-962; distThreshold 0.0, step_coeff 0.999f (:967-968) ----
//
// rm_step: one iteration of the reference's while loop for one ray (one dependent read of the
// float distance transform).  Returns true when the ray has ended and `result` is final.
__device__ __forceinline__ bool rm_step(const MapView& mv, float max_range, float x0, float y0, float dx, float dy,
                                        float& t, float& result) {
  const int px = __float2int_rz(fadd(x0, fmul(dx, t)));
  const int py = __float2int_rz(fadd(y0, fmul(dy, t)));
  result = max_range;
  if ((unsigned)px >= (unsigned)mv.W || (unsigned)py >= (unsigned)mv.H) return true;
  const float d = __ldg(mv.dt + dt_index(px, py, mv.H));
  if (d <= 0.0f) {
    const float xd = fsub((float)px, x0), yd = fsub((float)py, y0);
    result = __fsqrt_rn(fadd(fmul(xd, xd), fmul(yd, yd)));
    return true;
  }
  t = fadd(t, fmaxf(fmul(d, 0.999f), 1.0f));
  return !(t < max_range);
}

// Cooperative march of ONE ray by all 32 lanes of a warp (same x0, y0, dx, dy, t in every lane).
//
// Sphere tracing is a chain of dependent reads: on B200 an L2 hit is ~140 ns and a lone warp
// needs another ~70 ns of dependent ALU work per step, so a ray that crawls along a wall for
// 150-300 steps holds its warp -- and, in a small launch such as a 4000 x 60 particle-filter
// update, the whole kernel -- for 30-60 us while 31 lanes idle.  Here the idle lanes turn the
// chain into batches: the ray is a straight line, so the cells it can visit over the next
// ~48 px are known before any distance is.  Every lane reads the distance at two parameters
// t + 0.75 (32 p + j) (one L2 round trip for all 64 probes); the warp then replays the reference's stepping
// sequence out of registers, finding the cell of each exact sample position
// (px, py) = (int)(x0 + dx t), (int)(y0 + dy t) among the lanes with a ballot.  A sample whose
// cell was not prefetched (a corner clipped between two probes, or a jump past the window) just
// starts the next batch there; lane 0 always probes the exact current sample, so every batch
// advances at least one step.  Arithmetic and step sequence are the reference's, untouched:
// the probes are a register-resident cache, never a source of different values.
#ifndef RL_COOP_SPACING
#define RL_COOP_SPACING 0.75f
#endif
#ifndef RL_COOP_PROBES
#define RL_COOP_PROBES 2  // probes per lane and batch: 64 probes = 48 px of ray per L2 round trip
#endif
#ifndef RL_COOP_UNROLL
#define RL_COOP_UNROLL 2  // replay steps per exit test (2: 21.0 us, 4: 21.3, 8: 22.4 per 4000 x 60 update)
#endif
#define RL_STEP_INF 0x7f800000u  // +inf in the per-lane step table: "obstacle here" -- and the value of "no probe"

// Conservative parameter interval of one probed cell (round 2).  The replay loop of rm_march_coop has to know, for the
// current parameter t, which probed cell the reference's sample (int)(x0 + dx t), (int)(y0 + dy t) falls into.  Round 1
// recomputed that cell on every replay step (FMUL, FADD, F2I, key build, compare: half of a ~100-cycle dependent chain,
// tools/replay_bench.cu).  A cell is a parameter INTERVAL of the ray, so each lane can instead carry [lo, hi] for its
// probe and the replay step becomes two compares, a select and the warp-wide minimum (46 cycles measured).  Exactness:
// the interval is made a SUBSET of the cell's true float preimage --
//   fl(x0 + fl(dx t)) differs from the real number X(t) = x0 + dx t by at most E0 = 2^-24 (|dx t| + |X|) <=
//   2^-24 (max_range + side + 1); a parameter with X(t) in [cx + margin, cx + 1 - margin], margin = 8 E0, therefore
//   truncates to cx (for cx = 0 the reference's cell also covers (-1, 0); not claiming that part is conservative).
//   The interval ends are formed with three float operations each; their rounding (relative 2^-23 of the end, plus
//   2^-23 side in X, well inside the margin) is covered by moving the ends inwards by a relative 2^-20 --
// and a parameter no lane claims (a margin zone, a clipped corner, a jump past the window, the map edge) simply takes the
// exact round-1 test.  The claim therefore never decides differently from the exact computation; it only skips it.
struct CoopRay {
  float x0, y0, dx, dy, inv_dx, inv_dy, margin;
  float below_max;  // the largest float below max_range: no interval claims a parameter the reference's loop never samples
};

__device__ __forceinline__ CoopRay coop_ray(const MapView& mv, float max_range, float x0, float y0, float dx, float dy) {
  CoopRay r;
  r.x0 = x0; r.y0 = y0; r.dx = dx; r.dy = dy;
  r.inv_dx = __fdiv_rn(1.0f, dx);
  r.inv_dy = __fdiv_rn(1.0f, dy);
  r.below_max = nextafterf(max_range, -INFINITY);
  r.margin = fmul(4.76837158203125e-07f /* 2^-21 */, fadd(fadd(max_range, (float)max(mv.W, mv.H)), 2.0f));
  return r;
}

// [lo, hi] for one axis: parameters whose coordinate lies in [c + margin, c + 1 - margin]
__device__ __forceinline__ void coop_axis_interval(float c, float o, float d, float inv_d, float margin, float* lo,
                                                   float* hi) {
  const float a = fsub(c, o);                        // c - x0
  const float near = fmul(fadd(a, margin), inv_d);   // parameter of the edge at c (+ margin)
  const float far = fmul(fsub(fadd(a, 1.0f), margin), inv_d);  // parameter of the edge at c + 1 (- margin)
  if (d > 0.0f) { *lo = near; *hi = far; }
  else if (d < 0.0f) { *lo = far; *hi = near; }
  else { *lo = -INFINITY; *hi = INFINITY; }          // the coordinate never changes: fl(o + fl(0 t)) == o
}

__device__ __forceinline__ float rm_march_coop(const MapView& mv, float max_range, float x0, float y0, float dx,
                                               float dy, float t) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned W = (unsigned)mv.W, H = (unsigned)mv.H;
  const CoopRay cr = coop_ray(mv, max_range, x0, y0, dx, dy);
#ifdef RL_TRACE
  unsigned n_batches = 0, n_fast = 0, n_exact = 0, n_groups = 0;
  long long c_replay = 0, c_mark = 0;
  const long long c_enter = clock64();
#define RL_COOP_COUNT(x) (++(x))
#define RL_COOP_CLOCK_BEGIN() (c_mark = clock64())
#define RL_COOP_CLOCK_END() (c_replay += clock64() - c_mark)
  // slots 11-13: totals over the CTA's rays; 14 / 15: the CTA's longest single ray -- cycles in the tail packed with
  // its cycles inside replay loops (<< 32), and its steps packed with its batches (<< 32)
#define RL_COOP_REPORT()                                                                                   \
  do {                                                                                                     \
    if (lane == 0) {                                                                                       \
      RL_TRACE_ADD(11, n_batches);                                                                         \
      RL_TRACE_ADD(12, n_fast + n_groups * RL_COOP_UNROLL);                                                                            \
      RL_TRACE_ADD(13, n_exact);                                                                           \
      rl_trace_max(14, ((unsigned long long)(clock64() - c_enter) << 32) | (unsigned)c_replay);            \
      rl_trace_max(15, ((unsigned long long)(n_fast + n_groups * RL_COOP_UNROLL) << 32) | n_batches);                                    \
    }                                                                                                      \
  } while (0)
#else
#define RL_COOP_COUNT(x)
#define RL_COOP_CLOCK_BEGIN()
#define RL_COOP_CLOCK_END()
#define RL_COOP_REPORT()
#endif
  while (true) {
    RL_COOP_COUNT(n_batches);
    // ---- probe batch: probe p of lane j reads the cell at parameter t + 0.75 (32 p + j) and keeps the
    // step that cell implies.  The loads of a lane are independent (one L2 round trip for the batch).
    // Cells are identified by key = cx << 16 | cy; this path is only taken when W, H and max_range are
    // below 32768, so a sample outside the map (coordinate negative or >= the side) can never produce
    // the key of a probed cell; unused probes carry key -1, a step of +inf and an empty interval.
    int key[RL_COOP_PROBES];
    unsigned stepbits[RL_COOP_PROBES];
    float lo[RL_COOP_PROBES], hi[RL_COOP_PROBES];
    // Branch-free, in three sweeps, so that all loads of the batch are in flight together: a probe outside the
    // map reads cell 0 and is discarded (a branch around the load made ptxas serialise the two probes of a lane:
    // two L2 round trips per batch, measured +2.3 us on the 4000 x 60 update).
    float d[RL_COOP_PROBES];
    bool inb[RL_COOP_PROBES];
    int cxs[RL_COOP_PROBES], cys[RL_COOP_PROBES];
#pragma unroll
    for (int p = 0; p < RL_COOP_PROBES; ++p) {
      const float s = t + (float)(32 * p + lane) * RL_COOP_SPACING;  // p = 0, lane 0: exactly t
      cxs[p] = __float2int_rz(fadd(x0, fmul(dx, s)));
      cys[p] = __float2int_rz(fadd(y0, fmul(dy, s)));
      inb[p] = (unsigned)cxs[p] < W && (unsigned)cys[p] < H;
      d[p] = __ldg(mv.dt + (inb[p] ? dt_index(cxs[p], cys[p], mv.H) : 0u));
    }
#pragma unroll
    for (int p = 0; p < RL_COOP_PROBES; ++p) {
      float lx, hx, ly, hy;
      coop_axis_interval((float)cxs[p], x0, dx, cr.inv_dx, cr.margin, &lx, &hx);
      coop_axis_interval((float)cys[p], y0, dy, cr.inv_dy, cr.margin, &ly, &hy);
      // inwards by a relative 2^-20 (t >= 0: a lower end below 0 is 0)
      lo[p] = inb[p] ? fmul(fmaxf(fmaxf(lx, ly), 0.0f), 1.00000095367431640625f) : INFINITY;
      hi[p] = inb[p] ? fminf(fmul(fminf(hx, hy), 0.99999904632568359375f), cr.below_max) : -INFINITY;
      key[p] = inb[p] ? ((cxs[p] << 16) | cys[p]) : -1;
    }
#pragma unroll
    for (int p = 0; p < RL_COOP_PROBES; ++p)
      stepbits[p] = (inb[p] && !(d[p] <= 0.0f)) ? __float_as_uint(fmaxf(fmul(d[p], 0.999f), 1.0f)) : RL_STEP_INF;
    // ---- replay the reference's steps out of the probes: one warp-wide min-reduction per step.
    // A lane offers the step of its probe if the sample is certainly in the probe's cell (interval test), +inf
    // otherwise.  +inf from that reduction (obstacle, nobody sure) takes the exact test: a lane offers its step if
    // the probe IS the sampled cell.  +inf from the exact test (obstacle, cell not probed, or sample outside the
    // map) ends the loop through its only exit test and the three cases are told apart afterwards.
    float tp;
    RL_COOP_CLOCK_BEGIN();
    while (true) {
      // fast loop: RL_COOP_UNROLL steps of straight-line code per exit test (measured in place, round 2: of the
      // ~105 cycles of a replay step ~25 were the `claimed?` branch and ~25 the loop branch).  No interval reaches
      // max_range (hi is clamped below it), so the first parameter that is not `< max_range` -- a real step past
      // max_range, or +inf because nobody claimed the sample -- turns every later one into +inf: the parameters of
      // a group form a valid prefix followed by the stop, which is picked out afterwards.
      float ts[RL_COOP_UNROLL + 1];
      do {
        ts[0] = t;
#pragma unroll
        for (int u = 0; u < RL_COOP_UNROLL; ++u) {
          const float tu = ts[u];
          unsigned mine = (tu >= lo[0] && tu <= hi[0]) ? stepbits[0] : RL_STEP_INF;
#pragma unroll
          for (int p = 1; p < RL_COOP_PROBES; ++p) mine = (tu >= lo[p] && tu <= hi[p]) ? stepbits[p] : mine;
          ts[u + 1] = fadd(tu, __uint_as_float(__reduce_min_sync(FULL, mine)));
        }
        t = ts[RL_COOP_UNROLL];
        RL_COOP_COUNT(n_groups);
      } while (t < max_range);
      tp = ts[0];
      t = ts[1];
#pragma unroll
      for (int u = 1; u < RL_COOP_UNROLL; ++u) {
        if (ts[u] < max_range) {
          tp = ts[u];
          t = ts[u + 1];
        }
      }
      if (t != __uint_as_float(RL_STEP_INF)) break;  // a real step carried t to max_range
      // nobody was sure about the sample at tp: the exact test
      RL_COOP_COUNT(n_exact);
      const int ex = __float2int_rz(fadd(x0, fmul(dx, tp)));
      const int ey = __float2int_rz(fadd(y0, fmul(dy, tp)));
      const int k0 = (ex << 16) | ey;
      unsigned mine = (key[0] == k0) ? stepbits[0] : RL_STEP_INF;
#pragma unroll
      for (int p = 1; p < RL_COOP_PROBES; ++p) mine = (key[p] == k0) ? stepbits[p] : mine;
      t = fadd(tp, __uint_as_float(__reduce_min_sync(FULL, mine)));
      if (!(t < max_range)) break;  // +inf (obstacle, cell not probed, outside the map) or max_range reached
    }
    RL_COOP_CLOCK_END();
    const int px = __float2int_rz(fadd(x0, fmul(dx, tp)));
    const int py = __float2int_rz(fadd(y0, fmul(dy, tp)));
    if ((unsigned)px >= W || (unsigned)py >= H) { RL_COOP_REPORT(); return max_range; }  // left the map (RangeLib.h:942-944)
    if (t != __uint_as_float(RL_STEP_INF)) { RL_COOP_REPORT(); return max_range; }  // a real step carried t to max_range (:938)
    const int k0 = (px << 16) | py;
    bool probed = false;
#pragma unroll
    for (int p = 0; p < RL_COOP_PROBES; ++p) probed = probed || key[p] == k0;
    if (__any_sync(FULL, probed)) {  // probed and +inf: d <= distThreshold (:952-956)
      const float xd = fsub((float)px, x0), yd = fsub((float)py, y0);
      RL_COOP_REPORT();
      return __fsqrt_rn(fadd(fmul(xd, xd), fmul(yd, yd)));
    }
    t = tp;  // cell not probed: next batch starts at this sample
  }
}

// ---- predicated form of the step (used by every marching loop that runs a converged warp) ----
// A lane whose ray has ended keeps its t, which by construction reproduces the end condition (t >= max_range,
// or the out-of-map / obstacle cell at (int)(x0 + dx t), (int)(y0 + dy t)), so a burst of steps is straight-line
// code -- no branch, no reconvergence point -- and rm_result classifies the ray afterwards from t alone.
struct RmSlot {
  float x0, y0, dx, dy, t;
  unsigned cell = 0;  // index of the last distance-map cell read (always inside the map)
  int id;
  bool busy;   // the slot holds a ray (marching, or ended and not yet written out)
  bool alive;  // ... and it is still marching
};

#ifndef RL_RM_PRED_LOAD
#define RL_RM_PRED_LOAD 0  // 1: idle lanes re-read their last cell (17-instruction step); measured slower, see rm_step_pred
#endif
template <bool COND_LOAD>
__device__ __forceinline__ void rm_step_pred(const float* __restrict__ dt, unsigned W, unsigned H, float max_range,
                                             RmSlot& r) {
  const int px = __float2int_rz(fadd(r.x0, fmul(r.dx, r.t)));
  const int py = __float2int_rz(fadd(r.y0, fmul(r.dy, r.t)));
  const bool go = r.alive && (unsigned)px < W && (unsigned)py < H;
  // COND_LOAD: lanes that are not marching issue no load (ptxas makes it a BSSY / BRA / BSYNC triple, three
  // more instructions per step); otherwise they read cell 0 and ignore it (one more 128-byte line in the
  // request -- the gather rate of this kernel is bounded by L1 tag lookups, one line per clock per SM)
  float d = 0.0f;
  if (COND_LOAD) {
    if (go) d = __ldg(dt + ((unsigned)px * H + (unsigned)py));
  } else {
#if RL_RM_PRED_LOAD
    // Measured and dropped (round 2): lanes that are not marching re-read the last cell they visited -- `cell = go ? new
    // : cell` is one predicated IMAD and ptxas then forms the address with one IMAD.WIDE, 17 instructions per step
    // instead of 19 (IMAD, SEL, LEA, LEA.HI.X).  Slower everywhere (RM random 38.4 -> 35.6 G rays/s, 100000 x 60 fused
    // 23.5 -> 18.2, judged C2 step 23.6 -> 25.5 us): 32 idle lanes now touch up to 32 different lines per request where
    // they all hit cell 0's line, and L1 tag lookups are what these kernels run out of.  (A predicated load in inline
    // PTX is turned back into a branch by ptxas.)
    r.cell = go ? (unsigned)px * H + (unsigned)py : r.cell;
    d = __ldg(dt + r.cell);
#else
    d = __ldg(dt + (go ? (unsigned)px * H + (unsigned)py : 0u));
#endif
  }
  const bool adv = go && !(d <= 0.0f);
  const float tn = fadd(r.t, fmaxf(fmul(d, 0.999f), 1.0f));
  r.t = adv ? tn : r.t;
  r.alive = adv && (tn < max_range);
}

// result of a ray that ended with parameter t (see above); RangeLib.h:938-961
__device__ __forceinline__ float rm_result(unsigned W, unsigned H, float max_range, const RmSlot& r) {
  if (!(r.t < max_range)) return max_range;
  const int px = __float2int_rz(fadd(r.x0, fmul(r.dx, r.t)));
  const int py = __float2int_rz(fadd(r.y0, fmul(r.dy, r.t)));
  if ((unsigned)px >= W || (unsigned)py >= H) return max_range;
  const float xd = fsub((float)px, r.x0), yd = fsub((float)py, r.y0);
  return __fsqrt_rn(fadd(fmul(xd, xd), fmul(yd, yd)));
}

// Marches the rays held by the threads of one CTA (blockDim.x <= 256) to completion.  Must be called by
// ALL threads of the CTA (it synchronises).
//   phase 1  every thread steps its own ray (rm_step_pred) in bursts of RL_BLOCK_BURST iterations; after each burst
//            the CTA counts the rays still alive.  This finishes the bulk of the rays (mean ~6 steps) at one
//            dependent L2 read per step and lane;
//   phase 2  once at most mv.coop_threshold rays are left (default 8: one per warp; 4..16 measure the same, profiles/tune_r01_fused.log) they are parked in shared
//            memory and the warps of the CTA take them one at a time and finish each with all 32 lanes
//            (rm_march_coop).  The long crawls along walls -- the rays that decide the duration of a small
//            launch -- thus run concurrently on different warps, each at the cooperative rate, instead of
//            holding a nearly empty warp each.
// mv.coop_threshold == 0 (or a map / range too large for the 16-bit cell keys) keeps everything in phase 1.
#ifndef RL_BLOCK_BURST
#define RL_BLOCK_BURST 12
#endif
__device__ __forceinline__ float rm_march_block(const MapView& mv, float max_range, bool active, float x0, float y0,
                                                float dx, float dy) {
  __shared__ float4 s_ray[256];
  __shared__ float s_t[256];
  __shared__ float s_res[256];
  __shared__ short s_slot[256];
  __shared__ int s_n, s_next;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const float* __restrict__ dt = mv.dt;
  const unsigned W = (unsigned)mv.W, H = (unsigned)mv.H;
  RmSlot r;
  r.x0 = x0; r.y0 = y0; r.dx = dx; r.dy = dy;
  r.t = 0.0f;
  r.id = 0;
  r.busy = r.alive = active;
  const int handoff = (mv.W < 32768 && mv.H < 32768 && max_range < 32768.0f) ? min(mv.coop_threshold, 256) : 0;
  if (handoff <= 0) {
    while (__any_sync(FULL, r.alive)) {
#pragma unroll
      for (int k = 0; k < 4; ++k) rm_step_pred<false>(dt, W, H, max_range, r);
    }
    return active ? rm_result(W, H, max_range, r) : max_range;
  }
  if (threadIdx.x == 0) {
    s_n = 0;
    s_next = 0;
  }
  RL_TRACE_MARK(1);
  for (int steps = 2 * mv.block_burst_pairs;; steps += 2 * mv.block_burst_pairs) {
    if (__any_sync(FULL, r.alive)) {
#pragma unroll 1
      for (int b = 0; b < mv.block_burst_pairs; ++b) {
        rm_step_pred<false>(dt, W, H, max_range, r);
        rm_step_pred<false>(dt, W, H, max_range, r);
      }
    }
    const int alive = __syncthreads_count(r.alive);
    if (steps == 2 * mv.block_burst_pairs) {
      RL_TRACE_MARK(2);
      RL_TRACE_VALUE(9, alive);
    }
    RL_TRACE_MARK(3);
    RL_TRACE_VALUE(8, steps);
    RL_TRACE_VALUE(10, alive);
    if (alive == 0) return active ? rm_result(W, H, max_range, r) : max_range;
    // few rays left -- or, after 24 steps, a moderate number of rays that are evidently long ones
    if (alive <= handoff || (steps >= 24 && alive <= 4 * handoff)) break;
  }
  if (r.alive) {
    const int q = atomicAdd(&s_n, 1);
    s_ray[q] = make_float4(x0, y0, dx, dy);
    s_t[q] = r.t;
    s_slot[q] = (short)threadIdx.x;
  }
  __syncthreads();
  const int n_long = s_n;
  while (true) {
    int q = 0;
    if (lane == 0) q = atomicAdd(&s_next, 1);
    q = __shfl_sync(FULL, q, 0);
    if (q >= n_long) break;
    const float4 ray = s_ray[q];
    const float res = rm_march_coop(mv, max_range, ray.x, ray.y, ray.z, ray.w, s_t[q]);
    if (lane == 0) s_res[s_slot[q]] = res;
  }
  __syncthreads();
  RL_TRACE_MARK(5);
  if (r.alive) return s_res[threadIdx.x];
  return active ? rm_result(W, H, max_range, r) : max_range;
}

// pose -> ray; `ok` false for non-finite poses and for max_range <= 0 (the reference's loop body
// never runs / (int)NaN leaves the map: both return max_range)
__device__ __forceinline__ bool rm_setup(float max_range, float x, float y, float theta, float* dx, float* dy) {
  if (!finite3(x, y, theta) || !(0.0f < max_range)) return false;
  rl_sincosf(theta, dy, dx);
  return true;
}

// single-thread form (kept for callers that cannot guarantee a converged warp)
__device__ __forceinline__ float rm_cast(const MapView& mv, float max_range, float x0, float y0, float theta) {
  float dx, dy;
  if (!rm_setup(max_range, x0, y0, theta, &dx, &dy)) return max_range;
  float t = 0.0f, result;
  while (!rm_step(mv, max_range, x0, y0, dx, dy, t, result)) {
  }
  return result;
}

__device__ __forceinline__ bool occ_at(const MapView& mv, int x, int y) {  // OMap::isOccupied :204-210
  if ((unsigned)x >= (unsigned)mv.W || (unsigned)y >= (unsigned)mv.H) return false;
  return (__ldg(mv.bits_t + (size_t)(x >> 3) * mv.tiles8_y + (y >> 3)) >> ((x & 7) * 8 + (y & 7))) & 1ULL;
}

// BresenhamsLine::calc_range, RangeLib.h:696-769.  All state is float, as in the reference; the
// walk is the reference's recurrence (_x += +-1, error += deltay, conditional _y += +-1) step for
// step, so the visited cells and the returned distance are bit-identical.  What differs is how a
// step is evaluated:
//  * cell test = one bit of a register-cached 64-bit word holding an 8x8-cell tile of the map; a walk
//    crosses a tile in ~8-11 steps whatever its direction, so ~270 cell tests per ray cost ~30 loads
//    (bit-packing along one axis needed a load on almost every step of a diagonal walk, and the kernel
//    was bound by L2 sector traffic: ncu, profiles/);
//  * the float bounds tests `0 <= v && v < limit` (:755/:761) become one unsigned compare of
//    floor(v) -- identical for every finite v -- and floor(v) is also the cell index;
//  * the loop test `(int)_x != (int)(x1 + xstep)` becomes an interval test on _x (trunc(v) == T is an
//    interval of v), no conversion;
//  * the reference keeps walking after the ray has left the map; coordinates are monotone along
//    the walk, so once a coordinate has left on the side it is moving towards no cell can be hit
//    any more and the result (max_range) is returned at once.
struct BlState {
  float _x, _y, error, deltax, deltay, xstep, ystep, lo, hi, x0, y0;
  float half_dx;         // smallest float e with 2 e >= deltax: `error * 2 >= deltax` is `error >= half_dx`
  unsigned lim_a, lim_b; // map extent along the major / minor coordinate of the walk
  int target;            // the walk ends when trunc(_x) == target
  unsigned swap_mask;    // 0 when steep (map cell = (a, b)), ~0 otherwise (map cell = (b, a)): one LOP3 per coordinate
  float stop_u;  // the walk is over (or has jumped its target) once xstep * _x >= stop_u
  int cur_tile;
  unsigned long long cur;
  bool steep, skipped;
};

// everything before the walk (:698-745).  Returns true when the result is already known.
__device__ __forceinline__ bool bl_setup(const MapView& mv, float max_range, float x, float y, float heading,
                                         BlState& st, float* result) {
  *result = max_range;
  if (!finite3(x, y, heading)) return true;  // the reference does not terminate on these
  if (occ_at(mv, f2i(x), f2i(y))) {
    *result = 0.0f;
    return true;
  }
  float sn, cs;
  rl_sincosf(heading, &sn, &cs);
  float x0 = y, y0 = x;
  float x1 = fadd(y, fmul(max_range, sn));
  float y1 = fadd(x, fmul(max_range, cs));
  st.steep = fabsf(fsub(y1, y0)) > fabsf(fsub(x1, x0));
  if (st.steep) {
    float tmp = x0; x0 = y0; y0 = tmp;
    tmp = x1; x1 = y1; y1 = tmp;
  }
  st.deltax = fabsf(fsub(x1, x0));
  st.deltay = fabsf(fsub(y1, y0));
  // 2 e is exact for every float e (no overflow at these magnitudes), so 2 e >= deltax  <=>  e >= deltax / 2 as real
  // numbers  <=>  e >= the float just at or above deltax / 2
  st.half_dx = __fmul_ru(st.deltax, 0.5f);
  // not steep: _y indexes map x (< width), _x indexes map y (< height); steep: the other way round (:755/:761)
  st.lim_a = st.steep ? (unsigned)mv.W : (unsigned)mv.H;
  st.lim_b = st.steep ? (unsigned)mv.H : (unsigned)mv.W;
  st.swap_mask = st.steep ? 0u : 0xffffffffu;
  st.error = 0.0f;
  st._x = st.x0 = x0;
  st._y = st.y0 = y0;
  st.xstep = (x0 < x1) ? 1.0f : -1.0f;
  st.ystep = (y0 < y1) ? 1.0f : -1.0f;
  // the loop ends when trunc(_x) == target, i.e. lo <= _x < hi
  const int target = f2i(fadd(x1, st.xstep));
  if (target > 0) { st.lo = (float)target; st.hi = (float)target + 1.0f; }
  else if (target < 0) { st.lo = nextafterf((float)target - 1.0f, 0.0f); st.hi = nextafterf((float)target, 0.0f); }
  else { st.lo = nextafterf(-1.0f, 0.0f); st.hi = 1.0f; }
  if (target == INT_MIN || fabsf(x0) > 8388608.0f || fabsf(x1) > 8388608.0f) return true;  // far outside any map
  st.target = target;
  st.cur_tile = -1;
  st.cur = 0;
  st.skipped = false;
  // _x moves monotonically by xstep, so "trunc(_x) == target for the first time" is a one-sided test on
  // u = xstep * _x (exact: a sign flip): u >= lo when walking up, u > -hi when walking down.
  st.stop_u = (st.xstep > 0.0f) ? st.lo : nextafterf(-st.hi, INFINITY);
  return st._x >= st.lo && st._x < st.hi;  // zero-length walk
}

// one iteration of the walk (:745-767).  Returns true when the ray has ended.
// *result is written only when the walk ends on an obstacle; the caller presets it to max_range
__device__ __forceinline__ bool bl_step(const MapView& mv, float max_range, BlState& st, float* result) {
  st._x = fadd(st._x, st.xstep);
  st.error = fadd(st.error, st.deltay);
  if (st.error >= st.half_dx) {  // (double)error*2.0 >= (double)deltax (:750), see half_dx
    st._y = fadd(st._y, st.ystep);
    st.error = fsub(st.error, st.deltax);
  }
  const unsigned lim_a = st.lim_a, lim_b = st.lim_b;  // major coordinate _x, minor coordinate _y
  const int a = __float2int_rd(st._x), b = __float2int_rd(st._y);
  if ((unsigned)a < lim_a && (unsigned)b < lim_b) {
    const unsigned sm = st.swap_mask;  // map cell: (a, b) when steep, (b, a) otherwise -- a bitwise select each
    const int cx = (int)(((unsigned)a & ~sm) | ((unsigned)b & sm)), cy = (int)(((unsigned)b & ~sm) | ((unsigned)a & sm));
    const int tile = (cx >> 3) * mv.tiles8_y + (cy >> 3);
    if (tile != st.cur_tile) {
      st.cur_tile = tile;
      st.cur = __ldg(mv.bits_t + tile);
    }
    unsigned occ_bit;  // bit (cx & 7) * 8 + (cy & 7) of the tile word, as a 32-bit test (SHF.R.U64 + LOP3 instead of a 64-bit compare)
    asm("{\n\t.reg .u64 t;\n\tshr.u64 t, %1, %2;\n\tcvt.u32.u64 %0, t;\n\t}" : "=r"(occ_bit) : "l"(st.cur), "r"((cx & 7) * 8 + (cy & 7)));
    if (occ_bit & 1u) {
      const float xd = fsub(st._x, st.x0), yd = fsub(st._y, st.y0);
      *result = __fsqrt_rn(fadd(fmul(xd, xd), fmul(yd, yd)));
      return true;
    }
    // End of the walk, inside the map: _x >= 0 here, so trunc(_x) is the cell index a and the reference's loop test
    // `(int)_x != (int)(x1 + xstep)` is one integer compare.  (a moves monotonically; if the float accumulation made it
    // jump over the target -- see below -- it never equals it and the walk goes on, as the reference's does.)
    return a == st.target;
  }
  // outside the map: left on the side the walk is heading to -> nothing can be hit any more
  if ((st.xstep > 0.0f) ? (a >= (int)lim_a) : (a < 0)) return true;
  if ((st.ystep > 0.0f) ? (b >= (int)lim_b) : (b < 0)) return true;
  if (fmul(st.xstep, st._x) >= st.stop_u) {
    // Normally this is the end of the walk (trunc(_x) == target).  The reference's `_x += xstep` is a float
    // accumulation, though: when _x crosses a power of two with low fraction bits set the sum rounds and
    // (int)_x can jump over the target, after which the reference keeps walking (and never returns once the
    // walk is outside the map).  We follow it for as long as a cell can still be hit -- the "left the map
    // for good" test above ends such walks -- with a backstop one map diagonal further on.
    if (st.skipped || (st._x >= st.lo && st._x < st.hi)) return true;
    st.skipped = true;
    st.stop_u = fadd(fmul(st.xstep, st._x), (float)(mv.W + mv.H) + max_range);
  }
  return false;
}

__device__ __forceinline__ float bl_cast(const MapView& mv, float max_range, float x, float y, float heading) {
  BlState st;
  float result;
  if (bl_setup(mv, max_range, x, y, heading, st, &result)) return result;
  result = max_range;
  while (!bl_step(mv, max_range, st, &result)) {
  }
  return result;
}

// CDDTCast::discretize_theta RangeLib.h:1287-1340 (_USE_ALTERNATE_MOD 1, _USE_CACHED_CONSTANTS 1,
// _USE_FAST_ROUND 0).  The wrap loops add/subtract the double constant and narrow to float each
// turn exactly as the reference does; they are capped (|theta| beyond ~6000 rad is wrapped with
// fmod instead -- the reference would spin for that many iterations).
__device__ __forceinline__ void cddt_discretize(const CddtView& cv, float theta, int* bin, bool* flipped) {
  if ((double)theta < 0.0) {
    int it = 0;
    while ((double)theta < 0.0) {
      theta = __double2float_rn((double)theta + RL_M_2PI);
      if (++it > 1000) { theta = __double2float_rn(fmod((double)theta, RL_M_2PI) + RL_M_2PI); }
    }
  } else if ((double)theta > RL_M_2PI) {
    int it = 0;
    while ((double)theta > RL_M_2PI) {
      theta = __double2float_rn((double)theta - RL_M_2PI);
      if (++it > 1000) { theta = __double2float_rn(fmod((double)theta, RL_M_2PI)); }
    }
  }
  bool f = false;
  if ((double)theta >= RL_PI) {
    f = true;
    theta = __double2float_rn((double)theta - RL_PI);
  }
  int rounded = (int)roundf(fmul(theta, cv.td_div_2pi));
  if ((unsigned)rounded == (cv.td >> 1)) {
    rounded = 0;
    f = !f;
  }
  *bin = (int)((unsigned)rounded % cv.td);
  *flipped = f;
}

// Search of one bin through the query index (rl_cddt.cu: cddt_index_build).  Returns the absolute position in
// values[] of the first element of the bin [o0, o0 + size) that does NOT satisfy the bin-monotone predicate
//   le ? v <= lx : v < lx        (o0 + size when all satisfy it)
// which is o0 + the `lo` the plain bisections below end with.  skip[k] is the 16-bit monotone code of values[16 k]
// (rl_internal.cuh: cddt_code), so the blocks k0+1 .. k1 that START inside the bin are bisected through skip[]
// (L2-resident; a probe whose code equals the query's reads the value itself), and the one 64-byte aligned block that holds the
// boundary is read with 16-byte loads and counted branch-free; elements of the block that belong to the neighbouring
// bins are masked out by position.
#ifndef RL_CDDT_BLOCK_UNROLL
#define RL_CDDT_BLOCK_UNROLL 4  // 16-byte loads of the block in flight at once (4: all; 1: one at a time, 12 registers less)
#endif
__device__ __forceinline__ unsigned cddt_search_indexed(const CddtView& cv, unsigned o0, unsigned size, float lx,
                                                        bool le, float first, float last) {
  const unsigned o1 = o0 + size;  // size >= 1
  const unsigned k0 = o0 >> 4, k1 = (o1 - 1) >> 4;
  const uint16_t* __restrict__ S = cv.skip + k0 + 1;
  int lo = 0, n = (int)(k1 - k0);
  if (n > 0) {
    const float scale = cddt_code_scale(first, last);
    const unsigned cq = cddt_code(lx, first, scale);  // first <= lx <= last here
    while (n > 0) {
      const int half = n >> 1;
      const unsigned c = __ldg(S + lo + half);
      bool go_right = c < cq;  // code order implies value order; equal codes: compare the values
      if (c == cq) {
        const float v = __ldg(cv.values + ((size_t)(k0 + 1 + lo + half) << 4));
        go_right = le ? !(lx < v) : (v < lx);
      }
      lo = go_right ? lo + half + 1 : lo;
      n = go_right ? n - half - 1 : half;
    }
  }
  const unsigned kk = k0 + (unsigned)lo;  // the last block whose first element satisfies the predicate, or k0
  const float4* __restrict__ blk = reinterpret_cast<const float4*>(cv.values + ((size_t)kk << 4));
  const unsigned base = kk << 4;
  const unsigned from = max(o0, base);
  unsigned count = 0;
  constexpr int kBlockUnroll = RL_CDDT_BLOCK_UNROLL;
#pragma unroll kBlockUnroll
  for (int j = 0; j < 4; ++j) {
    const float4 q = __ldg(blk + j);
    const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned idx = base + 4 * j + i;
      const bool in = idx >= from && idx < o1;
      const bool sat = le ? !(lx < e[i]) : (e[i] < lx);
      count += (in && sat) ? 1u : 0u;
    }
  }
  return from + count;
}

// the part of CDDTCast::calc_range that does not touch the table: theta -> slice, rotation, bin.  false: max_range.
__device__ __forceinline__ bool cddt_locate(const CddtView& cv, float x, float y, float heading, bool* flipped,
                                            float* lx, int64_t* bin) {
  if (!finite3(x, y, heading)) return false;
  int a;
  cddt_discretize(cv, -heading, &a, flipped);
  const float ca = __ldg(cv.cosv + a), sa = __ldg(cv.sinv + a);
  *lx = fsub(fmul(x, ca), fmul(y, sa));
  const float ly = fadd(fadd(fmul(x, sa), fmul(y, ca)), __ldg(cv.trans + a));
  const unsigned li = (unsigned)f2i(ly);
  if (li >= (unsigned)__ldg(cv.widths + a)) return false;
  *bin = __ldg(cv.slice0 + a) + li;
  return true;
}

// CDDTCast::calc_range through the query index (tables larger than L2): the same decisions as cddt_cast below, taken
// from the bin's 16-byte record, the skip entries and one 64-byte block of zero points.  The occupancy word is
// requested together with the record (its address only depends on the pose), not after the early exits.
__device__ __forceinline__ float cddt_cast_indexed(const MapView& mv, const CddtView& cv, float max_range, float x,
                                                   float y, float heading) {
  bool flipped;
  float lx;
  int64_t b;
  if (!cddt_locate(cv, x, y, heading, &flipped, &lx, &b)) return max_range;
  const int cx = f2i(x), cy = f2i(y);
  const bool inside = (unsigned)cx < (unsigned)mv.W && (unsigned)cy < (unsigned)mv.H;
  const unsigned long long word = inside ? __ldg(mv.bits_t + (size_t)(cx >> 3) * mv.tiles8_y + (cy >> 3)) : 0ULL;
  const uint4 mt = __ldg(reinterpret_cast<const uint4*>(cv.meta) + b);
  const unsigned o0 = mt.x, size = mt.y;
  if (size == 0) return max_range;
  const float first = __uint_as_float(mt.z), last = __uint_as_float(mt.w);
  bool le = true;
  if (flipped) {
    if (first > lx) return max_range;
    if (last < lx) return fsub(lx, last);
  } else {
    if (last < lx) return max_range;
    if (first > lx) return fsub(first, lx);
    le = (int)size - 1 > RL_BINARY_SEARCH_THRESHOLD;
  }
  if ((word >> ((cx & 7) * 8 + (cy & 7))) & 1ULL) return 0.0f;  // map.grid[x][y] :1413 / :1475
  const unsigned p = cddt_search_indexed(cv, o0, size, lx, le, first, last);
  return flipped ? fsub(lx, __ldg(cv.values + p - 1)) : fsub(__ldg(cv.values + p), lx);
}

// CDDTCast::calc_range RangeLib.h:1342-1516.  cos/sin of the discrete angle come from the
// host-tabulated libm values (cv.cosv/sinv), so trig is out of the parity question.
__device__ __forceinline__ float cddt_cast(const MapView& mv, const CddtView& cv, float max_range, float x, float y,
                                           float heading) {
  if (cv.meta) return cddt_cast_indexed(mv, cv, max_range, x, y, heading);  // warp-uniform
  bool flipped;
  float lx;
  int64_t b;
  if (!cddt_locate(cv, x, y, heading, &flipped, &lx, &b)) return max_range;
  // The reference scans bins of <= 65 entries linearly and uses std::upper_bound on larger ones.  A
  // sorted, duplicate-free bin makes the linear scans equal to a binary search (first element >= lx,
  // resp. last element <= lx), so every case below is one branch-free binary search; `strict` selects
  // upper_bound (first element > lx) or lower_bound (first element >= lx).
  const int64_t o0 = __ldg(cv.offsets + b), o1 = __ldg(cv.offsets + b + 1);
  const float* __restrict__ B = cv.values + o0;
  const int size = (int)(o1 - o0);
  const int high = size - 1;
  if (high == -1) return max_range;
  const float first = __ldg(B), last = __ldg(B + high);
  if (flipped) {
    if (first > lx) return max_range;
    if (last < lx) return fsub(lx, last);
    if (occ_at(mv, f2i(x), f2i(y))) return 0.0f;  // map.grid[x][y] :1413
    // :1416-1418 upper_bound - 1; :1429-1443 last element <= lx == upper_bound - 1
    int lo = 0, n = size;
    while (n > 0) {
      const int half = n >> 1;
      const bool go_right = !(lx < __ldg(B + lo + half));
      lo = go_right ? lo + half + 1 : lo;
      n = go_right ? n - half - 1 : half;
    }
    return fsub(lx, __ldg(B + lo - 1));
  } else {
    if (last < lx) return max_range;
    if (first > lx) return fsub(first, lx);
    if (occ_at(mv, f2i(x), f2i(y))) return 0.0f;  // :1475
    // :1480-1481 upper_bound (first > lx) for large bins; :1494-1510 first element >= lx for small ones
    const bool strict = high > RL_BINARY_SEARCH_THRESHOLD;
    int lo = 0, n = size;
    while (n > 0) {
      const int half = n >> 1;
      const float v = __ldg(B + lo + half);
      const bool go_right = strict ? !(lx < v) : (v < lx);
      lo = go_right ? lo + half + 1 : lo;
      n = go_right ? n - half - 1 : half;
    }
    return fsub(__ldg(B + lo), lx);  // values[] is padded: lo == size reads the next bin's first value / the pad
  }
  return -1.0f;  // the reference's assert(0) fall-through (:1514)
}

// CDDTCast::calc_range_pair RangeLib.h:1521-1649: the range along `heading` and the range along heading + pi from
// one bin (the two neighbours of lx in the sorted bin).  Quirks kept: the non-flipped branch has no `first > lx`
// early return and its occupied-cell test is a no-op (:1612, the pair is built and dropped); large bins use
// lower_bound here where calc_range uses upper_bound (:1619 vs :1481).  The reference indexes the slice without a
// bounds check (:1538-1539, undefined for a pose whose rotated y leaves the table); here that returns
// (max_range, max_range).
__device__ __forceinline__ float2 cddt_cast_pair(const MapView& mv, const CddtView& cv, float max_range, float x,
                                                 float y, float heading) {
  if (!finite3(x, y, heading)) return make_float2(max_range, max_range);
  int a;
  bool flipped;
  cddt_discretize(cv, -heading, &a, &flipped);
  const float ca = __ldg(cv.cosv + a), sa = __ldg(cv.sinv + a);
  const float lx = fsub(fmul(x, ca), fmul(y, sa));
  const float ly = fadd(fadd(fmul(x, sa), fmul(y, ca)), __ldg(cv.trans + a));
  const unsigned li = (unsigned)f2i(ly);
  if (li >= (unsigned)__ldg(cv.widths + a)) return make_float2(max_range, max_range);
  const int64_t b = __ldg(cv.slice0 + a) + li;
  const int64_t o0 = __ldg(cv.offsets + b), o1 = __ldg(cv.offsets + b + 1);
  const float* __restrict__ B = cv.values + o0;
  const int size = (int)(o1 - o0);
  if (size == 0) return make_float2(max_range, max_range);
  const float first = __ldg(B), last = __ldg(B + size - 1);
  if (flipped) {
    if (first > lx) return make_float2(max_range, fminf(max_range, fsub(first, lx)));  // :1552-1553
    if (last < lx) return make_float2(fsub(lx, last), max_range);                      // :1554-1555
    if (occ_at(mv, f2i(x), f2i(y))) return make_float2(0.0f, 0.0f);                    // :1559
    int lo = 0, n = size;  // last element <= lx == upper_bound - 1 (:1566 and the scan :1569-1576)
    while (n > 0) {
      const int half = n >> 1;
      const bool go_right = !(lx < __ldg(B + lo + half));
      lo = go_right ? lo + half + 1 : lo;
      n = go_right ? n - half - 1 : half;
    }
    const int index = lo - 1;
    const float r = fsub(lx, __ldg(B + index));
    if (index + 1 == size) return make_float2(r, max_range);
    return make_float2(r, fsub(__ldg(B + index + 1), lx));
  }
  if (last < lx) return make_float2(max_range, fminf(max_range, fsub(lx, last)));  // :1605-1606
  int lo = 0, n = size;  // first element >= lx (:1619 lower_bound and the scan :1623-1631)
  while (n > 0) {
    const int half = n >> 1;
    const bool go_right = __ldg(B + lo + half) < lx;
    lo = go_right ? lo + half + 1 : lo;
    n = go_right ? n - half - 1 : half;
  }
  const float r = fsub(__ldg(B + lo), lx);
  if (lo == 0) return make_float2(r, max_range);
  return make_float2(r, fsub(lx, __ldg(B + lo - 1)));
}

// GiantLUTCast::calc_range RangeLib.h:1869-1880 with discretize_theta :1833-1867 (no flip, unlike CDDT)
__device__ __forceinline__ float glt_cast(const MapView& mv, float max_range, float x, float y, float theta) {
  if (!finite3(x, y, theta)) return max_range;
  if (x < 0.0f || x >= (float)(unsigned)mv.W || y < 0.0f || y >= (float)(unsigned)mv.H) return max_range;
  if ((double)theta < 0.0) {
    int it = 0;
    while ((double)theta < 0.0) {
      theta = __double2float_rn((double)theta + RL_M_2PI);
      if (++it > 1000) { theta = __double2float_rn(fmod((double)theta, RL_M_2PI) + RL_M_2PI); }
    }
  } else if ((double)theta > RL_M_2PI) {
    int it = 0;
    while ((double)theta > RL_M_2PI) {
      theta = __double2float_rn((double)theta - RL_M_2PI);
      if (++it > 1000) { theta = __double2float_rn(fmod((double)theta, RL_M_2PI)); }
    }
  }
  const int rounded = (int)roundf(fmul(theta, mv.glt_td_div_2pi));
  const int i = rounded % (int)mv.glt_td;
  const size_t cell = (size_t)__float2int_rz(x) * (unsigned)mv.H + __float2int_rz(y);
  return fmul((float)(int)__ldg(mv.glt + cell * mv.glt_td + i), mv.glt_max_div_limits);
}

template <int KIND>
__device__ __forceinline__ float cast_one(const MapView& mv, const CddtView& cv, float max_range, float x, float y,
                                          float th) {
  if (KIND == RL_BL) return bl_cast(mv, max_range, x, y, th);
  if (KIND == RL_RM) return rm_cast(mv, max_range, x, y, th);
  if (KIND == RL_GLT) return glt_cast(mv, max_range, x, y, th);
  return cddt_cast(mv, cv, max_range, x, y, th);
}

// world -> grid pose, RangeLib.h:464-473.  Returns (x, y, theta) with the reference's names;
// the caller passes them to calc_range as (y, x, theta) (:475).
__device__ __forceinline__ void world_to_grid(const WorldXform& xf, float xw, float yw, float thw, float* x, float* y,
                                              float* th) {
  float xx = fmul(fsub(xw, xf.ox), xf.inv_scale);
  float yy = fmul(fsub(yw, xf.oy), xf.inv_scale);
  float tmp = xx;
  xx = fsub(fmul(xf.cos_a, xx), fmul(xf.sin_a, yy));
  yy = fadd(fmul(xf.sin_a, tmp), fmul(xf.cos_a, yy));
  *x = xx;
  *y = yy;
  *th = fadd(-thw, xf.rot);
}

// clamp + truncate of RangeLib.h:547-551 / 603-606: std::min<float>(std::max<float>(v,0),K-1)
__device__ __forceinline__ int sensor_index(float v, float kmax) {
  v = (v < 0.0f) ? 0.0f : v;      // std::max<float>(v, 0.0): (v < 0) ? 0 : v
  v = (kmax < v) ? kmax : v;      // std::min<float>(v, kmax): (kmax < v) ? kmax : v
  return __float2int_rz(v);
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------

// t / m and t % m for t < 2^22 * m, t < 2^24 (ray indices inside a chunk or a particle group) without the integer
// division sequence (ncu source view, round 2: the 64-bit `r / M` of the pose load was 12 % of all instructions issued by
// the lidar-fan kernel).  rcp_m = 1 / m rounded towards zero; the truncated float product is the quotient or one
// below it (both roundings go down, relative error < 2^-22), which one multiply-subtract and a compare settle.
__device__ __forceinline__ float rcp_floor(unsigned m) { return __frcp_rz((float)m); }
__device__ __forceinline__ void divmod_small(unsigned t, unsigned m, float rcp_m, unsigned* q, unsigned* r) {
  unsigned qq = __float2uint_rz(__fmul_rz((float)t, rcp_m));
  unsigned rr = t - qq * m;
  if (rr >= m) {
    ++qq;
    rr -= m;
  }
  *q = qq;
  *r = rr;
}

// resolves ray r of a batch into the pose handed to calc_range, per entry point
template <int MODE>
__device__ __forceinline__ void load_pose(const WorldXform& xf, const float* __restrict__ ins,
                                          const float* __restrict__ angles, long long r, int M, float* gx, float* gy,
                                          float* gth) {
  if (MODE == MODE_GLT_BUILD) {
    // table entry r = (x*H + y)*td + i: RM seeded from the pixel corner (x, y) at angle i * 2pi/td
    // (RangeLib.h:1791-1801); M = td, xf.rot carries (float)(M_2PI / (float)td), xf.oy carries H
    const long long cell = r / M;
    const int Hh = (int)xf.oy;
    *gx = (float)(int)(cell / Hh);
    *gy = (float)(int)(cell % Hh);
    *gth = fmul((float)(int)(r - cell * M), xf.rot);
  } else if (MODE == MODE_GRID) {
    *gx = __ldg(ins + 3 * r);
    *gy = __ldg(ins + 3 * r + 1);
    *gth = __ldg(ins + 3 * r + 2);
  } else {
    const long long i = (MODE == MODE_ANGLES) ? r / M : r;
    float x, y, th;
    world_to_grid(xf, __ldg(ins + 3 * i), __ldg(ins + 3 * i + 1), __ldg(ins + 3 * i + 2), &x, &y, &th);
    if (MODE == MODE_ANGLES) th = fsub(th, __ldg(angles + (int)(r - i * M)));
    *gx = y;  // calc_range(y, x, theta): RangeLib.h:475 / :518
    *gy = x;
    *gth = th;
  }
}

// one ray per thread, grid-stride; the loop is CTA-uniform so that RM can march at CTA level (rm_march_block)
template <int KIND, int MODE>
__global__ void __launch_bounds__(256, KIND == RL_RM ? 7 : 4)
cast_kernel(MapView mv, CddtView cv, WorldXform xf, float max_range, const float* __restrict__ ins,
            const float* __restrict__ angles, float* __restrict__ outs, long long total, int M) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += stride) {  // CTA-uniform trip count
    const long long r = base + threadIdx.x;
    const bool valid = r < total;
    float gx = 0.f, gy = 0.f, gth = 0.f;
    if (valid) load_pose<MODE>(xf, ins, angles, r, M, &gx, &gy, &gth);
    float range;
    if (KIND == RL_RM) {
      float dx = 0.f, dy = 0.f;
      const bool ok = valid && rm_setup(max_range, gx, gy, gth, &dx, &dy);
      range = rm_march_block(mv, max_range, ok, gx, gy, dx, dy);
    } else {
      range = valid ? cast_one<KIND>(mv, cv, max_range, gx, gy, gth) : 0.0f;
    }
    if (valid) outs[r] = (MODE == MODE_GRID) ? range : fmul(range, xf.scale);
  }
}

// CDDT / PCDDT batches.  ncu on cast_kernel<CDDT> with the query index (profiles/r02/ncu_c3_cddt_index_first.txt):
// nothing is saturated -- DRAM 36 %, L1->L2 requests 41 %, issue slots 28 % -- and 20 warps per scheduler-issue wait on
// a load: a query is a chain of dependent misses (pose, bin record, skip entry, block of zero points; or pose, offsets,
// bin ends, ~5 bisection probes without the index) and only 1024 of them fit an SM at cast_kernel's 56 registers.
// This kernel carries nothing but the one CDDT path it is instantiated for and is bounded for 32 registers, so that
// 2048 queries are in flight per SM.
#ifndef RL_CDDT_IDX_MINB
#define RL_CDDT_IDX_MINB 6  // resident CTAs per SM the register allocation is bounded for
#endif
template <int MODE, bool INDEXED>
__global__ void __launch_bounds__(256, RL_CDDT_IDX_MINB)
cddt_batch_kernel(MapView mv, CddtView cv, WorldXform xf, float max_range, const float* __restrict__ ins,
                    const float* __restrict__ angles, float* __restrict__ outs, long long total, int M) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += stride) {
    float gx, gy, gth;
    load_pose<MODE>(xf, ins, angles, r, M, &gx, &gy, &gth);
    const float range = INDEXED ? cddt_cast_indexed(mv, cv, max_range, gx, gy, gth) : cddt_cast(mv, cv, max_range, gx, gy, gth);
    outs[r] = (MODE == MODE_GRID) ? range : fmul(range, xf.scale);
  }
}

// RangeMethod::calc_range_many_radial_optimized RangeLib.h:616-676.  Row i of outs (num_rays floats) gets, for
// a <= max_pair, the pair (range at beam a, range at beam a + index_offset) from one calc_range_pair, and for
// max_pair < a < index_offset a plain calc_range; beams the reference never writes are left untouched.
// `beam_angles[a]` is the reference's float accumulation min_angle + a * step, tabulated on the host.  A pair
// whose second beam lands on a slot a later iteration of the reference overwrites (a + index_offset <= max_pair)
// or outside the row (a + index_offset >= num_rays; the reference writes into the next row / past the buffer)
// is dropped.  Methods without calc_range_pair return (-1, -1) from it (:419).
template <int KIND>
__global__ void __launch_bounds__(256, 4)
radial_kernel(MapView mv, CddtView cv, WorldXform xf, float max_range, const float* __restrict__ ins,
              const float* __restrict__ beam_angles, float* __restrict__ outs, long long total, int num_rays, int count,
              int max_pair, int index_offset) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += stride) {
    const long long i = r / count;
    const int a = (int)(r - i * count);
    float x, y, th;
    world_to_grid(xf, __ldg(ins + 3 * i), __ldg(ins + 3 * i + 1), __ldg(ins + 3 * i + 2), &x, &y, &th);
    th = fsub(th, __ldg(beam_angles + a));
    float* row = outs + i * num_rays;
    if (a <= max_pair) {
      float2 pr = make_float2(-1.0f, -1.0f);
      if (KIND == RL_CDDT) pr = cddt_cast_pair(mv, cv, max_range, y, x, th);
      if (a < num_rays) row[a] = fmul(pr.x, xf.scale);
      const int b = a + index_offset;
      if (b > max_pair && b >= 0 && b < num_rays) row[b] = fmul(pr.y, xf.scale);
    } else if (a < num_rays) {
      row[a] = fmul(cast_one<KIND>(mv, cv, max_range, y, x, th), xf.scale);
    }
  }
}

// One CTA handles `ppb` consecutive particles per iteration (grid-stride over particle groups).
// Beams are processed in chunks of at most `chunk` so shared memory stays bounded for any M.
// smem: double vals[ppb * chunk].
// PARAM_BEAMS: angles and observation arrive as kernel parameters (rl::BeamParams) and are staged in shared memory.
template <int KIND, bool PARAM_BEAMS>
__global__ void __launch_bounds__(256, KIND == RL_RM ? 7 : 4)
fused_kernel(MapView mv, CddtView cv, WorldXform xf, SensorView sv, float max_range, const float* __restrict__ ins,
             const float* __restrict__ angles, const float* __restrict__ obs, double* __restrict__ weights, int N,
             int M, int ppb, int chunk, PeerOut peers,
             typename std::conditional<PARAM_BEAMS, BeamParams, NoBeamParams>::type beams,
             const int* __restrict__ perm) {
  extern __shared__ double vals[];
  RL_TRACE_MARK(0);
  __shared__ float s_beams[PARAM_BEAMS ? 2 * RL_PARAM_BEAMS : 1];
  if (PARAM_BEAMS) {
    const BeamParams& bp = reinterpret_cast<const BeamParams&>(beams);
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
      s_beams[i] = bp.angles[i];
      s_beams[RL_PARAM_BEAMS + i] = bp.obs[i];
    }
    __syncthreads();
  }
  const float kmax = (float)((double)(float)sv.K - 1.0);
  const int groups = (N + ppb - 1) / ppb;
  long long epoch = 0;
  if (peers.sig) {
    // signalled multi-GPU mode: this launch is epoch e; wait until every rank has finished e-1
    epoch = *(volatile long long*)peers.epoch + 1;
    if (threadIdx.x < peers.n) {
      volatile long long* mine = peers.flags[peers.rank];
      while (mine[threadIdx.x] < epoch - 1) {
      }
    }
    __syncthreads();
  }
  double* const* out_ptrs = (peers.sig && (epoch & 1)) ? peers.ptr1 : peers.ptr;
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const int p0 = g * ppb;
    const int np = min(ppb, N - p0);
    double w = 1.0;  // running product, owned by thread p < np
    for (int c0 = 0; c0 < M; c0 += chunk) {
      const int cm = min(chunk, M - c0);
      const int rays = np * cm;
      const float rcp_cm = rcp_floor((unsigned)cm);
      for (int k0 = 0; k0 < rays; k0 += blockDim.x) {  // block-uniform trip count
        const int k = k0 + threadIdx.x;
        const bool valid = k < rays;
        int p = 0, a = c0;
        float gx = 0.f, gy = 0.f, gth = 0.f;
        if (valid) {
          unsigned up, ua;
          divmod_small((unsigned)k, (unsigned)cm, rcp_cm, &up, &ua);  // k < ppb * cm <= 65536
          p = (int)up;
          a = c0 + (int)ua;
          const size_t i = perm ? (size_t)__ldg(perm + p0 + p) : (size_t)(p0 + p);  // processing order only
          float x, y, th;
          world_to_grid(xf, __ldg(ins + 3 * i), __ldg(ins + 3 * i + 1), __ldg(ins + 3 * i + 2), &x, &y, &th);
          gx = y;
          gy = x;
          gth = fsub(th, PARAM_BEAMS ? s_beams[a] : __ldg(angles + a));
        }
        float d;
        if (KIND == RL_RM) {
          float dx = 0.f, dy = 0.f;
          const bool ok = valid && rm_setup(max_range, gx, gy, gth, &dx, &dy);
          d = rm_march_block(mv, max_range, ok, gx, gy, dx, dy);
        } else {
          d = valid ? cast_one<KIND>(mv, cv, max_range, gx, gy, gth) : 0.0f;
        }
        if (valid) {
          const int di = sensor_index(d, kmax);                                   // :602-603 (no scaling)
          const float ob = PARAM_BEAMS ? s_beams[RL_PARAM_BEAMS + a] : __ldg(obs + a);
          const int ri = sensor_index(fmul(ob, xf.inv_scale), kmax);  // :605-606
          vals[p * cm + (a - c0)] = __ldg(sv.table + (size_t)ri * sv.K + di);
        }
      }
      __syncthreads();
      RL_TRACE_MARK(6);
      if (threadIdx.x < np) {
        const double* v = vals + threadIdx.x * cm;
        for (int a = 0; a < cm; ++a) w = __dmul_rn(w, v[a]);  // reference order: beam ascending
      }
      __syncthreads();
      RL_TRACE_MARK(7);
    }
    if (threadIdx.x < np) {
      const size_t i = perm ? (size_t)__ldg(perm + p0 + threadIdx.x) : (size_t)(p0 + threadIdx.x);
      if (peers.n == 0) {
        weights[i] = w;
      } else {  // all-gather by direct peer stores (NVLink): every GPU gets this rank's slice
        for (int r = 0; r < peers.n; ++r) out_ptrs[r][peers.offset + i] = w;
      }
    }
  }
  if (peers.sig) {
    // publish: once every CTA of this launch has made its peer stores visible system-wide, the last one
    // to arrive tells every rank that this rank's slice of epoch e is complete
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned done = atomicAdd(peers.counter, 1u);
      if (done == gridDim.x - 1) {
        *peers.counter = 0u;
        __threadfence_system();
        *(volatile long long*)peers.epoch = epoch;
        for (int r = 0; r < peers.n; ++r) ((volatile long long*)peers.flags[r])[peers.rank] = epoch;
      }
    }
  }
}

// Deep fused launches (many particle groups per CTA) of fused_kernel's kinds: the same work with the PRODUCT taken off
// the marching warps' path.  fused_kernel ends every group with barrier / product / barrier: with 1080 beams the
// product is a chain of 1080 dependent DMULs (~9 us) during which 255 threads wait (ncu, config 5: 9.8 warps per issued
// instruction stalled at the barrier, issue slots 59 % busy).  Here a ninth warp owns the products: the eight marching
// warps write a group's table values into one of two shared-memory buffers, ARRIVE at a named barrier and go on with the
// next group; the product warp waits on that barrier, multiplies in beam order (lane p = particle p of the group),
// stores the weights and frees the buffer through a second named barrier.  Barrier ids 1..4 (0 is __syncthreads):
// FULL[b] = buffer b holds a whole group (256 arrive, 32 wait), FREE[b] = buffer b has been consumed (32 arrive, 256
// wait, from the buffer's second use on).  Weights are bit-identical: same values, same order.
// Measured (B200, profiles/r02/fused_overlap_r02.log): 20000 x 1080 on the 5 cm map 33.9 -> 37.6 G rays/s; on the
// 8192^2 map (distance transform 268 MB > L2) 26.6 -> 24.9 -- nine warps per CTA leave room for 7 CTAs = 56 marching
// warps per SM where fused_kernel has 64, and a launch that waits on L2 misses needs them -- so the launcher uses this
// kernel only while the structure the rays read fits L2.  Also measured there and dropped: a barrier that keeps the
// marching warps on the same group (no change), and rotating which warps get the partial last round of a group
// (slower: a warp that keeps its beams looks the same way from particle to particle, and the particles are neighbours);
// a 512-thread form for the big map (15 marching warps + the product warp, four CTAs = 60 marching warps per SM, 1 to 4
// particles per group): 25.4 / 24.5 / 24.5 / 23.7 against fused_kernel's 25.9 -- hiding the product buys nothing there.
// (immediate barrier ids: with an id in a register ptxas reserves all 16 named barriers for the CTA)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  switch (id) {
    case 1: asm volatile("barrier.sync 1, %0;" ::"r"(count) : "memory"); break;
    case 2: asm volatile("barrier.sync 2, %0;" ::"r"(count) : "memory"); break;
    case 3: asm volatile("barrier.sync 3, %0;" ::"r"(count) : "memory"); break;
    default: asm volatile("barrier.sync 4, %0;" ::"r"(count) : "memory"); break;
  }
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  switch (id) {
    case 1: asm volatile("barrier.arrive 1, %0;" ::"r"(count) : "memory"); break;
    case 2: asm volatile("barrier.arrive 2, %0;" ::"r"(count) : "memory"); break;
    case 3: asm volatile("barrier.arrive 3, %0;" ::"r"(count) : "memory"); break;
    default: asm volatile("barrier.arrive 4, %0;" ::"r"(count) : "memory"); break;
  }
}
#define RL_OVERLAP_MARCHERS 256
#define RL_OVERLAP_THREADS (RL_OVERLAP_MARCHERS + 32)
template <int KIND>
__global__ void __launch_bounds__(RL_OVERLAP_THREADS, 7)
fused_overlap_kernel(MapView mv, CddtView cv, WorldXform xf, SensorView sv, float max_range,
                     const float* __restrict__ ins, const float* __restrict__ angles, const float* __restrict__ obs,
                     double* __restrict__ weights, int N, int M, int ppb, int chunk, PeerOut peers,
                     const int* __restrict__ perm) {
  extern __shared__ double vals[];  // two buffers of ppb * chunk doubles
  const unsigned FULL = 0xffffffffu;
  const float kmax = (float)((double)(float)sv.K - 1.0);
  const int groups = (N + ppb - 1) / ppb;
  const int buf_elems = ppb * chunk;
  long long epoch = 0;
  if (peers.sig) {  // signalled multi-GPU mode, as in fused_kernel
    epoch = *(volatile long long*)peers.epoch + 1;
    if (threadIdx.x < peers.n) {
      volatile long long* mine = peers.flags[peers.rank];
      while (mine[threadIdx.x] < epoch - 1) {
      }
    }
    __syncthreads();
  }
  double* const* out_ptrs = (peers.sig && (epoch & 1)) ? peers.ptr1 : peers.ptr;
  int use = 0;  // (group, chunk) iterations so far: buffer use & 1
  if (threadIdx.x < RL_OVERLAP_MARCHERS) {
    const float* __restrict__ dt = mv.dt;
    const unsigned W = (unsigned)mv.W, H = (unsigned)mv.H;
    for (int g = blockIdx.x; g < groups; g += gridDim.x) {
      const int p0 = g * ppb;
      const int np = min(ppb, N - p0);
      for (int c0 = 0; c0 < M; c0 += chunk, ++use) {
        const int cm = min(chunk, M - c0);
        const int rays = np * cm;
        const float rcp_cm = rcp_floor((unsigned)cm);
        double* v = vals + (use & 1) * buf_elems;
        if (use >= 2) named_bar_sync(3 + (use & 1), RL_OVERLAP_THREADS);  // FREE[b]
        for (int k0 = 0; k0 < rays; k0 += RL_OVERLAP_MARCHERS) {  // warp-uniform trip count
          const int k = k0 + threadIdx.x;
          const bool valid = k < rays;
          int p = 0, a = c0;
          float gx = 0.f, gy = 0.f, gth = 0.f;
          if (valid) {
            unsigned up, ua;
            divmod_small((unsigned)k, (unsigned)cm, rcp_cm, &up, &ua);  // k < ppb * cm <= 65536
            p = (int)up;
            a = c0 + (int)ua;
            const size_t i = perm ? (size_t)__ldg(perm + p0 + p) : (size_t)(p0 + p);  // processing order only
            float x, y, th;
            world_to_grid(xf, __ldg(ins + 3 * i), __ldg(ins + 3 * i + 1), __ldg(ins + 3 * i + 2), &x, &y, &th);
            gx = y;
            gy = x;
            gth = fsub(th, __ldg(angles + a));
          }
          float d;
          if (KIND == RL_RM) {
            RmSlot r;
            r.x0 = gx; r.y0 = gy; r.dx = 0.f; r.dy = 0.f;
            r.t = 0.0f;
            r.id = 0;
            const bool ok = valid && rm_setup(max_range, gx, gy, gth, &r.dx, &r.dy);
            r.busy = r.alive = ok;
            while (__any_sync(FULL, r.alive)) {
#pragma unroll
              for (int q = 0; q < 4; ++q) rm_step_pred<false>(dt, W, H, max_range, r);
            }
            d = ok ? rm_result(W, H, max_range, r) : max_range;
          } else {
            d = valid ? cast_one<KIND>(mv, cv, max_range, gx, gy, gth) : 0.0f;
          }
          if (valid) {
            const int di = sensor_index(d, kmax);                                   // :602-603 (no scaling)
            const int ri = sensor_index(fmul(__ldg(obs + a), xf.inv_scale), kmax);  // :605-606
            v[p * cm + (a - c0)] = __ldg(sv.table + (size_t)ri * sv.K + di);
          }
        }
        named_bar_arrive(1 + (use & 1), RL_OVERLAP_THREADS);  // FULL[b]
      }
    }
  } else {
    const int lane = threadIdx.x & 31;
    for (int g = blockIdx.x; g < groups; g += gridDim.x) {
      const int p0 = g * ppb;
      const int np = min(ppb, N - p0);
      double w = 1.0;  // running product, owned by lane p < np
      for (int c0 = 0; c0 < M; c0 += chunk, ++use) {
        const int cm = min(chunk, M - c0);
        const double* v = vals + (use & 1) * buf_elems + lane * cm;
        named_bar_sync(1 + (use & 1), RL_OVERLAP_THREADS);  // FULL[b]
        if (lane < np)
          for (int a = 0; a < cm; ++a) w = __dmul_rn(w, v[a]);  // reference order: beam ascending
        named_bar_arrive(3 + (use & 1), RL_OVERLAP_THREADS);  // FREE[b]
      }
      if (lane < np) {
        const size_t i = perm ? (size_t)__ldg(perm + p0 + lane) : (size_t)(p0 + lane);
        if (peers.n == 0) {
          weights[i] = w;
        } else {  // all-gather by direct peer stores (NVLink): every GPU gets this rank's slice
          for (int r = 0; r < peers.n; ++r) out_ptrs[r][peers.offset + i] = w;
        }
      }
    }
  }
  if (peers.sig) {  // publish, as in fused_kernel
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned done = atomicAdd(peers.counter, 1u);
      if (done == gridDim.x - 1) {
        *peers.counter = 0u;
        __threadfence_system();
        *(volatile long long*)peers.epoch = epoch;
        for (int r = 0; r < peers.n; ++r) ((volatile long long*)peers.flags[r])[peers.rank] = epoch;
      }
    }
  }
}

// ---- deep fused updates as two kernels: cast to memory, then this one ------------------------------------------------
// Measured on BASELINE config 5's map (8192^2, 200000 x 1080, tile-ordered particles): the lidar-fan cast alone runs at
// 42.7 G rays/s where fused_kernel reaches 26.1 -- the fused launch pays for its per-particle structure (rounds of 256
// rays that end with the slowest ray of the round, a barrier, the product), while the ranges cost 8 bytes of HBM
// traffic per ray there and back, 5 % of the bandwidth at that rate.  Fusing is a latency device for launches that fit
// the chip once; a deep update is the persistent lane-re-queuing cast into a scratch array followed by a streaming
// evaluation (launch_fused_twostep).  Measured and dropped: the table lookup in the cast's epilogue, the second kernel
// only multiplying 8-byte values -- 35.9 -> 33.1 G rays/s on the 8192^2 map, 38.6 -> 36.2 on the 5 cm map: the gather and
// its index arithmetic cost the issue-bound cast more than the streaming kernel pays for them.
// eval_overlap_kernel: the sensor-model half (RangeLib.h:533-555 / :596-610) in fused_overlap_kernel's form -- eight
// loader warps turn ranges into table values (coalesced range loads four deep, one table gather each) for a group of up
// to 32 particles and a chunk of beams, the ninth warp multiplies, lane p = particle p, beam order, running product
// carried over the chunks.  Rows of a value buffer are padded to an odd length (bank conflicts of the product warp).
// range_scale = inv_world_scale for eval_sensor_model (:549-550), 1 for the fused semantics (:602-603, exact).
__global__ void gather_poses_kernel(const float* __restrict__ ins, const int* __restrict__ perm, float* __restrict__ out,
                                    int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t s = (size_t)__ldg(perm + i);
  out[3 * (size_t)i] = __ldg(ins + 3 * s);
  out[3 * (size_t)i + 1] = __ldg(ins + 3 * s + 1);
  out[3 * (size_t)i + 2] = __ldg(ins + 3 * s + 2);
}

#ifndef RL_EVAL_MINB
#define RL_EVAL_MINB 4
#endif
__global__ void __launch_bounds__(RL_OVERLAP_THREADS, RL_EVAL_MINB)
eval_overlap_kernel(SensorView sv, float obs_scale, float range_scale, const float* __restrict__ obs,
                    const float* __restrict__ ranges, double* __restrict__ weights, int N, int M, int ppb, int chunk,
                    long long out_base, const int* __restrict__ perm, PeerOut peers, int sig_first, int sig_last) {
  extern __shared__ double vals[];  // two buffers of ppb * (chunk | 1) doubles
  const float kmax = (float)((double)(float)sv.K - 1.0);
  const int groups = (N + ppb - 1) / ppb;
  const int stride = chunk | 1;
  const int buf_elems = ppb * stride;
  long long epoch = 0;
  if (peers.sig) {  // signalled multi-GPU mode (fused_kernel): the first kernel of a step waits, the last publishes
    epoch = *(volatile long long*)peers.epoch + 1;
    if (sig_first) {
      if (threadIdx.x < peers.n) {
        volatile long long* mine = peers.flags[peers.rank];
        while (mine[threadIdx.x] < epoch - 1) {
        }
      }
      __syncthreads();
    }
  }
  double* const* out_ptrs = (peers.sig && (epoch & 1)) ? peers.ptr1 : peers.ptr;
  int use = 0;
  if (threadIdx.x < RL_OVERLAP_MARCHERS) {
    // Loader warp w takes particles w, w + 8, w + 16, w + 24 of the group (ppb <= 32); its lanes take beams lane and
    // lane + 32 of the chunk (chunk <= 64).  The observation's table row is therefore fixed per lane and chunk: one
    // row pointer per half, no index division, and every range load is a 128-byte row segment.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = blockIdx.x; g < groups; g += gridDim.x) {
      const int p0 = g * ppb;
      const int np = min(ppb, N - p0);
      for (int c0 = 0; c0 < M; c0 += chunk, ++use) {
        const int cm = min(chunk, M - c0);
        const bool h0 = lane < cm, h1 = lane + 32 < cm;
        const double* __restrict__ row0 =
            sv.table + (size_t)sensor_index(fmul(__ldg(obs + c0 + (h0 ? lane : 0)), obs_scale), kmax) * sv.K;
        const double* __restrict__ row1 =
            sv.table + (size_t)sensor_index(fmul(__ldg(obs + c0 + (h1 ? lane + 32 : 0)), obs_scale), kmax) * sv.K;
        double* v = vals + (use & 1) * buf_elems;
        float r[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int p = warp + 8 * q;
          const float* __restrict__ base = ranges + (size_t)(p0 + min(p, np - 1)) * M + c0;
          r[2 * q] = h0 ? __ldg(base + lane) : 0.0f;
          r[2 * q + 1] = h1 ? __ldg(base + lane + 32) : 0.0f;
        }
        if (use >= 2) named_bar_sync(3 + (use & 1), RL_OVERLAP_THREADS);  // FREE[b]
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int p = warp + 8 * q;
          const double t0 = __ldg(row0 + sensor_index(fmul(r[2 * q], range_scale), kmax));
          const double t1 = __ldg(row1 + sensor_index(fmul(r[2 * q + 1], range_scale), kmax));
          if (p < np) {
            if (h0) v[p * stride + lane] = t0;
            if (h1) v[p * stride + lane + 32] = t1;
          }
        }
        named_bar_arrive(1 + (use & 1), RL_OVERLAP_THREADS);  // FULL[b]
      }
    }
  } else {
    const int lane = threadIdx.x & 31;
    for (int g = blockIdx.x; g < groups; g += gridDim.x) {
      const int p0 = g * ppb;
      const int np = min(ppb, N - p0);
      double w = 1.0;
      for (int c0 = 0; c0 < M; c0 += chunk, ++use) {
        const int cm = min(chunk, M - c0);
        const double* v = vals + (use & 1) * buf_elems + lane * stride;
        named_bar_sync(1 + (use & 1), RL_OVERLAP_THREADS);  // FULL[b]
        if (lane < np)
          for (int a = 0; a < cm; ++a) w = __dmul_rn(w, v[a]);  // reference order: beam ascending
        named_bar_arrive(3 + (use & 1), RL_OVERLAP_THREADS);  // FREE[b]
      }
      if (lane < np) {
        const size_t i = perm ? (size_t)__ldg(perm + p0 + lane) : (size_t)(out_base + p0 + lane);
        if (peers.n == 0) {
          weights[i] = w;
        } else {
          for (int r = 0; r < peers.n; ++r) out_ptrs[r][peers.offset + i] = w;
        }
      }
    }
  }
  if (peers.sig && sig_last) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned done = atomicAdd(peers.counter, 1u);
      if (done == gridDim.x - 1) {
        *peers.counter = 0u;
        __threadfence_system();
        *(volatile long long*)peers.epoch = epoch;
        for (int r = 0; r < peers.n; ++r) ((volatile long long*)peers.flags[r])[peers.rank] = epoch;
      }
    }
  }
}

// consumer-side wait of the signalled mode: returns (on the stream) once every rank's slice of the
// latest epoch launched on this rank has arrived
__global__ void peers_wait_kernel(PeerOut peers) {
  const long long epoch = *(volatile long long*)peers.epoch;
  if (threadIdx.x < peers.n) {
    volatile long long* mine = peers.flags[peers.rank];
    while (mine[threadIdx.x] < epoch) {
    }
  }
}

// eval_sensor_model: same epilogue, ranges come from memory (coalesced tile load).
__global__ void __launch_bounds__(256)
eval_sensor_kernel(SensorView sv, float inv_scale, const float* __restrict__ obs, const float* __restrict__ ranges,
                   double* __restrict__ outs, int N, int M, int ppb, int chunk) {
  extern __shared__ double vals[];
  const float kmax = (float)((double)(float)sv.K - 1.0);
  const int groups = (N + ppb - 1) / ppb;
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const int p0 = g * ppb;
    const int np = min(ppb, N - p0);
    double w = 1.0;
    for (int c0 = 0; c0 < M; c0 += chunk) {
      const int cm = min(chunk, M - c0);
      for (int k = threadIdx.x; k < np * cm; k += blockDim.x) {
        const int p = k / cm, a = c0 + (k - p * cm);
        const int ri = sensor_index(fmul(__ldg(obs + a), inv_scale), kmax);                            // :547-548
        const int di = sensor_index(fmul(__ldg(ranges + (size_t)(p0 + p) * M + a), inv_scale), kmax);  // :549-550
        vals[p * cm + (a - c0)] = __ldg(sv.table + (size_t)ri * sv.K + di);
      }
      __syncthreads();
      if (threadIdx.x < np) {
        const double* v = vals + threadIdx.x * cm;
        for (int a = 0; a < cm; ++a) w = __dmul_rn(w, v[a]);
      }
      __syncthreads();
    }
    if (threadIdx.x < np) outs[p0 + threadIdx.x] = w;
  }
}

__global__ void sincosf_kernel(const float* x, float* s, float* c, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rl_sincosf(x[i], s + i, c + i);
}

// ------------------------------------------------------------------------------------------
// BL, large independent batches: persistent warps with lane re-queuing.  A Bresenham walk takes
// anything from 1 to ~500 steps (it stops at the first obstacle or at the map border); with one
// ray per thread ncu shows 10.9 of 32 threads active per instruction on an instruction-issue
// bound kernel (87 % issue slots busy; profiles/ncu_bl_r01.txt).  Here a warp owns a chunk of
// rays; the lanes walk in bursts of RL_BL_BURST steps, and between bursts the lanes whose walk
// has ended are handed the next rays of the chunk (their set-up -- bounds test, sin/cos, deltas,
// ~120 instructions -- runs for those lanes only, so it waits until RL_BL_REFILL lanes are idle).
// ------------------------------------------------------------------------------------------
#define RL_BL_REFILL 4
#define RL_BL_BURST 32
template <int MODE>
__global__ void __launch_bounds__(256, 5)
bl_persist_kernel(MapView mv, WorldXform xf, float max_range, const float* __restrict__ ins,
                  const float* __restrict__ angles, float* __restrict__ outs, long long total, int M, int chunk,
                  int refill_at, int burst_len, unsigned long long* __restrict__ work) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  // the warps claim pieces of `chunk` rays from a global cursor (see rm_persist_kernel): walks differ in length by two
  // orders of magnitude, and with fixed slices 27 % of the resident warp slots idled (ncu, round 2)
  long long begin = 0;
  int count = 0, next = 0;  // current piece, next ray of it to hand out (warp-uniform)
  bool exhausted = false;
  bool active = false;
  long long id = 0;
  BlState st;
  while (true) {
    const unsigned idle = __ballot_sync(FULL, !active);
    const int n_idle = __popc(idle);
    if (next >= count && !exhausted && (n_idle >= refill_at || n_idle == 32)) {
      unsigned long long b = 0;
      if (lane == 0) b = atomicAdd(work, (unsigned long long)chunk);
      b = __shfl_sync(FULL, b, 0);
      next = 0;
      if ((long long)b >= total) {
        exhausted = true;
        count = 0;
      } else {
        begin = (long long)b;
        count = (int)min((long long)chunk, total - begin);
      }
    }
    if (next < count && (n_idle >= refill_at || n_idle == 32)) {
      const int rank = __popc(idle & ((1u << lane) - 1u));
      if (!active && next + rank < count) {
        id = begin + next + rank;
        float gx, gy, gth, result;
        load_pose<MODE>(xf, ins, angles, id, M, &gx, &gy, &gth);
        if (bl_setup(mv, max_range, gx, gy, gth, st, &result)) {
          outs[id] = (MODE == MODE_GRID) ? result : fmul(result, xf.scale);
        } else {
          active = true;
        }
      }
      next = min(count, next + n_idle);
      continue;  // lanes whose new ray ended during set-up are refilled before stepping
    }
    if (n_idle == 32) break;  // nothing left to hand out and nobody walking
    if (active) {
      float result = max_range;
      bool done;
      int burst = burst_len;
      do {
        done = bl_step(mv, max_range, st, &result);
      } while (!done && --burst);
      if (done) {
        outs[id] = (MODE == MODE_GRID) ? result : fmul(result, xf.scale);
        active = false;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// RM, large independent batches: persistent warps with lane re-queuing.
//
// Sphere tracing needs 2..300 dependent distance-map reads per ray (mean ~6.5 on the basement
// maps) so one-ray-per-thread leaves two thirds of the lanes of a warp idle while its longest
// ray finishes (ncu, round 1: 10.7 of 32 threads active per instruction).  Here a warp owns a
// contiguous chunk of rays.  It alternates between
//   setup   all 32 lanes convergent: load pose, world->grid, (double-precision) sin/cos for the
//           next 128 rays of the chunk, parked as (x0, y0, dx, dy) in shared memory, and
//   march   a burst of sphere-tracing steps for every lane that holds a ray; a lane whose ray ended
//           writes its range and takes the next parked ray.
// In-flight rays keep their state in registers across a setup phase, so nothing drains between
// batches and the marching loop runs with most lanes busy until the chunk is exhausted.
// ------------------------------------------------------------------------------------------
#ifndef RL_QB
#define RL_QB 4  // parked rays per lane and setup phase (shared-memory variant)
#endif
#ifndef RL_FUSED_GROUP_RAYS
#define RL_FUSED_GROUP_RAYS 4096  // rays of one particle group in fused_rm_persist_kernel (2048: -8 %, 6144: -30 %)
#endif
#ifndef RL_RM_BURST_PAIRS
#define RL_RM_BURST_PAIRS 4  // sphere-tracing steps per refill round = 2 * this (3 -> 4: lidar fans 39.5 -> 40.3 G rays/s, others unchanged)
#endif

// ------------------------------------------------------------------------------------------
// Shape of the marching loop (round-1 ncu source view of the first version of this kernel: 19 % of all issued
// warp-instructions were the hit epilogue -- int->float, squares, sqrt -- run inside the stepping loop for an
// average of 2.4 lanes; another ~8 % were BSSY/BSYNC/branch bookkeeping of the loop's three exits; and 45 % of
// the stall samples sat on the instruction after the distance load).
//  * The stepping burst is straight-line predicated code: a lane whose ray has ended keeps its t (which, by
//    construction, reproduces the end condition: t >= max_range, or the out-of-map / obstacle cell at
//    (int)(x0 + dx t), (int)(y0 + dy t)) and simply stops loading; no branch, no reconvergence point.
//  * The epilogue runs once per refill round for all lanes that ended in the burst, classifying from t alone.
//  * SLOTS rays per lane are marched interleaved (independent dependent-load chains), which doubles the loads
//    in flight per warp at the same occupancy.
// Arithmetic per step is rm_step's, i.e. the reference's.
// ------------------------------------------------------------------------------------------
// PARK_REGS: the rays set up ahead of the march are parked one per lane in registers and handed out with
// shuffles, instead of RL_QB per lane in shared memory: the kernel then uses no shared memory at all and the
// whole 256 KB of the SM serves as L1 for the distance-map gathers (which are what bounds it: ncu shows the
// L1 -> crossbar miss-request interface busy 85 % of the active cycles, one missing sector per clock per SM).
#ifndef RL_RM_PERSIST_MINB
#define RL_RM_PERSIST_MINB 6
#endif
template <int MODE, int SLOTS, bool COND_LOAD, bool PARK_REGS>
__global__ void __launch_bounds__(256, SLOTS == 1 ? RL_RM_PERSIST_MINB : 4)
rm_persist_kernel(MapView mv, WorldXform xf, float max_range, const float* __restrict__ ins,
                  const float* __restrict__ angles, float* __restrict__ outs, long long total, int M, int chunk,
                  int burst_pairs, unsigned long long* __restrict__ work) {
  constexpr int PARKED = PARK_REGS ? 32 : RL_QB * 32;  // rays per setup phase
  __shared__ float4 q_all[PARK_REGS ? 1 : 8 * RL_QB * 32];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float4* q = q_all + (PARK_REGS ? 0 : wib * (RL_QB * 32));
  float4 parked = make_float4(0.f, 0.f, 0.f, 0.f);
  const float out_scale = (MODE == MODE_GRID) ? 1.0f : xf.scale;
  const float* __restrict__ dt = mv.dt;
  const unsigned W = (unsigned)mv.W, H = (unsigned)mv.H;

  // Work distribution (round 2): the warps CLAIM pieces of `chunk` rays from a global cursor instead of owning a fixed
  // slice of the batch.  With fixed slices the launch ended when the slowest warp did, and ncu showed 59 % (lidar fans)
  // / 70 % (random rays) of the warp slots occupied on average where 75 % were resident: warps that had finished their
  // slice idled.  A warp asks for the next piece as soon as the rays of the current one have all been handed to lanes,
  // so rays of two pieces march side by side and there is no per-piece tail.
  long long begin = 0;  // first ray of the current piece
  int count = 0;        // its length (bookkeeping relative to `begin`)
  bool exhausted = false;
  // lidar fans: ray begin + k is beam (a0 + k) % M of particle p0 + (a0 + k) / M -- one 64-bit division per piece
  long long fan_p0 = 0;
  unsigned fan_a0 = 0u;
  const float fan_rcp = (MODE == MODE_ANGLES) ? rcp_floor((unsigned)M) : 0.0f;
  const bool fan_fast = (MODE == MODE_ANGLES) && M < (1 << 22);
  int next_setup = 0, batch_base = 0, batch_n = 0, batch_pos = 0;
  RmSlot r[SLOTS];
  long long rid[SLOTS];  // absolute index of the ray a slot holds
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    r[s].x0 = r[s].y0 = r[s].dx = r[s].dy = r[s].t = 0.f;
    r[s].id = 0;
    rid[s] = 0;
    r[s].busy = r[s].alive = false;
  }

  while (true) {
    bool any_busy = false;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      // with 32 parked rays a refill can exhaust them and still leave lanes idle: a second pass sets up more
#pragma unroll
      for (int pass = 0; pass < (PARK_REGS ? 2 : 1); ++pass) {
        const unsigned idle = __ballot_sync(FULL, !r[s].busy);
        if (idle) {
          if (batch_pos == batch_n && next_setup >= count && !exhausted) {  // next piece (warp-uniform)
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(work, (unsigned long long)chunk);
            b = __shfl_sync(FULL, b, 0);
            next_setup = 0;
            if ((long long)b >= total) {
              exhausted = true;
              count = 0;
            } else {
              begin = (long long)b;
              count = (int)min((long long)chunk, total - begin);
              if (MODE == MODE_ANGLES) {
                fan_p0 = begin / M;
                fan_a0 = (unsigned)(begin - fan_p0 * M);
              }
            }
          }
          if (batch_pos == batch_n && next_setup < count) {
            // ---- setup phase (warp-uniform branch): pose -> (x0, y0, cos, sin) for the next PARKED rays ----
            __syncwarp();
            batch_base = next_setup;
            batch_n = min(PARKED, count - next_setup);
            batch_pos = 0;
            for (int e = lane; e < batch_n; e += 32) {
              float gx, gy, gth;
              if (MODE == MODE_ANGLES && fan_fast) {
                unsigned dp, a;
                divmod_small(fan_a0 + (unsigned)(batch_base + e), (unsigned)M, fan_rcp, &dp, &a);
                const long long i = fan_p0 + dp;
                float x, y, th;
                world_to_grid(xf, __ldg(ins + 3 * i), __ldg(ins + 3 * i + 1), __ldg(ins + 3 * i + 2), &x, &y, &th);
                gx = y;  // calc_range(y, x, theta): RangeLib.h:518
                gy = x;
                gth = fsub(th, __ldg(angles + a));
              } else {
                load_pose<MODE>(xf, ins, angles, begin + batch_base + e, M, &gx, &gy, &gth);
              }
              float sn = 0.f, cs = 0.f;
              const bool ok = finite3(gx, gy, gth);
              if (ok) rl_sincosf(gth, &sn, &cs);
              // a non-finite pose leaves the map on its first step -> max_range
              const float4 ray = ok ? make_float4(gx, gy, cs, sn) : make_float4(-1e30f, 0.f, 0.f, 0.f);
              if (PARK_REGS) parked = ray; else q[e] = ray;
            }
            next_setup += batch_n;
            __syncwarp();
          }
          const int avail = batch_n - batch_pos;
          if (avail > 0) {
            const int rank = __popc(idle & ((1u << lane) - 1u));
            const bool take = !r[s].busy && rank < avail;
            float4 ray;
            if (PARK_REGS) {
              const int src = (batch_pos + rank) & 31;
              ray.x = __shfl_sync(FULL, parked.x, src);
              ray.y = __shfl_sync(FULL, parked.y, src);
              ray.z = __shfl_sync(FULL, parked.z, src);
              ray.w = __shfl_sync(FULL, parked.w, src);
            } else if (take) {
              ray = q[batch_pos + rank];
            }
            if (take) {
              r[s].x0 = ray.x; r[s].y0 = ray.y; r[s].dx = ray.z; r[s].dy = ray.w;
              r[s].t = 0.0f;
              rid[s] = begin + batch_base + batch_pos + rank;
              r[s].busy = r[s].alive = true;
            }
            batch_pos += min(avail, __popc(idle));
          }
        }
      }
      any_busy = any_busy || r[s].busy;
    }
    if (!__any_sync(FULL, any_busy)) break;

#pragma unroll 1
    for (int b = 0; b < burst_pairs; ++b) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) rm_step_pred<COND_LOAD>(dt, W, H, max_range, r[s]);
      }
    }

#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      if (r[s].busy && !r[s].alive) {
        const float result = rm_result(W, H, max_range, r[s]);
        if (MODE == MODE_GLT_BUILD) {
          // r = min(max_range, r); uint16 val = r * limits_div_max (RangeLib.h:1804-1806); xf.scale carries the factor
          const float rr = (result < max_range) ? result : max_range;
          ((uint16_t*)outs)[rid[s]] = (uint16_t)__float2int_rz(fmul(rr, xf.scale));
        } else {
          outs[rid[s]] = (MODE == MODE_GRID) ? result : fmul(result, out_scale);
        }
        r[s].busy = false;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fused RM sensor update for launches many waves deep (a global-localisation cloud, BASELINE config 5).
// fused_kernel marches one ray per thread to completion, which is what a launch that fits the chip once wants
// (with its cooperative tail); deep launches are throughput problems and get the marching loop of
// rm_persist_kernel instead: a CTA takes a group of particles whose rays (<= 2048) it numbers 0..R-1; its warps
// draw batches of 32 rays from a shared counter, set them up convergently, park them in registers, and every
// lane whose ray has ended takes the next parked one.  A finished ray goes straight through the sensor table
// into shared memory; when the group's rays are done one thread per particle forms the product in beam order
// (bit-identical to the reference's sequential product).  Ranges never reach HBM.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 5)
fused_rm_persist_kernel(MapView mv, WorldXform xf, SensorView sv, float max_range, const float* __restrict__ ins,
                        const float* __restrict__ angles, const float* __restrict__ obs,
                        double* __restrict__ weights, int N, int M, int ppb, int chunk, PeerOut peers,
                        const int* __restrict__ perm, int burst_pairs, unsigned long long* __restrict__ work) {
  extern __shared__ double vals[];
  __shared__ int s_next, s_group;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const float kmax = (float)((double)(float)sv.K - 1.0);
  const float* __restrict__ dt = mv.dt;
  const unsigned W = (unsigned)mv.W, H = (unsigned)mv.H;
  const int groups = (N + ppb - 1) / ppb;
  while (true) {
    // the CTAs claim particle groups from a global cursor (groups differ in how long their rays are)
    if (threadIdx.x == 0) s_group = (int)atomicAdd(work, 1ULL);
    __syncthreads();
    const int g = s_group;
    if (g >= groups) break;
    const int p0 = g * ppb;
    const int np = min(ppb, N - p0);
    double w = 1.0;  // running product, owned by thread p < np
    for (int c0 = 0; c0 < M; c0 += chunk) {
      const int cm = min(chunk, M - c0);
      const int rays = np * cm;
      const float rcp_cm = rcp_floor((unsigned)cm);
      if (threadIdx.x == 0) s_next = 0;
      __syncthreads();
      // ---- this warp's share of the group's rays: lane re-queuing as in rm_persist_kernel ----
      RmSlot r;
      r.x0 = r.y0 = r.dx = r.dy = r.t = 0.f;
      r.id = 0;
      r.busy = r.alive = false;
      float4 parked = make_float4(0.f, 0.f, 0.f, 0.f);
      int batch_base = 0, batch_n = 0, batch_pos = 0;
      bool more = true;
      while (true) {
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const unsigned idle = __ballot_sync(FULL, !r.busy);
          if (idle) {
            if (batch_pos == batch_n && more) {
              int b = 0;
              if (lane == 0) b = atomicAdd(&s_next, 32);
              b = __shfl_sync(FULL, b, 0);
              batch_pos = 0;
              if (b >= rays) {
                more = false;
                batch_n = 0;
              } else {
                batch_base = b;
                batch_n = min(32, rays - b);
                if (lane < batch_n) {
                  const int k = b + lane;
                  unsigned up, ua;
                  divmod_small((unsigned)k, (unsigned)cm, rcp_cm, &up, &ua);  // k < group rays
                  const int p = (int)up;
                  const int a = c0 + (int)ua;
                  const size_t i = perm ? (size_t)__ldg(perm + p0 + p) : (size_t)(p0 + p);
                  float x, y, th;
                  world_to_grid(xf, __ldg(ins + 3 * i), __ldg(ins + 3 * i + 1), __ldg(ins + 3 * i + 2), &x, &y, &th);
                  const float gth = fsub(th, __ldg(angles + a));
                  float sn = 0.f, cs = 0.f;
                  const bool ok = finite3(y, x, gth);
                  if (ok) rl_sincosf(gth, &sn, &cs);
                  // calc_range(y, x, theta) RangeLib.h:594; a non-finite pose leaves the map at once -> max_range
                  parked = ok ? make_float4(y, x, cs, sn) : make_float4(-1e30f, 0.f, 0.f, 0.f);
                }
              }
            }
            const int avail = batch_n - batch_pos;
            if (avail > 0) {
              const int rank = __popc(idle & ((1u << lane) - 1u));
              const bool take = !r.busy && rank < avail;
              const int src = (batch_pos + rank) & 31;
              const float px = __shfl_sync(FULL, parked.x, src), py = __shfl_sync(FULL, parked.y, src);
              const float pz = __shfl_sync(FULL, parked.z, src), pw = __shfl_sync(FULL, parked.w, src);
              if (take) {
                r.x0 = px; r.y0 = py; r.dx = pz; r.dy = pw;
                r.t = 0.0f;
                r.id = batch_base + batch_pos + rank;
                r.busy = r.alive = true;
              }
              batch_pos += min(avail, __popc(idle));
            }
          }
        }
        if (!__any_sync(FULL, r.busy)) break;
#pragma unroll 1
        for (int b = 0; b < burst_pairs; ++b) {
          rm_step_pred<false>(dt, W, H, max_range, r);
          rm_step_pred<false>(dt, W, H, max_range, r);
        }
        if (r.busy && !r.alive) {
          const float d = rm_result(W, H, max_range, r);
          const int k = r.id;
          unsigned up, ua;
          divmod_small((unsigned)k, (unsigned)cm, rcp_cm, &up, &ua);
          const int a = c0 + (int)ua;
          const int di = sensor_index(d, kmax);                                   // :602-603 (no scaling)
          const int ri = sensor_index(fmul(__ldg(obs + a), xf.inv_scale), kmax);  // :605-606
          vals[k] = __ldg(sv.table + (size_t)ri * sv.K + di);
          r.busy = false;
        }
      }
      __syncthreads();
      if (threadIdx.x < np) {
        const double* v = vals + threadIdx.x * cm;
        for (int a = 0; a < cm; ++a) w = __dmul_rn(w, v[a]);  // reference order: beam ascending
      }
      __syncthreads();
    }
    if (threadIdx.x < np) {
      const size_t i = perm ? (size_t)__ldg(perm + p0 + threadIdx.x) : (size_t)(p0 + threadIdx.x);
      if (peers.n == 0) {
        weights[i] = w;
      } else {  // all-gather by direct peer stores (NVLink): every GPU gets this rank's slice
        for (int rr = 0; rr < peers.n; ++rr) peers.ptr[rr][peers.offset + i] = w;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
// SM count of the CURRENT device (the ABI layer has made the handle's device current), cached per device
static int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = cached[dev & 63];
  if (!c) {
    cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
    if (c <= 0) c = 148;
  }
  return c;
}

// CTAs of a persistent kernel that are resident per SM (its grid must be exactly one wave: the launch bound only caps
// the register count, what fits is decided by the count ptxas ended up with).  RL_PERSIST_CTAS overrides.
template <class K>
static int resident_ctas(K kernel, int threads, int fallback, size_t dyn_smem = 0) {
  static const int forced = getenv("RL_PERSIST_CTAS") ? atoi(getenv("RL_PERSIST_CTAS")) : 0;
  if (forced > 0) return forced;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, dyn_smem) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = fallback;
  }
  return n;
}

// the ray cursor of the persistent kernels, zeroed on the handle's stream before the launch that uses it
static int fresh_work_cursor(rl_method* m, unsigned long long** out) {
  if (!m->d_work) RL_CUDA(cudaMalloc(&m->d_work, sizeof(unsigned long long)));
  RL_CUDA(cudaMemsetAsync(m->d_work, 0, sizeof(unsigned long long), m->stream));
  *out = m->d_work;
  return RL_OK;
}
// rays a warp of a persistent kernel claims at a time: RL_CLAIM_RAYS (default 128; measured 128 / 256 / 512 / 1024 on
// 2^24 random rays 38.5 / 38.9 / 38.3 / 34.8 G rays/s, BL 2^22 rays 2.46 / 2.28 / 2.42 / -), less when the batch is
// small, so that every resident warp gets about four pieces
// (lidar fans of >= 256 beams take 512 at a time -- half a particle's fan, neighbouring beams in one warp: 42.9 -> 43.6 G
// rays/s on config 5's map, 200000 x 1080)
static int claim_rays(long long total, long long resident_warps, int fan_beams = 0) {
  static const int env = getenv("RL_CLAIM_RAYS") ? atoi(getenv("RL_CLAIM_RAYS")) : 0;
  const int v = env > 0 ? env : (fan_beams >= 256 ? 512 : 128);
  static const int pieces = getenv("RL_CLAIM_PIECES") ? max(1, atoi(getenv("RL_CLAIM_PIECES"))) : 8;
  const long long even = (total / (resident_warps * pieces) + 31) / 32 * 32;
  return (int)max(32LL, min((long long)max(32, (v / 32) * 32), even));
}

// CTA size of SMALL fused launches (the ones that fit the chip about once and run the cooperative tail)
static int small_fused_threads() {
  static const int v = getenv("RL_FUSED_SMALL_THREADS") ? atoi(getenv("RL_FUSED_SMALL_THREADS")) : 256;
  return (v == 64 || v == 128 || v == 192) ? v : 256;
}

static int block_burst_pairs() {  // tuning knob of rm_march_block (RL_BLOCK_BURST_PAIRS), default RL_BLOCK_BURST / 2
  static const int v = getenv("RL_BLOCK_BURST_PAIRS") ? max(1, atoi(getenv("RL_BLOCK_BURST_PAIRS"))) : RL_BLOCK_BURST / 2;
  return v;
}

// eval_overlap_kernel over n particles whose ranges start at `ranges`; weight of particle i goes to perm[i] if perm,
// else to out_base + i
static int launch_eval_overlap(rl_method* m, float range_scale, const float* obs, const float* ranges, double* weights,
                               int n, int M, long long out_base, const int* perm, const PeerOut& po, bool sig_first,
                               bool sig_last) {
  const int pieces = (M + 63) / 64;
  const int chunk = max(1, (M + pieces - 1) / pieces);  // <= 64 beams at a time, pieces of equal length
  const int ppb = 32;
  const size_t smem = 2 * (size_t)ppb * (chunk | 1) * sizeof(double);
  const int groups = (n + ppb - 1) / ppb;
  const int grid = max(1, min(groups, sm_count() * resident_ctas(eval_overlap_kernel, RL_OVERLAP_THREADS, 6, smem)));
  eval_overlap_kernel<<<grid, RL_OVERLAP_THREADS, smem, m->stream>>>(m->sensor_view(), m->xf.inv_scale, range_scale, obs,
                                                                    ranges, weights, n, M, ppb, chunk, out_base, perm, po,
                                                                    sig_first ? 1 : 0, sig_last ? 1 : 0);
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

template <int KIND>
static int launch_cast_kind(rl_method* m, int mode, const float* ins, const float* angles, const float* obs,
                            float* outs, double* weights, int n, int M, const PeerOut* peers);

// Scratch of a two-kernel deep update: ranges of one chunk of particles (RL_TWOSTEP_SCRATCH_MB, default 1024; a quarter
// and a sixteenth of it are tried when the allocation fails) and, for tile-ordered clouds, the chunk's poses in
// processing order.  Returns the particles per chunk, or 0 when there is no memory for it -- the caller then keeps the
// fused kernels, which need none.  (Allocates on first use, like the handle's other scratch: capture a deep update into
// a CUDA graph only after one warm-up call.)
static int twostep_chunk(rl_method* m, int n, int M, bool ordered) {
  static const size_t scratch_mb = getenv("RL_TWOSTEP_SCRATCH_MB") ? (size_t)max(16, atoi(getenv("RL_TWOSTEP_SCRATCH_MB"))) : 1024;
  const size_t per_particle = (size_t)M * sizeof(float);
  for (size_t budget = scratch_mb << 20; budget >= ((size_t)16 << 20); budget >>= 2) {
    int cap = (int)max((size_t)1024, min((size_t)n, budget / per_particle));
    const int nchunks = (n + cap - 1) / cap;
    cap = (n + nchunks - 1) / nchunks;  // chunks of equal size
    const size_t want = (size_t)cap * per_particle;
    bool ok = true;
    if (want > m->ts_ranges_bytes) {
      cudaFree(m->d_ts_ranges);
      m->d_ts_ranges = nullptr;
      m->ts_ranges_bytes = 0;
      ok = cudaMalloc(&m->d_ts_ranges, want) == cudaSuccess;
      if (ok) m->ts_ranges_bytes = want;
    }
    if (ok && ordered && (size_t)cap * 12 > m->ts_poses_bytes) {
      cudaFree(m->d_ts_poses);
      m->d_ts_poses = nullptr;
      m->ts_poses_bytes = 0;
      ok = cudaMalloc(&m->d_ts_poses, (size_t)cap * 12) == cudaSuccess;
      if (ok) m->ts_poses_bytes = (size_t)cap * 12;
    }
    if (ok) return cap;
    cudaGetLastError();  // out of memory: try a smaller chunk
  }
  return 0;
}

// A deep fused update as cast-to-memory + evaluation, in chunks of `cap` particles (twostep_chunk).  perm (tile order,
// or null) is the processing order of the whole cloud.
template <int KIND>
static int launch_fused_twostep(rl_method* m, const float* ins, const float* angles, const float* obs, double* weights,
                                int n, int M, const PeerOut& po, const int* perm, int cap) {
  for (int c0 = 0; c0 < n; c0 += cap) {
    const int np = min(cap, n - c0);
    const float* poses = ins + 3 * (size_t)c0;
    if (perm) {
      gather_poses_kernel<<<(np + 255) / 256, 256, 0, m->stream>>>(ins, perm + c0, m->d_ts_poses, np);
      count_launch();
      RL_CHECK_LAUNCH();
      poses = m->d_ts_poses;
    }
    // the fused update indexes the table with the range in PIXELS (RangeLib.h:602-603): the cast must not apply the
    // world scale that numpy_calc_range_angles applies to its outputs (x 1.0f is exact)
    const float world_scale = m->xf.scale;
    m->xf.scale = 1.0f;
    const int rc = launch_cast_kind<KIND>(m, MODE_ANGLES, poses, angles, nullptr, m->d_ts_ranges, nullptr, np, M, nullptr);
    m->xf.scale = world_scale;
    if (rc) return rc;
    const int rc2 = launch_eval_overlap(m, 1.0f, obs, m->d_ts_ranges, weights, np, M, c0, perm ? perm + c0 : nullptr, po,
                                        c0 == 0, c0 + cap >= n);
    if (rc2) return rc2;
  }
  return RL_OK;
}

template <int KIND>
static int launch_cast_kind(rl_method* m, int mode, const float* ins, const float* angles, const float* obs,
                            float* outs, double* weights, int n, int M, const PeerOut* peers) {
  MapView mv = m->map_view();
  const CddtView cv = m->cddt_view();
  const int threads = 256;
  // The cooperative tail is a latency device for launches that fit the chip about once (a particle-filter
  // update): there the longest rays decide the kernel time.  When the launch is many waves deep other CTAs
  // hide that latency and the hand-off only costs barriers and idle warps (measured 34 vs 21 G rays/s on
  // 20000 x 1080), so it is switched off.
  {
    const long long rays = (mode >= MODE_ANGLES) ? (long long)n * M : (long long)n;
    if (rays > 2LL * sm_count() * 7 * threads) mv.coop_threshold = 0;
    mv.block_burst_pairs = block_burst_pairs();
  }
  if (mode == MODE_FUSED) {
    if (!m->d_table) {
      set_error("calc_range_repeat_angles_eval_sensor_model: set_sensor_model has not been called");
      return RL_E_STATE;
    }
    const int chunk = min(M, 2048);
    // Deep launches of many-beam particles (BASELINE config 5: 10^6 x 1080): a CTA handles one particle at a time and
    // stops at a barrier while one thread forms the 1080-term product.  Measured alternatives, all bit-exact, none
    // faster (round 2, 8192^2 map / 5 cm map 20000 x 1080, G rays/s): 256-thread CTAs 26.5 / 32.2 (kept);
    // 128-thread CTAs (RL_FUSED_DEEP_THREADS=128) 23.6 / 33.3; 64-thread 16.8 / 23.5; one warp per particle with the
    // product slipped into the next batch's march, no barrier at all, 25.3 / 31.4 (ncu: issue slots 70 % busy instead
    // of 60 %, but 12 % more instructions -- the launch is bound by instruction issue either way); producer warps +
    // one consumer warp streaming rays through shared-memory rings 8.1 / 7.0.
    static const int deep_threads = getenv("RL_FUSED_DEEP_THREADS") ? atoi(getenv("RL_FUSED_DEEP_THREADS")) : 256;
    const bool many_beams = mv.coop_threshold == 0 && M >= 512 && deep_threads >= 32 && deep_threads < 256 && deep_threads % 8 == 0;
    const int fthreads = many_beams ? deep_threads : (mv.coop_threshold != 0 ? small_fused_threads() : threads);
    const int ppb = max(1, min(fthreads / max(M, 1), 32));
    const int groups = (n + ppb - 1) / ppb;
    const size_t smem = (size_t)ppb * chunk * sizeof(double);
    int per_sm = 8;  // resident CTAs per SM: the grid-stride loop wants exactly one wave
    if (many_beams &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_kernel<KIND, false>, fthreads, smem) != cudaSuccess) {
      cudaGetLastError();
      per_sm = 8;
    }
    const int grid = max(1, min(groups, sm_count() * max(per_sm, 1) * (many_beams ? 1 : 256 / fthreads)));
    PeerOut po{};
    if (peers) po = *peers;
    // big clouds on structures larger than L2: process the particles tile by tile (rl_sort.cu)
    static const bool spatial = !(getenv("RL_SPATIAL_SORT") && atoi(getenv("RL_SPATIAL_SORT")) == 0);
    const int* perm = nullptr;
    const size_t struct_bytes = (KIND == RL_RM) ? m->dt_elems() * sizeof(float)
                                : (KIND == RL_GLT) ? m->dt_elems() * (size_t)m->td * sizeof(uint16_t) : 0;
    if (spatial && m->spatial_sort && n >= 32768 && struct_bytes > ((size_t)48 << 20)) {
      const int rc = spatial_order(m, ins, n, &perm);
      if (rc) return rc;
    }
    // deep updates: cast to memory + evaluation (see eval_overlap_kernel).  RL_FUSED_TWOSTEP=0 keeps the fused kernels.
    static const int twostep_env = getenv("RL_FUSED_TWOSTEP") ? atoi(getenv("RL_FUSED_TWOSTEP")) : -1;
    // Measured per kind (B200, G rays/s, fused kernels -> two kernels; profiles/r02/twostep_r02.log): RM on the 8192^2
    // map 200000 x 1080 25.9 -> 35.9; RM on the 5 cm map 20000 x 1080 37.7 -> 38.6, 50000 x 360 33.5 -> 33.9, 100000 x 60
    // 26.7 -> 25.8 (the re-queuing fused kernel keeps short fans on structures that fit L2); CDDT 100000 x 60 37.4 ->
    // 43.0, 50000 x 360 51.3 -> 62.3, 20000 x 1080 58.5 -> 75.3; GiantLUT 42.8 -> 44.3 / 42.6 -> 65.4 / 58.4 -> 94.4;
    // BL 2000 x 1080 2.42 -> 2.79.  RL_FUSED_TWOSTEP = bit mask of kinds (1 << kind) overrides.
    const bool twostep_kind = twostep_env >= 0 ? (twostep_env & (1 << KIND)) != 0
                                               : (KIND != RL_RM || M > 128 || struct_bytes > ((size_t)48 << 20));
    if (twostep_kind && mv.coop_threshold == 0 && m->persist && m->max_range > 0.0f &&
        (long long)n * M >= (long long)sm_count() * 48 * 64 * 4) {
      const int cap = twostep_chunk(m, n, M, perm != nullptr);
      if (cap > 0) return launch_fused_twostep<KIND>(m, ins, angles, obs, weights, n, M, po, perm, cap);
    }
    static const bool deep = !(getenv("RL_FUSED_PERSIST") && atoi(getenv("RL_FUSED_PERSIST")) == 0);
    static const int group_rays = getenv("RL_FUSED_GROUP_RAYS") ? atoi(getenv("RL_FUSED_GROUP_RAYS")) : RL_FUSED_GROUP_RAYS;
    // many waves deep and at least six particles per group: the re-queuing kernel (measured, basement 5 cm map:
    // 100000 x 60 16.8 -> 25.0 G rays/s, 50000 x 360 22.2 -> 30.2; 8192^2 map, 10^6 x 60 14.2 -> 22.0).  With few
    // ~1000-beam particles per group the neighbouring beams of fused_kernel's warps already finish together and
    // the per-group tail of the re-queuing loop costs more than it saves (20000 x 1080: 33.9 vs 30.0).
    if (KIND == RL_RM && deep && mv.coop_threshold == 0 && m->max_range > 0.0f && !po.sig && m->persist &&
        6 * M <= group_rays) {
      const int ppb2 = max(1, min(group_rays / max(M, 1), 128));
      const int groups2 = (n + ppb2 - 1) / ppb2;
      const int grid2 = max(1, min(groups2, sm_count() * resident_ctas(fused_rm_persist_kernel, 256, 5,
                                                                         (size_t)ppb2 * chunk * sizeof(double))));
      unsigned long long* work = nullptr;
      const int rcw = fresh_work_cursor(m, &work);
      if (rcw) return rcw;
      fused_rm_persist_kernel<<<grid2, 256, (size_t)ppb2 * chunk * sizeof(double), m->stream>>>(
          mv, m->xf, m->sensor_view(), m->max_range, ins, angles, obs, weights, n, M, ppb2, chunk, po, perm,
          RL_RM_BURST_PAIRS, work);
    } else {
      // deep launches: the product of a group overlaps the march of the next one (fused_overlap_kernel);
      // RL_FUSED_OVERLAP=0 keeps fused_kernel (A/B)
      static const bool overlap = !(getenv("RL_FUSED_OVERLAP") && atoi(getenv("RL_FUSED_OVERLAP")) == 0);
      const int ppb_o = max(1, min(RL_OVERLAP_MARCHERS / max(M, 1), 32));
      const int groups_o = (n + ppb_o - 1) / ppb_o;
      const size_t smem_o = 2 * (size_t)ppb_o * chunk * sizeof(double);
      const int grid_o = sm_count() * resident_ctas(fused_overlap_kernel<KIND>, RL_OVERLAP_THREADS, 7, smem_o);
      // measured per kind (5 cm map, G rays/s, fused_kernel -> this kernel): RM 20000 x 1080 34.2 -> 37.8; CDDT
      // 100000 x 60 29.4 -> 37.7, 50000 x 360 36.8 -> 51.5, 20000 x 1080 44.7 -> 59.0; BL 20000 x 60 1.35 -> 1.13 (its
      // long walks are bound by issue slots and want the eighth CTA per SM more than the hidden product) -> not for BL
      const bool fits_l2 = struct_bytes <= ((size_t)48 << 20) && (KIND != RL_CDDT || !cv.meta);
      if (overlap && KIND != RL_BL && mv.coop_threshold == 0 && !many_beams && fits_l2 && groups_o >= 2 * grid_o) {
        fused_overlap_kernel<KIND><<<grid_o, RL_OVERLAP_THREADS, smem_o, m->stream>>>(
            mv, cv, m->xf, m->sensor_view(), m->max_range, ins, angles, obs, weights, n, M, ppb_o, chunk, po, perm);
      } else {
        fused_kernel<KIND, false><<<grid, fthreads, smem, m->stream>>>(mv, cv, m->xf, m->sensor_view(), m->max_range,
                                                                       ins, angles, obs, weights, n, M, ppb, chunk, po,
                                                                       NoBeamParams{}, perm);
      }
    }
  } else {
    const long long total = (mode == MODE_ANGLES) ? (long long)n * M : (long long)n;
    // RM with enough rays to give every resident warp more than one ray per lane: persistent
    // warps with lane re-queuing.  (max_range <= 0 never enters the marching loop.)
    if (KIND == RL_BL && total >= (long long)sm_count() * 40 * 64 && m->persist) {
      static const int bl_refill = getenv("RL_BL_REFILL") ? atoi(getenv("RL_BL_REFILL")) : RL_BL_REFILL;
      static const int bl_burst = getenv("RL_BL_BURST") ? atoi(getenv("RL_BL_BURST")) : RL_BL_BURST;
      const int ctas = mode == MODE_GRID    ? resident_ctas(bl_persist_kernel<MODE_GRID>, 256, 5)
                       : mode == MODE_WORLD ? resident_ctas(bl_persist_kernel<MODE_WORLD>, 256, 5)
                                            : resident_ctas(bl_persist_kernel<MODE_ANGLES>, 256, 5);
      const long long rw = (long long)sm_count() * 8 * ctas;
      const int chunk = claim_rays(total, rw);
      const long long warps = min(rw, (total + chunk - 1) / chunk);
      const int grid = (int)((warps + 7) / 8);
      unsigned long long* work = nullptr;
      const int rcw = fresh_work_cursor(m, &work);
      if (rcw) return rcw;
      if (mode == MODE_GRID)
        bl_persist_kernel<MODE_GRID><<<grid, 256, 0, m->stream>>>(mv, m->xf, m->max_range, ins, angles, outs, total, M, chunk, bl_refill, bl_burst, work);
      else if (mode == MODE_WORLD)
        bl_persist_kernel<MODE_WORLD><<<grid, 256, 0, m->stream>>>(mv, m->xf, m->max_range, ins, angles, outs, total, M, chunk, bl_refill, bl_burst, work);
      else
        bl_persist_kernel<MODE_ANGLES><<<grid, 256, 0, m->stream>>>(mv, m->xf, m->max_range, ins, angles, outs, total, M, chunk, bl_refill, bl_burst, work);
      count_launch();
      RL_CHECK_LAUNCH();
      return RL_OK;
    }
    // RL_RM_PERSIST (tuning / A-B; measured on 2^24 random rays, basement_hallways_5cm):
    //   0 off | 1 default: parked rays in registers, unconditional load (39.6 G rays/s) | 2 the same with a
    //   conditional load (39.1) | 3 parked rays in shared memory (37.3) | 4 = 3 with two rays per lane (36.9)
    static const int rm_persist_env = getenv("RL_RM_PERSIST") ? atoi(getenv("RL_RM_PERSIST")) : -1;
    static const int rm_burst_pairs = getenv("RL_RM_BURST_PAIRS") ? max(1, atoi(getenv("RL_RM_BURST_PAIRS"))) : RL_RM_BURST_PAIRS;
    const int variant = rm_persist_env >= 0 ? rm_persist_env : m->persist;
    int rm_ctas = variant == 4 ? 4 : RL_RM_PERSIST_MINB;
    if (KIND == RL_RM && variant == 1)
      rm_ctas = mode == MODE_GRID    ? resident_ctas(rm_persist_kernel<MODE_GRID, 1, false, true>, 256, RL_RM_PERSIST_MINB)
                : mode == MODE_WORLD ? resident_ctas(rm_persist_kernel<MODE_WORLD, 1, false, true>, 256, RL_RM_PERSIST_MINB)
                                     : resident_ctas(rm_persist_kernel<MODE_ANGLES, 1, false, true>, 256, RL_RM_PERSIST_MINB);
    const long long resident_warps = (long long)sm_count() * 8 * rm_ctas;
    if (KIND == RL_RM && m->max_range > 0.0f && total >= (long long)sm_count() * 48 * 64 && variant) {
      const int chunk = claim_rays(total, resident_warps, mode == MODE_ANGLES ? M : 0);
      const long long warps = min(resident_warps, (total + chunk - 1) / chunk);
      const int grid = (int)((warps + 7) / 8);
      unsigned long long* work = nullptr;
      const int rcw = fresh_work_cursor(m, &work);
      if (rcw) return rcw;
#define RL_PERSIST_ARGS <<<grid, 256, 0, m->stream>>>(mv, m->xf, m->max_range, ins, angles, outs, total, M, chunk, rm_burst_pairs, work)
#define RL_LAUNCH_PERSIST(MD)                                                      \
  do {                                                                             \
    if (variant == 2) rm_persist_kernel<MD, 1, true, true> RL_PERSIST_ARGS;        \
    else if (variant == 3) rm_persist_kernel<MD, 1, false, false> RL_PERSIST_ARGS; \
    else if (variant == 4) rm_persist_kernel<MD, 2, true, false> RL_PERSIST_ARGS;  \
    else rm_persist_kernel<MD, 1, false, true> RL_PERSIST_ARGS;                    \
  } while (0)
      if (mode == MODE_GRID) RL_LAUNCH_PERSIST(MODE_GRID);
      else if (mode == MODE_WORLD) RL_LAUNCH_PERSIST(MODE_WORLD);
      else RL_LAUNCH_PERSIST(MODE_ANGLES);
#undef RL_LAUNCH_PERSIST
#undef RL_PERSIST_ARGS
      count_launch();
      RL_CHECK_LAUNCH();
      return RL_OK;
    }
    const long long blocks = (total + threads - 1) / threads;
    const int grid = (int)max(1LL, min(blocks, (long long)sm_count() * 8 * 64));
    if (KIND == RL_CDDT && (cv.meta || total >= 65536)) {  // big batches, and every batch on a table larger than L2
#define RL_LAUNCH_CDDT(MD)                                                                                              \
  do {                                                                                                                  \
    if (cv.meta)                                                                                                        \
      cddt_batch_kernel<MD, true><<<grid, threads, 0, m->stream>>>(mv, cv, m->xf, m->max_range, ins, angles, outs, total, M);  \
    else                                                                                                                \
      cddt_batch_kernel<MD, false><<<grid, threads, 0, m->stream>>>(mv, cv, m->xf, m->max_range, ins, angles, outs, total, M); \
  } while (0)
      if (mode == MODE_GRID) RL_LAUNCH_CDDT(MODE_GRID);
      else if (mode == MODE_WORLD) RL_LAUNCH_CDDT(MODE_WORLD);
      else RL_LAUNCH_CDDT(MODE_ANGLES);
#undef RL_LAUNCH_CDDT
      count_launch();
      RL_CHECK_LAUNCH();
      return RL_OK;
    }
    if (mode == MODE_GRID)
      cast_kernel<KIND, MODE_GRID><<<grid, threads, 0, m->stream>>>(mv, cv, m->xf, m->max_range, ins, angles, outs, total, M);
    else if (mode == MODE_WORLD)
      cast_kernel<KIND, MODE_WORLD><<<grid, threads, 0, m->stream>>>(mv, cv, m->xf, m->max_range, ins, angles, outs, total, M);
    else
      cast_kernel<KIND, MODE_ANGLES><<<grid, threads, 0, m->stream>>>(mv, cv, m->xf, m->max_range, ins, angles, outs, total, M);
  }
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

int launch_cast(rl_method* m, int mode, const float* ins, const float* angles, const float* obs, float* outs,
                double* weights, int n, int M, const PeerOut* peers) {
  if (n <= 0 || (mode >= MODE_ANGLES && M <= 0)) return RL_OK;
  switch (m->kind) {
    case RL_BL: return launch_cast_kind<RL_BL>(m, mode, ins, angles, obs, outs, weights, n, M, peers);
    case RL_RM: return launch_cast_kind<RL_RM>(m, mode, ins, angles, obs, outs, weights, n, M, peers);
    case RL_GLT: return launch_cast_kind<RL_GLT>(m, mode, ins, angles, obs, outs, weights, n, M, peers);
    default: return launch_cast_kind<RL_CDDT>(m, mode, ins, angles, obs, outs, weights, n, M, peers);
  }
}

template <int KIND>
static int launch_fused_beam_params_kind(rl_method* m, const float* ins, const BeamParams& beams, double* weights, int n,
                                         int M) {
  MapView mv = m->map_view();
  int threads = 256;
  if ((long long)n * M > 2LL * sm_count() * 7 * threads) mv.coop_threshold = 0;  // as in launch_cast_kind
  else threads = small_fused_threads();
  mv.block_burst_pairs = block_burst_pairs();
  const int chunk = min(M, 2048);
  const int ppb = max(1, min(threads / max(M, 1), 32));
  const int groups = (n + ppb - 1) / ppb;
  const int grid = max(1, min(groups, sm_count() * 8 * 256 / threads));
  const size_t smem = (size_t)ppb * chunk * sizeof(double);
  fused_kernel<KIND, true><<<grid, threads, smem, m->stream>>>(mv, m->cddt_view(), m->xf, m->sensor_view(), m->max_range,
                                                               ins, nullptr, nullptr, weights, n, M, ppb, chunk,
                                                               PeerOut{}, beams, nullptr);
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

// the fused call with angles / observation passed inside the launch; n, M > 0, M <= RL_PARAM_BEAMS
int launch_fused_beam_params(rl_method* m, const float* ins, const BeamParams& beams, double* weights, int n, int M) {
  if (!m->d_table) {
    set_error("calc_range_repeat_angles_eval_sensor_model: set_sensor_model has not been called");
    return RL_E_STATE;
  }
  switch (m->kind) {
    case RL_BL: return launch_fused_beam_params_kind<RL_BL>(m, ins, beams, weights, n, M);
    case RL_RM: return launch_fused_beam_params_kind<RL_RM>(m, ins, beams, weights, n, M);
    case RL_GLT: return launch_fused_beam_params_kind<RL_GLT>(m, ins, beams, weights, n, M);
    default: return launch_fused_beam_params_kind<RL_CDDT>(m, ins, beams, weights, n, M);
  }
}

int launch_radial(rl_method* m, const float* ins, const float* beam_angles, float* outs, int n, int num_rays, int count,
                  int max_pair, int index_offset) {
  const long long total = (long long)n * count;
  if (total <= 0) return RL_OK;
  const MapView mv = m->map_view();
  const CddtView cv = m->cddt_view();
  const int grid = (int)min((total + 255) / 256, (long long)sm_count() * 32);
#define RL_LAUNCH_RADIAL(K) \
  radial_kernel<K><<<grid, 256, 0, m->stream>>>(mv, cv, m->xf, m->max_range, ins, beam_angles, outs, total, num_rays, \
                                               count, max_pair, index_offset)
  switch (m->kind) {
    case RL_BL: RL_LAUNCH_RADIAL(RL_BL); break;
    case RL_RM: RL_LAUNCH_RADIAL(RL_RM); break;
    case RL_GLT: RL_LAUNCH_RADIAL(RL_GLT); break;
    default: RL_LAUNCH_RADIAL(RL_CDDT); break;
  }
#undef RL_LAUNCH_RADIAL
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

int launch_eval_sensor(rl_method* m, const float* obs, const float* ranges, double* outs, int M, int n) {
  if (n <= 0) return RL_OK;
  if (!m->d_table) {
    set_error("eval_sensor_model: set_sensor_model has not been called");
    return RL_E_STATE;
  }
  static const bool overlap = !(getenv("RL_EVAL_OVERLAP") && atoi(getenv("RL_EVAL_OVERLAP")) == 0);
  if (overlap && n >= 64 * sm_count())  // at least two groups of 32 particles per SM: the streaming form
    return launch_eval_overlap(m, m->xf.inv_scale, obs, ranges, outs, n, M, 0, nullptr, PeerOut{}, false, false);
  const int threads = 256;
  const int chunk = max(1, min(M, 2048));
  const int ppb = max(1, min(threads / max(M, 1), 32));
  const int groups = (n + ppb - 1) / ppb;
  const int grid = max(1, min(groups, sm_count() * 8));
  const size_t smem = (size_t)ppb * chunk * sizeof(double);
  eval_sensor_kernel<<<grid, threads, smem, m->stream>>>(m->sensor_view(), m->xf.inv_scale, obs, ranges, outs, n, M,
                                                         ppb, chunk);
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

// GiantLUTCast::GiantLUTCast (RangeLib.h:1781-1823): W*H*td RM casts, written as uint16
int glt_build(rl_method* m) {
  const long long total = (long long)m->W * m->H * m->td;
  if (!m->d_glt) RL_CUDA(cudaMalloc(&m->d_glt, sizeof(uint16_t) * (size_t)(total > 0 ? total : 1)));
  if (total == 0) return RL_OK;
  MapView mv = m->map_view();
  WorldXform xf{};
  xf.rot = mv.glt_twopi_div_td;
  xf.oy = (float)m->H;
  xf.scale = mv.glt_limits_div_max;
  if (!(m->max_range > 0.0f)) {  // every cast returns max_range without entering the loop
    set_error("GiantLUTCast needs max_range > 0");
    return RL_E_INVALID;
  }
  const long long resident_warps =
      (long long)sm_count() * 8 * resident_ctas(rm_persist_kernel<MODE_GLT_BUILD, 1, false, true>, 256, RL_RM_PERSIST_MINB);
  const int chunk = claim_rays(total, resident_warps);
  const long long warps = min(resident_warps, (total + chunk - 1) / chunk);
  const int grid = (int)((warps + 7) / 8);
  unsigned long long* work = nullptr;
  const int rcw = fresh_work_cursor(m, &work);
  if (rcw) return rcw;
  rm_persist_kernel<MODE_GLT_BUILD, 1, false, true><<<grid, 256, 0, m->stream>>>(mv, xf, m->max_range, nullptr, nullptr,
                                                                        (float*)m->d_glt, total, (int)m->td, chunk,
                                                                        RL_RM_BURST_PAIRS, work);
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

int launch_peers_wait(rl_method* m) {
  peers_wait_kernel<<<1, 32, 0, m->stream>>>(m->peer_cfg);
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

int launch_sincosf(const float* x, float* s, float* c, int n, cudaStream_t st) {
  if (n <= 0) return RL_OK;
  sincosf_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, s, c, n);
  count_launch();
  RL_CHECK_LAUNCH();
  return RL_OK;
}

}  // namespace rl

#ifdef RL_TRACE
extern "C" int rl_debug_set_trace(void* device_buffer) {
  return cudaMemcpyToSymbol(rl::g_rl_trace, &device_buffer, sizeof(void*)) == cudaSuccess ? 0 : -2;
}
#endif
