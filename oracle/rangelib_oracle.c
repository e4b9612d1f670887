/* TEST INFRASTRUCTURE ONLY -- the checker, never the thing measured or shipped.
 *
 * CPU restatement, in plain C, of the reference's batched 2-D ray casting hot path
 * (kctess5/range_libc).  Every function cites the reference file:line it follows.
 * Arithmetic mirrors the reference compiled with STRICT IEEE flags
 * (-O2 -fno-fast-math -ffp-contract=off): float ops in source order, no contraction.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function below
 * bit-for-bit against oracle/_ref/libref_strict.so (the unmodified reference compiled
 * from /root/reference by oracle/Makefile) and tests/test_golden.py checks it against the
 * fixtures under tests/golden/ that were generated from that library
 * (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EPSILON 0.00001          /* RangeLib.h:60  _EPSILON  */
#define ORC_M_2PI 6.28318530718      /* RangeLib.h:61  M_2PI     */
#define ORC_BINARY_SEARCH_THRESHOLD 64 /* RangeLib.h:62 */

/* ------------------------------------------------------------------------------------------
 * sinf / cosf: restatement of the algorithm glibc >= 2.28 uses (sysdeps/ieee754/flt-32/
 * s_sinf.c, s_cosf.c, sincosf.h -- the ARM "optimized routines" single-precision sincos,
 * pinned here to glibc 2.39, the libm the reference links in this image).  The reference
 * calls libm sinf/cosf at RangeLib.h:713-714, 931-932, 1012, 1051-1053, 1095-1096, 1199-1200,
 * 1368-1369.  The oracle's ray casters below call libm directly, as the reference does;
 * orc_sinf/orc_cosf exist so that tests can pin the DEVICE trig (which restates the same
 * double-precision polynomial) against libm on the CPU.
 * ------------------------------------------------------------------------------------------ */
static const double SC_HPI_INV = 0x1.45F306DC9C883p+23; /* 2/pi * 2^24 */
static const double SC_HPI = 0x1.921FB54442D18p0;       /* pi/2 */
static const double SC_C0 = 0x1p0, SC_C1 = -0x1.ffffffd0c621cp-2, SC_C2 = 0x1.55553e1068f19p-5,
                    SC_C3 = -0x1.6c087e89a359dp-10, SC_C4 = 0x1.99343027bf8c3p-16;
static const double SC_S1 = -0x1.555545995a603p-3, SC_S2 = 0x1.1107605230bc4p-7,
                    SC_S3 = -0x1.994eb3774cf24p-13;
static const double SC_PI63 = 0x1.921FB54442D18p-62;
/* 4/pi as overlapping 32-bit windows, stepping 8 bits (192 bits total) */
static const uint32_t SC_INV_PIO4[24] = {
    0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};

static inline uint32_t sc_asuint(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline uint32_t sc_abstop12(float f) { return (sc_asuint(f) >> 20) & 0x7ff; }

/* polynomial on [-pi/4, pi/4]; odd n -> cosine series, even n -> sine series; `neg`
 * selects the table whose cosine coefficients are negated */
static inline float sc_poly(double x, double x2, int n, int neg) {
  if ((n & 1) == 0) {
    double x3 = x * x2;
    double s1 = SC_S2 + x2 * SC_S3;
    double x7 = x3 * x2;
    double s = x + x3 * SC_S1;
    return (float)(s + x7 * s1);
  } else {
    double sg = neg ? -1.0 : 1.0;
    double x4 = x2 * x2;
    double c2 = sg * SC_C3 + x2 * (sg * SC_C4);
    double c1 = sg * SC_C0 + x2 * (sg * SC_C1);
    double x6 = x4 * x2;
    double c = c1 + x4 * (sg * SC_C2);
    return (float)(c + x6 * c2);
  }
}

static inline double sc_reduce_fast(double x, int* np) {
  double r = x * SC_HPI_INV;
  int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return x - n * SC_HPI;
}

static inline double sc_reduce_large(uint32_t xi, int* np) {
  const uint32_t* arr = &SC_INV_PIO4[(xi >> 26) & 15];
  int shift = (xi >> 23) & 7;
  uint64_t n, res0, res1, res2;
  xi = (xi & 0xffffff) | 0x800000;
  xi <<= shift;
  res0 = xi * arr[0];
  res1 = (uint64_t)xi * arr[4];
  res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  double x = (double)(int64_t)res0;
  *np = (int)n;
  return x * SC_PI63;
}

static const double SC_SIGN[4] = {1.0, -1.0, -1.0, 1.0};

static float sc_sincos_one(float y, int want_cos) {
  double x = y;
  int n;
  if (sc_abstop12(y) < 0x3f4 /* abstop12(pi/4) */) {
    double x2 = x * x;
    if (sc_abstop12(y) < 0x398 /* abstop12(2^-12) */) return want_cos ? 1.0f : y;
    return sc_poly(x, x2, want_cos, 0);
  } else if (sc_abstop12(y) < 0x42f /* abstop12(120.0f) */) {
    x = sc_reduce_fast(x, &n);
    double s = SC_SIGN[n & 3];
    return sc_poly(x * s, x * x, want_cos ? (n ^ 1) : n, (n & 2) != 0);
  } else if (sc_abstop12(y) < 0x7f8 /* abstop12(inf) */) {
    uint32_t xi = sc_asuint(y);
    int sign = xi >> 31;
    x = sc_reduce_large(xi, &n);
    double s = SC_SIGN[(n + sign) & 3];
    return sc_poly(x * s, x * x, want_cos ? (n ^ 1) : n, ((n + sign) & 2) != 0);
  }
  return y - y; /* inf/nan -> nan */
}

float orc_sinf(float y) { return sc_sincos_one(y, 0); }
float orc_cosf(float y) { return sc_sincos_one(y, 1); }

/* count of floats in bit-pattern range [lo, hi) stepping `step` whose orc_sinf/orc_cosf differ
 * from libm.  out[0] = sin mismatches, out[1] = cos mismatches. */
void orc_trig_compare(uint32_t lo, uint32_t hi, uint32_t step, uint64_t* out) {
  uint64_t ms = 0, mc = 0;
  for (uint64_t u = lo; u < hi; u += step) {
    uint32_t b = (uint32_t)u;
    float f;
    memcpy(&f, &b, 4);
    if (!(f == f) || isinf(f)) continue;
    float a = orc_sinf(f), r = sinf(f);
    if (memcmp(&a, &r, 4)) ms++;
    a = orc_cosf(f);
    r = cosf(f);
    if (memcmp(&a, &r, 4)) mc++;
  }
  out[0] = ms;
  out[1] = mc;
}

/* ------------------------------------------------------------------------------------------
 * Distance transform: RangeLib.h:345-373 feeding vendor/distance_transform.h:873-910 (pass
 * order), :1054-1095 (1-D lower envelope), :1117-1121 (sqrt).
 * occ is x-major occ[x*H+y]; out is x-major out[x*H+y] like DistanceTransform::grid[x][y].
 * ------------------------------------------------------------------------------------------ */
static void edt_1d(const float* f, float* D, size_t n, size_t stride, size_t* v, double* z) {
  /* distance_transform.h:1056-1062 */
  if (n == 0) return;
  if (n == 1) {
    D[0] = f[0];
    return;
  }
  size_t k = 0;
  double s = 0.0;
  v[0] = 0;
  z[0] = -DBL_MAX;
  z[1] = DBL_MAX;
  for (size_t q = 1; q < n; ++q) { /* :1072-1083 */
    ++k;
    do {
      --k;
      float fq = f[q * stride] + (float)(q * q);            /* float sum, size_t q*q -> float */
      float fv = f[v[k] * stride] + (float)(v[k] * v[k]);
      s = ((double)fq - (double)fv) / ((double)(2 * q) - (double)(2 * v[k])); /* :1077 */
    } while (s <= z[k]);
    ++k;
    v[k] = q;
    z[k] = s;
    z[k + 1] = DBL_MAX;
  }
  k = 0;
  for (size_t q = 0; q < n; ++q) { /* :1086-1091 */
    while (z[k + 1] < (double)q) ++k;
    float dq = (float)q - (float)v[k];
    D[q * stride] = f[v[k] * stride] + dq * dq;
  }
}

void orc_edt(const uint8_t* occ, int W, int H, float* out) {
  size_t n = (size_t)W * H;
  float* f = (float*)malloc(n * sizeof(float));
  float* g = (float*)malloc(n * sizeof(float));
  size_t m = (size_t)(W > H ? W : H);
  size_t* v = (size_t*)malloc(m * sizeof(size_t));
  double* z = (double*)malloc((m + 1) * sizeof(double));
  for (size_t i = 0; i < n; ++i) f[i] = occ[i] ? 0.0f : FLT_MAX; /* RangeLib.h:353-356 */
  /* d = 0: slices of the first dimension: for each x a scanline along y (distance_transform.h:893-900) */
  for (int x = 0; x < W; ++x) edt_1d(f + (size_t)x * H, g + (size_t)x * H, (size_t)H, 1, v, z);
  /* d = 1: for each y a scanline along x */
  for (int y = 0; y < H; ++y) edt_1d(g + y, f + y, (size_t)W, (size_t)H, v, z);
  for (size_t i = 0; i < n; ++i) out[i] = (float)sqrt((double)f[i]); /* :1117-1121 */
  free(f);
  free(g);
  free(v);
  free(z);
}

/* ------------------------------------------------------------------------------------------
 * Map context shared by the ray casters
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int W, H;
  const uint8_t* occ; /* x-major, borrowed */
  float max_range;
  /* world transform, RangeLib.h:134-141; defaults RangeLibc.pyx:174-180 */
  float world_scale, world_angle, world_origin_x, world_origin_y, world_sin_angle, world_cos_angle;
  /* RM */
  float* dt;
  /* CDDT */
  unsigned td;
  int* widths;       /* td */
  float* trans;      /* td */
  int64_t* slice0;   /* td+1: first bin of slice a in the flat bin numbering */
  int64_t* offsets;  /* nbins+1 */
  float* values;
  int64_t prune_unassigned; /* times prune hit the reference's unassigned-index path */
  /* sensor model */
  int K;
  double* table;
  int kind; /* 0 BL 1 RM 2 CDDT 4 GLT */
  /* GiantLUTCast */
  uint16_t* glt; /* [(x*H + y)*td + i] */
  float max_div_limits, limits_div_max;
} orc_ctx;

static inline int occ_at(const orc_ctx* c, int x, int y) { /* OMap::isOccupied RangeLib.h:204-210 */
  if (x < 0 || x >= c->W || y < 0 || y >= c->H) return 0;
  return c->occ[(size_t)x * c->H + y] != 0;
}

/* RayMarching::calc_range, RangeLib.h:927-962 (distThreshold 0.0, step_coeff 0.999f :967-968) */
static float rm_calc_range(const orc_ctx* c, float x, float y, float heading) {
  float x0 = x, y0 = y;
  float dx = cosf(heading), dy = sinf(heading);
  float t = 0.0f;
  while (t < c->max_range) {
    int px = (int)(x0 + dx * t);
    int py = (int)(y0 + dy * t);
    if (px >= c->W || px < 0 || py < 0 || py >= c->H) return c->max_range;
    float d = c->dt[(size_t)px * c->H + py];
    if (d <= 0.0f) {
      float xd = (float)px - x0;
      float yd = (float)py - y0;
      return sqrtf(xd * xd + yd * yd);
    }
    float step = d * 0.999f;
    t += (step > 1.0f ? step : 1.0f);
  }
  return c->max_range;
}

/* number of distance-map reads RayMarching::calc_range performs for this ray (analysis aid for
 * DESIGN.md / bench: the sphere-tracing step count is what bounds the GPU kernels) */
static int rm_step_count(const orc_ctx* c, float x, float y, float heading) {
  float x0 = x, y0 = y;
  float dx = cosf(heading), dy = sinf(heading);
  float t = 0.0f;
  int steps = 0;
  while (t < c->max_range) {
    int px = (int)(x0 + dx * t);
    int py = (int)(y0 + dy * t);
    if (px >= c->W || px < 0 || py < 0 || py >= c->H) return steps;
    float d = c->dt[(size_t)px * c->H + py];
    ++steps;
    if (d <= 0.0f) return steps;
    float step = d * 0.999f;
    t += (step > 1.0f ? step : 1.0f);
  }
  return steps;
}

void orc_rm_step_counts(const orc_ctx* c, const float* ins, int* counts, int n) {
  for (int i = 0; i < n; ++i) counts[i] = rm_step_count(c, ins[3 * i], ins[3 * i + 1], ins[3 * i + 2]);
}

/* BresenhamsLine::calc_range, RangeLib.h:696-769 */
static float bl_calc_range(const orc_ctx* c, float x, float y, float heading) {
  if (occ_at(c, (int)x, (int)y)) return 0.0f;
  float x0 = y, y0 = x;
  float x1 = y + c->max_range * sinf(heading);
  float y1 = x + c->max_range * cosf(heading);
  int steep = fabsf(y1 - y0) > fabsf(x1 - x0);
  if (steep) {
    float tmp = x0; x0 = y0; y0 = tmp;
    tmp = x1; x1 = y1; y1 = tmp;
  }
  float deltax = fabsf(x1 - x0), deltay = fabsf(y1 - y0);
  float error = 0.0f, _x = x0, _y = y0;
  int xstep = (x0 < x1) ? 1 : -1;
  int ystep = (y0 < y1) ? 1 : -1;
  float width = (float)(unsigned)c->W, height = (float)(unsigned)c->H; /* float-vs-unsigned compares :755,761 */
  int target = (int)(x1 + (float)xstep);
  while ((int)_x != target) {
    _x += (float)xstep;
    error += deltay;
    if ((double)error * 2.00 >= (double)deltax) {
      _y += (float)ystep;
      error -= deltax;
    }
    /* TERMINATION GUARD (not in the reference).  `_x += xstep` is a float accumulation: when _x crosses a
     * power of two with low fraction bits set the sum rounds, (int)_x can jump over `target`, and the
     * reference then walks on for ever (e.g. x=1188.2903 y=547.99994 heading=1.2550871 on the 1200^2 map).
     * _x and _y move monotonically, so once a coordinate has left the map on the side it is moving
     * towards no cell can be hit any more: whenever the reference terminates from here it returns
     * max_range, so returning it now changes no result and only ends the walks the reference never ends. */
    {
      float lim_x = steep ? width : height, lim_y = steep ? height : width;
      if ((xstep > 0) ? (_x >= lim_x) : (_x < 0.0f)) return c->max_range;
      if ((ystep > 0) ? (_y >= lim_y) : (_y < 0.0f)) return c->max_range;
    }
    if (!steep) {
      if (0 <= _y && _y < width && 0 <= _x && _x < height && occ_at(c, (int)_y, (int)_x)) {
        float xd = _x - x0, yd = _y - y0;
        return sqrtf(xd * xd + yd * yd);
      }
    } else {
      if (0 <= _x && _x < width && 0 <= _y && _y < height && occ_at(c, (int)_x, (int)_y)) {
        float xd = _x - x0, yd = _y - y0;
        return sqrtf(xd * xd + yd * yd);
      }
    }
  }
  return c->max_range;
}

/* ------------------------------------------------------------------------------------------
 * CDDT: constants RangeLib.h:974-1061, edge map :293-312 + RangeUtils.h:36-52, projection
 * :1083-1129, sort+unique :1132-1142
 * ------------------------------------------------------------------------------------------ */
static float cddt_M_2PI_div_td(unsigned td) { return (float)(ORC_M_2PI / (double)((float)td)); } /* :977 */
static float cddt_td_div_M_2PI(unsigned td) { return (float)((double)td / ORC_M_2PI); }          /* :976 */

void orc_edge_map(const uint8_t* occ, int W, int H, uint8_t* edge) {
  static const int ox[8] = {1, -1, 0, 0, 1, -1, 1, -1};
  static const int oy[8] = {0, 0, 1, -1, 1, 1, -1, -1};
  memset(edge, 0, (size_t)W * H);
  for (int x = 0; x < W; ++x)
    for (int y = 0; y < H; ++y) {
      if (!occ[(size_t)x * H + y]) continue;
      for (int i = 0; i < 8; ++i) {
        int cx = x + ox[i], cy = y + oy[i];
        if (0 <= cx && 0 <= cy && cx < W && cy < H && !occ[(size_t)cx * H + cy]) {
          edge[(size_t)x * H + y] = 1;
          break;
        }
      }
    }
}

static int cmp_float(const void* a, const void* b) {
  float fa = *(const float*)a, fb = *(const float*)b;
  return (fa > fb) - (fa < fb);
}

/* projection of one pixel centre for slice a: returns lut_space_x, sets lower/upper bins (:1099-1105) */
static inline float cddt_project(float pcx, float pcy, float cosangle, float sinangle, float trans, int* lower,
                                 int* upper) {
  float half = (float)((double)(fabsf(sinangle) + fabsf(cosangle)) / 2.0);
  float lx = pcx * cosangle - pcy * sinangle;
  float ly = (pcx * sinangle + pcy * cosangle) + trans;
  *upper = (int)((double)(ly + half) - ORC_EPSILON);
  *lower = (int)((double)(ly - half) + ORC_EPSILON);
  return lx;
}

static void cddt_build(orc_ctx* c) {
  unsigned td = c->td;
  int W = c->W, H = c->H;
  float step = cddt_M_2PI_div_td(td);
  c->widths = (int*)calloc(td, sizeof(int));
  c->trans = (float*)calloc(td, sizeof(float));
  c->slice0 = (int64_t*)calloc(td + 1, sizeof(int64_t));
  float* cosv = (float*)malloc(td * sizeof(float));
  float* sinv = (float*)malloc(td * sizeof(float));
  for (unsigned i = 0; i < td; ++i) { /* :991-1061 */
    float angle = (float)(int)i * step;
    float ca = cosf(angle), sa = sinf(angle);
    cosv[i] = ca;
    sinv[i] = sa;
    float rotated_height = fabsf((float)(unsigned)W * sa) + fabsf((float)(unsigned)H * ca);
    c->widths[i] = (int)(unsigned)ceil((double)rotated_height - ORC_EPSILON);
    float ltc = (float)(unsigned)H * ca;
    float rtc = (float)(unsigned)W * sa + (float)(unsigned)H * ca;
    float rbc = (float)(unsigned)W * sa;
    /* std::min(ltc, std::min(rtc, rbc)) (:1057); std::min(a, b) is b < a ? b : a */
    float inner = (rbc < rtc) ? rbc : rtc;
    float mn = (inner < ltc) ? inner : ltc;
    double tr = -1.0 * (double)mn - ORC_EPSILON;
    c->trans[i] = (float)(0.0 < tr ? tr : 0.0); /* std::max(0.0, tr) */
    c->slice0[i + 1] = c->slice0[i] + c->widths[i];
  }
  int64_t nbins = c->slice0[td];
  uint8_t* edge = (uint8_t*)malloc((size_t)W * H);
  orc_edge_map(c->occ, W, H, edge);
  int64_t* count = (int64_t*)calloc(nbins + 1, sizeof(int64_t));
  /* the reference iterates a < td / 2.0 (:1089) */
  int na = 0;
  while ((double)na < (double)td / 2.0) ++na;
  for (int pass = 0; pass < 2; ++pass) {
    int64_t* cursor = NULL;
    if (pass == 1) {
      /* exclusive scan */
      int64_t acc = 0;
      for (int64_t b = 0; b <= nbins; ++b) {
        int64_t t = count[b];
        count[b] = acc;
        acc += t;
      }
      c->values = (float*)malloc((size_t)(count[nbins] > 0 ? count[nbins] : 1) * sizeof(float));
      cursor = (int64_t*)malloc((nbins + 1) * sizeof(int64_t));
      memcpy(cursor, count, (nbins + 1) * sizeof(int64_t));
    }
    for (int x = 0; x < W; ++x)
      for (int y = 0; y < H; ++y) {
        if (!edge[(size_t)x * H + y]) continue;
        float pcx = (float)((double)x + 0.5), pcy = (float)((double)y + 0.5); /* :1088 */
        for (int a = 0; a < na; ++a) {
          int lower, upper;
          float lx = cddt_project(pcx, pcy, cosv[a], sinv[a], c->trans[a], &lower, &upper);
          for (int i = lower; i <= upper; ++i) {
            if (i < 0 || i >= c->widths[a]) continue; /* reference would write out of bounds */
            int64_t b = c->slice0[a] + i;
            if (pass == 0) count[b]++;
            else c->values[cursor[b]++] = lx;
          }
        }
      }
    if (cursor) free(cursor);
  }
  /* sort + unique per bin, compacting in place (:1132-1142) */
  c->offsets = (int64_t*)malloc((nbins + 1) * sizeof(int64_t));
  int64_t w = 0;
  for (int64_t b = 0; b < nbins; ++b) {
    int64_t lo = count[b], hi = count[b + 1];
    qsort(c->values + lo, (size_t)(hi - lo), sizeof(float), cmp_float);
    c->offsets[b] = w;
    for (int64_t i = lo; i < hi; ++i)
      if (i == lo || c->values[i] != c->values[i - 1]) c->values[w++] = c->values[i];
  }
  c->offsets[nbins] = w;
  free(count);
  free(edge);
  free(cosv);
  free(sinv);
}

/* CDDTCast::prune, RangeLib.h:1176-1283.  The linear-search branch of the reference leaves
 * `index` unassigned when no element >= lut_space_x exists (:1237-1248).  That is formally
 * undefined, but every build of the reference tried here (gcc 13 -O0, -O2, -O3 -ffast-math)
 * behaves the same way: `index` lives in one stack slot / register for the whole call, so the
 * unassigned read sees the value the most recent earlier pixel (in the loop order angle, x, y)
 * stored there.  The restatement pins exactly that ("stale index") and counts the occurrences
 * in prune_unassigned.  Before the first assignment the slot holds garbage; we use a value
 * that marks nothing. */
static void cddt_prune(orc_ctx* c, float max_range) {
  unsigned td = c->td;
  int W = c->W, H = c->H;
  float step = cddt_M_2PI_div_td(td);
  int64_t nbins = c->slice0[td];
  int64_t total = c->offsets[nbins];
  uint8_t* used = (uint8_t*)calloc((size_t)(total > 0 ? total : 1), 1);
  int na = 0;
  while ((double)na < (double)td / 2.0) ++na;
  int index = -2; /* one slot for the whole call, see above */
  for (int a = 0; a < na; ++a) {
    float angle = (float)a * step;
    float ca = cosf(angle), sa = sinf(angle);
    float tr = c->trans[a];
    for (int x = 0; x < W; ++x) {
      float _x = (float)(0.5 + (double)x);
      for (int y = 0; y < H; ++y) {
        float _y = (float)(0.5 + (double)y);
        float lx = _x * ca - _y * sa;
        float ly = (_x * sa + _y * ca) + tr;
        unsigned li = (unsigned)(int)ly;
        if (li >= (unsigned)c->widths[a]) continue; /* reference indexes unchecked */
        int64_t b = c->slice0[a] + li;
        const float* B = c->values + c->offsets[b];
        uint8_t* U = used + c->offsets[b];
        int size = (int)(c->offsets[b + 1] - c->offsets[b]);
        int high = size - 1;
        if (high == -1) continue;
        if (c->occ[(size_t)x * H + y]) continue;
        if (B[high] < lx && lx - B[high] < max_range) {
          U[high] = 1;
          continue;
        }
        if (high > ORC_BINARY_SEARCH_THRESHOLD) { /* std::lower_bound: first >= lx */
          int lo = 0, hi = size;
          while (lo < hi) {
            int mid = lo + (hi - lo) / 2;
            if (B[mid] < lx) lo = mid + 1; else hi = mid;
          }
          index = lo;
        } else {
          int found = 0;
          for (int i = 0; i < size; ++i)
            if (B[i] >= lx) { index = i; found = 1; break; }
          if (!found) c->prune_unassigned++; /* index keeps its stale value */
        }
        /* std::set<int> inserts of {index, index-1}; only entries in [0,size) survive the rebuild */
        if (index >= 0 && index < size) U[index] = 1;
        if (index - 1 >= 0 && index - 1 < size) U[index - 1] = 1;
      }
    }
  }
  /* rebuild bins of slices [0, na) with the marked entries (:1268-1278); slices >= na are empty anyway */
  int64_t w = 0;
  int64_t first_untouched = c->slice0[na];
  for (int64_t b = 0; b < nbins; ++b) {
    int64_t lo = c->offsets[b], hi = c->offsets[b + 1];
    c->offsets[b] = w;
    for (int64_t i = lo; i < hi; ++i)
      if (b >= first_untouched || used[i]) c->values[w++] = c->values[i];
  }
  c->offsets[nbins] = w;
  free(used);
}

/* CDDTCast::discretize_theta, RangeLib.h:1287-1340 with _USE_ALTERNATE_MOD 1,
 * _USE_CACHED_CONSTANTS 1, _USE_FAST_ROUND 0 */
static void cddt_discretize_theta(unsigned td, float theta, int* binned, float* discrete, int* flipped) {
  if ((double)theta < 0.0) {
    while ((double)theta < 0.0) theta = (float)((double)theta + ORC_M_2PI);
  } else if ((double)theta > ORC_M_2PI) {
    while ((double)theta > ORC_M_2PI) theta = (float)((double)theta - ORC_M_2PI);
  }
  int is_flipped = 0;
  if ((double)theta >= M_PI) {
    is_flipped = 1;
    theta = (float)((double)theta - M_PI);
  }
  int rounded = (int)roundf(theta * cddt_td_div_M_2PI(td));
  if ((unsigned)rounded == (td >> 1)) {
    rounded = 0;
    is_flipped = !is_flipped;
  }
  int b = (int)((unsigned)rounded % td);
  *binned = b;
  *discrete = (float)b * cddt_M_2PI_div_td(td);
  *flipped = is_flipped;
}

/* CDDTCast::calc_range, RangeLib.h:1342-1516 */
static float cddt_calc_range(const orc_ctx* c, float x, float y, float heading) {
  int a, flipped;
  float dth;
  cddt_discretize_theta(c->td, (float)(-1.0 * (double)heading), &a, &dth, &flipped);
  float ca = cosf(dth), sa = sinf(dth);
  float lx = x * ca - y * sa;
  float ly = (x * sa + y * ca) + c->trans[a];
  unsigned li = (unsigned)(int)ly;
  if (li >= (unsigned)c->widths[a]) return c->max_range;
  int64_t b = c->slice0[a] + li;
  const float* B = c->values + c->offsets[b];
  int size = (int)(c->offsets[b + 1] - c->offsets[b]);
  int high = size - 1;
  if (high == -1) return c->max_range;
  if (flipped) {
    if (B[0] > lx) return c->max_range;
    if (B[high] < lx) return lx - B[high];
    if (occ_at(c, (int)x, (int)y)) return 0.0f; /* map.grid[x][y], float -> index (:1413) */
    if (high > ORC_BINARY_SEARCH_THRESHOLD) {   /* std::upper_bound: first > lx */
      int lo = 0, hi = size;
      while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (!(lx < B[mid])) lo = mid + 1; else hi = mid;
      }
      return lx - B[lo - 1];
    }
    for (int i = high; i >= 0; --i)
      if (B[i] <= lx) return lx - B[i];
  } else {
    if (B[high] < lx) return c->max_range;
    if (B[0] > lx) return B[0] - lx;
    if (occ_at(c, (int)x, (int)y)) return 0.0f; /* :1475 */
    if (high > ORC_BINARY_SEARCH_THRESHOLD) {
      int lo = 0, hi = size;
      while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (!(lx < B[mid])) lo = mid + 1; else hi = mid;
      }
      return B[lo] - lx; /* lo == size cannot happen: B[high] >= lx ... unless B[high] == lx */
    }
    for (int i = 0; i < size; ++i)
      if (B[i] >= lx) return B[i] - lx;
  }
  return -1.0f; /* the reference's assert(0) fall-through (:1514) */
}

/* CDDTCast::calc_range_pair, RangeLib.h:1521-1649: (range along heading, range along heading + pi).
 * The reference does not bounds-check lut_index (:1538-1539; undefined outside the table) -- here
 * (max_range, max_range).  Its non-flipped occupied-cell test is a no-op (:1612) and is omitted. */
static void cddt_calc_range_pair(const orc_ctx* c, float x, float y, float heading, float* r, float* r_inv) {
  int a, flipped;
  float dth;
  cddt_discretize_theta(c->td, (float)(-1.0 * (double)heading), &a, &dth, &flipped);
  float ca = cosf(dth), sa = sinf(dth);
  float lx = x * ca - y * sa;
  float ly = (x * sa + y * ca) + c->trans[a];
  unsigned li = (unsigned)(int)ly;
  float mr = c->max_range;
  *r = mr;
  *r_inv = mr;
  if (li >= (unsigned)c->widths[a]) return;
  int64_t b = c->slice0[a] + li;
  const float* B = c->values + c->offsets[b];
  int size = (int)(c->offsets[b + 1] - c->offsets[b]);
  int high = size - 1;
  if (high == -1) return;
  int index = 0;
  if (flipped) {
    if (B[0] > lx) { *r_inv = fminf(mr, B[0] - lx); return; }  /* std::min(max_range, .) :1553 */
    if (B[high] < lx) { *r = lx - B[high]; return; }
    if (occ_at(c, (int)x, (int)y)) { *r = 0.0f; *r_inv = 0.0f; return; }
    if (high > ORC_BINARY_SEARCH_THRESHOLD) { /* upper_bound - 1 :1566 */
      int lo = 0, hi = size;
      while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (!(lx < B[mid])) lo = mid + 1; else hi = mid;
      }
      index = lo - 1;
    } else {
      for (int i = high; i >= 0; --i)
        if (B[i] <= lx) { index = i; break; }
    }
    *r = lx - B[index];
    if (index + 1 != size) *r_inv = B[index + 1] - lx;
  } else {
    if (B[high] < lx) { *r_inv = fminf(mr, lx - B[high]); return; }
    if (high > ORC_BINARY_SEARCH_THRESHOLD) { /* lower_bound :1619 */
      int lo = 0, hi = size;
      while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (B[mid] < lx) lo = mid + 1; else hi = mid;
      }
      index = lo;
    } else {
      for (int i = 0; i < size; ++i)
        if (B[i] >= lx) { index = i; break; }
    }
    *r = B[index] - lx;
    if (index - 1 != -1) *r_inv = lx - B[index - 1];
  }
}

/* ------------------------------------------------------------------------------------------
 * GiantLUTCast, RangeLib.h:1772-1904 (_GIANT_LUT_SHORT_DATATYPE 1, _USE_CACHED_CONSTANTS 1,
 * _USE_ALTERNATE_MOD 1, _USE_FAST_ROUND 0): a uint16 range for every (x, y, theta bin), seeded by
 * RayMarching from the pixel CORNER (x, y) (:1801)
 * ------------------------------------------------------------------------------------------ */
static void glt_build(orc_ctx* c) {
  unsigned td = c->td;
  float step = cddt_M_2PI_div_td(td);                       /* :1784, same expression as CDDT's */
  c->max_div_limits = c->max_range / (float)65535;          /* :1785 */
  c->limits_div_max = (float)65535 / c->max_range;          /* :1786 */
  c->glt = (uint16_t*)malloc((size_t)c->W * c->H * td * sizeof(uint16_t));
  size_t k = 0;
  for (int x = 0; x < c->W; ++x)
    for (int y = 0; y < c->H; ++y)
      for (unsigned i = 0; i < td; ++i) {
        float angle = (float)(int)i * step;                  /* :1797 */
        float r = rm_calc_range(c, (float)x, (float)y, angle);
        r = (r < c->max_range) ? r : c->max_range;           /* std::min(max_range, r) :1804 */
        c->glt[k++] = (uint16_t)(int)(r * c->limits_div_max); /* :1806 */
      }
}

static int glt_discretize_theta(unsigned td, float theta) { /* :1833-1867 */
  if ((double)theta < 0.0) {
    while ((double)theta < 0.0) theta = (float)((double)theta + ORC_M_2PI);
  } else if ((double)theta > ORC_M_2PI) {
    while ((double)theta > ORC_M_2PI) theta = (float)((double)theta - ORC_M_2PI);
  }
  int rounded = (int)roundf(theta * cddt_td_div_M_2PI(td));
  return rounded % (int)td;
}

static float glt_calc_range(const orc_ctx* c, float x, float y, float heading) { /* :1869-1880 */
  if (x < 0 || x >= (float)(unsigned)c->W || y < 0 || y >= (float)(unsigned)c->H) return c->max_range;
  int i = glt_discretize_theta(c->td, heading);
  return (float)(int)c->glt[((size_t)(int)x * c->H + (int)y) * c->td + i] * c->max_div_limits;
}

static float calc_range_any(const orc_ctx* c, float x, float y, float th) {
  switch (c->kind) {
    case 0: return bl_calc_range(c, x, y, th);
    case 1: return rm_calc_range(c, x, y, th);
    case 4: return glt_calc_range(c, x, y, th);
    default: return cddt_calc_range(c, x, y, th);
  }
}

/* ------------------------------------------------------------------------------------------
 * public C surface (loaded with ctypes by oracle/port.py)
 * ------------------------------------------------------------------------------------------ */
orc_ctx* orc_create(int kind, const uint8_t* occ, int W, int H, float max_range, unsigned td) {
  orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
  c->kind = kind == 3 ? 2 : kind;
  c->W = W;
  c->H = H;
  uint8_t* o = (uint8_t*)malloc((size_t)W * H);
  memcpy(o, occ, (size_t)W * H);
  c->occ = o;
  c->max_range = max_range;
  c->world_scale = 1.0f;
  c->world_cos_angle = 1.0f;
  c->td = td;
  if (c->kind == 1 || c->kind == 4) {
    c->dt = (float*)malloc((size_t)W * H * sizeof(float));
    orc_edt(c->occ, W, H, c->dt);
    if (c->kind == 4) glt_build(c);
  } else if (c->kind == 2) {
    cddt_build(c);
    if (kind == 3) cddt_prune(c, max_range);
  }
  return c;
}

void orc_destroy(orc_ctx* c) {
  free((void*)c->occ);
  free(c->dt);
  free(c->widths);
  free(c->trans);
  free(c->slice0);
  free(c->offsets);
  free(c->values);
  free(c->table);
  free(c->glt);
  free(c);
}

void orc_set_world(orc_ctx* c, float scale, float angle, float ox, float oy, float sin_a, float cos_a) {
  c->world_scale = scale;
  c->world_angle = angle;
  c->world_origin_x = ox;
  c->world_origin_y = oy;
  c->world_sin_angle = sin_a;
  c->world_cos_angle = cos_a;
}

void orc_prune(orc_ctx* c, float max_range) {
  if (c->kind == 2) cddt_prune(c, max_range);
}
int64_t orc_prune_unassigned(const orc_ctx* c) { return c->prune_unassigned; }

float orc_calc_range(const orc_ctx* c, float x, float y, float th) { return calc_range_any(c, x, y, th); }

const float* orc_dt(const orc_ctx* c) { return c->dt; }
const uint16_t* orc_glt(const orc_ctx* c) { return c->glt; }
int64_t orc_cddt_nbins(const orc_ctx* c) { return c->slice0 ? c->slice0[c->td] : 0; }
int64_t orc_cddt_nvalues(const orc_ctx* c) { return c->offsets ? c->offsets[c->slice0[c->td]] : 0; }
const int* orc_cddt_widths(const orc_ctx* c) { return c->widths; }
const float* orc_cddt_trans(const orc_ctx* c) { return c->trans; }
const int64_t* orc_cddt_offsets(const orc_ctx* c) { return c->offsets; }
const float* orc_cddt_values(const orc_ctx* c) { return c->values; }

/* ---- batched loops, sliced over pthreads like oracle/ref_shim.cpp slices the reference ---- */
typedef struct {
  const orc_ctx* c;
  int op;
  const float *ins, *angles, *obs, *ranges;
  float* outs;
  double* wts;
  int lo, hi, m;
} job_t;

/* world -> grid, RangeLib.h:442-475 / 485-518 / 561-607 */
typedef struct { float inv, scale, ox, oy, s, co, rot; } xform_t;
static xform_t make_xform(const orc_ctx* c) {
  xform_t t;
  t.inv = (float)(1.0 / (double)c->world_scale);
  t.scale = c->world_scale;
  t.ox = c->world_origin_x;
  t.oy = c->world_origin_y;
  t.s = c->world_sin_angle;
  t.co = c->world_cos_angle;
  t.rot = (float)(-1.0 * (double)c->world_angle - 3.0 * M_PI / 2.0);
  return t;
}

static void* job_run(void* p) {
  job_t* j = (job_t*)p;
  const orc_ctx* c = j->c;
  xform_t t = make_xform(c);
  float Kf = (float)((double)(float)c->K - 1.0); /* (float)sensor_model.size()-1.0 narrowed by std::min<float> */
  for (int i = j->lo; i < j->hi; ++i) {
    if (j->op == 0) { /* grid coordinates */
      j->outs[i] = calc_range_any(c, j->ins[3 * i], j->ins[3 * i + 1], j->ins[3 * i + 2]);
      continue;
    }
    if (j->op == 3) { /* eval_sensor_model RangeLib.h:533-555 */
      double w = 1.0;
      for (int k = 0; k < j->m; ++k) {
        float r = j->obs[k] * t.inv;
        r = fminf(fmaxf(r, 0.0f), Kf);
        float d = j->ranges[(size_t)i * j->m + k] * t.inv;
        d = fminf(fmaxf(d, 0.0f), Kf);
        w *= c->table[(size_t)(int)r * c->K + (int)d];
      }
      j->wts[i] = w;
      continue;
    }
    float xw = j->ins[3 * i], yw = j->ins[3 * i + 1], thw = j->ins[3 * i + 2];
    float x = (xw - t.ox) * t.inv;
    float y = (yw - t.oy) * t.inv;
    float tmp = x;
    x = t.co * x - t.s * y;
    y = t.s * tmp + t.co * y;
    float th = -thw + t.rot;
    if (j->op == 1) { /* numpy_calc_range :439-480 (note the x/y swap :475) */
      j->outs[i] = calc_range_any(c, y, x, th) * t.scale;
    } else if (j->op == 2) { /* numpy_calc_range_angles :482-520 */
      for (int a = 0; a < j->m; ++a)
        j->outs[(size_t)i * j->m + a] = calc_range_any(c, y, x, th - j->angles[a]) * t.scale;
    } else { /* fused :558-612 */
      double w = 1.0;
      for (int a = 0; a < j->m; ++a) {
        float d = calc_range_any(c, y, x, th - j->angles[a]);
        d = fminf(fmaxf(d, 0.0f), Kf);
        float r = j->obs[a] * t.inv;
        r = fminf(fmaxf(r, 0.0f), Kf);
        w *= c->table[(size_t)(int)r * c->K + (int)d];
      }
      j->wts[i] = w;
    }
  }
  return NULL;
}

static void run_jobs(job_t proto, int n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (n < 2 * nthreads) nthreads = 1;
  pthread_t th[256];
  job_t jobs[256];
  int per = (n + nthreads - 1) / nthreads;
  int used = 0;
  for (int t = 0; t < nthreads; ++t) {
    int lo = t * per, hi = lo + per > n ? n : lo + per;
    if (lo >= hi) break;
    jobs[t] = proto;
    jobs[t].lo = lo;
    jobs[t].hi = hi;
    if (nthreads == 1) job_run(&jobs[t]);
    else pthread_create(&th[t], NULL, job_run, &jobs[t]);
    used++;
  }
  if (nthreads > 1)
    for (int t = 0; t < used; ++t) pthread_join(th[t], NULL);
}

void orc_calc_range_many(const orc_ctx* c, const float* ins, float* outs, int n, int nthreads) {
  job_t j = {c, 0, ins, NULL, NULL, NULL, outs, NULL, 0, 0, 0};
  run_jobs(j, n, nthreads);
}
void orc_numpy_calc_range(const orc_ctx* c, const float* ins, float* outs, int n, int nthreads) {
  job_t j = {c, 1, ins, NULL, NULL, NULL, outs, NULL, 0, 0, 0};
  run_jobs(j, n, nthreads);
}
void orc_numpy_calc_range_angles(const orc_ctx* c, const float* ins, const float* angles, float* outs, int n,
                                 int m, int nthreads) {
  job_t j = {c, 2, ins, angles, NULL, NULL, outs, NULL, 0, 0, m};
  run_jobs(j, n, nthreads);
}
/* replace semantics (the reference appends, RangeLib.h:523-532; calling it once is identical) */
void orc_set_sensor_model(orc_ctx* c, const double* table, int k) {
  free(c->table);
  c->table = (double*)malloc((size_t)k * k * sizeof(double));
  memcpy(c->table, table, (size_t)k * k * sizeof(double));
  c->K = k;
}
void orc_eval_sensor_model(const orc_ctx* c, const float* obs, const float* ranges, double* outs, int m, int n,
                           int nthreads) {
  job_t j = {c, 3, NULL, NULL, obs, ranges, NULL, outs, 0, 0, m};
  run_jobs(j, n, nthreads);
}
void orc_calc_range_repeat_angles_eval_sensor_model(const orc_ctx* c, const float* ins, const float* angles,
                                                    const float* obs, double* weights, int n, int m,
                                                    int nthreads) {
  job_t j = {c, 4, ins, angles, obs, NULL, NULL, weights, 0, 0, m};
  run_jobs(j, n, nthreads);
}

/* RangeMethod::calc_range_many_radial_optimized, RangeLib.h:616-676.  Sequential like the reference, so that a
 * pair's second beam overwritten by a later iteration ends with the later value; writes that would leave the
 * particle's row (the reference lets them run into the next row or past the buffer) are dropped. */
void orc_calc_range_many_radial_optimized(const orc_ctx* c, const float* ins, float* outs, int n, int num_rays,
                                          float min_angle, float max_angle) {
  xform_t t = make_xform(c);
  float step = (max_angle - min_angle) / (num_rays - 1);
  int max_pairwise_index = (float)num_rays / 3.0;
  float index_offset_float = (num_rays - 1.0) * M_PI / (max_angle - min_angle);
  int index_offset = roundf(index_offset_float);
  int is_cddt = c->kind == 2; /* CDDT, pruned or not */
  for (int i = 0; i < n; ++i) {
    float xw = ins[3 * i], yw = ins[3 * i + 1], thw = ins[3 * i + 2];
    float theta = -thw + t.rot;
    float x = (xw - t.ox) * t.inv;
    float y = (yw - t.oy) * t.inv;
    float tmp = x;
    x = t.co * x - t.s * y;
    y = t.s * tmp + t.co * y;
    float angle = min_angle;
    float* row = outs + (size_t)i * num_rays;
    int a;
    for (a = 0; a <= max_pairwise_index; ++a) {
      float r = -1.0f, r_inv = -1.0f; /* RangeMethod::calc_range_pair default :419 */
      if (is_cddt) cddt_calc_range_pair(c, y, x, theta - angle, &r, &r_inv);
      if (a < num_rays) row[a] = r * t.scale;
      if (a + index_offset >= 0 && a + index_offset < num_rays) row[a + index_offset] = r_inv * t.scale;
      angle += step;
    }
    for (a = max_pairwise_index + 1; a < index_offset; ++a) {
      if (a < num_rays) row[a] = calc_range_any(c, y, x, theta - angle) * t.scale;
      angle += step;
    }
  }
}
