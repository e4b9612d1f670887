// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// extern "C" shim around the UNMODIFIED reference header.  It is compiled from the
// reference sources where they lie (-I/root/reference, see oracle/Makefile) into
// oracle/_ref/libref_{strict,shipped}.so.  Nothing of the reference is copied into this
// repository: this file only #includes it and forwards calls.
//
// Used by: tests/ (to pin oracle/rangelib_oracle.c and to generate tests/golden/*),
// bench.py's cpu_baseline leg and `bench.py --impl reference` (kind = "reference").
//
// With -DUSE_CUDA=1 (oracle/_ref/libref_cuda.so, linked with the reference's own includes/kernels.cu
// compiled for sm_100a) kind 5 is ranges::RayMarchingGPU, the reference's CUDA path: the baseline "recompiled
// reference kernels" figure in bench.py.  Its batched methods are non-virtual and shadow the base class's
// (RangeLib.h:819-911), so they are called through the concrete pointer.
//
// Threading: the reference is single-threaded.  The *_mt entry points split the batch
// into contiguous slices and run the reference's own batched loop on each slice from a
// std::thread; calc_range is read-only with the reference's default flags
// (_USE_LRU_CACHE 0, _MAKE_TRACE_MAP 0, _TRACK_COLLISION_INDEXES 0 -- RangeLib.h:53-71).

#include "includes/RangeLib.h"

#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// RayMarching::distImage is protected (RangeLib.h:965-968); expose it read-only.
struct RMOpen : public ranges::RayMarching {
  RMOpen(ranges::OMap m, float mr) : ranges::RayMarching(m, mr) {}
  float dt_at(int x, int y) { return distImage.get(x, y); }
};

// GiantLUTCast::giant_lut is protected (RangeLib.h:1895-1903); expose it read-only.
struct GLTOpen : public ranges::GiantLUTCast {
  GLTOpen(ranges::OMap m, float mr, int td) : ranges::GiantLUTCast(m, mr, td) {}
  uint16_t at(int x, int y, int i) { return giant_lut[x][y][i]; }
};

struct RefMethod {
  int kind;  // 0 BL, 1 RM, 2 CDDT (prune() turns it into PCDDT), 4 GLT
  ranges::RangeMethod* base = nullptr;
  ranges::BresenhamsLine* bl = nullptr;
  RMOpen* rm = nullptr;
  ranges::CDDTCast* cddt = nullptr;
  GLTOpen* glt = nullptr;
#if USE_CUDA == 1
  ranges::RayMarchingGPU* gpu = nullptr;
#endif
};

template <class F>
void run_sliced(int n, int nthreads, F f) {
  if (nthreads <= 1 || n < 2 * nthreads) {
    f(0, n);
    return;
  }
  std::vector<std::thread> th;
  int per = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    int lo = t * per, hi = std::min(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back([=] { f(lo, hi); });
  }
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

int ref_has_cuda() { return USE_CUDA == 1; }

// occ is x-major: occ[x*H + y] != 0 <=> OMap::grid[x][y] (RangeLib.h:126).
void* ref_map_create(const uint8_t* occ, int W, int H) {
  ranges::OMap* m = new ranges::OMap(W, H);
  for (int x = 0; x < W; ++x)
    for (int y = 0; y < H; ++y) m->grid[x][y] = occ[(size_t)x * H + y] != 0;
  // the defaults the Cython layer installs (RangeLibc.pyx:174-180)
  m->world_scale = 1.0f;
  m->world_angle = 0.0f;
  m->world_origin_x = 0.0f;
  m->world_origin_y = 0.0f;
  m->world_sin_angle = 0.0f;
  m->world_cos_angle = 1.0f;
  return m;
}

// PNG ingest through the reference's own OMap(filename, threshold) (RangeLib.h:159-201).
void* ref_map_load_png(const char* path, float threshold) {
  ranges::OMap* m = new ranges::OMap(std::string(path), threshold);
  if (m->error()) {
    delete m;
    return nullptr;
  }
  m->world_scale = 1.0f;
  m->world_angle = 0.0f;
  m->world_origin_x = 0.0f;
  m->world_origin_y = 0.0f;
  m->world_sin_angle = 0.0f;
  m->world_cos_angle = 1.0f;
  return m;
}

int ref_map_width(void* mp) { return (int)((ranges::OMap*)mp)->width; }
int ref_map_height(void* mp) { return (int)((ranges::OMap*)mp)->height; }

void ref_map_get(void* mp, uint8_t* out) {
  ranges::OMap* m = (ranges::OMap*)mp;
  for (unsigned x = 0; x < m->width; ++x)
    for (unsigned y = 0; y < m->height; ++y) out[(size_t)x * m->height + y] = m->grid[x][y] ? 1 : 0;
}

void ref_map_edge(void* mp, uint8_t* out) {
  ranges::OMap* m = (ranges::OMap*)mp;
  ranges::OMap e = m->make_edge_map(true);
  for (unsigned x = 0; x < m->width; ++x)
    for (unsigned y = 0; y < m->height; ++y) out[(size_t)x * m->height + y] = e.grid[x][y] ? 1 : 0;
}

void ref_map_set_world(void* mp, float scale, float angle, float ox, float oy, float sin_a, float cos_a) {
  ranges::OMap* m = (ranges::OMap*)mp;
  m->world_scale = scale;
  m->world_angle = angle;
  m->world_origin_x = ox;
  m->world_origin_y = oy;
  m->world_sin_angle = sin_a;
  m->world_cos_angle = cos_a;
}

void ref_map_destroy(void* mp) { delete (ranges::OMap*)mp; }

void* ref_method_create(int kind, void* mp, float max_range, unsigned td) {
  ranges::OMap* m = (ranges::OMap*)mp;
  RefMethod* r = new RefMethod();
  r->kind = kind;
  if (kind == 0) {
    r->bl = new ranges::BresenhamsLine(*m, max_range);
    r->base = r->bl;
  } else if (kind == 1) {
    r->rm = new RMOpen(*m, max_range);
    r->base = r->rm;
  } else if (kind == 2 || kind == 3) {
    r->cddt = new ranges::CDDTCast(*m, max_range, td);
    if (kind == 3) r->cddt->prune(max_range);
    r->base = r->cddt;
  } else if (kind == 4) {
    r->glt = new GLTOpen(*m, max_range, (int)td);
    r->base = r->glt;
#if USE_CUDA == 1
  } else if (kind == 5) {
    r->gpu = new ranges::RayMarchingGPU(*m, max_range);
    r->base = r->gpu;
#endif
  } else {
    delete r;
    return nullptr;
  }
  return r;
}

// GiantLUTCast table dump: out[(x*H + y)*td + i]
int ref_glt_dump(void* rp, uint16_t* out, int td) {
  RefMethod* r = (RefMethod*)rp;
  if (!r->glt) return -1;
  ranges::OMap* m = r->glt->getMap();
  size_t k = 0;
  for (unsigned x = 0; x < m->width; ++x)
    for (unsigned y = 0; y < m->height; ++y)
      for (int i = 0; i < td; ++i) out[k++] = r->glt->at(x, y, i);
  return 0;
}

void ref_method_destroy(void* rp) {
  RefMethod* r = (RefMethod*)rp;
  delete r->base;
  delete r;
}

void ref_prune(void* rp, float max_range) {
  RefMethod* r = (RefMethod*)rp;
  if (r->cddt) r->cddt->prune(max_range);
}

float ref_calc_range(void* rp, float x, float y, float heading) {
  return ((RefMethod*)rp)->base->calc_range(x, y, heading);
}

// grid coordinates, no conversion: what RayMarchingGPU::calc_range_many computes
// (RangeLib.h:819-831) and what main.cpp's benchmarks call per ray.
void ref_calc_range_many(void* rp, const float* ins, float* outs, int n, int nthreads) {
#if USE_CUDA == 1
  if (((RefMethod*)rp)->gpu) return ((RefMethod*)rp)->gpu->calc_range_many(const_cast<float*>(ins), outs, n);
#endif
  ranges::RangeMethod* b = ((RefMethod*)rp)->base;
  run_sliced(n, nthreads, [=](int lo, int hi) {
    for (int i = lo; i < hi; ++i) outs[i] = b->calc_range(ins[3 * i], ins[3 * i + 1], ins[3 * i + 2]);
  });
}

void ref_numpy_calc_range(void* rp, const float* ins, float* outs, int n, int nthreads) {
#if USE_CUDA == 1
  if (((RefMethod*)rp)->gpu) return ((RefMethod*)rp)->gpu->numpy_calc_range(const_cast<float*>(ins), outs, n);
#endif
  ranges::RangeMethod* b = ((RefMethod*)rp)->base;
  run_sliced(n, nthreads, [=](int lo, int hi) {
    b->numpy_calc_range(const_cast<float*>(ins) + 3 * (size_t)lo, outs + lo, hi - lo);
  });
}

void ref_numpy_calc_range_angles(void* rp, const float* ins, const float* angles, float* outs, int n, int m,
                                 int nthreads) {
#if USE_CUDA == 1
  if (((RefMethod*)rp)->gpu)
    return ((RefMethod*)rp)->gpu->numpy_calc_range_angles(const_cast<float*>(ins), const_cast<float*>(angles), outs, n, m);
#endif
  ranges::RangeMethod* b = ((RefMethod*)rp)->base;
  run_sliced(n, nthreads, [=](int lo, int hi) {
    b->numpy_calc_range_angles(const_cast<float*>(ins) + 3 * (size_t)lo, const_cast<float*>(angles),
                               outs + (size_t)lo * m, hi - lo, m);
  });
}

void ref_set_sensor_model(void* rp, const double* table, int k) {
#if USE_CUDA == 1
  if (((RefMethod*)rp)->gpu) return ((RefMethod*)rp)->gpu->set_sensor_model(const_cast<double*>(table), k);
#endif
  ((RefMethod*)rp)->base->set_sensor_model(const_cast<double*>(table), k);
}

void ref_eval_sensor_model(void* rp, const float* obs, const float* ranges_in, double* outs, int m, int n,
                           int nthreads) {
  ranges::RangeMethod* b = ((RefMethod*)rp)->base;
  run_sliced(n, nthreads, [=](int lo, int hi) {
    b->eval_sensor_model(const_cast<float*>(obs), const_cast<float*>(ranges_in) + (size_t)lo * m, outs + lo, m,
                         hi - lo);
  });
}

void ref_calc_range_repeat_angles_eval_sensor_model(void* rp, const float* ins, const float* angles,
                                                    const float* obs, double* weights, int n, int m,
                                                    int nthreads) {
  ranges::RangeMethod* b = ((RefMethod*)rp)->base;
  run_sliced(n, nthreads, [=](int lo, int hi) {
    b->calc_range_repeat_angles_eval_sensor_model(const_cast<float*>(ins) + 3 * (size_t)lo,
                                                  const_cast<float*>(angles), const_cast<float*>(obs),
                                                  weights + lo, hi - lo, m);
  });
}

// RangeMethod::calc_range_many_radial_optimized (RangeLib.h:616-676), single-threaded as shipped
void ref_calc_range_many_radial_optimized(void* rp, const float* ins, float* outs, int n, int num_rays,
                                          float min_angle, float max_angle) {
  ((RefMethod*)rp)->base->calc_range_many_radial_optimized(const_cast<float*>(ins), outs, n, num_rays, min_angle,
                                                           max_angle);
}

void ref_calc_range_pair(void* rp, float x, float y, float heading, float* r, float* r_inv) {
  std::pair<float, float> p = ((RefMethod*)rp)->base->calc_range_pair(x, y, heading);
  *r = p.first;
  *r_inv = p.second;
}

// distance transform dump, x-major out[x*H+y]; RM only.
int ref_get_dt(void* rp, float* out) {
  RefMethod* r = (RefMethod*)rp;
  if (!r->rm) return -1;
  ranges::OMap* m = r->rm->getMap();
  for (unsigned x = 0; x < m->width; ++x)
    for (unsigned y = 0; y < m->height; ++y) out[(size_t)x * m->height + y] = r->rm->dt_at(x, y);
  return 0;
}

// CDDT table dump in CSR form.  Slices a in [0, td): bins lut_widths[a].
// pass offsets == nullptr to get sizes: returns total number of bins, *n_values = total floats.
int64_t ref_cddt_dims(void* rp, int64_t* n_values, int* widths /*td*/, float* translations /*td*/) {
  RefMethod* r = (RefMethod*)rp;
  if (!r->cddt) return -1;
  int64_t bins = 0, vals = 0;
  for (size_t a = 0; a < r->cddt->compressed_lut.size(); ++a) {
    if (widths) widths[a] = (int)r->cddt->compressed_lut[a].size();
    if (translations) translations[a] = r->cddt->lut_translations[a];
    bins += (int64_t)r->cddt->compressed_lut[a].size();
    for (auto& b : r->cddt->compressed_lut[a]) vals += (int64_t)b.size();
  }
  if (n_values) *n_values = vals;
  return bins;
}

// offsets has (total bins + 1) entries, bins ordered slice-major; values has n_values floats.
void ref_cddt_dump(void* rp, int64_t* offsets, float* values) {
  RefMethod* r = (RefMethod*)rp;
  int64_t bi = 0, vi = 0;
  for (auto& slice : r->cddt->compressed_lut) {
    for (auto& b : slice) {
      offsets[bi++] = vi;
      for (float v : b) values[vi++] = v;
    }
  }
  offsets[bi] = vi;
}

}  // extern "C"
