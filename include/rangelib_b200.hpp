// rangelib_b200.hpp -- header-only C++ mirror of the reference's class surface
// (/root/reference/includes/RangeLib.h: ranges::OMap :121, RangeMethod :412, BresenhamsLine :691,
// RayMarchingGPU :774, RayMarching :922, CDDTCast :971, GiantLUTCast :1772) on top of the C ABI
// (rangelib_b200.h).  Same class names, constructor arguments and batched member functions, so C++ code
// written against RangeLib.h -- e.g. the reference's own Cython binding, pywrapper/RangeLibc.pyx:25-92 --
// compiles against this header by switching the include and the namespace.  Every method runs on the GPU.
//
// Differences from the reference, all at the edges:
//  * failures throw std::runtime_error (the reference prints, throws std::string or is undefined);
//  * OMap is built from an occupancy array; PNG decoding is not part of this backend;
//  * set_sensor_model replaces the table (the reference appends rows on a second call);
//  * RayMarchingGPU::calc_range works (the reference prints a message and returns -1).
#ifndef RANGELIB_B200_HPP
#define RANGELIB_B200_HPP

#include <stdexcept>
#include <string>
#include <vector>

#include "rangelib_b200.h"

namespace ranges_b200 {

inline void check(int rc) {
  if (rc != RL_OK) throw std::runtime_error(std::string("rangelib_b200: ") + rl_last_error());
}

// OMap: grid[x][y] semantics, x-major storage; world parameters as plain fields like the reference (:134-141)
struct OMap {
  unsigned width = 0, height = 0;
  std::vector<unsigned char> grid;  // grid[x * height + y] != 0 <=> occupied
  float world_scale = 1.0f, world_angle = 0.0f, world_origin_x = 0.0f, world_origin_y = 0.0f;
  float world_sin_angle = 0.0f, world_cos_angle = 1.0f;

  OMap(int w, int h) : width(w), height(h), grid((size_t)w * h, 0) {}
  bool get(int x, int y) const { return grid[(size_t)x * height + y] != 0; }
  void set(int x, int y, bool occupied) { grid[(size_t)x * height + y] = occupied ? 1 : 0; }
  bool isOccupied(int x, int y) const {
    if (x < 0 || x >= (int)width || y < 0 || y >= (int)height) return false;
    return get(x, y);
  }
  bool error() const { return false; }
};

class RangeMethod {
 public:
  virtual ~RangeMethod() {
    if (h_) rl_method_destroy(h_);
  }
  RangeMethod(const RangeMethod&) = delete;
  RangeMethod& operator=(const RangeMethod&) = delete;

  float calc_range(float x, float y, float heading) {
    float out = 0.0f;
    check(rl_calc_range(h_, x, y, heading, &out));
    return out;
  }
  void numpy_calc_range(float* ins, float* outs, int num_casts) { check(rl_numpy_calc_range(h_, ins, outs, num_casts)); }
  void numpy_calc_range_angles(float* ins, float* angles, float* outs, int num_particles, int num_angles) {
    check(rl_numpy_calc_range_angles(h_, ins, angles, outs, num_particles, num_angles));
  }
  void set_sensor_model(double* table, int table_width) { check(rl_set_sensor_model(h_, table, table_width)); }
  void eval_sensor_model(float* obs, float* ranges, double* outs, int rays_per_particle, int particles) {
    check(rl_eval_sensor_model(h_, obs, ranges, outs, rays_per_particle, particles));
  }
  void calc_range_repeat_angles_eval_sensor_model(float* ins, float* angles, float* obs, double* weights,
                                                  int num_particles, int num_angles) {
    check(rl_calc_range_repeat_angles_eval_sensor_model(h_, ins, angles, obs, weights, num_particles, num_angles));
  }
  void calc_range_many_radial_optimized(float* ins, float* outs, int num_particles, int num_rays, float min_angle,
                                        float max_angle) {
    check(rl_calc_range_many_radial_optimized(h_, ins, outs, num_particles, num_rays, min_angle, max_angle));
  }
  // extensions (not in RangeLib.h): particle-filter steps either side of the sensor update, see rangelib_b200.h
  double normalize_weights(double* weights, int n, double inv_squash = 1.0) {
    double sum = 0.0;
    check(rl_pf_normalize_weights(h_, weights, n, inv_squash, &sum));
    return sum;
  }
  void resample(const float* particles, const double* weights, float* out_particles, int n, double u0) {
    check(rl_pf_resample(h_, particles, weights, out_particles, n, u0));
  }
  void motion_update(float* particles, int n, float dx, float dy, float dtheta, const float* noise = nullptr) {
    check(rl_pf_motion_update(h_, particles, n, dx, dy, dtheta, noise));
  }
  float maxRange() const { return max_range_; }
  long long memory() const { return (long long)rl_method_memory(h_); }
  rl_method* handle() { return h_; }

 protected:
  RangeMethod(int kind, const OMap& m, float mr, unsigned td) : max_range_(mr) {
    rl_map* map = nullptr;
    check(rl_map_create(m.grid.data(), (int)m.width, (int)m.height, &map));
    int rc = rl_map_set_world(map, m.world_scale, m.world_angle, m.world_origin_x, m.world_origin_y, m.world_sin_angle,
                              m.world_cos_angle);
    if (rc == RL_OK) rc = rl_method_create(kind, map, mr, td, -1, &h_);
    rl_map_destroy(map);
    check(rc);
  }
  rl_method* h_ = nullptr;
  float max_range_;
};

class BresenhamsLine : public RangeMethod {
 public:
  BresenhamsLine(const OMap& m, float mr) : RangeMethod(RL_BL, m, mr, 0) {}
};

class RayMarching : public RangeMethod {
 public:
  RayMarching(const OMap& m, float mr) : RangeMethod(RL_RM, m, mr, 0) {}
};

class RayMarchingGPU : public RangeMethod {
 public:
  RayMarchingGPU(const OMap& m, float mr) : RangeMethod(RL_RM, m, mr, 0) {}
  // grid coordinates, no world conversion (RangeLib.h:819-831)
  void calc_range_many(float* ins, float* outs, int num_casts) { check(rl_calc_range_many(h_, ins, outs, num_casts)); }
};

class CDDTCast : public RangeMethod {
 public:
  CDDTCast(const OMap& m, float mr, unsigned int td) : RangeMethod(RL_CDDT, m, mr, td) {}
  void prune(float max_range) { check(rl_method_prune(h_, max_range)); }
  // extension: binary checkpoint of the (pruned) table; load with rl_method_create_from_cddt (rangelib_b200.h)
  void save(const char* path) { check(rl_method_save_cddt(h_, path)); }
};

class GiantLUTCast : public RangeMethod {
 public:
  GiantLUTCast(const OMap& m, float mr, int td) : RangeMethod(RL_GLT, m, mr, (unsigned)td) {}
};

}  // namespace ranges_b200

#endif  // RANGELIB_B200_HPP
