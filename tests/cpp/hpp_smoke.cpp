// Compile-and-run check of include/rangelib_b200.hpp (the C++ mirror of the reference's class surface).
// Without a GPU it must fail loudly (no CPU fallback); with one it casts a few rays.
#include <cstdio>
#include <vector>

#include "rangelib_b200.hpp"

int main() {
  ranges_b200::OMap map(64, 48);
  for (int x = 0; x < 64; ++x) {
    map.set(x, 0, true);
    map.set(x, 47, true);
  }
  for (int y = 0; y < 48; ++y) {
    map.set(0, y, true);
    map.set(63, y, true);
  }
  try {
    ranges_b200::RayMarchingGPU rm(map, 100.0f);
    ranges_b200::BresenhamsLine bl(map, 100.0f);
    ranges_b200::CDDTCast cddt(map, 100.0f, 36);
    cddt.prune(100.0f);
    std::vector<float> ins = {32.f, 24.f, 0.f, 32.f, 24.f, 1.5707963f, 10.f, 10.f, 3.1415927f};
    std::vector<float> a(3), b(3), c(3);
    rm.calc_range_many(ins.data(), a.data(), 3);
    bl.numpy_calc_range(ins.data(), b.data(), 3);
    cddt.numpy_calc_range(ins.data(), c.data(), 3);
    std::printf("GPU ok: rm %.3f %.3f %.3f | single %.3f\n", a[0], a[1], a[2], rm.calc_range(32.f, 24.f, 0.f));
    return (a[0] > 29.f && a[0] < 33.f) ? 0 : 2;
  } catch (const std::runtime_error& e) {
    std::printf("no device: %s\n", e.what());
    return 3;
  }
}
