// geomtest_ref_shim.cpp — builds the reference's OWN C++ AABB slab test
// (test/geomtest.cpp:51-78 in the nim-raytracer tree) into oracle/_ref/ from
// where it lies under $(NIMRT_REFERENCE); nothing is copied into this repo.
// TEST INFRASTRUCTURE ONLY: used by tests/test_oracle_golden.py to pin the
// oracle's restatement of AABB.intersect (geom.nim:76-96) against compiled
// reference code.  The reference file is a stand-alone benchmark with its own
// main(); it is renamed away here and two C entry points are added.
#ifndef NIMRT_GEOMTEST_CPP
#error "compile with -DNIMRT_GEOMTEST_CPP='\"<reference>/test/geomtest.cpp\"'"
#endif
#define main nimrt_geomtest_benchmark_main
#include NIMRT_GEOMTEST_CPP
#undef main
#undef min
#undef max

extern "C" double geomtest_ref_aabb_intersect(const double* vmin, const double* vmax,
                                              const double* orig, const double* dir) {
  Ray ray;
  Init(&ray, Vector{{orig[0], orig[1], orig[2]}}, Vector{{dir[0], dir[1], dir[2]}});
  AABB box;
  Init(&box, Vector{{vmin[0], vmin[1], vmin[2]}}, Vector{{vmax[0], vmax[1], vmax[2]}});
  return intersect(&box, &ray);
}
