// Test helper: runs the C++ host mirror's loaders (nim_raytracer_b200/host/nrt_host.hpp) and dumps what they
// produced, so tests/test_loaders.py can compare them with loaders.py bit for bit.  No GPU, no libnrt.so.
//   loader_tool in.obj out.geom dump.bin        loadObj  -> writeGeom + raw arrays
//   loader_tool --geom in.geom out.geom dump.bin loadGeom -> writeGeom + raw arrays
#include <cstdio>
#include <cstring>
#include <fstream>

#include "../../nim_raytracer_b200/host/nrt_host.hpp"

int main(int argc, char** argv) {
  using namespace nimrt;
  const bool geom = argc > 1 && std::strcmp(argv[1], "--geom") == 0;
  if (argc != (geom ? 5 : 4)) { std::fprintf(stderr, "usage: loader_tool [--geom] in out.geom dump.bin\n"); return 2; }
  const char* in = argv[geom ? 2 : 1];
  const char* out = argv[geom ? 3 : 2];
  const char* dump = argv[geom ? 4 : 3];
  try {
    auto m = geom ? loadGeom(in) : loadObj(in);
    writeGeom(out, *m);
    std::ofstream f(dump, std::ios::binary);
    const int64_t nv = int64_t(m->vertices.size()), nf = int64_t(m->faces.size());
    f.write(reinterpret_cast<const char*>(&nv), 8);
    f.write(reinterpret_cast<const char*>(&nf), 8);
    for (const auto& v : m->vertices) { const double a[4] = {v.x, v.y, v.z, v.w}; f.write(reinterpret_cast<const char*>(a), 32); }
    for (const auto& v : m->normals) { const double a[4] = {v.x, v.y, v.z, v.w}; f.write(reinterpret_cast<const char*>(a), 32); }
    for (const auto& t : m->faces) f.write(reinterpret_cast<const char*>(t.vertexIdx), 24);
  } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
  return 0;
}
