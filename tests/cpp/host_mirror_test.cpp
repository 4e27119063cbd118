// host_mirror_test.cpp — exercises nim_raytracer_b200/host/nrt_host.hpp the way the reference's
// own front-end and tests drive the renderer:
//   * src/raytracer.nim:42-124  main(): Options, the mesh-bunny scene, one renderLine per scanline
//     (the worker-pool work items of :67-70), Stats accumulation, writePpm
//   * test/meshperftest.nim:22-44  1-triangle TriangleMesh hit from the origin along -z (t = 5)
// Usage: host_mirror_test <bunny.geom> <out.raw> <out.ppm>   |   host_mirror_test --expect-no-device
#include <cstring>
#include <iostream>

#include "../../nim_raytracer_b200/host/nrt_host.hpp"

using namespace nimrt;

static Scene bunnyScene(const std::string& geom) {
  // data/scenes/mesh-bunny.nim with test/bunny.geom in the teapot's place, prepared as frozen in
  // y_min subtracted, scaled x20 about the origin, v1/v2 swapped.
  auto mesh = loadGeom(geom);
  double ymin = 1e300;
  for (auto& v : mesh->vertices) ymin = std::min(ymin, v.y);
  for (auto& v : mesh->vertices) { v.y -= ymin; v.x *= 20.0; v.y *= 20.0; v.z *= 20.0; }
  for (auto& t : mesh->faces) std::swap(t.vertexIdx[1], t.vertexIdx[2]);
  calcNormals(*mesh);
  mesh->objectToWorld = translate(mat4(1.0), vec3(0.0, 0.0001, -12.0));
  mesh->worldToObject = inverse(mesh->objectToWorld);

  Scene scene;
  scene.objects = {
      Object{"mesh", mesh, Material{vec3(0.6, 0.9, 0.2), 0.0}},
      Object{"ground", initPlane(mat4(1.0)), Material{vec3(0.4), 0.0}},
  };
  scene.lights = {
      distantLight(vec3(1.0), 4.0, normalize(vec(-2.0, -0.8, -0.3))),
      distantLight(vec3(0.8, 0.3, 0.0), 1.0, normalize(vec(2.0, -0.8, -1.3))),
  };
  scene.bgColor = vec3(0.01, 0.03, 0.05);
  scene.fov = 50.0;
  scene.cameraToWorld = translate(rotate(mat4(1.0), X_AXIS, degToRad(-12.0)), vec3(0.0, 5.5, 1.5));
  return scene;
}

int main(int argc, char** argv) {
  if (argc == 2 && std::strcmp(argv[1], "--expect-no-device") == 0) {
    // host logic that needs no device: the units of raytracer.nim:67-70 as two partitions would share them
    // (serpentine deal: 0 | 1 1 | 0 0 | 1 1 | 0 ...)
    if (partitionRows(8, 0, 2, 0, 8) != std::vector<int>{0, 3, 4, 7} || partitionRows(8, 1, 2, 0, 8) != std::vector<int>{1, 2, 5, 6} ||
        unitOwner(5, 2) != 1) { std::cerr << "partitionRows / unitOwner\n"; return 3; }
    try {
      initRenderer();
    } catch (const Error& e) {
      std::cout << "no device: code " << e.code << " (" << e.what() << ")\n";
      return e.code == NRT_ERR_NO_DEVICE ? 0 : 2;
    }
    std::cout << "a device is present\n";
    return 0;
  }
  if (argc < 4) { std::cerr << "usage: host_mirror_test <bunny.geom> <out.raw> <out.ppm>\n"; return 64; }
  try {
    initRenderer();

    // --- test/meshperftest.nim:22-44 ---
    {
      std::vector<Vec4> v = {point(0.0, 1.0, -5.0), point(-2.0, -1.0, -5.0), point(2.0, -1.0, -5.0)};
      std::vector<Vec4> n = {vec(0.0, 0.0, 1.0)};
      Triangle t{{0, 1, 2}, {0, 0, 0}};
      Scene s;
      s.objects = {Object{"m", initTriangleMesh(v, n, {t}, mat4(1.0)), Material{vec3(1.0), 0.0}}};
      s.fov = 90.0;
      DeviceScene ds(s);
      Options o; o.width = 2; o.height = 2;
      std::vector<int32_t> obj(4), tri(4); std::vector<double> th(4);
      nrt_aov aov{obj.data(), tri.data(), th.data()};
      Framebuf fb = newFramebuf(2, 2);
      PageLock lock(fb);   // optional page-locking of the caller's framebuffer (released at scope exit)
      const nrt_options c = toC(o);
      check(nrt_render(ds.handle(), &c, 0, 2, 1, 1, fb.data.data(), nullptr, &aov), "nrt_render");
      if (!(obj[3] == 0 && tri[3] == 0 && th[3] == 5.0)) { std::cerr << "meshperftest: t = " << th[3] << "\n"; return 3; }
      std::cout << "meshperftest: t = " << th[3] << "\n";
    }

    // --- src/raytracer.nim:42-124 ---
    Options opts; opts.width = 300; opts.height = 200;   // raytracer.nim:47-51
    Scene scene = bunnyScene(argv[1]);
    DeviceScene ds(scene);
    Framebuf framebuf = newFramebuf(opts.width, opts.height);
    Stats totalStats;
    for (int line = 0; line < opts.height; ++line) totalStats += renderLine(ds, opts, framebuf, line);
    Framebuf whole = newFramebuf(opts.width, opts.height);
    const Stats frameStats = renderFrame(ds, opts, whole);
    if (whole.data != framebuf.data) { std::cerr << "renderLine x height != renderFrame\n"; return 4; }
    if (frameStats.numPrimaryRays != totalStats.numPrimaryRays || frameStats.numIntersectionTests != totalStats.numIntersectionTests ||
        frameStats.numIntersectionHits != totalStats.numIntersectionHits) { std::cerr << "stats differ\n"; return 5; }
    std::cout << "numPrimaryRays " << totalStats.numPrimaryRays << " numIntersectionTests " << totalStats.numIntersectionTests
              << " numIntersectionHits " << totalStats.numIntersectionHits << "\n";
    std::ofstream raw(argv[2], std::ios::binary);
    raw.write(reinterpret_cast<const char*>(framebuf.data.data()), std::streamsize(framebuf.data.size() * sizeof(float)));
    if (!writePpm(framebuf, argv[3], 8, true)) { std::cerr << "writePpm failed\n"; return 6; }
    {   // renderFrame + writePpm in one call: the same samples, 8-bit and big-endian 16-bit
      for (int bits : {8, 16}) {
        std::vector<unsigned char> fused, staged(size_t(opts.width) * opts.height * (bits <= 8 ? 3 : 6));
        const Stats qs = renderFrameQuantized(ds, opts, fused, bits, true);
        check(nrt_framebuf_quantize(framebuf.data.data(), opts.width, opts.height, bits, 1, staged.data()), "nrt_framebuf_quantize");
        if (fused != staged) { std::cerr << "renderFrameQuantized != renderFrame + writePpm at " << bits << " bits\n"; return 7; }
        if (qs.numPrimaryRays != frameStats.numPrimaryRays || qs.numIntersectionTests != frameStats.numIntersectionTests) { std::cerr << "quantized stats differ\n"; return 8; }
      }
      std::vector<unsigned char> rgba;
      if (!toRGBA8(framebuf, rgba, 200) || rgba.size() != size_t(opts.width) * opts.height * 4 || rgba[3] != 200) { std::cerr << "toRGBA8 failed\n"; return 9; }
      std::cout << "output stage: fused == staged (8, 16 bits), rgba8 ok\n";
    }
    shutdown();
  } catch (const Error& e) {
    std::cerr << "error " << e.code << ": " << e.what() << "\n";
    return 1;
  }
  return 0;
}
