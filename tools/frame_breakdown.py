"""Dev tool: frame time and prefilter time per workload (device-resident output)."""
import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import bench
from nim_raytracer_b200 import api
L = api.lib()
api.initRenderer(devices=[0])
import os
if os.environ.get('NRT_PART'):
    i, c = map(int, os.environ['NRT_PART'].split(','))
    api.setPartition(i, c)
for wl in sys.argv[1:] or ["config2"]:
    sc, o, desc = bench.workload(wl)
    co = o.to_c()
    ds = api.DeviceScene(sc)
    fb = C.c_void_p(); api.check(L.nrt_device_alloc(o.width * o.height * 12, C.byref(fb)), "alloc")
    cs = api.nrt_stats()
    def step():
        api.check(L.nrt_render_device(ds.handle, C.byref(co), 0, o.height, 1, 1, fb, C.byref(cs), None), "r")
    for _ in range(3): step()
    t0 = time.perf_counter(); n = 5
    for _ in range(n): step()
    wall = (time.perf_counter() - t0) / n * 1e3
    p = ds.profile()
    print(f"{wl}: wall {wall:.3f} ms frame {p.total_ms:.3f} ms prefilter {p.mesh_filter_ms:.3f} ms by_mode {[round(x,3) for x in p.mesh_ms_by_mode[:3]]} "
          f"tests {list(p.mesh_tests_by_mode)[:3]} pre {p.pre_candidates} cand {p.candidates} launches {p.kernel_launches} Mrays/s {cs.num_rays / wall / 1e3:.0f}")
    api.setKernelTiming(True); step(); kt = ds.kernelTimes(); api.setKernelTiming(False)
    tot = sum(ms for ms, _ in kt.values())
    print(f"   active/bounce {list(p.active_samples)[:4]} wavefront/bounce {list(p.wavefront_samples)[:4]} tail {p.tail_samples} lanes {p.lanes}")
    print("   " + " | ".join(f"{k.split(' ')[0]} {ms:.2f}" for k, (ms, n) in sorted(kt.items(), key=lambda kv: -kv[1][0])) + f" | sum {tot:.2f}")
    import hashlib, numpy as np
    host = np.empty(o.width * o.height * 3, dtype=np.float32)
    api.check(L.nrt_copy_to_host(host.ctypes.data_as(C.c_void_p), fb, host.nbytes), "copy")
    print(f"   fb sha256 {hashlib.sha256(host.tobytes()).hexdigest()[:16]} stats {cs.num_primary_rays} {cs.num_intersection_tests} {cs.num_intersection_hits} {cs.num_rays} {cs.num_capped_samples}"
          f"  NRT_PATH={os.environ.get('NRT_PATH', '1')} lib={os.path.basename(api.LIB_PATH)}")
    L.nrt_device_free(fb); ds.close()
