import sys, os, ctypes as C
sys.path.insert(0, '/root/repo')
import bench
from nim_raytracer_b200 import api
L = api.lib()
api.initRenderer(devices=[0])
if os.environ.get('NRT_PART'):
    i, c = map(int, os.environ['NRT_PART'].split(',')); api.setPartition(i, c)
sc, o, desc = bench.workload("config4")
co = o.to_c(); ds = api.DeviceScene(sc)
fb = C.c_void_p(); api.check(L.nrt_device_alloc(o.width * o.height * 12, C.byref(fb)), "alloc")
cs = api.nrt_stats()
for k in range(3):
    if k == 2: os.environ['NRT_TRACE_PREFILTER'] = '1'
    api.check(L.nrt_render_device(ds.handle, C.byref(co), 0, o.height, 1, 1, fb, C.byref(cs), None), "r")
