import sys, time, ctypes as C
sys.path.insert(0,'/root/repo')
import torch
from nim_raytracer_b200 import api, scenes, distributed as D
L = api.lib()
api.initRenderer(devices=[0])
sc = scenes.bunny(); o = api.Options(1920,1080); co = o.to_c()
ds = api.DeviceScene(sc)
W,H = 1920,1080
peer = D.PeerFramebuffer(W*H*12, 0, 1, None)
cs = api.nrt_stats()
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
def step():
    api.check(L.nrt_render_device(ds.handle, C.byref(co), 0, H, 1, 1, peer.ptr, C.byref(cs), None), "r")
for _ in range(3): step()
for mode in ("noflush", "flush", "flush_nosync", "sleep5ms"):
    tw=[]; tf=[]
    for _ in range(10):
        if mode.startswith("flush"):
            flush.fill_(1)
            if mode == "flush": torch.cuda.synchronize()
        if mode == "sleep5ms": time.sleep(0.005)
        t0=time.perf_counter(); step(); t1=time.perf_counter()
        tw.append((t1-t0)*1e3); tf.append(ds.profile().total_ms)
    p = ds.profile()
    print(mode, "wall ms", [round(x,2) for x in tw[2:6]], "frame ms", [round(x,2) for x in tf[2:6]], "filter ms", round(p.mesh_filter_ms,3), "launches", p.kernel_launches, list(p.mesh_ms_by_mode), list(p.mesh_tests_by_mode))

import threading, pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for name, fn in (("clock", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), ("maxclock", lambda: pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)),
                 ("reasons", lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)), ("power", lambda: pynvml.nvmlDeviceGetPowerUsage(h))):
    t0=time.perf_counter()
    for _ in range(5): fn()
    print("nvml", name, "ms/call", (time.perf_counter()-t0)/5*1e3)
def run_loop(n=20):
    t0=time.perf_counter()
    for _ in range(n): step()
    return (time.perf_counter()-t0)/n*1e3
print("baseline loop ms/step", run_loop())
for label, queries, period in (("clock only 50ms", ["clock"], 0.05), ("clock+reasons 50ms", ["clock","reasons"], 0.05), ("all4 50ms", ["clock","maxclock","reasons","power"], 0.05), ("clock+reasons 20ms", ["clock","reasons"], 0.02)):
    stop=[False]; cnt=[0]
    def poll():
        while not stop[0]:
            if "clock" in queries: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            if "maxclock" in queries: pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            if "reasons" in queries: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            if "power" in queries: pynvml.nvmlDeviceGetPowerUsage(h)
            cnt[0]+=1
            time.sleep(period)
    th=threading.Thread(target=poll, daemon=True); th.start()
    r=run_loop(40); stop[0]=True; th.join()
    print(label, "loop ms/step", round(r,3), "polls", cnt[0])

sys.path.insert(0, '/root/repo')
import bench
for use_sampler in (False, True, False, True):
    s = bench.ClockSampler(0)
    if use_sampler: s.start()
    torch.cuda.synchronize()
    t0=time.perf_counter()
    L.nrt_timer_begin()
    for _ in range(10):
        flush.fill_(1); torch.cuda.synchronize(); step(); p = ds.profile()
    ms = C.c_double(); L.nrt_timer_end(C.byref(ms))
    wall=(time.perf_counter()-t0)*1e3
    r = s.stop() if use_sampler else None
    print("bench-like loop sampler", use_sampler, "ms/step dev", ms.value/10, "wall", wall/10, r)

# ---- e2e breakdown: scene update (H2D + record build) vs host-buffer render (D2H) ----
hp = C.c_void_p(); L.nrt_host_alloc_pinned(W*H*12, C.byref(hp))
import numpy as np
pageable = np.zeros(W*H*3, dtype=np.float32)
for label, target in (("pinned fb", hp), ("pageable fb", pageable.ctypes.data_as(C.c_void_p))):
    tu=[]; tr=[]
    for _ in range(8):
        t0=time.perf_counter(); ds.update(); t1=time.perf_counter()
        api.check(L.nrt_render(ds.handle, C.byref(co), 0, H, 1, 1, target, C.byref(cs), None), "r"); t2=time.perf_counter()
        tu.append((t1-t0)*1e3); tr.append((t2-t1)*1e3)
    print("e2e", label, "update ms", [round(x,2) for x in tu[2:6]], "render ms", [round(x,2) for x in tr[2:6]], "frame", round(ds.profile().total_ms,2))
