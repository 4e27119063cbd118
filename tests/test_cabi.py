"""The C-ABI library loads on a CPU-only machine, exports every entry point declared in
include/nrt.h, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from nim_raytracer_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "nrt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nrt_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as g
    g.build_cuda()
    return api.lib()


def test_exports_every_declared_symbol(built_lib):
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/nrt.h but not exported by libnrt.so"
    assert built_lib.nrt_abi_version() == 2


def test_built_for_sm_100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_struct_layout_matches_header():
    # sizes the C compiler gives the POD structs == the ctypes mirror (api.py)
    code = r'''
    #include "nrt.h"
    #include <stdio.h>
    int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(nrt_object), sizeof(nrt_mesh), sizeof(nrt_light),
      sizeof(nrt_scene_desc), sizeof(nrt_options), sizeof(nrt_stats), sizeof(nrt_aov), sizeof(nrt_profile));return 0;}
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(code)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(td, "s.c"), "-o", os.path.join(td, "s")], check=True)
        sizes = [int(x) for x in subprocess.run([os.path.join(td, "s")], capture_output=True, text=True).stdout.split()]
    mirror = [api.nrt_object, api.nrt_mesh, api.nrt_light, api.nrt_scene_desc, api.nrt_options, api.nrt_stats,
              api.nrt_aov, api.nrt_profile]
    assert sizes == [C.sizeof(m) for m in mirror]


def test_no_cpu_fallback_without_gpu(built_lib):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    assert built_lib.nrt_init(1, None) == -3                       # NRT_ERR_NO_DEVICE
    assert b"no CPU fallback" in built_lib.nrt_last_error()
    h = C.c_void_p()
    from nim_raytracer_b200 import scenes
    d = api.SceneDesc(scenes.boxtest())
    assert built_lib.nrt_scene_create(d.ref(), C.byref(h)) == -4   # NRT_ERR_NOT_INIT: nothing renders on the CPU
    with pytest.raises(api.NrtError):
        api.initRenderer(1)


def test_page_locking_degrades_without_gpu(built_lib):
    # nrt_host_register is optional: without a device it reports an error code (never aborts) and
    # api.pinSceneArrays() simply pins nothing, so callers keep working with pageable memory
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    import numpy as np
    from nim_raytracer_b200 import scenes
    buf = np.zeros(1 << 16, dtype=np.float32)
    assert built_lib.nrt_host_register(buf.ctypes.data_as(C.c_void_p), buf.nbytes) != 0
    assert b"cudaHostRegister" in built_lib.nrt_last_error()
    assert built_lib.nrt_host_register(None, 16) == -1 and built_lib.nrt_host_unregister(None) == 0
    assert api.pinSceneArrays(scenes.bunny(stride=16)) == []


def test_product_never_touches_the_oracle():
    # nothing under nim_raytracer_b200/ or include/ may import, link or name oracle/ or the emulation
    bad = []
    for base in ("nim_raytracer_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".h", ".cu", ".cuh", ".hpp", ".cpp")):
                    s = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"^\s*(import|from)\s+oracle\b", s, flags=re.M) or "liboracle" in s or "libnrt_emu" in s:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
    syms = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in syms and "emu_" not in syms


def test_unit_owner_matches_the_python_mirror():
    # the serpentine deal of bands to partitions: the library's definition (nrt_unit_owner) and distributed.py's
    import ctypes as C
    import numpy as np
    from nim_raytracer_b200 import api, distributed as D
    L = api.lib()
    L.nrt_unit_owner.argtypes = [C.c_longlong, C.c_int]
    L.nrt_unit_owner.restype = C.c_int
    for world in (1, 2, 3, 4, 8):
        units = np.arange(0, 100)
        mine = D.unit_owner(units, world)
        for u in units:
            assert L.nrt_unit_owner(int(u), world) == int(mine[u]) == D.unit_owner(int(u), world)
        # every partition gets the same number of units per two rounds, at the same mean position
        for r in range(world):
            pos = [u % (2 * world) for u in range(2 * world) if D.unit_owner(u, world) == r]
            assert len(pos) == 2 and sum(pos) == 2 * world - 1
    assert L.nrt_unit_owner(-1, 4) == -1 and L.nrt_unit_owner(3, 0) == -1


def test_partition_rows_is_a_partition_and_equals_the_python_mirror():
    # nrt_partition_rows (the library's own enumeration of the units renderImpl deals out) against distributed.rows_of;
    # over all partitions the units are disjoint and complete, whatever y0 / y1 / step / band — progressive passes with
    # step >= count included (every partition still gets rows).  Host logic: no device needed.
    from nim_raytracer_b200 import distributed as D
    rng = np.random.default_rng(7)
    for _ in range(300):
        height = int(rng.integers(1, 400))
        y0 = int(rng.integers(-3, height))
        y1 = int(rng.integers(y0, height + 5))
        step = 1 << int(rng.integers(0, 4))
        band = 1 if step > 1 else 1 << int(rng.integers(0, 5))
        count = int(rng.integers(1, 10))
        seen = []
        for k in range(count):
            rows = api.partitionRows(height, k, count, y0, y1, step, band)
            assert rows == D.rows_of(k, count, height, y0, y1, step, band)
            seen += rows
        want = [y for y in range(max(0, y0), min(y1, height)) if (y - y0) % (step * band) == 0]
        assert sorted(seen) == want and len(set(seen)) == len(seen)
        if len(want) >= 2 * count:
            assert all(api.partitionRows(height, k, count, y0, y1, step, band) for k in range(count))
    L = api.lib()
    assert L.nrt_partition_rows(0, 0, 1, 1, 1, 0, 1, None, 0) == -1 and L.nrt_partition_rows(8, 0, 8, 1, 1, 2, 2, None, 0) == -1
    buf = (C.c_int * 2)()
    assert L.nrt_partition_rows(16, 0, 16, 1, 1, 0, 2, buf, 2) == 8 and list(buf) == [0, 3]    # truncated write, full count
