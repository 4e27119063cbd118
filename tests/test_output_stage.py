"""Output stage: writePpm's sample conversion (utils/framebuf.nim:55-93 with
utils/color.nim:17-22) and ImageRGBA.copyFrom (utils/image.nim:45-54).  Byte work: every comparison is `==`.

The device does not evaluate pow(): it counts the entries of a cut-point table (built on the host by libnrt.so with the
reference's own libm call) that are <= the input.  The CPU tests prove that table against the oracle's literal
restatement of outvalue over EVERY float32 input of the pow branch; the GPU tests check the kernels and the fused
epilogue of the pixel store."""
import numpy as np
import pytest

from nim_raytracer_b200 import api, scenes

LO = np.float32(0.0031308)


def _branch_floats(lo_bits, hi_bits):
    return np.arange(lo_bits, hi_bits, dtype=np.uint32).view(np.float32)


@pytest.mark.parametrize("bits", [8, 16])
def test_cut_points_equal_the_oracle_on_every_float_of_the_pow_branch(oracle_mod, bits):
    thr = api.outputCutPoints(bits)
    assert thr.shape == ((1 << bits) - 1,) and (np.diff(thr) >= 0).all()
    lo = int(LO.view(np.uint32)) + 1
    hi = int(np.float32(1.0).view(np.uint32)) + 1
    step = 1 << 23
    for a in range(lo, hi, step):                      # ~70 M floats, 8 M at a time
        v = _branch_floats(a, min(a + step, hi))
        got = np.searchsorted(thr, v, side="right")
        want = oracle_mod.outvalues(v, bits, True)
        assert (got == want).all(), f"first difference at {v[np.nonzero(got != want)[0][0]]!r}"


@pytest.mark.parametrize("bits", [1, 5, 10, 12])
def test_cut_points_other_depths(oracle_mod, bits):
    thr = api.outputCutPoints(bits)
    rng = np.random.default_rng(bits)
    v = rng.uniform(float(LO), 1.0, 2_000_000).astype(np.float32)
    v = v[v > LO]
    near = np.concatenate([(thr.view(np.uint32).astype(np.int64) + d).astype(np.uint32).view(np.float32) for d in range(-3, 4)])
    near = near[(near > LO) & (near <= 1.0)]
    v = np.concatenate([v, near])
    assert (np.searchsorted(thr, v, side="right") == oracle_mod.outvalues(v, bits, True)).all()


def test_oracle_outvalue_known_answers(oracle_mod, tmp_path):
    x = np.array([0.0, 1.0, 0.5, -3.0, 7.0, np.nan, 0.0031308, 1e-9], dtype=np.float32)
    # linear 0.5 -> sRGB 0.7354 -> 187.5.. -> 188; the linear segment: 12.92 * 0.0031308 * 255 = 10.3 -> 10
    assert oracle_mod.outvalues(x, 8, True).tolist() == [0, 255, 188, 0, 255, 0, 10, 0]
    assert oracle_mod.outvalues(x, 8, False).tolist() == [0, 255, 128, 0, 255, 0, 1, 0]
    assert oracle_mod.outvalues(x, 1, False).tolist() == [0, 1, 1, 0, 1, 0, 0, 0]          # round half away from zero: 0.5 -> 1
    assert oracle_mod.outvalues(x, 16, False).tolist() == [0, 65535, 32768, 0, 65535, 0, 205, 0]
    # monotonic, full range
    ramp = np.linspace(0, 1, 100001, dtype=np.float32)
    for bits in (8, 16):
        o = oracle_mod.outvalues(ramp, bits, True).astype(np.int64)
        assert o[0] == 0 and o[-1] == (1 << bits) - 1 and (np.diff(o) >= 0).all()
    # the PPM container (framebuf.nim:61-71): "P6 w h maxval " then samples, 16-bit big-endian
    fb = api.newFramebuf(3, 2)
    fb.data[:] = np.linspace(0, 1, 18, dtype=np.float32)
    p = str(tmp_path / "t.ppm")
    assert oracle_mod.write_ppm(fb, p, bits=16, sRGB=False)
    raw = open(p, "rb").read()
    assert raw.startswith(b"P6 3 2 65535 ") and len(raw) == len(b"P6 3 2 65535 ") + 36
    assert raw[-2:] == bytes([255, 255]) and raw[len(b"P6 3 2 65535 "):][:2] == bytes([0, 0])
    assert oracle_mod.rgba8(np.array([0.5, -1.0, 2.0, np.nan, 0.2, 1.0], dtype=np.float32), 7).tolist() == [128, 0, 255, 7, 0, 51, 255, 7]


def _special_floats():
    rng = np.random.default_rng(3)
    v = [rng.uniform(-0.2, 1.2, 300_000), rng.uniform(0.0, 0.01, 50_000), 10.0 ** rng.uniform(-12, 0.1, 50_000)]
    sp = [0.0, -0.0, 1.0, 0.0031308, np.nextafter(np.float32(0.0031308), np.float32(1)), np.inf, -np.inf, np.nan, 1e-45, 0.5, 0.99999994]
    for bits in (1, 5, 8, 10, 16):                     # either side of every cut point (up to 4096 of them per depth)
        t = api.outputCutPoints(bits)[:: max(1, ((1 << bits) - 1) // 4096)]
        for d in (-1, 0, 1):
            v.append((t.view(np.uint32).astype(np.int64) + d).astype(np.uint32).view(np.float32))
        mv = float((1 << bits) - 1)                    # and of the linear (sRGB off) rounding points
        k = np.arange(0, min(int(mv), 4096)) + 0.5
        v.append((k / mv).astype(np.float32))
    a = np.concatenate([np.asarray(x, dtype=np.float32) for x in v] + [np.asarray(sp, dtype=np.float32)])
    return a[: (a.size // 3) * 3]


@pytest.mark.gpu
def test_output_stage_bit_exact(oracle_mod):
    import ctypes as C
    api.initRenderer(1)
    L = api.lib()
    a = _special_floats()
    fb = api.newFramebuf(a.size // 3, 1)
    fb.data[:] = a
    for bits in (1, 5, 8, 10, 16):
        for srgb in (False, True):
            got = api.framebufQuantize(fb, bits, srgb).reshape(-1)
            assert (got == oracle_mod.outvalues(a, bits, srgb)).all(), (bits, srgb)
    assert (api.framebufToRgba8(fb, 0x7F).reshape(-1) == oracle_mod.rgba8(a, 0x7F)).all()
    # the same on device pointers: no staging, no copies
    npix = a.size // 3
    d_fb, d_out = C.c_void_p(), C.c_void_p()
    api.check(L.nrt_device_alloc(a.nbytes, C.byref(d_fb)), "alloc")
    api.check(L.nrt_device_alloc(npix * 6, C.byref(d_out)), "alloc")
    L.nrt_copy_to_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    api.check(L.nrt_copy_to_device(d_fb, a.ctypes.data_as(C.c_void_p), a.nbytes), "h2d")
    L.nrt_framebuf_quantize_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    api.check(L.nrt_framebuf_quantize_device(d_fb, npix, 1, 16, 1, d_out), "quantize_device")
    out = np.zeros(npix * 3, dtype=">u2")
    api.check(L.nrt_copy_to_host(out.ctypes.data_as(C.c_void_p), d_out, out.nbytes), "d2h")
    assert (out == oracle_mod.outvalues(a, 16, True)).all()
    L.nrt_device_free(d_fb); L.nrt_device_free(d_out)


@pytest.mark.gpu
def test_render_quantized_is_render_then_writePpm(oracle_mod):
    api.initRenderer(1)
    sc = scenes.bunny_spheres(stride=8)
    o = api.Options(200, 112, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=4)
    rfb, rst, _ = oracle_mod.render(sc, o)
    ds = api.DeviceScene(sc)
    for bits, srgb in ((8, True), (8, False), (16, True), (5, True)):
        img, st = api.renderFrameQuantized(ds, o, bits=bits, sRGB=srgb)
        assert st == rst
        assert (img.reshape(-1) == oracle_mod.outvalues(rfb.data, bits, srgb)).all(), (bits, srgb)
    img, st = api.renderFrameQuantized(ds, o, rgba=True, alpha=0xC0)
    assert (img.reshape(-1) == oracle_mod.rgba8(rfb.data, 0xC0)).all() and st == rst
    # progressive passes into the same integer image (renderer.nim:174-178,204-207)
    img = np.zeros((o.height, o.width, 3), dtype=np.uint8)
    ref = api.newFramebuf(o.width, o.height)
    step = 4
    while step >= 1:
        api.renderFrameQuantized(ds, o, bits=8, sRGB=True, step=step, maxStep=4, image=img)
        oracle_mod.render(sc, o, fb=ref, step=step, maxStep=4)
        assert (img.reshape(-1) == oracle_mod.outvalues(ref.data, 8, True)).all(), step
        step //= 2
    ds.close()


def test_framebuf_nim_self_test_replay(oracle_mod, tmp_path):
    """utils/framebuf.nim:98-129, the reference's own self-test of this stage: store / load round trip of two pixels
    (float32 storage: `eq` at 1e-15 cannot hold — its "TODO fix tests" — the float32 roundings do), then the 1024 x 768
    gradient written as 8- and 16-bit PPMs.  The reference checks only that writePpm returns true; here the oracle's
    files are checked against first principles and the product's cut-point tables must give the same samples."""
    W, H = 1024, 768
    fb = api.newFramebuf(W, H)
    img = fb.data.reshape(H, W, 3)
    img[0, 0] = (0.2, 0.6, 0.5)
    assert img[0, 0].tolist() == [float(np.float32(0.2)), float(np.float32(0.6)), 0.5]
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)
    img[..., 0], img[..., 1], img[..., 2] = y / (H - 1), x / (W - 1), (H - 1 - y) / (H - 1)    # framebuf.nim:121-126
    for bits in (8, 16):
        path = str(tmp_path / f"test-{bits}bit.ppm")
        assert oracle_mod.write_ppm(fb, path, bits=bits)
        raw = open(path, "rb").read()
        head = f"P6 {W} {H} {(1 << bits) - 1} ".encode()
        assert raw.startswith(head) and len(raw) == len(head) + W * H * 3 * (bits // 8)
        s = np.frombuffer(raw[len(head):], dtype=np.uint8 if bits == 8 else ">u2").reshape(H, W, 3).astype(np.int64)
        top = (1 << bits) - 1
        assert s[0, 0].tolist() == [0, 0, top] and s[H - 1, W - 1].tolist() == [top, top, 0]
        assert (np.diff(s[:, 0, 0]) >= 0).all() and (np.diff(s[0, :, 1]) >= 0).all() and (s[:, :, 0] + s[::-1, :, 2] == 2 * s[:, :, 0]).all()
        # the product's conversion of the same framebuffer (what the device evaluates: count of cut points <= v on the
        # pow branch, the exactly rounded linear segment below it)
        v = fb.data
        want = oracle_mod.outvalues(v, bits, True).astype(np.int64)
        assert (want == s.reshape(-1)).all()
        thr = api.outputCutPoints(bits)
        hi = v > LO
        assert (np.searchsorted(thr, v[hi], side="right") == want[hi]).all()
