"""The reference's FAST version (TDecGop.cpp:225-333, 486-656 + salmat.cpp) on the GPU (use_rbf = CDS_FUSION_FAST) against
  * the reference's own golden outputs bin/HMdata/0..38.png (tests/golden/fast_BasketballDrill.npz), and
  * the fast-version oracle (oracle/cds_oracle_fast.c), which is pinned by the same PNGs.
The kernels reproduce the reference's float32 operation order, so the bar against the oracle is bit-exact: the 8-bit maps, the
FP32 maps before x255 and the camera-motion votes must be identical, not merely within 1e-4."""
import os

import numpy as np
import pytest

import _oracle as o

pytestmark = pytest.mark.gpu
GOLD = os.path.join(o.ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def cds():
    import compressd_domain_saliencyprediction_b200 as pkg
    assert pkg.device_ok(), pkg.lib().cds_last_error()
    return pkg


def _oracle_clip(w, h, fps, hdr, tok, base, byts, frames=None):
    fo = o.FastOracle(w, h, fps)
    u8s, f32s, cams = [], [], []
    for f in range(hdr.shape[0] if frames is None else frames):
        mvx, mvy, dep = o.orc_getframe(w, h, hdr[f], tok[base[f]:base[f + 1]])
        u8, f32, cam, _ = fo.frame(mvx, mvy, dep, byts[f])
        u8s.append(u8); f32s.append(f32); cams.append(cam.copy())
    fo.close()
    return np.stack(u8s), np.stack(f32s), np.stack(cams)


def test_fast_version_reproduces_the_reference_golden_pngs(cds):
    d = np.load(os.path.join(GOLD, "fast_BasketballDrill.npz"))
    w, h, fps = int(d["w"]), int(d["h"]), float(d["fps"])
    clip = cds.SaliencyClip(w, h, fps, fast=True, out_u8=True, max_batch=16)
    got, cam = clip.process(0, d["hdr"], d["tok"], d["tok_base"], d["ctu_bytes"])
    clip.close()
    diff = np.abs(got.astype(np.int16) - d["png"].astype(np.int16))
    per = (diff > 0).reshape(len(got), -1).sum(1)
    print(f"GPU fast version vs the reference's PNGs: {int((per == 0).sum())} of {len(got)} pictures byte-identical, "
          f"{int(per.sum())} pixels differ in total (max |diff| {int(diff.max())})")
    assert diff.max() <= 1 and per.max() <= 1 and per.sum() <= 3 and (per == 0).sum() >= 36
    assert not cam.any()
    want, _, _ = _oracle_clip(w, h, fps, d["hdr"], d["tok"], d["tok_base"], d["ctu_bytes"])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name,batch", [("BasketballDrive", 5), ("BasketballPass", 32)])
def test_fast_version_bit_exact_vs_oracle_on_real_clips(cds, name, batch):
    """BasketballDrive 126..137 pans: CameraDect fires on every picture (vote_alg + Umean + translateTransform shifts)."""
    d = np.load(os.path.join(GOLD, f"real_{name}.npz"))
    w, h, fps = int(d["w"]), int(d["h"]), 50.0
    hdr, tok, base, byts = d["hdr"], d["tok"], d["tok_base"], d["ctu_bytes"]
    want_u8, want_f32, want_cam = _oracle_clip(w, h, fps, hdr, tok, base, byts)
    for u8 in (True, False):
        clip = cds.SaliencyClip(w, h, fps, fast=True, out_u8=u8, max_batch=batch)
        got, cam = clip.process(0, hdr, tok, base, byts)
        clip.close()
        assert np.array_equal(cam, want_cam)
        assert np.array_equal(got, want_u8 if u8 else want_f32)
    if name == "BasketballDrive":
        assert np.all(np.abs(want_cam).sum(1) > 0), want_cam


@pytest.mark.parametrize("w,h,fps,F,seed", [(416, 240, 50.0, 60, 5), (832, 480, 24.0, 30, 9), (1024, 768, 30.0, 12, 2), (1280, 720, 60.0, 8, 4),
                                            (416, 240, 24.0, 100, 6)])      # ring of 2*7 + 7 + 17 = 38 pictures: wraps twice
def test_fast_version_bit_exact_vs_oracle_on_synthetic_pans(cds, w, h, fps, F, seed):
    """Synthetic syntax with a global pan per picture: the camera branch fires, the clip runs past PS + PT pictures (ring wrap),
    and batches of 7 split the clip at positions unrelated to PS / PT."""
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, seed, F)
    want_u8, want_f32, want_cam = _oracle_clip(w, h, fps, hdr, tok, base, byts)
    clip = cds.SaliencyClip(w, h, fps, fast=True, out_u8=False, max_batch=7)
    got, cam = clip.process(0, hdr, tok, base, byts)
    clip.close()
    assert np.array_equal(cam, want_cam)
    assert np.abs(want_cam).sum() > 0
    bad = [f for f in range(F) if not np.array_equal(got[f], want_f32[f])]
    assert not bad, (bad, float(np.max(np.abs(got - want_f32))))


def test_fast_version_hook_surface_and_frame_range_halo(cds):
    """cds_submit_frame / cds_get_map (the decoder hook's calls) and a frame-range shard with halo give the same bytes."""
    w, h, fps, F = 416, 240, 30.0, 48
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, 11, F)
    ref = cds.SaliencyClip(w, h, fps, fast=True, out_u8=True, max_batch=16)
    want, wcam = ref.process(0, hdr, tok, base, byts)
    hook = cds.SaliencyClip(w, h, fps, fast=True, out_u8=True, max_batch=4)
    for f in range(10):
        hook.submit_frame(f, byts[f], hdr[f], tok[base[f]:base[f + 1]])
    hook.flush()
    m, cam = hook.get_map(9)
    assert np.array_equal(m, want[9]) and np.array_equal(cam, wcam[9])
    hook.close()
    first = 40
    halo = ref.PS + ref.PT - 2
    shard = cds.SaliencyClip(w, h, fps, fast=True, out_u8=True, max_batch=16)
    shard.seek(first - halo)
    sl = slice(first - halo, first)
    shard.context(first - halo, hdr[sl], tok[base[first - halo]:base[first]], base[first - halo:first + 1] - base[first - halo], byts[sl])
    sl = slice(first, F)
    got, cam = shard.process(first, hdr[sl], tok[base[first]:base[F]], base[first:F + 1] - base[first], byts[sl])
    assert np.array_equal(got, want[first:]) and np.array_equal(cam, wcam[first:])
    ref.close(); shard.close()


def test_fast_version_rejects_unsupported_sizes(cds):
    with pytest.raises(cds.CdsError):
        cds.SaliencyClip(424, 240, 30.0, fast=True)          # width not a multiple of 16


def test_reference_decoder_with_cds_hook_reproduces_its_own_golden_pngs(cds, tmp_path):
    """The reference's HM-16.0 decoder rebuilt with include/cds_hm_hook.h in place of its OpenCV block (oracle/_ref/hm_cds_decoder),
    CDS_MODE=fast, decodes HEVCfiles/BasketballDrill/str.bin: the first 39 maps it hands back are the authors' own
    bin/HMdata/0..38.png (36 byte-identical, 3 with one pixel off by one), and all 500 equal the library fed from the fixture."""
    import subprocess
    dec = os.path.join(o.ROOT, "oracle", "_ref", "hm_cds_decoder")
    stream = os.path.join(o.ROOT, "oracle", "_ref", "streams", "BasketballDrill.bin")
    if not (os.path.exists(dec) and os.path.exists(stream)):
        pytest.skip("hooked HM decoder / bitstream not built (needs the reference tree at build time)")
    d = np.load(os.path.join(GOLD, "fast_BasketballDrill.npz"))
    w, h = int(d["w"]), int(d["h"])
    out = str(tmp_path / "maps.u8")
    env = dict(os.environ, CDS_LIB=cds.lib_path(), CDS_OUT=out, CDS_FPS="50", CDS_MODE="fast", CDS_BATCH="32")
    r = subprocess.run([dec, "-b", stream], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "[cds hook] 500 pictures" in r.stderr, r.stderr[-2000:]
    print(r.stderr.strip().splitlines()[-1])
    maps = np.fromfile(out, np.uint8).reshape(-1, h, w)
    # the parse-only producer (no reconstruction / deblocking / SAO) must hand over the same syntax, hence the same maps
    out2 = str(tmp_path / "maps_po.u8")
    r2 = subprocess.run([dec, "-b", stream], env=dict(env, CDS_OUT=out2, CDS_PARSE_ONLY="1"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r2.returncode == 0, r2.stderr[-2000:]
    print("parse-only:", r2.stderr.strip().splitlines()[-1])
    assert np.array_equal(np.fromfile(out2, np.uint8).reshape(-1, h, w), maps)
    assert maps.shape[0] == 500
    n = d["png"].shape[0]
    per = (maps[:n] != d["png"]).reshape(n, -1).sum(1)
    diff = np.abs(maps[:n].astype(np.int16) - d["png"].astype(np.int16))
    print(f"decoder + hook vs bin/HMdata/*.png: {int((per == 0).sum())} of {n} byte-identical, {int(per.sum())} pixels differ (max {int(diff.max())})")
    assert diff.max() <= 1 and per.sum() <= 3 and (per == 0).sum() >= 36
    clip = cds.SaliencyClip(w, h, 50.0, fast=True, out_u8=True, max_batch=13)
    got, _ = clip.process(0, d["hdr"], d["tok"], d["tok_base"], d["ctu_bytes"])
    clip.close()
    assert np.array_equal(maps[:n], got)


# ---------------------------------------------------------------------------------------------------------------------------------
# The sequential-float-sum emulation on adversarial data: ties on the running sum's ulp grid, binade crossings, mixed signs, terms
# that vanish against the sum, terms larger than the sum.  Checker: the plain loops in oracle/cds_oracle_fast.c.
def _seq_ref(v, sel):
    import ctypes as C
    L = o.oracle(); L.orc_seq_sum_f32.restype = C.c_float
    L.orc_seq_sum_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
    cnt = C.c_int(0)
    return np.float32(L.orc_seq_sum_f32(v.ctypes.data, sel.ctypes.data, v.size, C.byref(cnt))), cnt.value


def _seq_gpu(cds, v, sel):
    import ctypes as C
    s = C.c_float(0); cnt = C.c_int(0)
    cds._lib.check(cds.lib().cds_probe_seq_sum(v.ctypes.data_as(C.c_void_p), sel.ctypes.data_as(C.c_void_p), v.size, C.byref(s), C.byref(cnt)), "probe")
    return np.float32(s.value), cnt.value


SEQ_CASES = {
    "uniform01": lambda r, n: r.random(n, dtype=np.float32),
    "dyadic_ties": lambda r, n: (r.integers(1, 64, n) * 0.125).astype(np.float32),              # every term a tie candidate once ulp(S) = 0.25
    "quarter_pel_mixed_sign": lambda r, n: (r.integers(-4000, 4001, n) * 0.25).astype(np.float32),
    "large_then_tiny": lambda r, n: np.concatenate([np.full(64, 3.0e6, np.float32), r.random(n - 64, dtype=np.float32) * 0.2]),
    "tiny_then_large": lambda r, n: np.concatenate([r.random(n - 64, dtype=np.float32) * 1e-3, np.full(64, 7.7e5, np.float32)]),
    "constant_1000.25": lambda r, n: np.full(n, 1000.25, np.float32),
    "wide_dynamic_range": lambda r, n: np.exp(r.uniform(-20, 12, n)).astype(np.float32),
    "cancelling": lambda r, n: (np.where(np.arange(n) % 2 == 0, 1.0, -1.0) * r.random(n)).astype(np.float32),
    "negative_only": lambda r, n: -r.random(n, dtype=np.float32) * 37,
    "zeros_and_ones": lambda r, n: r.integers(0, 2, n).astype(np.float32),
}


@pytest.mark.parametrize("case", sorted(SEQ_CASES))
def test_sequential_float_sum_emulation_is_bit_exact(cds, case):
    r = np.random.default_rng(abs(hash(case)) % 1000)
    for n, dens in ((4096, 1.0), (129600, 0.3), (400000, 0.9), (518400, 0.02)):
        v = np.ascontiguousarray(SEQ_CASES[case](r, n)[:n], np.float32)
        sel = (r.random(n) < dens).astype(np.uint8)
        want, wc = _seq_ref(v, sel)
        got, gc = _seq_gpu(cds, v, sel)
        assert gc == wc
        assert got.tobytes() == want.tobytes(), (case, n, dens, float(got), float(want))


@pytest.mark.parametrize("seed", range(6))
def test_replicated_block_sum_emulation_is_bit_exact(cds, seed):
    import ctypes as C
    L = o.oracle(); L.orc_run_sum_f32.restype = C.c_float
    L.orc_run_sum_f32.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_double, C.c_void_p]
    r = np.random.default_rng(seed)
    for (rows, cols, ms) in ((1080, 1920, 32), (1080, 1920, 64), (270, 480, 16), (480, 832, 32), (2160, 3840, 32)):
        gh, gw = -(-rows // ms), -(-cols // ms)
        kinds = [r.random((gh, gw), dtype=np.float32),
                 (r.integers(0, 17, (gh, gw)) * 0.0625).astype(np.float32),                       # ties once ulp(S) = 0.125
                 np.where(r.random((gh, gw)) < 0.9, np.float32(1.0), r.random((gh, gw), dtype=np.float32)).astype(np.float32),
                 np.full((gh, gw), 0.5625, np.float32)]
        for vals in kinds:
            vals = np.ascontiguousarray(vals, np.float32)
            for t in (-1.0, 0.3, float(vals.max()) - float(vals.mean())):
                wc = C.c_int(0)
                want = np.float32(L.orc_run_sum_f32(vals.ctypes.data, gh, gw, rows, cols, ms, t, C.byref(wc)))
                s = C.c_float(0); gc = C.c_int(0)
                cds._lib.check(cds.lib().cds_probe_run_sum(vals.ctypes.data_as(C.c_void_p), gh, gw, rows, cols, ms, C.c_double(t), C.byref(s), C.byref(gc)), "probe")
                assert gc.value == wc.value
                assert np.float32(s.value).tobytes() == want.tobytes(), (rows, cols, ms, t, float(s.value), float(want))


def _oracle_maps(w, h, fps, mvx, mvy, dep, byts):
    fo = o.FastOracle(w, h, fps)
    res = [fo.frame(mvx[f], mvy[f], dep[f], byts[f]) for f in range(mvx.shape[0])]
    fo.close()
    return np.stack([x[0] for x in res]), np.stack([x[1] for x in res]), np.stack([x[2] for x in res])


@pytest.mark.parametrize("case", ["huge_mv", "uniform_pan_ties", "static_zero", "all_intra_flat", "sparse_extremes"])
def test_fast_version_from_maps_on_adversarial_pictures(cds, case):
    """cds_clip_process_maps (stage-(a) maps in) on inputs no bitstream fixture reaches: motion so large that Umean's first loop
    is itself inexact (sum of |mv| >= 2^24 quarter-pels), a uniform pan whose terms tie on the sum's grid, static and flat pictures
    (every MatrixNorm / MapThreshold takes its copy branch), maps with single extreme cells."""
    w, h, fps, F = (1920, 1088, 50.0, 5) if case in ("huge_mv", "uniform_pan_ties") else (416, 240, 30.0, 14)
    h4, w4, nctu = h // 4, w // 4, ((w + 63) // 64) * ((h + 63) // 64)
    r = np.random.default_rng(42)
    if case == "huge_mv":
        mvx = r.integers(-30000, 30001, (F, h4, w4)).astype(np.int16); mvy = r.integers(-30000, 30001, (F, h4, w4)).astype(np.int16)
        mvx[:, :, : w4 // 2] = 28000 + (mvx[:, :, : w4 // 2] % 7)                          # a dominant direction bin with a huge mean
        dep = r.integers(0, 4, (F, h4, w4)).astype(np.uint8); byts = r.integers(0, 300, (F, nctu)).astype(np.uint16)
    elif case == "uniform_pan_ties":
        mvx = np.full((F, h4, w4), 4001, np.int16); mvy = np.full((F, h4, w4), -4001, np.int16)   # 1000.25 px: ties once ulp(S) = 0.5
        mvx[:, ::7, ::5] += 8
        dep = np.full((F, h4, w4), 2, np.uint8); byts = np.full((F, nctu), 57, np.uint16)
    elif case == "static_zero":
        mvx = np.zeros((F, h4, w4), np.int16); mvy = np.zeros((F, h4, w4), np.int16)
        dep = r.integers(0, 4, (F, h4, w4)).astype(np.uint8); byts = r.integers(0, 4095, (F, nctu)).astype(np.uint16)
    elif case == "all_intra_flat":
        mvx = np.zeros((F, h4, w4), np.int16); mvy = np.zeros((F, h4, w4), np.int16)
        dep = np.zeros((F, h4, w4), np.uint8); byts = np.zeros((F, nctu), np.uint16)
    else:
        mvx = np.zeros((F, h4, w4), np.int16); mvy = np.zeros((F, h4, w4), np.int16)
        mvx[:, 5, 7] = 32767; mvy[:, 50, 90] = -32768
        dep = np.zeros((F, h4, w4), np.uint8); dep[:, 0, 0] = 3
        byts = np.full((F, nctu), 10, np.uint16); byts[:, 3] = 4095
    want_u8, want_f32, want_cam = _oracle_maps(w, h, fps, mvx, mvy, dep, byts)
    clip = cds.SaliencyClip(w, h, fps, fast=True, out_u8=False, max_batch=4)
    got, cam = clip.process_maps(0, mvx, mvy, dep, byts)
    clip.close()
    assert np.array_equal(cam, want_cam), (cam, want_cam)
    assert got.tobytes() == want_f32.tobytes(), float(np.nanmax(np.abs(got - want_f32)))
    if case in ("huge_mv", "uniform_pan_ties"):
        assert np.abs(want_cam).sum() > 0


def test_process_maps_equals_process_from_syntax_in_the_svm_mode(cds):
    """cds_clip_process_maps in the nine-feature + SVM mode: feeding the stage-(a) maps gives the bytes the syntax path gives."""
    w, h, fps, F = 416, 240, 24.0, 9
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, 3, F)
    SV, coef, rho, gamma, labels = o.synth_model(64)
    gm = cds.ops.SvmModel(SVs=SV * 0.5 + 1.0, sv_coef=coef, rho=rho, gamma=gamma, Label=labels)
    a = cds.SaliencyClip(w, h, fps, model=gm, meanVec=np.full(9, 0.3), stdVec=np.full(9, 0.2), max_batch=4)
    want, wcam = a.process(0, hdr, tok, base, byts)
    a.close()
    maps = [cds.ops.getframe(w, h, hdr[f], tok[base[f]:base[f + 1]]) for f in range(F)]
    b = cds.SaliencyClip(w, h, fps, model=gm, meanVec=np.full(9, 0.3), stdVec=np.full(9, 0.2), max_batch=4)
    got, cam = b.process_maps(0, np.stack([m[0] for m in maps]), np.stack([m[1] for m in maps]), np.stack([m[2] for m in maps]), byts)
    b.close()
    assert np.array_equal(cam, wcam) and got.tobytes() == want.tobytes()


def test_fast_version_bit_exact_on_every_bundled_geometry(cds):
    """Two pictures of each of the 24 bundled streams the reference handles (416x240, 832x480, 1024x768, 1280x800, 1920x1080;
    tests/golden/all_streams.npz): 8-bit maps and camera votes identical to the oracle."""
    d = np.load(os.path.join(GOLD, "all_streams.npz"))
    seen = set()
    for k, name in enumerate(d["names"]):
        w, h = int(d[f"w{k}"]), int(d[f"h{k}"])
        hdr, tok, base, byts = d[f"hdr{k}"], d[f"tok{k}"], d[f"base{k}"], d[f"bytes{k}"]
        want, _, wcam = _oracle_clip(w, h, 30.0, hdr, tok, base, byts)
        clip = cds.SaliencyClip(w, h, 30.0, fast=True, out_u8=True, max_batch=2)
        got, cam = clip.process(0, hdr, tok, base, byts)
        clip.close()
        assert np.array_equal(cam, wcam) and np.array_equal(got, want), str(name)
        seen.add((w, h))
    assert len(seen) >= 4


@pytest.mark.parametrize("w,h,F", [(2560, 1600, 3), (3840, 2160, 2)])
def test_fast_version_bit_exact_at_large_sizes(cds, w, h, F):
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, 17, F)
    want_u8, want_f32, want_cam = _oracle_clip(w, h, 50.0, hdr, tok, base, byts)
    clip = cds.SaliencyClip(w, h, 50.0, fast=True, out_u8=False, max_batch=2)
    got, cam = clip.process(0, hdr, tok, base, byts)
    clip.close()
    assert np.array_equal(cam, want_cam) and got.tobytes() == want_f32.tobytes()


def test_feeder_many_streams_one_gpu_process(cds):
    """parse-only decoder PROCESSES (the reference's hooked HM, packed syntax on a pipe) feed one GPU-owning process.
    Two streams of different sizes decode concurrently; the maps equal the direct path's and the reference's golden PNGs."""
    dumper = os.path.join(o.ROOT, "oracle", "_ref", "hm_syntax_dump")
    sdir = os.path.join(o.ROOT, "oracle", "_ref", "streams")
    if not (os.path.exists(dumper) and os.path.exists(os.path.join(sdir, "BasketballDrill.bin"))):
        pytest.skip("hooked HM decoder / bitstreams not built (needs the reference tree at build time)")
    res = cds.feeder.run_streams(dumper, [(os.path.join(sdir, "BasketballDrill.bin"), 832, 480), (os.path.join(sdir, "BasketballPass.bin"), 416, 240)],
                                 fps=50.0, batch=32, collect=True, fast=True, out_u8=True)
    assert res["per_stream"] == [500, 500]
    print(f"feeder: {res['pictures']} pictures of 2 streams in {res['seconds']:.2f} s")
    d = np.load(os.path.join(GOLD, "fast_BasketballDrill.npz"))
    n = d["png"].shape[0]
    per = (res["maps"][0][:n] != d["png"]).reshape(n, -1).sum(1)
    assert per.sum() <= 3 and (per == 0).sum() >= 36
    clip = cds.SaliencyClip(832, 480, 50.0, fast=True, out_u8=True, max_batch=32)
    got, cam = clip.process(0, d["hdr"], d["tok"], d["tok_base"], d["ctu_bytes"])
    clip.close()
    assert np.array_equal(res["maps"][0][:n], got) and np.array_equal(res["cams"][0][:n], cam)
    p = np.load(os.path.join(GOLD, "real_BasketballPass.npz"))
    clip = cds.SaliencyClip(416, 240, 50.0, fast=True, out_u8=True, max_batch=32)
    got, _ = clip.process(0, p["hdr"], p["tok"], p["tok_base"], p["ctu_bytes"])
    clip.close()
    assert np.array_equal(res["maps"][1][:got.shape[0]], got)


def test_cpp_feeder_process_matches_the_direct_path(cds, tmp_path):
    """SURVEY 8(f).3 at box scale: the C++ feeder (one process per GPU; csrc_host/cds_feeder.cpp) spawns parse-only decoder processes,
    reads their packed syntax straight into the library's pinned staging buffers and drains the maps asynchronously.  Two streams of
    different sizes, four producers: every video's maps equal the direct path's bytes (and through it the golden PNGs)."""
    import json, subprocess
    feeder = os.path.join(os.path.dirname(cds.lib_path()), "cds_feeder")
    dumper = os.path.join(o.ROOT, "oracle", "_ref", "hm_syntax_dump")
    sdir = os.path.join(o.ROOT, "oracle", "_ref", "streams")
    if not (os.path.exists(feeder) and os.path.exists(dumper) and os.path.exists(os.path.join(sdir, "BasketballDrill.bin"))):
        pytest.skip("cds_feeder / hooked HM decoder / bitstreams not built (needs the reference tree at build time)")
    drill, bpass = os.path.join(sdir, "BasketballDrill.bin"), os.path.join(sdir, "BasketballPass.bin")
    pre = str(tmp_path) + "/"
    r = subprocess.run([feeder, "--producer", dumper, "--fast", "--fps", "50", "--batch", "32", "--out-prefix", pre,
                        drill + ":832x480", bpass + ":416x240"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert [v["pictures"] for v in res["videos"]] == [500, 500] and all(v["error"] == "" for v in res["videos"])
    print(f"cds_feeder: {res['pictures']} pictures of 2 streams in {res['seconds']:.2f} s = {res['pictures_per_s']:.0f} pictures/s")
    maps = np.fromfile(pre + "BasketballDrill.bin.u8", np.uint8).reshape(500, 480, 832)
    d = np.load(os.path.join(GOLD, "fast_BasketballDrill.npz"))
    n = d["png"].shape[0]
    clip = cds.SaliencyClip(832, 480, 50.0, fast=True, out_u8=True, max_batch=32)
    got, _ = clip.process(0, d["hdr"], d["tok"], d["tok_base"], d["ctu_bytes"])
    clip.close()
    assert np.array_equal(maps[:n], got)
    per = (maps[:n] != d["png"]).reshape(n, -1).sum(1)
    assert per.sum() <= 3 and (per == 0).sum() >= 36
    maps2 = np.fromfile(pre + "BasketballPass.bin.u8", np.uint8).reshape(500, 240, 416)
    p = np.load(os.path.join(GOLD, "real_BasketballPass.npz"))
    clip = cds.SaliencyClip(416, 240, 50.0, fast=True, out_u8=True, max_batch=32)
    got, _ = clip.process(0, p["hdr"], p["tok"], p["tok_base"], p["ctu_bytes"])
    clip.close()
    assert np.array_equal(maps2[:got.shape[0]], got)
    # the same two streams twice over (four producers): same checksums per video
    r2 = subprocess.run([feeder, "--producer", dumper, "--fast", "--fps", "50", "--batch", "32",
                         drill + ":832x480", bpass + ":416x240", drill + ":832x480", bpass + ":416x240"], capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stderr[-2000:]
    res2 = json.loads(r2.stdout.strip().splitlines()[-1])
    cs = [v["checksum"] for v in res2["videos"]]
    assert cs[0] == cs[2] == res["videos"][0]["checksum"] and cs[1] == cs[3] == res["videos"][1]["checksum"]
    print(f"cds_feeder: {res2['pictures']} pictures of 4 streams in {res2['seconds']:.2f} s = {res2['pictures_per_s']:.0f} pictures/s")


def test_fast_version_random_geometries_rates_and_batches(cds):
    """A seeded sweep over picture sizes (widths k*16, heights k*8 from 64 up), frame rates (PS / PT from cvRound, incl. 29.97 and 25 whose
    halves round to even) and batch sizes: every case bit-identical to the oracle (tools/fast_fuzz.py runs larger sweeps)."""
    r = np.random.default_rng(20)
    for i in range(24):
        w = int(r.integers(4, 60)) * 16; h = int(r.integers(8, 90)) * 8
        fps = float(r.choice([7.0, 10.0, 15.0, 24.0, 25.0, 29.97, 30.0, 50.0, 60.0]))
        F = int(r.integers(1, 12)); batch = int(r.integers(1, 9)); seed = int(r.integers(0, 10000))
        hdr, tok, base, byts = cds.ops.synth_clip(w, h, seed, F)
        want_u8, want_f32, want_cam = _oracle_clip(w, h, fps, hdr, tok, base, byts)
        clip = cds.SaliencyClip(w, h, fps, fast=True, out_u8=False, max_batch=batch)
        got, cam = clip.process(0, hdr, tok, base, byts)
        clip.close()
        assert np.array_equal(cam, want_cam) and got.tobytes() == want_f32.tobytes(), (i, w, h, fps, F, batch, seed)


@pytest.mark.parametrize("fps", [120.0, 240.0])
def test_fast_version_high_frame_rates(cds, fps):
    """PS = cvRound(0.3 fps) = 36 / 72: the smoothing tile needs more than the default 48 KB of dynamic shared memory."""
    w, h, F = 208, 120, 9
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, 12, F)
    want_u8, want_f32, want_cam = _oracle_clip(w, h, fps, hdr, tok, base, byts)
    clip = cds.SaliencyClip(w, h, fps, fast=True, out_u8=False, max_batch=4)
    assert clip.PS == int(np.rint(0.3 * fps))
    got, cam = clip.process(0, hdr, tok, base, byts)
    clip.close()
    assert np.array_equal(cam, want_cam) and got.tobytes() == want_f32.tobytes()
