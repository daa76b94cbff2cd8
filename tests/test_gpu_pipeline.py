"""GPU parity tests for the whole hot path (main.m:77-305): packed syntax -> saliency maps through
cds_clip_process / the hook-style cds_submit_frame API, checked against oracle/cds_oracle.c
(orc_process_clip) on the same seeded synthetic clips.  Tolerance: 1e-4 relative to the map range
(maps are pm_norm'ed to [0, 1]; north_star: "saliency maps must be within 1e-4 relative")."""
import numpy as np
import pytest

import _oracle as o

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4
MEAN9 = np.full(9, 0.3)
STD9 = np.full(9, 0.2)


@pytest.fixture(scope="module")
def cds():
    import compressd_domain_saliencyprediction_b200 as pkg
    assert pkg.device_ok(), pkg.lib().cds_last_error()
    return pkg


def _clip(cds, w, h, seed, F):
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, seed, F)
    maps = [o.orc_getframe(w, h, hdr[f], tok[base[f]:base[f + 1]]) for f in range(F)]
    mvx = np.stack([m[0] for m in maps]); mvy = np.stack([m[1] for m in maps]); dep = np.stack([m[2] for m in maps])
    return (hdr, tok, base, byts), (mvx, mvy, dep, byts)


def _model(cds, nsv):
    SV, coef, rho, gamma, labels = o.synth_model(nsv)
    SV = SV * 0.5 + 1.0           # the z-scored features live around (f - 0.3) / 0.2: keep the kernel responsive there
    return (SV, coef, rho, gamma, labels), cds.ops.SvmModel(SVs=SV, sv_coef=coef, rho=rho, gamma=gamma, Label=labels)


@pytest.mark.parametrize("cfg", [
    dict(w=416, h=240, fps=10.0, F=26, first=0, rbf=1, batch=8),     # PS 3, PT 21: full temporal window reached
    dict(w=416, h=240, fps=10.0, F=10, first=0, rbf=0, batch=4),     # linear fusion branch (main.m:258-272)
    dict(w=200, h=136, fps=7.0, F=9, first=0, rbf=1, batch=32),      # ragged CTU edge (136 = 2*64 + 8), tiny batch remainder
    dict(w=1280, h=720, fps=10.0, F=5, first=0, rbf=1, batch=4),     # 720p (720 = 11*64 + 16): the geometry of the bundled Vidyo / FourPeople clips,
                                                                     # which the reference's own getFrame mishandles (its CTU-row count)
])
def test_clip_maps_vs_oracle(cds, cfg):
    w, h, F = cfg["w"], cfg["h"], cfg["F"]
    packed, ints = _clip(cds, w, h, 1234, F)
    mdl, gm = _model(cds, 300)
    want, wcam, wfeat = o.orc_process_clip(w, h, cfg["fps"], *ints, mdl, MEAN9, STD9, first_out=cfg["first"], use_rbf=cfg["rbf"],
                                           want_feat=True)
    clip = cds.SaliencyClip(w, h, cfg["fps"], model=gm, meanVec=MEAN9, stdVec=STD9, use_rbf=bool(cfg["rbf"]), max_batch=cfg["batch"])
    got, cam = clip.process(0, *packed)
    assert np.array_equal(cam, wcam)
    assert got.shape == want.shape
    assert np.all(np.isfinite(got))
    worst = 0.0
    for f in range(F):
        wf = np.nan_to_num(want[f], nan=0.0)     # a flat final map is NaN in MATLAB (pm_norm 0/0); the GPU path writes 0
        err = np.max(np.abs(got[f] - wf))
        worst = max(worst, err)
        assert err <= REL_TOL, (f, err)
    print("worst abs error", worst)
    clip.close()


def test_clip_intermediate_features_vs_oracle(cds):
    """The nine 200x200 features (main.m:275-292) of the last picture of a batch."""
    w, h, F, fps = 416, 240, 12, 10.0
    packed, ints = _clip(cds, w, h, 77, F)
    mdl, gm = _model(cds, 64)
    want, wcam, wfeat = o.orc_process_clip(w, h, fps, *ints, mdl, MEAN9, STD9, use_rbf=1, want_feat=True)
    clip = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=F)
    clip.process(0, *packed)
    f = F - 1
    xs = clip.debug("xsvm", f, (40000, 9))
    feats = xs * STD9.astype(np.float32) + MEAN9.astype(np.float32)          # undo the z-score
    for i in range(9):
        wi = np.nan_to_num(wfeat[f, i], nan=0.0).reshape(-1)                   # row-major 200x200, like the GPU's pixel order
        err = np.max(np.abs(feats[:, i] - wi))
        assert err <= 2e-4, (i, err)                                           # features are in [0,1]; FP32 z-score round trip adds ~1e-7/0.2
    clip.close()


def test_hook_api_matches_batch_api(cds):
    w, h, F, fps = 416, 240, 11, 10.0
    packed, _ = _clip(cds, w, h, 5, F)
    hdr, tok, base, byts = packed
    mdl, gm = _model(cds, 64)
    a = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=4)
    ref_maps, ref_cam = a.process(0, *packed)
    b = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=4)
    for f in range(F):
        b.submit_frame(f, byts[f], hdr[f], tok[base[f]:base[f + 1]])
        if f % 4 == 3 or f == F - 1:                 # a batch completed (or the tail is flushed by get_map)
            lo = (f // 4) * 4
            for g in range(lo, f + 1):
                m, cam = b.get_map(g)
                assert np.array_equal(m, ref_maps[g]), g
                assert np.array_equal(cam, ref_cam[g])
    with pytest.raises(cds.CdsError):
        b.get_map(0)                                   # only the most recent batch stays retrievable
    with pytest.raises(cds.CdsError):
        b.submit_frame(F + 3, byts[0], hdr[0], tok[base[0]:base[1]])   # out of order
    a.close(); b.close()


def test_frame_range_shard_with_halo_is_bytewise_identical(cds):
    """a shard that seeks to first-(PS+PT-2), feeds the halo as context and then processes
    its own range must reproduce the single-context maps byte for byte."""
    w, h, F, fps = 416, 240, 40, 10.0          # PS 3, PT 21 -> halo 22
    packed, _ = _clip(cds, w, h, 9, F)
    hdr, tok, base, byts = packed
    mdl, gm = _model(cds, 64)
    one = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=16)
    full, fcam = one.process(0, *packed)
    first, halo = 30, 22
    s = first - halo
    sh = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=16)
    sh.seek(s)
    sh.context(s, hdr[s:first], tok, base[s:first + 1], byts[s:first])
    part, pcam = sh.process(first, hdr[first:], tok, base[first:], byts[first:])
    assert np.array_equal(part, full[first:])
    assert np.array_equal(pcam, fcam[first:])
    one.close(); sh.close()


def test_shard_planner_two_ranks_reproduce_single_context(cds):
    """plan_frame_ranges + process_frame_range for world = 2 (ranks run one after the other here) == one context."""
    w, h, F, fps = 416, 240, 33, 10.0
    packed, _ = _clip(cds, w, h, 21, F)
    hdr, tok, base, byts = packed
    mdl, gm = _model(cds, 64)
    one = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=8)
    full, fcam = one.process(0, *packed)
    one.close()
    plan = cds.shard.plan_frame_ranges(F, 2, cds.shard.halo_frames(fps))
    parts = []
    for p in plan:
        c = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=8)
        parts.append(cds.shard.process_frame_range(c, p, hdr, tok, base, byts))
        c.close()
    assert np.array_equal(np.concatenate([m for m, _ in parts]), full)
    assert np.array_equal(np.concatenate([c for _, c in parts]), fcam)


def test_u8_output(cds):
    w, h, F, fps = 416, 240, 5, 10.0
    packed, _ = _clip(cds, w, h, 3, F)
    mdl, gm = _model(cds, 32)
    a = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=8)
    f32, _ = a.process(0, *packed)
    b = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=8, out_u8=True)
    u8, _ = b.process(0, *packed)
    assert u8.dtype == np.uint8
    assert np.max(np.abs(u8.astype(np.float32) - np.rint(f32 * 255))) <= 1
    a.close(); b.close()


def test_malformed_stream_is_reported(cds):
    w, h = 416, 240
    packed, _ = _clip(cds, w, h, 3, 2)
    hdr, tok, base, byts = packed
    bad = tok.copy(); o0 = hdr[1, 0, 0] + base[1]; bad[o0:o0 + hdr[1, 0, 1]] = 99
    mdl, gm = _model(cds, 32)
    a = cds.SaliencyClip(w, h, 10.0, model=gm, meanVec=MEAN9, stdVec=STD9)
    with pytest.raises(cds.CdsError):
        a.process(0, hdr, bad, base, byts)
    a.close()


@pytest.mark.parametrize("bad_off", [-5, 10 ** 8, "past_picture"])
def test_corrupt_ctu_offset_is_reported(cds, bad_off):
    """A CTU header whose token offset is negative, past the end of the buffer or reaching into the next picture's
    tokens is a format error on every entry point (the kernel never reads outside the picture's own token slice)."""
    w, h = 416, 240
    packed, _ = _clip(cds, w, h, 3, 3)
    hdr, tok, base, byts = packed
    hb = hdr.copy()
    hb[1, 5, 0] = (base[2] - base[1]) - 2 if bad_off == "past_picture" else bad_off
    mdl, gm = _model(cds, 32)
    a = cds.SaliencyClip(w, h, 10.0, model=gm, meanVec=MEAN9, stdVec=STD9)
    with pytest.raises(cds.CdsError):
        a.process(0, hb, tok, base, byts)
    a.seek(0)
    got, _ = a.process(0, hdr, tok, base, byts)              # the context stays usable afterwards
    assert np.isfinite(got).all()
    a.close()


def test_svm_model_file_with_index_zero_is_rejected(cds, tmp_path):
    """libsvm indices are 1-based (svm.cpp:2641 svm_load_model); '0:' (precomputed-kernel syntax) or a negative index must
    be a format error, not a write before the SV row."""
    import os
    src = os.path.join(os.path.dirname(__file__), "golden", "svm_heart_scale.model")
    text = open(src).read()
    head, svs = text.split("SV\n", 1)
    lines = svs.splitlines()
    for repl in ("0:0.5", "-3:0.5"):
        bad = lines[:]
        first = bad[2].split()
        bad[2] = " ".join([first[0], repl] + first[2:])
        pth = tmp_path / "bad.model"
        pth.write_text(head + "SV\n" + "\n".join(bad) + "\n")
        with pytest.raises(cds.CdsError):
            cds.ops.SvmModel(path=str(pth))
    cds.ops.SvmModel(path=src)                                   # the untouched file still loads


def test_create_rejects_bad_parameters(cds):
    with pytest.raises(cds.CdsError):
        cds.SaliencyClip(417, 240, 25.0, use_rbf=False)
    with pytest.raises(cds.CdsError):
        cds.SaliencyClip(416, 240, 25.0, use_rbf=True, model=None)
    with pytest.raises(cds.CdsError):
        cds.SaliencyClip(416, 240, 25.0, use_rbf=False, stdVec=np.zeros(9))


@pytest.mark.parametrize("case", [(60, 104, 1, 11.0, 48), (60, 104, 2, 13.0, 48), (60, 104, 16, 3.0, 80), (33, 47, 4, 5.0, 40)])
def test_cu_contrast5_host_entry_vs_oracle(cds, case):
    rows, cols, cus, delta, patch = case
    a = np.random.default_rng(cus).uniform(0, 1, (rows, cols))
    got = cds.ops.cu_contrast5(a, cus, delta, patch)
    want = o.orc_cu_contrast5(a, cus, delta, patch)
    assert np.max(np.abs(got - want)) <= REL_TOL


def test_sudoku_host_entry_vs_oracle_signed_map(cds):
    a = np.random.default_rng(1).standard_normal((60, 104)) * 3
    got = cds.ops.Sudoku_contrastcpp5(a, 11.0, 1, 48)
    want = o.orc_sudoku(a, 11.0, 1, 48)
    assert np.max(np.abs(got - want)) <= REL_TOL


def test_random_geometries_rates_and_batches(cds):
    """A seeded sweep over picture sizes (multiples of 8 from 72 up), frame rates and batch sizes, RBF and linear fusion: camera votes
    exact, maps within 1e-4 (tools/main_fuzz.py runs larger sweeps)."""
    r = np.random.default_rng(3)
    mdl, gm = _model(cds, 200)
    o.oracle().orc_set_threads(16)
    worst = 0.0
    for i in range(12):
        w = int(r.integers(9, 90)) * 8; h = int(r.integers(9, 70)) * 8
        fps = float(r.choice([7.0, 10.0, 13.0, 24.0])); F = int(r.integers(2, 9)); batch = int(r.integers(1, 6)); seed = int(r.integers(0, 10000))
        rbf = int(r.integers(0, 4) > 0)
        packed, ints = _clip(cds, w, h, seed, F)
        want, wcam, _ = o.orc_process_clip(w, h, fps, *ints, mdl, MEAN9, STD9, first_out=0, use_rbf=rbf)
        clip = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, use_rbf=bool(rbf), max_batch=batch)
        got, cam = clip.process(0, *packed)
        clip.close()
        err = float(np.max(np.abs(got - np.nan_to_num(want))))
        worst = max(worst, err)
        assert np.array_equal(cam, wcam) and err <= REL_TOL, (i, w, h, fps, F, batch, rbf, seed, err)
    print("worst abs error over the sweep", worst)


def test_hook_zero_copy_acquire_and_early_drain(cds):
    """The decoder hook's own call sequence (include/cds_hm_hook.h): tokens written straight into the pinned buffers of
    cds_acquire_frame_buffers, cds_submit_frame with those pointers, maps of batch k-1 collected after batch k was launched.
    Same bytes as the batch entry point; a map leaves the result ring once two further batches were launched."""
    w, h, F, fps, B = 416, 240, 14, 10.0, 4
    packed, _ = _clip(cds, w, h, 8, F)
    hdr, tok, base, byts = packed
    mdl, gm = _model(cds, 64)
    a = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=B)
    ref_maps, ref_cam = a.process(0, *packed)
    a.close()
    b = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=B)
    drained = 0
    for f in range(F):
        pb, ph, pt = b.acquire_frame_buffers()
        n = int(base[f + 1] - base[f])
        assert pt.size >= n
        pb[:] = byts[f]; ph[:] = hdr[f]; pt[:n] = tok[base[f]:base[f + 1]]
        b.submit_frame_acquired(f, n)
        if f + 1 - drained >= 2 * B:                       # batch k was just launched: collect batch k-1
            for g in range(drained, drained + B):
                m, cam = b.get_map(g)
                assert np.array_equal(m, ref_maps[g]) and np.array_equal(cam, ref_cam[g]), g
            drained += B
    for g in range(drained, F):                            # the tail: get_map launches the partial batch itself
        m, cam = b.get_map(g)
        assert np.array_equal(m, ref_maps[g]) and np.array_equal(cam, ref_cam[g]), g
    with pytest.raises(cds.CdsError):
        b.get_map(1)                                       # batch 0 has left the two result batches
    b.close()


def test_hook_malformed_picture_is_reported_and_does_not_wedge_the_context(cds):
    """A corrupt token stream submitted through the hook surfaces as CDS_EFORMAT when its maps are collected; the context then
    accepts the next pictures in order (a failed flush must not leave pending_first behind)."""
    w, h, fps, B = 416, 240, 10.0, 4
    packed, _ = _clip(cds, w, h, 9, 8)
    hdr, tok, base, byts = packed
    mdl, gm = _model(cds, 32)
    c = cds.SaliencyClip(w, h, fps, model=gm, meanVec=MEAN9, stdVec=STD9, max_batch=B)
    for f in range(4):
        t = tok[base[f]:base[f + 1]].copy()
        if f == 2:
            t[hdr[f, 0, 0]:hdr[f, 0, 0] + hdr[f, 0, 1]] = 99       # an endless split chain in CTU 0
        c.submit_frame(f, byts[f], hdr[f], t)
    with pytest.raises(cds.CdsError):
        c.get_map(3)
    for f in range(4, 8):                                   # pictures 4..7 are still accepted in order
        c.submit_frame(f, byts[f], hdr[f], tok[base[f]:base[f + 1]])
    m, _ = c.get_map(7)
    assert np.isfinite(m).all()
    c.close()
