"""cds_clip_process_dev (device-resident syntax, several batches per call — what bench.py times as `value`): the library overlaps the
halves of consecutive batches on its own streams there (RBF mode: SVM -> map beside the next batch's front half; fast version: the camera
stages of batch k+1 beside the smoothing .. blur of batch k).  The bytes must equal the host-buffer entry point's, whatever the overlap."""
import ctypes as C

import numpy as np
import pytest

import _oracle as o

pytestmark = pytest.mark.gpu


def _dev_run(cds, clip, hdr, tok, base, byts, w, h, u8, calls):
    import torch
    from compressd_domain_saliencyprediction_b200._lib import check
    L = cds.lib(); nctu = cds.ops.nctu(w, h); dev = torch.device("cuda", 0)
    d_hdr, d_tok, d_base, d_byts = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (hdr, tok, base, byts))
    F = hdr.shape[0]
    out = torch.zeros((F, h, w), dtype=torch.uint8 if u8 else torch.float32, device=dev)
    cam = torch.zeros((F, 2), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    P = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
    px = h * w * (1 if u8 else 4)
    with torch.cuda.stream(stream):
        f = 0
        for n in calls:
            check(L.cds_clip_process_dev(clip.handle, f, n, P(d_byts, f * nctu * 2), P(d_hdr, f * nctu * 16), P(d_tok), P(d_base, f * 8),
                                         P(out, f * px), P(cam, f * 8), C.c_void_p(stream.cuda_stream)), "cds_clip_process_dev")
            f += n
        assert f == F
    stream.synchronize()
    return out.cpu().numpy(), cam.cpu().numpy()


@pytest.mark.parametrize("overlap", ["1", "0"])
def test_fast_version_device_entry_equals_host_entry(monkeypatch, overlap):
    import compressd_domain_saliencyprediction_b200 as cds
    monkeypatch.setenv("CDS_OVERLAP", overlap)
    w, h, fps, F = 416, 240, 50.0, 75
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, 21, F)
    host = cds.SaliencyClip(w, h, fps, fast=True, out_u8=True, max_batch=8)
    want, wcam = host.process(0, hdr, tok, base, byts)
    host.close()
    clip = cds.SaliencyClip(w, h, fps, fast=True, out_u8=True, max_batch=8)
    got, cam = _dev_run(cds, clip, hdr, tok, base, byts, w, h, True, [40, 3, 32])       # 5 batches, a single short one, 4 batches
    clip.close()
    assert np.array_equal(cam, wcam) and np.array_equal(got, want)
    assert np.abs(wcam).sum() > 0


def test_rbf_mode_device_entry_equals_host_entry():
    import compressd_domain_saliencyprediction_b200 as cds
    w, h, fps, F = 416, 240, 24.0, 30
    hdr, tok, base, byts = cds.ops.synth_clip(w, h, 8, F)
    SV, coef, rho, gamma, labels = o.synth_model(128)
    gm = cds.ops.SvmModel(SVs=SV * 0.5 + 1.0, sv_coef=coef, rho=rho, gamma=gamma, Label=labels)
    kw = dict(model=gm, meanVec=np.full(9, 0.3), stdVec=np.full(9, 0.2), max_batch=4)
    host = cds.SaliencyClip(w, h, fps, **kw)
    want, wcam = host.process(0, hdr, tok, base, byts)
    host.close()
    clip = cds.SaliencyClip(w, h, fps, **kw)
    got, cam = _dev_run(cds, clip, hdr, tok, base, byts, w, h, False, [18, 12])
    clip.close()
    assert np.array_equal(cam, wcam) and got.tobytes() == want.tobytes()
