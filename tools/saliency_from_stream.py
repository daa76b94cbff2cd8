#!/usr/bin/env python3
"""Bitstream(s) -> saliency maps on the GPU, end to end.

    saliency_from_stream.py --out DIR [--mode fast|svm] [--model svmRBFmodel_all.mat] [--fps 50] WxH:stream.bin [WxH:stream2.bin ...]

Every stream is decoded by its own parse-only producer process (the reference's hooked HM-16.0 decoder, oracle/_ref/hm_syntax_dump, built
from the reference tree by oracle/hm_dump/build_hm_dump.py); this process owns the GPU (compressd_domain_saliencyprediction_b200.feeder)
and writes DIR/<stream name>/<POC>.png like the reference decoder writes ../HMdata/<POC>.png (TDecGop.cpp:327-332).
--mode fast: the decoder's own arithmetic (TDecGop.cpp + salmat.cpp);  --mode svm: nine features + RBF-SVM (main.m), needs --model."""
import argparse, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import compressd_domain_saliencyprediction_b200 as cds  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("streams", nargs="+", help="WxH:path")
    ap.add_argument("--out", required=True)
    ap.add_argument("--mode", default="fast", choices=["fast", "svm"])
    ap.add_argument("--model", help="svmRBFmodel_all.mat (model, meanVec, stdVec) for --mode svm")
    ap.add_argument("--fps", type=float, default=50.0)
    ap.add_argument("--producer", default=os.path.join(ROOT, "oracle", "_ref", "hm_syntax_dump"))
    a = ap.parse_args()
    specs = []
    for s in a.streams:
        geo, path = s.split(":", 1); w, h = map(int, geo.lower().split("x"))
        specs.append((path, w, h))
    if a.mode == "svm":
        if not a.model:
            ap.error("--mode svm needs --model")
        model, mean, std = cds.ops.SvmModel.from_mat(a.model)
        kw = dict(model=model, meanVec=mean, stdVec=std, out_u8=True)
    else:
        kw = dict(fast=True, out_u8=True)
    res = cds.feeder.run_streams(a.producer, specs, fps=a.fps, collect=True, **kw)
    for (path, w, h), maps in zip(specs, res["maps"]):
        d = os.path.join(a.out, os.path.splitext(os.path.basename(os.path.dirname(path) if os.path.basename(path).startswith("str.") else path))[0])
        cds.formats.write_png_per_poc(d, maps)
        print(f"{path}: {maps.shape[0]} maps -> {d}")
    print(f"{res['pictures']} pictures in {res['seconds']:.2f} s ({res['pictures_per_s']:.0f} pictures/s)")


if __name__ == "__main__":
    main()
