#!/usr/bin/env python3
"""The reference's HM decoder with the cds hook on a bundled 1080p stream, parse-only: how much of the decoder thread's wall time
goes into the saliency hand-off + GPU wait (<= 10 %).  Runs both captures (direct int16 / legacy statics) and both
arithmetics (fast version, nine features + RBF-SVM).  GPU box only (oracle/_ref/hm_cds_decoder + streams travel with the snapshot)."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dec = os.path.join(ROOT, "oracle", "_ref", "hm_cds_decoder")
name = sys.argv[1] if len(sys.argv) > 1 else "BasketballDrive"
stream = os.path.join(ROOT, "oracle", "_ref", "streams", name + ".bin")
lib = os.path.join(ROOT, "compressd_domain_saliencyprediction_b200", "libcds_b200.so")
model = os.path.join(ROOT, "tests", "golden", "standin_svmRBFmodel.model")
for mode in ("fast", "svm"):
    for legacy in (False, True):
        for parse_only in (True, False):
            env = dict(os.environ, CDS_LIB=lib, CDS_FPS="50", CDS_BATCH="32", CDS_U8="1", CDS_PARSE_ONLY="1" if parse_only else "0")
            if mode == "fast":
                env["CDS_MODE"] = "fast"
            else:
                env.update(CDS_SVM_MODEL=model, CDS_MEAN=",".join(["0.3"] * 9), CDS_STD=",".join(["0.2"] * 9))
            if legacy:
                env["CDS_HOOK_LEGACY"] = "1"
            r = subprocess.run([dec, "-b", stream], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, cwd=tempfile.gettempdir())
            line = [l for l in r.stderr.splitlines() if l.startswith("[cds hook]")]
            print(f"{name} {mode:4s} capture={'legacy statics' if legacy else 'direct int16 '} {'parse-only ' if parse_only else 'full decode'}: " + (line[-1] if line else "FAILED " + r.stderr[-300:]))
