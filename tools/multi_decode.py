#!/usr/bin/env python3
"""Bitstream -> saliency maps end to end with several decoder processes feeding one GPU (second half).

Each process is the reference's HM-16.0 decoder rebuilt with include/cds_hm_hook.h (oracle/_ref/hm_cds_decoder): CABAC parse on a host
core, hooks -> packed syntax -> libcds_b200.so (its own context on the shared GPU) -> maps.  Prints pictures/s over all processes.
usage: multi_decode.py [processes] [stream] [mode fast|svm] [parse_only 0|1]"""
import os, subprocess, sys, tempfile, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import compressd_domain_saliencyprediction_b200 as cds  # noqa: E402


def main():
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    stream = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "oracle", "_ref", "streams", "BasketballDrill.bin")
    mode = sys.argv[3] if len(sys.argv) > 3 else "fast"
    parse_only = sys.argv[4] if len(sys.argv) > 4 else "1"
    dec = os.path.join(ROOT, "oracle", "_ref", "hm_cds_decoder")
    env = dict(os.environ, CDS_LIB=cds.lib_path(), CDS_FPS="50", CDS_BATCH="32", CDS_PARSE_ONLY=parse_only, CDS_U8="1")
    tmp = tempfile.mkdtemp()
    if mode == "fast":
        env["CDS_MODE"] = "fast"
    else:
        import _oracle as o
        SV, coef, rho, gamma, labels = o.synth_model(4096)
        path = os.path.join(tmp, "standin.model")
        npos = int(np.sum(coef > 0))
        with open(path, "w") as f:
            f.write(f"svm_type c_svc\nkernel_type rbf\ngamma {gamma:.17g}\nnr_class 2\ntotal_sv {len(coef)}\nrho {rho:.17g}\nlabel 1 -1\nnr_sv {npos} {len(coef) - npos}\nSV\n")
            for c, s in zip(coef, SV * 0.5 + 1.0):
                f.write(f"{c:.17g} " + " ".join(f"{i + 1}:{v:.17g}" for i, v in enumerate(s)) + " \n")
        env.update(CDS_SVM_MODEL=path, CDS_MEAN=",".join(["0.3"] * 9), CDS_STD=",".join(["0.2"] * 9))
    t0 = time.perf_counter()
    procs = [subprocess.Popen([dec, "-b", stream], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True) for _ in range(nproc)]
    pics = 0; lines = []
    for p in procs:
        err = p.communicate()[1]
        last = [l for l in err.splitlines() if l.startswith("[cds hook]")]
        assert p.returncode == 0 and last, err[-1000:]
        pics += int(last[-1].split()[2]); lines.append(last[-1])
    wall = time.perf_counter() - t0
    print(lines[0])
    print(f"{nproc} decoder processes x {os.path.basename(stream)} ({mode}, parse_only={parse_only}): {pics} pictures in {wall:.2f} s = {pics / wall:.0f} pictures/s "
          f"(includes process start-up and CUDA context creation; host cores: {os.cpu_count()})")


if __name__ == "__main__":
    main()
