#!/usr/bin/env python3
"""Bitstreams -> maps on every GPU of the box: one cds_feeder PROCESS per GPU (CUDA_VISIBLE_DEVICES), each fed by its
share of parse-only decoder processes.  Videos are independent, so the processes never talk to each other.

    python tools/feeder_box.py [--gpus N] [--copies K] [--mode fast|svm] [stream.bin:WxH ...]

Default streams: the bundled oracle/_ref/streams/*.bin (K copies of each per GPU).  Prints the per-GPU lines and the box total;
the host producer (HM parse) rate is printed beside it: it is the limit, not the GPUs."""
import json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "compressd_domain_saliencyprediction_b200")


def main():
    a = sys.argv[1:]
    def opt(name, default):
        if name in a:
            i = a.index(name); v = a[i + 1]; del a[i:i + 2]; return v
        return default
    gpus = int(opt("--gpus", "1")); copies = int(opt("--copies", "4")); mode = opt("--mode", "fast")
    sdir = os.path.join(ROOT, "oracle", "_ref", "streams")
    streams = a or [os.path.join(sdir, "BasketballDrill.bin") + ":832x480", os.path.join(sdir, "BasketballPass.bin") + ":416x240"]
    cmd = [os.path.join(PKG, "cds_feeder"), "--producer", os.path.join(ROOT, "oracle", "_ref", "hm_syntax_dump"), "--fps", "50", "--batch", "32"]
    if mode == "fast":
        cmd += ["--fast"]
    else:
        cmd += ["--svm-model", os.path.join(ROOT, "tests", "golden", "standin_svmRBFmodel.model"), "--mean", ",".join(["0.3"] * 9), "--std", ",".join(["0.2"] * 9)]
    t0 = time.time()
    procs = [subprocess.Popen(cmd + streams * copies, env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(g)), stdout=subprocess.PIPE, text=True) for g in range(gpus)]
    total = 0; steady = 0.0; startup = []
    for g, p in enumerate(procs):
        out, _ = p.communicate()
        r = json.loads(out.strip().splitlines()[-1])
        total += r["pictures"]; steady += r.get("steady_pictures_per_s", 0.0); startup.append(r.get("startup_seconds", 0.0))
        print(f"gpu {g}: {r['pictures']} pictures of {len(r['videos'])} streams in {r['seconds']:.2f} s = {r['pictures_per_s']:.0f} pictures/s"
              f" (start-up {r.get('startup_seconds', 0.0):.2f} s: CUDA, context, decoder; then {r.get('steady_pictures_per_s', 0.0):.0f} pictures/s)"
              + ("".join(" ERROR " + v["error"] for v in r["videos"] if v["error"])))
    sec = time.time() - t0
    print(json.dumps({"gpus": gpus, "producers": gpus * len(streams) * copies, "mode": mode, "pictures": total, "seconds": round(sec, 3),
                      "pictures_per_s": round(total / sec, 1), "steady_pictures_per_s": round(steady, 1), "startup_seconds_max": round(max(startup), 2),
                      "host_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
