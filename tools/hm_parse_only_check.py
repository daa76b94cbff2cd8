#!/usr/bin/env python3
"""Parse-only HM producer against the full decode, on every bundled bitstream (build container: needs /root/reference
and `make -C oracle ref`).  The hooked decoder with CDS_PARSE_ONLY=1 skips reconstruction, deblocking and SAO; the per-CTU syntax
its hooks capture must be byte-identical.  Prints seconds per 500-picture stream for both modes."""
import filecmp, glob, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP = os.path.join(ROOT, "oracle", "_ref", "hm_syntax_dump")


def run(stream, out, parse_only):
    env = dict(os.environ, HM_SYNTAX_DUMP=out)
    if parse_only:
        env["CDS_PARSE_ONLY"] = "1"
    t0 = time.perf_counter()
    r = subprocess.run([DUMP, "-b", stream], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0, r.returncode


def main():
    streams = sorted(glob.glob("/root/reference/HEVCfiles/*/str.bin")) if len(sys.argv) < 2 else sys.argv[1:]
    tot = [0.0, 0.0]; bad = 0
    for s in streams:
        name = os.path.basename(os.path.dirname(s))
        a, b = f"/tmp/po_full_{name}.bin", f"/tmp/po_parse_{name}.bin"
        tf, rf = run(s, a, False); tp, rp = run(s, b, True)
        same = rf == 0 and rp == 0 and filecmp.cmp(a, b, shallow=False)
        n = os.path.getsize(a) if os.path.exists(a) else 0
        print(f"{name:22s} full {tf:6.2f} s  parse-only {tp:6.2f} s  ({tf / max(tp, 1e-9):4.1f}x)  syntax {'identical' if same else 'DIFFERENT'} ({n} bytes)", flush=True)
        bad += not same; tot[0] += tf; tot[1] += tp
        for f in (a, b):
            if os.path.exists(f): os.remove(f)
    print(f"total: full {tot[0]:.1f} s, parse-only {tot[1]:.1f} s ({tot[0] / max(tot[1], 1e-9):.2f}x); {bad} streams differ")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
