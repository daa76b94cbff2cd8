#!/usr/bin/env python3
"""Write profiles/r02_sass_<kernel>.txt: for the tcgen05 / TMA kernels of libcds_b200.so, the counts of the SASS mnemonics that prove
the tensor-core / TMEM / TMA path (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, SYNCS = mbarrier; FADD2 / FFMA2 = the packed FP32 forms add.f32x2 / fma.f32x2) and the first occurrences with their addresses.  CPU only (cuobjdump)."""
import collections, os, re, subprocess, sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIB = os.path.join(ROOT, "compressd_domain_saliencyprediction_b200", "libcds_b200.so")
KERNELS = {"svm_tc": "svm_rbf9_tcp_kernelILin1", "contrast_packed": "contrast_sep_kernelILi48ELi8ELi6ELb1", "svm_tc_2cta": "svm_rbf9_tc_kernelILi2", "vpass_tc": "fullres_vpass_tc_kernel", "temporal_tma": "temporal_group_kernel"}
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|UTCATOM|SYNCS|MUFU\.EX2|MUFU\.SQRT|FFMA2|FADD2|LDGSTS)\b")


def main():
    names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.findall(r"Function : (\S+)", names)
    for tag, key in KERNELS.items():
        fn = next((f for f in funcs if key in f), None)
        if fn is None:
            print("missing", key); continue
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, LIB], capture_output=True, text=True).stdout
        lines = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
        cnt = collections.Counter(); first = {}
        for l in lines:
            m = PAT.search(l)
            if m:
                cnt[m.group(1)] += 1
                first.setdefault(m.group(1), []).append(l.strip()[:120])
        out = os.path.join(ROOT, "profiles", f"r02_sass_{tag}.txt")
        with open(out, "w") as f:
            f.write(f"# cuobjdump -sass -fun {fn} libcds_b200.so  ({len(lines)} instructions)\n# mnemonic counts: " +
                    ", ".join(f"{k} {v}" for k, v in sorted(cnt.items())) + "\n")
            for k in sorted(first):
                f.write(f"\n## {k} ({cnt[k]})\n" + "\n".join(first[k][:6]) + "\n")
        print(out, dict(cnt))


if __name__ == "__main__":
    sys.exit(main())
