"""In-tree build of libcds_b200.so (sm_100a only) with nvcc.  No JIT cache, no multi-arch."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcds_b200.so")
SYNTH_LIB = os.path.join(HERE, "libcds_synth.so")       # host-only synthetic syntax producer (no CUDA): csrc_host/synth.cpp
SYNTH_SRC = os.path.join(HERE, "csrc_host", "synth.cpp")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "cds.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_synth(force=False):
    """g++ only: the synthetic producer must stay loadable where no CUDA runtime is mapped."""
    hdr = os.path.join(HERE, "..", "include", "cds_synth.h")
    if not force and os.path.exists(SYNTH_LIB) and os.path.getmtime(SYNTH_LIB) > max(os.path.getmtime(SYNTH_SRC), os.path.getmtime(hdr)):
        return SYNTH_LIB
    r = subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SYNTH_LIB, SYNTH_SRC],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("libcds_synth.so build failed:\n" + r.stdout)
    return SYNTH_LIB


FEEDER = os.path.join(HERE, "cds_feeder")                # host binary: one process per GPU fed by N decoder processes (csrc_host/cds_feeder.cpp)
FEEDER_SRC = os.path.join(HERE, "csrc_host", "cds_feeder.cpp")


def build_feeder(force=False):
    """g++ only; links the C ABI of libcds_b200.so (rpath $ORIGIN), so it is built after the library."""
    deps = [FEEDER_SRC, os.path.join(HERE, "..", "include", "cds.h"), LIB]
    if not force and os.path.exists(FEEDER) and os.path.getmtime(FEEDER) > max(os.path.getmtime(d) for d in deps):
        return FEEDER
    r = subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-pthread", "-o", FEEDER, FEEDER_SRC, "-L" + HERE, "-lcds_b200",
                        "-Wl,-rpath,$ORIGIN", "-Wl,--allow-shlib-undefined"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("cds_feeder build failed:\n" + r.stdout)
    return FEEDER


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libcds_b200.so next to this file."""
    build_synth(force)
    if not force and not needs_build():
        build_feeder(force)
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
        hdrs.append(os.path.join(HERE, "..", "include", "cds.h"))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max([os.path.getmtime(src)] + [os.path.getmtime(x) for x in hdrs]):
            continue
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose:
            print(out)
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    build_feeder(True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
