"""Compile libbsnative.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libbsnative.so")
SOURCES = ["prims.cu", "plan.cu", "stage1.cu", "stage2.cu", "agglom_par.cu", "agglom_pq.cu", "stage3.cu", "cc.cu", "mws.cu", "affagglom.cu", "gauss.cu", "afferr.cu", "labelstats.cu", "synth.cu", "api.cu"]
HEADERS = ["common.cuh", "geom.h", "agglom.cuh", "front2d.cuh", os.path.join("..", "..", "include", "bsnative.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + (["-DBS_TRACE"] if os.environ.get("BS_TRACE") else []) + (["-DBS_PROBE=" + os.environ["BS_PROBE"]] if os.environ.get("BS_PROBE") else [])


STAMP = os.path.join(HERE, ".build_flags")


def needs_build():
    if not os.path.exists(LIB):
        return True
    if os.path.exists(STAMP) and open(STAMP).read() != " ".join(FLAGS):      # BS_TRACE / BS_PROBE changed since the last build
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen([NVCC] + FLAGS + ["-c", os.path.join(HERE, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    ok = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            ok = False
            sys.stderr.write(f"--- {src}\n{out}\n")
        elif verbose:
            sys.stderr.write(f"--- {src}\n{out}\n")
    if not ok:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"])
    with open(STAMP, "w") as f:
        f.write(" ".join(FLAGS))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
