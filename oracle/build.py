"""Compile the oracle's C/C++ restatements into oracle/_build/liboracle.so (gcc/g++ only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")
SOURCES = [
    os.path.join(HERE, "csrc", "skimage_restated.c"),
    os.path.join(HERE, "csrc", "waterz_restated.cpp"),
    os.path.join(HERE, "csrc", "mws_restated.cpp"),
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in SOURCES)


def build(force=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, os.path.basename(src) + ".o")
        cc = ["g++", "-std=c++17"] if src.endswith(".cpp") else ["gcc", "-std=c11"]
        # -ffp-contract=off: the oracle must not fuse a*b+c, the reference's builds do not
        subprocess.check_call(cc + ["-O2", "-fPIC", "-ffp-contract=off", "-c", src, "-o", obj])
        objs.append(obj)
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
