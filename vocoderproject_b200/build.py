"""Builds vocoderproject_b200/lib/libvp_engine.so in-tree with nvcc for sm_100a.
No torch, no JIT cache: the .so travels with the repo snapshot."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libvp_engine.so")
SOURCES = ["vp_engine.cu", "vp_voc.cu", "vp_pitch.cu", "vp_misc.cu"]
HEADERS = ["vp_common.cuh", "vp_synth.h", "vp_synth_host.hpp", os.path.join("..", "..", "include", "vp_engine.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path():
    p = shutil.which("nvcc")
    if p:
        return p
    p = "/usr/local/cuda/bin/nvcc"
    if os.path.exists(p):
        return p
    raise RuntimeError("nvcc not found: the engine has no CPU fallback and cannot be built without CUDA")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    extra = os.environ.get("VP_NVCC_EXTRA", "").split()  # developer A/B builds (e.g. -DAV_STAGE_UNROLL=6)
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


HOST = os.path.join(LIBDIR, "vp_host")


def build_host(force=False):
    """Headless C++ host (csrc/vp_host.cpp over csrc/vp_facade.hpp), linked to libvp_engine.so through the C ABI."""
    src = [os.path.join(CSRC, "vp_host.cpp"), os.path.join(CSRC, "vp_facade.hpp"), LIB]
    if not force and os.path.exists(HOST) and all(os.path.getmtime(f) <= os.path.getmtime(HOST) for f in src):
        return HOST
    cxx = shutil.which("g++") or "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-o", HOST, os.path.join(CSRC, "vp_host.cpp"), "-L" + LIBDIR, "-lvp_engine",
                           "-Wl,-rpath,$ORIGIN"], cwd=CSRC)
    return HOST


WAVBATCH = os.path.join(LIBDIR, "vp_wavbatch")


def build_wavbatch(force=False):
    """WAV front-end of the batch engine (csrc/vp_wavbatch.cpp: vp_wav.hpp + vp_facade.hpp over the C ABI)."""
    src = [os.path.join(CSRC, f) for f in ("vp_wavbatch.cpp", "vp_wav.hpp", "vp_facade.hpp")] + [LIB]
    if not force and os.path.exists(WAVBATCH) and all(os.path.getmtime(f) <= os.path.getmtime(WAVBATCH) for f in src):
        return WAVBATCH
    cxx = shutil.which("g++") or "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-o", WAVBATCH, os.path.join(CSRC, "vp_wavbatch.cpp"), "-L" + LIBDIR,
                           "-lvp_engine", "-Wl,-rpath,$ORIGIN"], cwd=CSRC)
    return WAVBATCH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
    print(build_wavbatch(force="--force" in sys.argv))
