"""ctypes binding for oracle/libvp_oracle.so -- the plain-C CPU restatement of
the reference's hot path (oracle/vp_oracle.c). Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from refbind import Params, PitchFrame, VocFrame, Sizes, default_params, sched_arrays, _fptr  # same struct layouts

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, "oracle", "libvp_oracle.so")

UB_NAMES = {1: "U2_closest_prev", 2: "U3_yin_end", 4: "U4_prev_empty", 8: "U5_interp", 16: "capacity", 32: "assert"}

_lib = None


def build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libvp_oracle.so"],
                          stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(ROOT, "oracle", "vp_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
        build()
    lib = C.CDLL(LIB)
    fp = C.POINTER(C.c_float)
    lib.vpo_process.restype = C.c_int
    lib.vpo_process.argtypes = [C.c_double, C.c_int, C.c_int, fp, fp, fp, C.POINTER(Params), fp, fp,
                                C.POINTER(Sizes), C.POINTER(PitchFrame), C.c_int, C.POINTER(C.c_int),
                                C.POINTER(VocFrame), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.vpo_process_sched.restype = C.c_int
    lib.vpo_process_sched.argtypes = [C.c_double, C.c_int, C.c_int, fp, fp, fp, C.POINTER(Params), C.POINTER(Params),
                                      C.POINTER(C.c_int), C.c_int, fp, fp,
                                      C.POINTER(Sizes), C.POINTER(PitchFrame), C.c_int, C.POINTER(C.c_int),
                                      C.POINTER(VocFrame), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.vpo_bench.restype = C.c_double
    lib.vpo_bench.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, fp, fp, fp, C.POINTER(Params), C.c_int, fp]
    lib.vpo_notes.restype = C.c_int
    lib.vpo_notes.argtypes = [C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.c_int,
                              C.POINTER(C.c_double)]
    lib.vpo_set_window.restype = None
    lib.vpo_set_window.argtypes = [C.c_int]
    lib.vpo_set_defined.restype = None
    lib.vpo_set_defined.argtypes = [C.c_int]
    lib.vpo_defined_deviations.restype = C.c_long
    lib.vpo_defined_deviations.argtypes = []
    lib.vpo_sizes_for.restype = None
    lib.vpo_sizes_for.argtypes = [C.c_double, C.c_int, C.c_int, C.POINTER(Sizes)]
    _lib = lib
    return lib


def run(fs, B, voice, synthL, synthR=None, params=None, log=False, schedule=None):
    """schedule: [(block, Params), ...] -- parameter automation, applied before that block."""
    lib = load()
    params = params or default_params()
    voice = np.ascontiguousarray(voice, np.float32)
    synthL = np.ascontiguousarray(synthL, np.float32)
    synthR = synthL if synthR is None else np.ascontiguousarray(synthR, np.float32)
    nBlocks = len(voice) // B
    n = nBlocks * B
    outL = np.zeros(n, np.float32)
    outR = np.zeros(n, np.float32)
    sizes = Sizes()
    pcap = n // 64 + 16
    vcap = n // 32 + 16
    plog = (PitchFrame * pcap)() if log else None
    vlog = (VocFrame * vcap)() if log else None
    nP, nV, ub = C.c_int(0), C.c_int(0), C.c_int(0)
    sp, sb, ns = sched_arrays(schedule)
    rc = lib.vpo_process_sched(fs, B, nBlocks, _fptr(voice), _fptr(synthL), _fptr(synthR), C.byref(params), sp, sb, ns,
                               _fptr(outL), _fptr(outR), C.byref(sizes), plog, pcap, C.byref(nP), vlog, vcap,
                               C.byref(nV), C.byref(ub))
    if rc != 0:
        raise RuntimeError("vpo_process failed: %d" % rc)
    res = {"outL": outL, "outR": outR, "ub": ub.value,
           "sizes": {k: getattr(sizes, k) for k, _ in Sizes._fields_}}
    if log:
        res["pitch"] = [plog[i] for i in range(min(nP.value, pcap))]
        res["voc"] = [vlog[i] for i in range(min(nV.value, vcap))]
    return res


def set_window(hann):
    """Vocoder window of the runs that follow: False = "sine", True = "hann" (process-global)."""
    load().vpo_set_window(1 if hann else 0)


def set_defined(on):
    """Defined-behaviour mode of the oracle (process-global; resets the deviation counter)."""
    load().vpo_set_defined(1 if on else 0)


def defined_deviations():
    return int(load().vpo_defined_deviations())


def bench(fs, B, voice, synthL, synthR=None, params=None, threads=1, want_out=False):
    lib = load()
    params = params or default_params()
    voice = np.ascontiguousarray(voice, np.float32)
    synthL = np.ascontiguousarray(synthL, np.float32)
    S, n = voice.shape
    nBlocks = n // B
    assert nBlocks * B == n
    out = np.zeros((S, 2, n), np.float32) if want_out else None
    sec = lib.vpo_bench(fs, B, nBlocks, S, _fptr(voice), _fptr(synthL),
                        None if synthR is None else _fptr(np.ascontiguousarray(synthR, np.float32)),
                        C.byref(params), threads, None if out is None else _fptr(out))
    return sec, out


def notes_table(key, fMin=100.0, fMax=800.0):
    lib = load()
    buf = (C.c_double * 128)()
    popped = C.c_double(0)
    n = lib.vpo_notes(key, fMin, fMax, buf, 128, C.byref(popped))
    return np.array(buf[:n]), popped.value


def sizes_for(fs, B, key=12):
    lib = load()
    s = Sizes()
    lib.vpo_sizes_for(fs, B, key, C.byref(s))
    return {k: getattr(s, k) for k, _ in Sizes._fields_}
