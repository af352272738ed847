"""ctypes binding for oracle/_ref/libvpref_*.so -- the reference's own C++
compiled in place (oracle/Makefile, oracle/ref_harness.cpp). Test
infrastructure only: imported by tests/, bench.py's cpu_baseline / --impl
reference legs and __graft_entry__.smoke(), never by the product package."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MAX_MARKS = 64


class Params(C.Structure):
    _fields_ = [("gainPitch", C.c_float), ("gainVoice", C.c_float), ("gainSynth", C.c_float),
                ("gainVoc", C.c_float), ("lpcVoice", C.c_int), ("lpcPitch", C.c_int),
                ("lpcSynth", C.c_int), ("keyPitch", C.c_int), ("pitchBool", C.c_int),
                ("vocBool", C.c_int)]


class PitchFrame(C.Structure):
    _fields_ = [("frame", C.c_int), ("startSample", C.c_int), ("block", C.c_int), ("gated", C.c_int),
                ("period", C.c_int), ("periodNew", C.c_int), ("prevVoicedPeriod", C.c_int),
                ("note", C.c_int), ("nAn", C.c_int), ("nSt", C.c_int), ("anStale", C.c_int),
                ("anMarks", C.c_int * MAX_MARKS), ("stMarks", C.c_int * MAX_MARKS),
                ("pitch", C.c_double), ("closestFreq", C.c_double), ("beta", C.c_double)]


class VocFrame(C.Structure):
    _fields_ = [("frame", C.c_int), ("startSample", C.c_int), ("block", C.c_int), ("gated", C.c_int),
                ("EeVoice", C.c_double), ("EeSynth", C.c_double), ("g", C.c_double)]


class Sizes(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("hopV", "wlenV", "hopP", "frameLenP", "chunk", "tauMax",
                                       "latency", "keep", "inSize", "outSize", "anCap", "nFreq")]


def default_params(**kw):
    p = Params(0.0, -60.0, -60.0, 0.0, 40, 15, 5, 12, 1, 1)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def lib_path(kind="strict"):
    return os.path.join(ROOT, "oracle", "_ref", "libvpref_%s.so" % kind)


def available(kind="strict"):
    return os.path.exists(lib_path(kind))


_libs = {}


def load(kind="strict"):
    if kind in _libs:
        return _libs[kind]
    lib = C.CDLL(lib_path(kind))
    fp = C.POINTER(C.c_float)
    lib.vpref_run.restype = C.c_int
    lib.vpref_run.argtypes = [C.c_double, C.c_int, C.c_int, fp, fp, fp, C.POINTER(Params), C.c_int, fp, fp,
                              C.POINTER(Sizes), C.POINTER(PitchFrame), C.c_int, C.POINTER(C.c_int),
                              C.POINTER(VocFrame), C.c_int, C.POINTER(C.c_int)]
    lib.vpref_run_sched.restype = C.c_int
    lib.vpref_run_sched.argtypes = [C.c_double, C.c_int, C.c_int, fp, fp, fp, C.POINTER(Params), C.POINTER(Params),
                                    C.POINTER(C.c_int), C.c_int, C.c_int, fp, fp,
                                    C.POINTER(Sizes), C.POINTER(PitchFrame), C.c_int, C.POINTER(C.c_int),
                                    C.POINTER(VocFrame), C.c_int, C.POINTER(C.c_int)]
    lib.vpref_bench.restype = C.c_double
    lib.vpref_bench.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, fp, fp, fp, C.POINTER(Params), C.c_int, fp]
    lib.vpref_notes.restype = C.c_int
    lib.vpref_notes.argtypes = [C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.c_int,
                                C.POINTER(C.c_double)]
    lib.vpref_set_window.restype = None
    lib.vpref_set_window.argtypes = [C.c_int]
    lib.vpref_closest_freq.restype = C.c_double
    lib.vpref_closest_freq.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    _libs[kind] = lib
    return lib


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def sched_arrays(schedule):
    """[(block, Params), ...] -> (Params array, int array, count) for the *_sched entry points."""
    if not schedule:
        return None, None, 0
    n = len(schedule)
    ps = (Params * n)()
    bs = (C.c_int * n)()
    for i, (b, q) in enumerate(schedule):
        C.memmove(C.byref(ps[i]), C.byref(q), C.sizeof(Params))
        bs[i] = int(b)
    return ps, bs, n


def run(fs, B, voice, synthL, synthR=None, params=None, log=False, kind="strict", schedule=None):
    """One stream through the reference's processBlock. Returns dict with
    outL, outR (float32[nBlocks*B]), sizes, and (log=True) pitch/voc frame logs.
    schedule: [(block, Params), ...] -- parameter automation, stored before that block."""
    lib = load(kind)
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
    nP = C.c_int(0)
    nV = C.c_int(0)
    sp, sb, ns = sched_arrays(schedule)
    rc = lib.vpref_run_sched(fs, B, nBlocks, _fptr(voice), _fptr(synthL), _fptr(synthR), C.byref(params), sp, sb, ns,
                             1 if log else 0, _fptr(outL), _fptr(outR), C.byref(sizes), plog, pcap, C.byref(nP),
                             vlog, vcap, C.byref(nV))
    if rc != 0:
        raise RuntimeError("vpref_run failed: %d" % rc)
    res = {"outL": outL, "outR": outR, "sizes": {k: getattr(sizes, k) for k, _ in Sizes._fields_}}
    if log:
        res["pitch"] = [plog[i] for i in range(min(nP.value, pcap))]
        res["voc"] = [vlog[i] for i in range(min(nV.value, vcap))]
    return res


def set_window(hann, kind="strict"):
    """Vocoder window of the instances created from now on: False = "sine" (prepareToPlay's), True = "hann"."""
    load(kind).vpref_set_window(1 if hann else 0)


def bench(fs, B, voice, synthL, synthR=None, params=None, threads=1, kind="fast", want_out=False):
    """voice/synthL/synthR: float32 [S][n]. Returns (seconds, out or None)."""
    lib = load(kind)
    params = params or default_params()
    voice = np.ascontiguousarray(voice, np.float32)
    synthL = np.ascontiguousarray(synthL, np.float32)
    S, n = voice.shape
    nBlocks = n // B
    assert nBlocks * B == n
    out = np.zeros((S, 2, n), np.float32) if want_out else None
    sec = lib.vpref_bench(fs, B, nBlocks, S, _fptr(voice), _fptr(synthL),
                          None if synthR is None else _fptr(np.ascontiguousarray(synthR, np.float32)),
                          C.byref(params), threads, None if out is None else _fptr(out))
    return sec, out


def notes_table(key, fMin=100.0, fMax=800.0):
    lib = load()
    buf = (C.c_double * 128)()
    popped = C.c_double(0)
    n = lib.vpref_notes(key, fMin, fMax, buf, 128, C.byref(popped))
    return np.array(buf[:n]), popped.value
