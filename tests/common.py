"""Shared helpers of the parity tests: the SURVEY App. E known-answer inputs,
error metrics, decision comparison, golden-fixture access."""
import json
import os
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

# Stated tolerances of BASELINE.json's north_star (audio vs the float32 reference).
SNR_MIN_DB = 80.0
MAXABS_MAX = 1e-4


def kat_inputs(fs=44100, seconds=4):
    """integer-arithmetic inputs, exactly representable in float32."""
    n = seconds * fs
    x = 12345
    noise = np.empty(n)
    for i in range(n):
        x = (1664525 * x + 1013904223) & 0xFFFFFFFF
        noise[i] = ((x >> 16) - 32768) / 2.0 ** 22
    i = np.arange(n)
    voice = (((i % 221) - 110) / 256.0 + noise).astype(np.float32)
    synth = ((((3 * i) % 401) - 200) / 1024.0).astype(np.float32)
    return voice, synth


def snr_db(ref, x):
    ref = np.asarray(ref, np.float64)
    x = np.asarray(x, np.float64)
    err = float(np.sum((ref - x) ** 2))
    sig = float(np.sum(ref ** 2))
    if err == 0.0:
        return 300.0
    if sig == 0.0:
        return -300.0
    return 10.0 * np.log10(sig / err)


def maxabs(ref, x):
    return float(np.max(np.abs(np.asarray(ref, np.float64) - np.asarray(x, np.float64)))) if len(ref) else 0.0


def crc(a):
    return "%08x" % zlib.crc32(np.ascontiguousarray(a).tobytes())


def stats(out):
    o = np.asarray(out, np.float64)
    nz = np.flatnonzero(o)
    return {"first_nonzero": int(nz[0]) if len(nz) else -1, "sum": float(o.sum()), "sum_abs": float(np.abs(o).sum()),
            "rms": float(np.sqrt(np.mean(o * o))), "max_abs": float(np.abs(o).max())}


def oracle_decisions(plog):
    """Per pitch frame decisions of an oracle / reference log as plain tuples."""
    rows = []
    for a in plog:
        rows.append({"gated": int(a.gated), "period": int(a.period), "periodNew": int(a.periodNew), "note": int(a.note),
                     "an": [int(v) for v in a.anMarks[:a.nAn]], "st": [int(v) for v in a.stMarks[:a.nSt]],
                     "stale": int(a.anStale), "beta": float(a.beta)})
    return rows


def compare_decisions(vp, ref_rows, eng_frames):
    """Bit-exact comparison of the integer decisions (period, snapped note, marks).
    Returns (n_frames, n_mismatch, n_flagged, first_mismatch_text). Frames the engine flags as
    within epsilon of a decision boundary (or UB in the reference) are excluded, and so is everything
    after the first such frame of a stream: the mark chain carries state from frame to frame."""
    n = min(len(ref_rows), len(eng_frames))
    bad = flagged = 0
    first = None
    tainted = False
    for i in range(n):
        a, b = ref_rows[i], eng_frames[i]
        if b.flags & (vp.PF_NEAR_YIN | vp.PF_NEAR_GATE | vp.PF_UB):
            flagged += 1
            tainted = True
        if tainted:
            continue
        gated = 1 if (b.flags & vp.PF_GATED) else 0
        ok = a["gated"] == gated
        if ok and not gated:
            ok = (a["period"] == b.period and a["an"] == list(b.anMarks[:b.nAn]) and a["st"] == list(b.stMarks[:b.nSt]) and
                  a["note"] == b.note and a["stale"] == b.anStale)
            if ok and a["an"]:
                ok = a["periodNew"] == b.periodNew and a["beta"] == b.beta
        if not ok:
            bad += 1
            if first is None:
                first = "frame %d ref %r | eng flags=%d period=%d pnew=%d note=%d an=%r st=%r stale=%d beta=%r" % (
                    i, a, b.flags, b.period, b.periodNew, b.note, list(b.anMarks[:b.nAn]), list(b.stMarks[:b.nSt]),
                    b.anStale, b.beta)
    return n, bad, flagged, first


def golden_index():
    with open(os.path.join(GOLDEN, "index.json")) as f:
        return json.load(f)


def golden_load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))
