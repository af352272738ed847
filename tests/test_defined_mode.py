"""the defined-behaviour mode -- the bounds-correct reading at every place where the reference's C++ has
undefined behaviour (App. B U1-U6). Parity mode stays the default. The CPU oracle and the engine implement the same rules
(oracle/vp_oracle.h vpo_set_defined, include/vp_engine.h VP_MODE_DEFINED); here the oracle side (no GPU), the engine side is
tests/test_gpu_parity.py::test_defined_mode_*."""
import numpy as np
import pytest

import refbind
from cases import CASES, case_inputs, case_schedule
from common import golden_load, kat_inputs, oracle_decisions


@pytest.fixture()
def defined(oracle):
    oracle.set_defined(True)
    yield oracle
    oracle.set_defined(False)


def u4_input(fs=44100.0, B=1024, off=0):
    """Loud periodic voice, 0.6 s at -100 dB (still periodic: the gate closes while the last pitch is > 1), loud again: the
    first un-gated frame is voiced with no previous marks -> prevAnMarks.back() on an empty vector (PitchProcess.cpp:487)."""
    v, s = kat_inputs(int(fs), 4)
    n = int(fs * 3.0) // B * B
    voice = v[:n].copy()
    a, b = int(fs * 1.0), int(fs * 1.6) + off
    voice[a:b] *= 1e-5
    return voice, s[:n].copy()


def u6_input(fs=44100.0, B=1024):
    """Pulse train with a period of 55 samples = 801.8 Hz: above the last entry of every note table (Notes.cpp:99)."""
    _, s = kat_inputs(int(fs), 4)
    n = int(fs * 2.0) // B * B
    i = np.arange(n)
    voice = ((i % 55) < 3).astype(np.float32) * 0.8 - 0.05
    voice += 0.001 * np.random.RandomState(1).randn(n).astype(np.float32)
    return voice.astype(np.float32), s[:n].copy()


@pytest.mark.parametrize("name", sorted(CASES))
def test_defined_mode_equals_the_reference_where_no_site_decides_differently(vp, defined, name):
    case = CASES[name]
    g = golden_load(name)
    voice, sl, sr = case_inputs(vp, case)
    sched = [(b, refbind.default_params(**d)) for b, d in case_schedule(case)]
    defined.set_window(case.get("window") == "hann")
    try:
        r = defined.run(case["fs"], case["B"], voice, sl, synthR=sr, params=refbind.default_params(**case["params"]), log=True, schedule=sched)
    finally:
        defined.set_window(False)
    dev = defined.defined_deviations()
    assert r["ub"] == 0
    assert np.isfinite(r["outL"]).all()
    if dev == 0:
        assert np.array_equal(r["outL"], g["outL"])  # bit-identical to the reference's own output
        assert [x["period"] for x in oracle_decisions(r["pitch"])] == list(g["period"])
    else:
        # the stale-slot read of PitchProcess.cpp:818 decided differently for `dev` grains: same periods and marks (the
        # site only selects which analysis grain a synthesis mark copies), audio close
        rows = oracle_decisions(r["pitch"])
        assert [x["period"] for x in rows] == list(g["period"])
        assert [x["an"] for x in rows] == [[int(v) for v in a if v >= 0] for a in g["anMarks"]]
        assert dev < 0.05 * max(1, sum(len(x["st"]) for x in rows))


def test_defined_mode_previous_frame_voiced_without_marks(oracle, defined):
    voice, synth = u4_input()
    prm = refbind.default_params(keyPitch=3)
    r1 = defined.run(44100.0, 1024, voice, synth, params=prm, log=True)
    assert r1["ub"] == 0 and defined.defined_deviations() >= 1
    r2 = defined.run(44100.0, 1024, voice, synth, params=prm, log=True)
    assert np.array_equal(r1["outL"], r2["outL"])  # deterministic
    oracle.set_defined(False)
    r0 = oracle.run(44100.0, 1024, voice, synth, params=prm, log=True)
    assert r0["ub"] & 4  # the reference itself: U4
    # same decisions up to the frame where the gate re-opens
    k = next(i for i, (a, b) in enumerate(zip(oracle_decisions(r0["pitch"]), oracle_decisions(r1["pitch"]))) if a != b)
    assert r1["pitch"][k - 1].gated and not r1["pitch"][k].gated and r1["pitch"][k].period > 0


@pytest.mark.parametrize("key,popped_wins", [(0, True), (1, False), (2, True), (3, False), (12, False)])
def test_defined_mode_notes_upper_edge(oracle, defined, key, popped_wins):
    voice, synth = u6_input()
    prm = refbind.default_params(keyPitch=key)
    r1 = defined.run(44100.0, 1024, voice, synth, params=prm, log=True)
    dev = defined.defined_deviations()
    oracle.set_defined(False)
    r0 = oracle.run(44100.0, 1024, voice, synth, params=prm, log=True)
    voiced = [i for i, p in enumerate(r0["pitch"]) if p.period == 55 and p.nAn > 0]
    assert len(voiced) > 20
    n_table = oracle.notes_table(key)[0].size
    if popped_wins:  # the reference snaps to the popped slot freq[size]; defined mode to the last table entry
        assert all(r0["pitch"][i].note == n_table and r1["pitch"][i].note == n_table - 1 for i in voiced) and dev > 0
    else:
        assert all(r0["pitch"][i].note == r1["pitch"][i].note == n_table - 1 for i in voiced)
        assert np.array_equal(r0["outL"], r1["outL"]) and dev == 0
