"""Parity tests proper (B200): the CUDA engine, called through the C ABI
(include/vp_engine.h via the ctypes mirror), against
  * the committed golden fixtures = the reference's own C++ output (tests/golden/),
  * the CPU oracle on the same seeded inputs (oracle/vp_oracle.c),
  * SURVEY.md App. E's known-answer table,
and, at larger sizes, through size-independent properties (shard / pass / host-vs-device
invariance, determinism, pure-delay of the dry path, latency).

Tolerances (BASELINE.json north_star): audio SNR >= 80 dB and max |err| <= 1e-4 vs the float32
reference; detected period, snapped note index and pitch marks bit-exact except frames the engine
flags as within epsilon of a decision boundary."""
import numpy as np
import pytest

import refbind
from cases import CASES, KAT, KAT_BETA, KAT_MARKS, KAT_NOTE, KAT_PERIOD, KAT_PERIODNEW, case_inputs, case_schedule
from common import MAXABS_MAX, SNR_MIN_DB, assert_decisions, compare_decisions, reference_runs, golden_index, golden_load, kat_inputs, maxabs, oracle_decisions, snr_db, stats

pytestmark = pytest.mark.gpu


def run_engine(vp, fs, B, voice, sl, sr, params, **kw):
    """voice/sl/sr: [S][n]; returns (outL, outR, engine) -- caller closes the engine."""
    S, n = voice.shape
    eng = vp.Engine(fs, B, S, n // B, params=vp.default_params(**params), **kw)
    outL, outR = eng.process(voice, sl, sr)
    return outL, outR, eng


def assert_audio(ref, got, what):
    s, m = snr_db(ref, got), maxabs(ref, got)
    assert s >= SNR_MIN_DB and m <= MAXABS_MAX, "%s: SNR %.1f dB, max abs err %.3e" % (what, s, m)
    return s, m


def golden_rows(g):
    rows = []
    for i in range(len(g["period"])):
        rows.append({"gated": int(g["gated"][i]), "period": int(g["period"][i]), "periodNew": int(g["periodNew"][i]),
                     "note": int(g["note"][i]), "an": [int(v) for v in g["anMarks"][i] if v >= 0],
                     "st": [int(v) for v in g["stMarks"][i] if v >= 0], "stale": int(g["stale"][i]), "beta": float(g["beta"][i])})
    return rows


class _FrameRows:
    """pitch / vocoder frame records of one stream accumulated over several calls"""
    def __init__(self, pf, vf):
        self.pf, self.vf = pf, vf

    def pitch_frames(self, s):
        return self.pf

    def voc_frames(self, s):
        return self.vf

    def close(self):
        pass


def run_engine_scheduled(vp, case, voice, sl, sr, extra_cuts=()):
    """One stream of an automation case: process calls split at the schedule's blocks (+ extra cuts), vp_engine_set_params
    in between -- the C-ABI equivalent of a DAW writing the plug-in's parameter atomics between processBlock calls."""
    fs, B = case["fs"], case["B"]
    nb = len(voice) // B
    sched = dict(case_schedule(case))
    cuts = sorted(set([0, nb]) | set(sched) | set(extra_cuts))
    eng = vp.Engine(fs, B, 1, max(b - a for a, b in zip(cuts, cuts[1:])), params=vp.default_params(**case["params"]),
                    reserve_orders=case.get("reserve"), window=vp.VP_WINDOW_HANN if case.get("window") == "hann" else vp.VP_WINDOW_SINE)
    try:
        outL, outR, pf = [], [], []
        vf = {"gated": [], "EeVoice": [], "EeSynth": [], "g": []}
        for a, b in zip(cuts, cuts[1:]):
            if a in sched:
                eng.set_params(vp.default_params(**sched[a]))
            l, r = eng.process(voice[None, a * B:b * B], sl[None, a * B:b * B], sr[None, a * B:b * B])
            outL.append(l[0]); outR.append(r[0])
            pf += eng.pitch_frames(0)
            v = eng.voc_frames(0)
            for k in vf:
                vf[k].append(v[k])
    finally:
        eng.close()
    return np.concatenate(outL), np.concatenate(outR), _FrameRows(pf, {k: np.concatenate(v) for k, v in vf.items()})


@pytest.mark.parametrize("name", sorted(CASES))
def test_engine_matches_reference_golden(vp, name):
    case = CASES[name]
    g = golden_load(name)
    voice, sl, sr = case_inputs(vp, case)
    if "schedule" in case:
        l, r, eng = run_engine_scheduled(vp, case, voice, sl, sr)
        outL, outR = l[None], r[None]
    else:
        outL, outR, eng = run_engine(vp, case["fs"], case["B"], voice[None], sl[None], sr[None], case["params"],
                                     window=vp.VP_WINDOW_HANN if case.get("window") == "hann" else vp.VP_WINDOW_SINE)
    try:
        assert_audio(g["outL"], outL[0], name + " L")
        assert_audio(g["outR"] if len(g["outR"]) else g["outL"], outR[0], name + " R")
        if case["params"].get("pitchBool", 1):
            dec = compare_decisions(vp, golden_rows(g), eng.pitch_frames(0))
            assert dec.n == golden_index()[name]["pitch_frames"]
            assert_decisions(dec, name)
            assert dec.flagged <= max(1, dec.n // 50)
        if case["params"].get("vocBool", 1):
            vf = eng.voc_frames(0)
            assert list(vf["gated"]) == list(g["vocGated"])
            live = g["vocGated"] == 0
            for k in ("EeVoice", "EeSynth", "g"):
                rel = np.abs(vf[k][live] - g[k][live]) / np.maximum(np.abs(g[k][live]), 1e-300)
                assert rel.max() < 1e-6, (k, rel.max())
    finally:
        eng.close()


@pytest.mark.parametrize("which", ["chain", "voc", "pitch"])
def test_engine_known_answer_table(vp, which):
    voice, synth = kat_inputs()
    k = KAT[which]
    outL, outR, eng = run_engine(vp, 44100.0, 1024, voice[None], synth[None], synth[None], k["params"])
    try:
        out = outL[0]
        assert np.array_equal(out, outR[0])
        s = stats(out)
        assert s["first_nonzero"] == k["first_nonzero"]
        assert abs(s["rms"] - k["rms"]) < 1e-6 and abs(s["max_abs"] - k["max_abs"]) < 1e-5
        assert abs(s["sum_abs"] - k["sum_abs"]) < 1e-5 * k["sum_abs"]
        for i, v in k["samples"].items():
            assert abs(float(out[i]) - v) < 1e-5
        if which != "voc":
            pf = eng.pitch_frames(0)
            assert len(pf) == 230 and pf[0].period == 0 and pf[1].period == 0
            for f in pf[2:]:
                assert f.period == KAT_PERIOD and f.note == KAT_NOTE and f.periodNew == KAT_PERIODNEW and f.beta == KAT_BETA
            for f, (_, an, st) in KAT_MARKS.items():
                assert list(pf[f].anMarks[:pf[f].nAn]) == an and list(pf[f].stMarks[:pf[f].nSt]) == st
    finally:
        eng.close()


@pytest.mark.parametrize("fs,B,S,secs,flavour,params", [
    (44100.0, 1024, 6, 3.0, 0, dict()),                       # chain, chromatic
    (48000.0, 1024, 4, 3.0, 0, dict(keyPitch=3)),             # chain 48 kHz, C major
    (44100.0, 1024, 4, 3.0, 1, dict(pitchBool=0)),            # vocoder only, clean voice (FP64 stress)
    (44100.0, 1024, 6, 3.0, 0, dict(vocBool=0)),              # pitch only
    (44100.0, 64, 2, 2.0, 0, dict()),                         # small host block (block-size dependent semantics)
    (44100.0, 1000, 2, 2.0, 0, dict(keyPitch=9)),             # ragged block
    (44100.0, 4096, 2, 3.0, 0, dict()),                       # block larger than every frame
    (88200.0, 1024, 1, 1.5, 0, dict()),                       # double sample rate: all sizes x2
    (96000.0, 1024, 1, 1.2, 0, dict(keyPitch=3)),             # 96 kHz: tauMax 960 -> generic YIN decision kernel
])
def test_engine_matches_oracle(vp, oracle, fs, B, S, secs, flavour, params):
    n = int(fs * secs) // B * B
    voice, sl, sr = vp.synth_host(fs, S, n, flavour=flavour, first_stream=100)
    outL, outR, eng = run_engine(vp, fs, B, voice, sl, sr, params)
    try:
        tot = bad = flg = 0
        for s in range(S):
            r = oracle.run(fs, B, voice[s], sl[s], synthR=sr[s], params=refbind.default_params(**params), log=True)
            assert r["ub"] == 0
            assert_audio(r["outL"], outL[s], "stream %d L" % s)
            assert_audio(r["outR"], outR[s], "stream %d R" % s)
            if params.get("pitchBool", 1):
                dec = compare_decisions(vp, oracle_decisions(r["pitch"]), eng.pitch_frames(s))
                assert_decisions(dec, "stream %d" % s)
                tot += dec.n; bad += dec.bad; flg += dec.flagged
        assert flg <= max(1, tot // 100)
    finally:
        eng.close()


def test_all_zero_and_single_block(vp):
    z = np.zeros((2, 8 * 1024), np.float32)
    outL, outR, eng = run_engine(vp, 44100.0, 1024, z, z, z, dict())
    try:
        assert not outL.any() and not outR.any()
        assert all(f.flags & vp.PF_GATED for f in eng.pitch_frames(0))
        assert eng.voc_frames(1)["gated"].all()
    finally:
        eng.close()
    voice, synth = kat_inputs(44100, 1)
    outL, outR, eng = run_engine(vp, 44100.0, 1024, voice[None, :1024], synth[None, :1024], synth[None, :1024], dict())
    eng.close()
    assert not outL.any()  # a single block only fills the latency


def test_latency_and_dry_path_is_a_pure_delay(vp):
    """vocBool = pitchBool = 0, gainVoice = 0 dB: output = input delayed by `latency` (MyBuffer.cpp:309-373)."""
    fs, B = 44100.0, 512
    voice, sl, sr = vp.synth_host(fs, 2, 40 * B, flavour=0, first_stream=7)
    outL, outR, eng = run_engine(vp, fs, B, voice, sl, sr, dict(vocBool=0, pitchBool=0, gainVoice=0.0))
    lat = eng.sizes["latency"]
    eng.close()
    assert not outL[:, :lat].any()
    assert np.array_equal(outL[:, lat:], voice[:, :-lat]) and np.array_equal(outR, outL)


def test_shard_pass_and_path_invariance_bitwise(vp):
    """Streams are independent: a stream's output must not depend on which batch, pass or path carried it."""
    fs, B, S = 44100.0, 1024, 40
    n = 48 * B
    voice, sl, sr = vp.synth_host(fs, S, n, flavour=0, first_stream=0)
    outL, outR, eng = run_engine(vp, fs, B, voice, sl, sr, dict(keyPitch=3))
    frames17 = [(f.period, f.note, list(f.anMarks[:f.nAn])) for f in eng.pitch_frames(17)]
    eng.close()
    # (1) the shard [16, 24) on its own engine
    a, _, e2 = run_engine(vp, fs, B, voice[16:24], sl[16:24], sr[16:24], dict(keyPitch=3))
    assert [(f.period, f.note, list(f.anMarks[:f.nAn])) for f in e2.pitch_frames(1)] == frames17
    e2.close()
    assert np.array_equal(a, outL[16:24])
    # (2) tiny workspace -> several passes of few streams
    b, _, e3 = run_engine(vp, fs, B, voice, sl, sr, dict(keyPitch=3), workspace_bytes=8 << 20)
    e3.close()
    assert np.array_equal(b, outL)
    # (3) device-resident path == host path, and a second run is bit-identical (deterministic accumulation)
    e4 = vp.Engine(fs, B, S, n // B, params=vp.default_params(keyPitch=3))
    try:
        nb = S * n * 4
        dv, dl, do = e4.device_alloc(nb), e4.device_alloc(nb), e4.device_alloc(nb)
        e4.h2d(dv, voice); e4.h2d(dl, sl)
        c = np.zeros((S, n), np.float32)
        for _ in range(2):
            e4.reset()  # consecutive calls continue the streams; reset = prepareToPlay again
            e4.process_device(n // B, dv, dl, None, do, None, n)
            e4.d2h(c, do)
            assert np.array_equal(c, outL)
        # (4) the on-device synthetic generator reproduces the host generator bit for bit
        e4.synth_device(0, 0, S, n, n, dv, dl, None)
        e4.d2h(c, dv)
        assert np.array_equal(c, voice)
        e4.d2h(c, dl)
        assert np.array_equal(c, sl)
        for p in (dv, dl, do):
            e4.device_free(p)
    finally:
        e4.close()


@pytest.mark.parametrize("fs,B,pieces,params", [
    (44100.0, 1024, (7, 1, 16, 24), dict(keyPitch=3)),
    (48000.0, 1024, (1, 1, 2, 3, 5, 36), dict()),
    (44100.0, 1000, (5, 43), dict(gainVoice=-6.0, gainSynth=-12.0)),
    (44100.0, 512, (30, 30, 36), dict(lpcVoice=24, lpcSynth=8, lpcPitch=20)),   # generic-order kernels
])
def test_consecutive_calls_continue_the_streams(vp, fs, B, pieces, params):
    """nBlocks = a then b then ... must give exactly the output of one call with the sum: the engine carries what
    MyBuffer / VocoderProcess / PitchProcess keep between processBlock calls (rings, energy histories, pitch marks)."""
    S, nb = 5, sum(pieces)
    n = nb * B
    voice, sl, sr = vp.synth_host(fs, S, n, flavour=0, first_stream=300)
    ref = vp.Engine(fs, B, S, nb, params=vp.default_params(**params))
    refL, refR = ref.process(voice, sl, sr)
    ref_frames = [(f.flags, f.period, f.note, list(f.anMarks[:f.nAn]), list(f.stMarks[:f.nSt])) for f in ref.pitch_frames(2)]
    ref.close()
    eng = vp.Engine(fs, B, S, max(pieces), params=vp.default_params(**params))
    try:
        outL, outR, frames, b0 = [], [], [], 0
        for k in pieces:
            sl_ = slice(b0 * B, (b0 + k) * B)
            l, r = eng.process(voice[:, sl_], sl[:, sl_], sr[:, sl_])
            outL.append(l); outR.append(r)
            frames += [(f.flags, f.period, f.note, list(f.anMarks[:f.nAn]), list(f.stMarks[:f.nSt])) for f in eng.pitch_frames(2)]
            b0 += k
        outL, outR = np.concatenate(outL, axis=1), np.concatenate(outR, axis=1)
        assert frames == ref_frames
        if "lpcVoice" in params:
            # generic-order fallback kernels: the overlap-add groups its float partial sums by tile, and the tiling
            # follows the call boundaries -> equal up to float rounding of the sum order
            assert np.abs(outL - refL).max() <= 4e-7 and np.abs(outR - refR).max() <= 4e-7
        else:
            assert np.array_equal(outL, refL) and np.array_equal(outR, refR)  # bit-exact
        eng.reset()  # and reset really is prepareToPlay
        l, r = eng.process(voice[:, :pieces[0] * B], sl[:, :pieces[0] * B], sr[:, :pieces[0] * B])
        assert np.abs(l - refL[:, :pieces[0] * B]).max() <= (4e-7 if "lpcVoice" in params else 0.0)
    finally:
        eng.close()


def test_automation_is_independent_of_call_splitting(vp):
    """Parameter changes between calls: the result depends on WHERE (which block) they happen, not on how the blocks in
    between are grouped into calls -- gains / key, and the LPC orders and enables too (frames in flight across a call
    boundary keep their order; orphaned vocoder frames and cut pitch frames come out the same)."""
    for name, cuts in (("chain44_automation", (1, 39, 41, 98, 149, 151, 200, 257)),
                       ("chain44_toggles", (1, 29, 31, 119, 121, 122, 161, 199, 201, 202, 203, 259, 261, 262, 400)),
                       ("chain48_b128_toggles", tuple(range(95, 150)) + (261, 262, 264, 265, 267, 268, 401, 499, 501, 519, 521, 599, 602, 603)),
                       ("chain48_b1024_toggles", (1, 19, 21, 23, 24, 42, 43, 61, 62, 64, 65, 91, 92))):
        case = CASES[name]
        voice, sl, sr = case_inputs(vp, case)
        aL, aR, fa = run_engine_scheduled(vp, case, voice, sl, sr)
        bL, bR, fb = run_engine_scheduled(vp, case, voice, sl, sr, extra_cuts=cuts)
        if name == "chain44_automation":
            assert np.array_equal(aL, bL) and np.array_equal(aR, bR), name  # default orders, grid kernels only: bit-exact
        else:
            # generic-order kernels and the orphan kernel group their float partial sums by tile / by call: the same
            # contributions in another order of float additions
            assert np.abs(aL - bL).max() <= 1e-6 and np.abs(aR - bR).max() <= 1e-6, (name, np.abs(aL - bL).max())
        key = lambda f: (f.flags, f.period, f.note, list(f.anMarks[:f.nAn]), list(f.stMarks[:f.nSt]))
        assert [key(f) for f in fa.pitch_frames(0)] == [key(f) for f in fb.pitch_frames(0)], name


def test_orders_beyond_the_reservation_are_refused_and_lpc_pitch_is_ignored_mid_stream(vp):
    case = CASES["chain44_automation"]
    voice, sl, sr = case_inputs(vp, case)
    n4 = 4 * 512
    eng = vp.Engine(case["fs"], case["B"], 1, 8, params=vp.default_params(**case["params"]))
    try:
        eng.process(voice[None, :n4], sl[None, :n4], sr[None, :n4])
        for bad in (dict(lpcVoice=41), dict(lpcSynth=6)):  # beyond the orders at prepare and nothing reserved
            with pytest.raises(vp.EngineError) as ei:
                eng.set_params(vp.default_params(**dict(case["params"], **bad)))
            assert ei.value.code == vp.VP_E_STATE
        for fine in (dict(lpcVoice=30), dict(lpcSynth=3), dict(vocBool=0), dict(pitchBool=0), dict(lpcPitch=9)):
            eng.set_params(vp.default_params(**dict(case["params"], **fine)))
            eng.process(voice[None, n4:2 * n4], sl[None, n4:2 * n4], sr[None, n4:2 * n4])
        eng.reset()
        eng.set_params(vp.default_params(**dict(case["params"], lpcVoice=64)))  # after prepareToPlay: re-sized by the next prepare
        with pytest.raises(vp.EngineError) as ei:
            eng.process(voice[None, :n4], sl[None, :n4], sr[None, :n4])
        assert ei.value.code == vp.VP_E_STATE
    finally:
        eng.close()
    # lpcPitch on a running stream changes nothing (PitchProcess.cpp:70: read in prepare only)
    a = vp.Engine(case["fs"], case["B"], 1, 8, params=vp.default_params(keyPitch=3))
    b = vp.Engine(case["fs"], case["B"], 1, 8, params=vp.default_params(keyPitch=3))
    try:
        outs = []
        for eng, change in ((a, False), (b, True)):
            l1, _ = eng.process(voice[None, :n4], sl[None, :n4], sr[None, :n4])
            if change:
                eng.set_params(vp.default_params(keyPitch=3, lpcPitch=9))
            l2, _ = eng.process(voice[None, n4:3 * n4], sl[None, n4:3 * n4], sr[None, n4:3 * n4])
            outs.append(np.concatenate([l1, l2], axis=1))
        assert np.array_equal(outs[0], outs[1])
    finally:
        a.close(); b.close()


def test_block_by_block_streaming_matches_oracle(vp, oracle):
    """Config-5 style: one call per host block of 128 samples (what processBlock is), 1.5 s, vs the oracle's run."""
    fs, B, S = 44100.0, 128, 3
    nb = int(fs * 1.5) // B
    voice, sl, sr = vp.synth_host(fs, S, nb * B, flavour=0, first_stream=500)
    eng = vp.Engine(fs, B, S, 1, params=vp.default_params())
    try:
        out = np.zeros((S, nb * B), np.float32)
        for b in range(nb):
            sl_ = slice(b * B, (b + 1) * B)
            l, _ = eng.process(voice[:, sl_], sl[:, sl_], None, want_right=False)
            out[:, sl_] = l
    finally:
        eng.close()
    for s in range(S):
        r = oracle.run(fs, B, voice[s], sl[s], params=refbind.default_params())
        assert_audio(r["outL"], out[s], "stream %d" % s)


def test_graph_streaming_equals_host_calls(vp):
    """vp_engine_stream_block (pinned buffers + one CUDA graph per block phase) == vp_engine_process_host per block."""
    fs, B, S, nb = 44100.0, 128, 9, 120
    voice, sl, _ = vp.synth_host(fs, S, nb * B, flavour=0, first_stream=900)
    ref = vp.Engine(fs, B, S, nb, params=vp.default_params(keyPitch=3))
    refL, _ = ref.process(voice, sl, None, want_right=False)
    ref.close()
    eng = vp.Engine(fs, B, S, 1, params=vp.default_params(keyPitch=3))
    try:
        hv, hs, ho = eng.stream_buffers()
        out = np.zeros_like(refL)
        for b in range(nb):
            hv[:] = voice[:, b * B:(b + 1) * B]
            hs[:] = sl[:, b * B:(b + 1) * B]
            eng.stream_block()
            out[:, b * B:(b + 1) * B] = ho
        st = eng.stream_stats()
        assert st["graph_launches"] == nb and st["graph_captures"] <= 16  # 6 block phases x 2 history buffers (+ first blocks)
        assert np.array_equal(out, refL)
    finally:
        eng.close()


def test_many_streams_spot_parity(vp, oracle):
    """A batch wide enough to fill the GPU (1024 streams x 2 s, chain); the oracle checks a spread of streams."""
    fs, B, S = 44100.0, 1024, 1024
    n = 86 * B
    voice, sl, sr = vp.synth_host(fs, S, n, flavour=0, first_stream=0)
    outL, outR, eng = run_engine(vp, fs, B, voice, sl, None, dict())
    try:
        assert np.isfinite(outL).all()
        tot = flg = 0
        for s in (0, 1, 255, 256, 511, 777, 1023):
            r = oracle.run(fs, B, voice[s], sl[s], params=refbind.default_params(), log=True)
            assert_audio(r["outL"], outL[s], "stream %d" % s)
            dec = compare_decisions(vp, oracle_decisions(r["pitch"]), eng.pitch_frames(s))
            assert_decisions(dec, "stream %d" % s)
            tot += dec.n; flg += dec.flagged
        assert flg <= 2
        st = eng.stats()
        assert st["kernel_launches"] >= 10 and st["yin_frames"] == S * len(eng.pitch_frames(0))
        assert st["yin_rechecked"] < 0.02 * st["yin_frames"]
    finally:
        eng.close()


@pytest.mark.parametrize("fs,S,secs,params,ws_mb", [
    (48000.0, 64, 60.0, dict(), 2048),             # the bench's own stream length (chain, 60 s @ 48 kHz), >= 3 passes
    (44100.0, 32, 30.0, dict(pitchBool=0), 512),   # BASELINE configs[1] flavour: vocoder only, 30 s
    (44100.0, 32, 30.0, dict(vocBool=0), 512),     # BASELINE configs[2] flavour: pitch corrector only, 30 s
])
def test_long_streams_match_oracle(vp, fs, S, secs, params, ws_mb):
    """Parity at the size that is benchmarked: full-length streams (20 716 vocoder / 3 453 pitch frames per stream at 60 s),
    a workspace small enough to force several passes, EVERY stream against the reference on the host cores -- audio within
    tolerance, every pitch decision compared (checked-frame floor), the mark chain over thousands of sequential frames."""
    B = 1024
    n = int(fs * secs) // B * B
    voice, sl, _ = vp.synth_host(fs, S, n, flavour=0, first_stream=2000, want_right=False)
    eng = vp.Engine(fs, B, S, n // B, params=vp.default_params(**params), workspace_bytes=ws_mb << 20)
    try:
        assert -(-S // eng.info()["streams_per_pass"]) >= 3, "workspace too large: fewer than 3 passes"
        outL, _ = eng.process(voice, sl, None, want_right=False)
        refs, kind = reference_runs([(fs, B, voice[s], sl[s], None, params) for s in range(S)])
        tot = chk = flg = 0
        worst = (1e9, 0.0)
        for s in range(S):
            sn, mx = assert_audio(refs[s]["outL"], outL[s], "stream %d (%s)" % (s, kind))
            worst = (min(worst[0], sn), max(worst[1], mx))
            if params.get("pitchBool", 1):
                dec = compare_decisions(vp, oracle_decisions(refs[s]["pitch"]), eng.pitch_frames(s))
                assert dec.n == len(refs[s]["pitch"]) == (n + eng.sizes["hopP"] - 1) // eng.sizes["hopP"]
                assert_decisions(dec, "stream %d" % s)
                tot += dec.n; chk += dec.checked; flg += dec.flagged
            if params.get("vocBool", 1):
                vf = eng.voc_frames(s)
                assert list(vf["gated"]) == [x.gated for x in refs[s]["voc"]]
        print("long streams %s: %d x %.0f s, worst SNR %.1f dB, max|err| %.2e, %d / %d pitch frames compared, %d flagged (%s)" % (
            params or "chain", S, n / fs, worst[0], worst[1], chk, tot, flg, kind))
        assert flg <= max(2, tot // 200)
    finally:
        eng.close()


@pytest.mark.parametrize("params", [dict(lpcVoice=24, lpcSynth=8, lpcPitch=20),   # side-chain order > 5: generic synthesis kernel
                                    dict(lpcVoice=20, lpcSynth=3, lpcPitch=9),    # below 40 / 5: tuned kernels on zero-padded rows
                                    dict(lpcVoice=64, lpcSynth=12, lpcPitch=30)])  # above 40: generic kernels throughout
def test_non_default_orders_at_scale(vp, params):
    """LPC orders other than the plug-in's defaults on a batch that fills the GPU (256 streams x 10 s, 2 passes): a spread of
    streams against the reference."""
    fs, B, S = 44100.0, 1024, 256
    n = int(fs * 10.0) // B * B
    voice, sl, _ = vp.synth_host(fs, S, n, flavour=0, first_stream=3000, want_right=False)
    eng = vp.Engine(fs, B, S, n // B, params=vp.default_params(**params), workspace_bytes=3 << 30)
    try:
        outL, _ = eng.process(voice, sl, None, want_right=False)
        assert np.isfinite(outL).all()
        picks = [0, 1, 63, 127, 128, 200, 254, 255]
        refs, kind = reference_runs([(fs, B, voice[s], sl[s], None, params) for s in picks])
        for s, r in zip(picks, refs):
            assert_audio(r["outL"], outL[s], "stream %d (%s)" % (s, kind))
            assert_decisions(compare_decisions(vp, oracle_decisions(r["pitch"]), eng.pitch_frames(s)), "stream %d" % s)
    finally:
        eng.close()


def test_two_phase_yin_equals_single_pass(vp, monkeypatch):
    """The two-lag-phase YIN (phase 1 decides on lags < 240, phase 2 only for the frames that need more) is exact: same
    periods, same marks, bit-identical audio as one pass over all lags (VP_YIN_PHASES=1), on low and high voices."""
    fs, B, S = 48000.0, 1024, 48
    n = int(fs * 2.5) // B * B
    voice, sl, sr = vp.synth_host(fs, S, n, flavour=0, first_stream=900)
    res = []
    for mode in ("1", "2"):
        monkeypatch.setenv("VP_YIN_PHASES", mode)
        outL, outR, eng = run_engine(vp, fs, B, voice, sl, None, dict(keyPitch=3))
        try:
            fr = [[(f.period, f.note, list(f.anMarks[:f.nAn]), list(f.stMarks[:f.nSt])) for f in eng.pitch_frames(s)] for s in range(S)]
            res.append((outL, fr))
        finally:
            eng.close()
    assert res[0][1] == res[1][1]
    assert np.array_equal(res[0][0], res[1][0])
    periods = [p[0] for s in res[0][1] for p in s if p[0] > 0]
    assert min(periods) < 200 and max(periods) > 280  # both phases decided frames


def test_voiced_silent_voiced_transitions(vp, oracle):
    """Gates closing and re-opening mid-stream (voiced 1.2 s, silence 0.8 s, voiced again): the pitch path's restart after a
    gated stretch (anMarks cleared, prevPitch = 0, PitchProcess.cpp:208-214) and the vocoder's frozen energy histories
    (VocoderProcess.cpp:199-204) -- audio and every decision against the oracle, which reports no UB on this input."""
    fs, B = 44100.0, 1024
    v, s_ = kat_inputs(int(fs), 4)
    n = int(fs * 3.2) // B * B
    voice = np.zeros(n, np.float32)
    a, b = int(fs * 1.2), int(fs * 2.0)
    voice[:a] = v[:a]
    voice[b:] = v[:n - b]
    synth = s_[:n].copy()
    r = oracle.run(fs, B, voice, synth, params=refbind.default_params(keyPitch=3), log=True)
    assert r["ub"] == 0
    outL, outR, eng = run_engine(vp, fs, B, voice[None], synth[None], synth[None], dict(keyPitch=3))
    try:
        assert_audio(r["outL"], outL[0], "voiced-silent-voiced")
        rows = oracle_decisions(r["pitch"])
        assert sum(x["gated"] for x in rows) > 20 and rows[-1]["period"] > 0
        dec = compare_decisions(vp, rows, eng.pitch_frames(0))
        assert_decisions(dec, "voiced-silent-voiced")
        assert dec.flagged <= 3
        vf = eng.voc_frames(0)
        assert list(vf["gated"]) == [x.gated for x in r["voc"]]
    finally:
        eng.close()


def test_pcm16_host_path_equals_host_conversion(vp, oracle):
    """vp_engine_process_host_pcm16 (int16 across the link, converted on the device) == converting on the host the way
    csrc/vp_wav.hpp does and calling the float entry point: bit-identical int16 output; and within one LSB of the oracle's
    output on the same quantised input."""
    fs, B, S = 44100.0, 1024, 5
    n = 60 * B + 0
    voice, sl, sr = vp.synth_host(fs, S, n, flavour=0, first_stream=40)
    q = lambda a: np.clip(np.rint(a * np.float32(32768.0)), -32768, 32767).astype(np.int16)
    qv, ql, qr = q(voice), q(sl), q(sr)
    fv, fl, fr = (a.astype(np.float32) * np.float32(1.0 / 32768.0) for a in (qv, ql, qr))
    for params, use_r in ((dict(keyPitch=3), False), (dict(gainSynth=-12.0, gainVoice=-9.0), True)):
        eng = vp.Engine(fs, B, S, n // B, params=vp.default_params(**params), workspace_bytes=64 << 20)  # several passes / slices
        try:
            oL, oR = eng.process_pcm16(qv, ql, qr if use_r else None)
            eng.reset()
            rL, rR = eng.process(fv, fl, fr if use_r else None)
            assert np.array_equal(oL, q(rL)) and np.array_equal(oR, q(rR))
            assert oL.any() and (np.array_equal(oL, oR) != use_r)
        finally:
            eng.close()
        r = oracle.run(fs, B, fv[2], fl[2], synthR=fr[2] if use_r else None, params=refbind.default_params(**params))
        assert np.abs(oL[2].astype(np.int32) - q(r["outL"]).astype(np.int32)).max() <= 1


def _run_mode(vp, fs, B, voice, synth, params, mode):
    eng = vp.Engine(fs, B, 1, len(voice) // B, params=vp.default_params(**params))
    try:
        eng.set_mode(mode)
        outL, _ = eng.process(voice[None], synth[None], None, want_right=False)
        return outL[0], eng.pitch_frames(0)
    finally:
        eng.close()


def test_defined_mode_matches_the_oracles_defined_mode(vp, oracle):
    """VP_MODE_DEFINED against the oracle's defined mode, on inputs that reach the undefined-behaviour sites:
    U4 (previous frame voiced, no previous marks), U6 (pitch above the note table) and -- on every voiced input -- U1 (the
    stale-slot read). No frame is flagged VP_PF_UB; decisions bit-exact, audio within tolerance; deterministic."""
    from test_defined_mode import u4_input, u6_input
    fs, B = 44100.0, 1024
    v6, s6 = u6_input()
    v4, s4 = u4_input()
    v0, s0, _ = vp.synth_host(fs, 1, 86 * B, flavour=0, first_stream=21, want_right=False)
    cases = [("U4", v4, s4, dict(keyPitch=3)), ("U6 key A", v6, s6, dict(keyPitch=0)), ("U6 key D", v6, s6, dict(keyPitch=5)),
             ("voice", v0[0], s0[0], dict())]
    try:
        for what, voice, synth, params in cases:
            oracle.set_defined(True)
            r = oracle.run(fs, B, voice, synth, params=refbind.default_params(**params), log=True)
            dev = oracle.defined_deviations()
            assert r["ub"] == 0
            out, frames = _run_mode(vp, fs, B, voice, synth, params, vp.VP_MODE_DEFINED)
            assert_audio(r["outL"], out, what)
            assert not any(f.flags & vp.PF_UB for f in frames), what
            dec = compare_decisions(vp, oracle_decisions(r["pitch"]), frames)
            assert_decisions(dec, what)
            out2, _ = _run_mode(vp, fs, B, voice, synth, params, vp.VP_MODE_DEFINED)
            assert np.array_equal(out, out2)
            # parity mode is still what the reference does, and differs exactly when a site decided differently
            oracle.set_defined(False)
            r0 = oracle.run(fs, B, voice, synth, params=refbind.default_params(**params), log=True)
            outp, framesp = _run_mode(vp, fs, B, voice, synth, params, vp.VP_MODE_PARITY)
            if r0["ub"] == 0:
                assert_audio(r0["outL"], outp, what + " (parity mode)")
            else:
                assert any(f.flags & vp.PF_UB for f in framesp), what  # the engine says where the reference is undefined
            if dev == 0:
                assert np.array_equal(out, outp), what
    finally:
        oracle.set_defined(False)


def test_errors_are_codes(vp):
    eng = vp.Engine(44100.0, 1024, 2, 4)
    try:
        z = np.zeros((2, 8 * 1024), np.float32)
        with pytest.raises(vp.EngineError) as ei:
            eng.process(z, z, z)  # nBlocks 8 > maxBlocks 4
        assert ei.value.code == vp.VP_E_ARG
        with pytest.raises(vp.EngineError) as ei:
            eng.set_params(vp.default_params(lpcVoice=101))
        assert ei.value.code == vp.VP_E_RANGE
        z4 = z[:, :4 * 1024]
        eng.process(z4, z4, z4)
        eng.set_params(vp.default_params(lpcPitch=16))  # read in prepare only (PitchProcess.cpp:70): ignored mid-stream
        eng.process(z4, z4, z4)
        # on a reset engine it is accepted, but the workspace has to be sized again before the next block
        eng.reset()
        eng.set_params(vp.default_params(lpcPitch=16))
        with pytest.raises(vp.EngineError) as ei:
            eng.process(z4, z4, z4)
        assert ei.value.code == vp.VP_E_STATE
        eng._check(eng.lib.vp_engine_prepare(eng.h, 44100.0, 1024, 2, 4, 0))
        eng.process(z4, z4, z4)
    finally:
        eng.close()
