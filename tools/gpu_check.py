"""Developer tool (GPU box): engine vs reference/oracle on a few streams with
detailed per-stage diagnostics. Not a test; tests/ holds the parity tests."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vocoderproject_b200 as vp  # noqa: E402
import oraclebind  # noqa: E402
import refbind  # noqa: E402


def snr_db(ref, x):
    ref = ref.astype(np.float64); x = x.astype(np.float64)
    err = np.sum((ref - x) ** 2)
    sig = np.sum(ref ** 2)
    if err == 0:
        return 300.0
    if sig == 0:
        return -300.0
    return 10 * np.log10(sig / err)


def check(fs, B, S, secs, flavour, label, **kw):
    n = int(fs * secs) // B * B
    voice, sl, sr = vp.synth_host(fs, S, n, flavour=flavour)
    prm = vp.default_params(**kw)
    eng = vp.Engine(fs, B, S, n // B, params=prm)
    t0 = time.time()
    outL, outR = eng.process(voice, sl, sr)
    t1 = time.time()
    use_ref = refbind.available()
    worst_snr, worst_abs = 1e9, 0.0
    nper = nbad = nflag = nre = 0
    first_report = True
    for s in range(S):
        rp = refbind.default_params(**kw)
        if use_ref:
            r = refbind.run(fs, B, voice[s], sl[s], synthR=sr[s], params=rp, log=True)
        else:
            r = oraclebind.run(fs, B, voice[s], sl[s], synthR=sr[s], params=rp, log=True)
        sn = snr_db(r["outL"], outL[s]); ab = np.abs(r["outL"].astype(np.float64) - outL[s]).max()
        snR = snr_db(r["outR"], outR[s])
        worst_snr = min(worst_snr, sn, snR); worst_abs = max(worst_abs, ab)
        pf = eng.pitch_frames(s) if prm.pitchBool else []
        for a, b in zip(r.get("pitch", []), pf):
            nper += 1
            flagged = b.flags & (vp.PF_NEAR_YIN | vp.PF_NEAR_GATE | vp.PF_UB)
            if b.flags & vp.PF_YIN_RECHECKED: nre += 1
            if flagged: nflag += 1
            gated = bool(b.flags & vp.PF_GATED)
            ok = (a.gated == int(gated))
            if ok and not gated:
                ok = (a.period == b.period and a.nAn == b.nAn and a.nSt == b.nSt and
                      list(a.anMarks[:a.nAn]) == list(b.anMarks[:b.nAn]) and list(a.stMarks[:a.nSt]) == list(b.stMarks[:b.nSt]) and
                      (a.note == b.note) and a.anStale == b.anStale and (a.nAn == 0 or a.beta == b.beta))
            if not ok:
                nbad += 1
                if first_report:
                    first_report = False
                    print("   first decision mismatch: stream", s, "frame", a.frame, "ref:", a.gated, a.period, a.note, list(a.anMarks[:a.nAn]), list(a.stMarks[:a.nSt]), a.anStale, a.beta,
                          "| eng:", b.flags, b.period, b.note, list(b.anMarks[:b.nAn]), list(b.stMarks[:b.nSt]), b.anStale, b.beta)
        if prm.vocBool and (sn < 80 or s == 0):
            vf = eng.voc_frames(s)
            rv = r["voc"]
            m = min(len(rv), len(vf["g"]))
            eV = np.array([x.EeVoice for x in rv[:m]]); eS = np.array([x.EeSynth for x in rv[:m]]); gg = np.array([x.g for x in rv[:m]])
            gt = np.array([x.gated for x in rv[:m]])
            live = gt == 0
            def rel(a, b):
                d = np.abs(a - b) / np.maximum(np.abs(a), 1e-300)
                return float(d[live].max()) if live.any() else 0.0
            print("   stream %d voc frames %d gate mismatches %d  rel err EeV %.2e EeS %.2e g %.2e" % (
                s, m, int((gt != vf["gated"][:m]).sum()), rel(eV, vf["EeVoice"][:m]), rel(eS, vf["EeSynth"][:m]), rel(gg, vf["g"][:m])))
        if sn < 80 and s < 3:
            d = np.abs(r["outL"].astype(np.float64) - outL[s])
            i = int(np.argmax(d))
            print("   stream %d snr %.1f dB maxabs %.3e at %d (ref %.6f eng %.6f), first |d|>1e-4 at %s" % (
                s, sn, d.max(), i, r["outL"][i], outL[s][i], np.argmax(d > 1e-4) if (d > 1e-4).any() else None))
    tot, stages = eng.last_timing()
    st = eng.stats()
    res = dict(label=label, fs=fs, B=B, S=S, secs=secs, worst_snr_db=round(worst_snr, 2), worst_maxabs=worst_abs,
               pitch_frames=nper, decision_mismatch=nbad, flagged=nflag, rechecked=nre, wall_s=round(t1 - t0, 3),
               stats=st, ref="reference" if use_ref else "oracle-port")
    print(json.dumps(res))
    eng.close()
    return res


def perf(fs, B, S, secs, **kw):
    n = int(fs * secs) // B * B
    prm = vp.default_params(**kw)
    os.environ["VP_STAGE_TIMING"] = "1"
    eng = vp.Engine(fs, B, S, n // B, params=prm)
    nb = S * n * 4
    dv, dl, do = eng.device_alloc(nb), eng.device_alloc(nb), eng.device_alloc(nb)
    eng.synth_device(0, 0, S, n, n, dv, dl, None)
    for it in range(3):
        eng.process_device(n // B, dv, dl, None, do, None, n)
        tot, stages = eng.last_timing()
        audio_s = S * n / fs
        print("perf S=%d secs=%g fs=%g: %.2f ms -> %.0f x RT ; stages(ms): %s" % (
            S, secs, fs, tot, audio_s / (tot * 1e-3), {k: round(v, 2) for k, v in stages.items() if v > 0}))
    print("stats", eng.stats())
    for p in (dv, dl, do): eng.device_free(p)
    eng.close()
    del os.environ["VP_STAGE_TIMING"]


if __name__ == "__main__":
    what = sys.argv[1:] or ["peaks", "voc", "pitch", "chain", "perf"]
    if "peaks" in what:
        e = vp.Engine(44100, 1024, 1, 4)
        print("peaks", json.dumps(e.measure_peaks()))
        e.close()
    if "voc" in what:
        check(44100, 1024, 4, 3.0, 0, "voc-only breathy", pitchBool=0)
        check(44100, 1024, 2, 3.0, 1, "voc-only clean", pitchBool=0)
    if "pitch" in what:
        check(44100, 1024, 4, 3.0, 0, "pitch-only chromatic", vocBool=0)
        check(44100, 1024, 2, 3.0, 0, "pitch-only C major", vocBool=0, keyPitch=3)
    if "chain" in what:
        check(44100, 1024, 4, 3.0, 0, "chain 44.1k B1024")
        check(48000, 1024, 4, 3.0, 0, "chain 48k B1024")
        check(44100, 128, 2, 3.0, 0, "chain 44.1k B128")
        check(44100, 1000, 2, 3.0, 2, "chain 44.1k B1000 gated", gainVoice=-6.0, gainSynth=-12.0)
    if "perf" in what:
        perf(44100, 1024, 256, 10.0)
        perf(48000, 1024, 256, 10.0)
