"""Generates tests/golden/*.npz + index.json by running the REFERENCE ITSELF
(oracle/_ref/libvpref_strict.so = /root/reference/Source/*.cpp compiled in place,
see oracle/Makefile) on the cases of tests/cases.py. Run in the authoring
container only (needs /root/reference to build oracle/_ref):

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures pin the CPU oracle (tests/test_oracle.py, no GPU) and the CUDA
engine (tests/test_gpu_parity.py) on boxes where /root/reference is absent."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path[:0] = [ROOT, TESTS]

import vocoderproject_b200 as vp  # noqa: E402  (host-side synthetic input generator only)
import refbind  # noqa: E402
from cases import CASES, case_inputs, case_schedule  # noqa: E402
from common import crc, oracle_decisions, stats  # noqa: E402


def main():
    assert refbind.available("strict"), "build oracle/_ref first: make -C oracle ref"
    only = set(sys.argv[1:])  # optional: regenerate the named cases only
    index = {}
    if only:
        with open(os.path.join(HERE, "index.json")) as f:
            index = json.load(f)
    for name, case in CASES.items():
        if only and name not in only:
            continue
        voice, sl, sr = case_inputs(vp, case)
        prm = refbind.default_params(**case["params"])
        sched = [(b, refbind.default_params(**d)) for b, d in case_schedule(case)]
        refbind.set_window(case.get("window") == "hann")
        r = refbind.run(case["fs"], case["B"], voice, sl, synthR=sr, params=prm, log=True, kind="strict", schedule=sched)
        refbind.set_window(False)
        rows = oracle_decisions(r["pitch"])
        mm = max([len(x["an"]) for x in rows] + [len(x["st"]) for x in rows] + [1])
        an = np.full((len(rows), mm), -1, np.int32)
        st = np.full((len(rows), mm), -1, np.int32)
        for i, x in enumerate(rows):
            an[i, :len(x["an"])] = x["an"]
            st[i, :len(x["st"])] = x["st"]
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), outL=r["outL"],
            outR=(r["outR"] if not np.array_equal(r["outL"], r["outR"]) else np.zeros(0, np.float32)),  # empty = same as outL
            gated=np.array([x["gated"] for x in rows], np.int32), period=np.array([x["period"] for x in rows], np.int32),
            periodNew=np.array([x["periodNew"] for x in rows], np.int32), note=np.array([x["note"] for x in rows], np.int32),
            stale=np.array([x["stale"] for x in rows], np.int32), beta=np.array([x["beta"] for x in rows], np.float64),
            anMarks=an, stMarks=st,
            vocGated=np.array([v.gated for v in r["voc"]], np.int32), EeVoice=np.array([v.EeVoice for v in r["voc"]]),
            EeSynth=np.array([v.EeSynth for v in r["voc"]]), g=np.array([v.g for v in r["voc"]]))
        index[name] = {"input_crc": {"voice": crc(voice), "synthL": crc(sl), "synthR": crc(sr)}, "out_crc": crc(r["outL"]),
                       "stats": stats(r["outL"]), "n": int(len(voice)), "pitch_frames": len(rows), "voc_frames": len(r["voc"]),
                       "voiced_frames": int(sum(1 for x in rows if x["period"] > 0)),
                       "gated_frames": int(sum(x["gated"] for x in rows)), "l_equals_r": bool(np.array_equal(r["outL"], r["outR"]))}
        print(name, index[name])
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
