"""the JUCE plug-in adapter (vocoderproject_b200/csrc/vp_juce_adapter.hpp). tests/adapter/b200_plugin.cpp
is a juce::AudioProcessor -- compiled against the oracle's JUCE stub -- whose prepareToPlay / processBlock forward to the
B200 engine; it is driven block by block with an in-place 3-channel buffer exactly like the reference plug-in
(PluginProcessor.cpp:203-234). CPU: it compiles and fails loudly without a device. GPU: its output equals the
REFERENCE's own output on the golden fixtures (incl. the parameter-automation one)."""
import json
import os
import subprocess

import numpy as np
import pytest

from cases import CASES, case_inputs
from common import MAXABS_MAX, SNR_MIN_DB, golden_load, maxabs, snr_db

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "adapter", "_build")


@pytest.fixture(scope="module")
def plugin(vp):
    from vocoderproject_b200 import build as b
    b.build()  # libvp_engine.so (no-op when up to date)
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "b200_plugin")
    src = os.path.join(ROOT, "tests", "adapter", "b200_plugin.cpp")
    deps = [src, os.path.join(ROOT, "vocoderproject_b200", "csrc", "vp_juce_adapter.hpp"),
            os.path.join(ROOT, "vocoderproject_b200", "csrc", "vp_facade.hpp"), os.path.join(ROOT, "include", "vp_engine.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        lib = os.path.join(ROOT, "vocoderproject_b200", "lib")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-include", "math.h", "-include", "stdlib.h",
                               "-I", os.path.join(ROOT, "oracle", "stub", "JuceLibraryCode"),
                               "-I", os.path.join(ROOT, "vocoderproject_b200", "csrc"), src,
                               "-L", lib, "-lvp_engine", "-Wl,-rpath," + lib, "-o", exe])
    return exe


def run_plugin(exe, fs, B, voice, sl, sr, args, tmp_path):
    n = len(voice) // B * B
    fin, fout = str(tmp_path / "in.f32"), str(tmp_path / "out.f32")
    np.concatenate([voice[:n], sl[:n], sr[:n]]).astype(np.float32).tofile(fin)
    r = subprocess.run([exe, fin, fout, repr(fs), str(B), str(n // B)] + args, capture_output=True, text=True)
    return r, (np.fromfile(fout, np.float32).reshape(2, n) if r.returncode == 0 else None)


def plugin_args(case):
    a = ["%s=%r" % kv for kv in case["params"].items()]
    for b, d in case.get("schedule", []):
        a += ["@%d" % b] + ["%s=%r" % kv for kv in d.items()]
    return a


def test_adapter_builds_and_has_no_cpu_fallback(vp, plugin, tmp_path):
    assert os.path.exists(plugin)
    if vp.load_library().vp_device_count() == 0:
        z = np.zeros(2048, np.float32)
        r, out = run_plugin(plugin, 44100.0, 1024, z, z, z, [], tmp_path)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr and out is None


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["chain44_cmaj", "chain44_b1000_mix", "chain48_chrom", "chain44_automation", "chain44_toggles",
                                  "chain48_b128_toggles", "chain48_b1024_toggles"])
def test_plugin_adapter_matches_reference_golden(vp, plugin, tmp_path, name):
    case = CASES[name]
    g = golden_load(name)
    voice, sl, sr = case_inputs(vp, case)
    r, out = run_plugin(plugin, case["fs"], case["B"], voice, sl, sr, plugin_args(case), tmp_path)
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout.strip().splitlines()[-1])["latency_samples"] == (1024 if case["fs"] == 44100.0 else 1112)
    gR = g["outR"] if len(g["outR"]) else g["outL"]
    for ref, got, ch in ((g["outL"], out[0], "L"), (gR, out[1], "R")):
        s, m = snr_db(ref, got), maxabs(ref, got)
        assert s >= SNR_MIN_DB and m <= MAXABS_MAX, "%s %s: SNR %.1f dB, max abs err %.3e" % (name, ch, s, m)
