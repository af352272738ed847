"""The headless C++ host (vocoderproject_b200/csrc/vp_host.cpp over vp_facade.hpp): the reference's class / method
names over the C ABI. CPU: it builds and fails loudly without a device. GPU: processBlock-by-processBlock output of the
C++ host == the Python mirror's single call, bit for bit (same library, same streams)."""
import json
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(vp):
    from vocoderproject_b200 import build as b
    return b.build_host()


def test_host_builds_and_has_no_cpu_fallback(vp, host):
    assert os.path.exists(host)
    out = subprocess.run([host, "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--blocks-per-call" in out.stdout
    if vp.load_library().vp_device_count() == 0:
        out = subprocess.run([host, "--streams", "1", "--seconds", "0.1"], capture_output=True, text=True)
        assert out.returncode == 1 and "no CPU fallback" in out.stderr


def test_facade_mirrors_the_reference_interface():
    """Class and method names of the reference headers (Source/MyBuffer.h:25-55, VocoderProcess.h:29-34,
    PitchProcess.h:40-48, Notes.h:27-28) exist in the facade."""
    src = open(os.path.join(ROOT, "vocoderproject_b200", "csrc", "vp_facade.hpp")).read()
    for cls, methods in {"MyBuffer": ["prepare", "fillInputBuffers", "fillOutputBuffer", "addDryVoice", "addSynth"],
                         "VocoderProcess": ["prepare", "getLatency", "process"],
                         "PitchProcess": ["prepare", "prepare2", "getLatency", "process", "silence"],
                         "Notes": ["prepare", "getClosestFreq"],
                         "VocoderBatchProcessor": ["prepareToPlay", "processBlock"]}.items():
        assert "class %s" % cls in src
        body = src[src.index("class %s" % cls):]
        for m in methods:
            assert "%s(" % m in body, (cls, m)


def test_facade_notes_equals_reference_notes(oracle, tmp_path):
    """vpb200::Notes (host side of the facade, the reference's Notes interface) against the oracle's table and -- where
    oracle/_ref is built -- the reference's own Notes::getClosestFreq, for every key, bit for bit."""
    import refbind
    exe = str(tmp_path / "notes_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "vocoderproject_b200", "csrc"),
                           os.path.join(ROOT, "tests", "adapter", "notes_check.cpp"), "-o", exe])
    pitches = [90.0, 100.0, 109.99, 110.0, 116.54, 123.0, 220.0, 233.08, 246.9, 440.0, 452.9, 783.99, 790.0, 801.8, 900.0]
    for key in range(13):
        out = subprocess.run([exe, str(key), "100", "800"] + [repr(p) for p in pitches], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        vals = out.stdout.split()
        n = int(vals[0])
        table, popped = oracle.notes_table(key)
        assert n == len(table)
        assert [float(v) for v in vals[1:1 + n]] == list(table)
        assert float(vals[1 + n]) == popped
        got = [float(v) for v in vals[2 + n:2 + n + len(pitches)]]
        if refbind.available("strict"):
            lib = refbind.load("strict")
            assert got == [lib.vpref_closest_freq(key, 100.0, 800.0, p) for p in pitches]
        # key change: the table of the new key
        k2 = (key + 5) % 13
        t2, _ = oracle.notes_table(k2)
        assert int(vals[-1]) == len(t2)
        assert float(vals[-2]) in list(t2)


@pytest.mark.gpu
@pytest.mark.parametrize("K", [1, 4])
def test_cpp_host_block_by_block_equals_python_batch(vp, host, K):
    fs, B, S, secs = 44100.0, 512, 6, 1.0
    out = subprocess.run([host, "--streams", str(S), "--seconds", str(secs), "--fs", str(fs), "--block", str(B),
                          "--blocks-per-call", str(K), "--key", "3", "--gain-voice", "-6", "--gain-synth", "-12"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    j = json.loads(out.stdout.strip().splitlines()[-1])
    nb = j["blocks"]
    n = nb * B
    voice, sl, sr = vp.synth_host(fs, S, n, flavour=0, first_stream=0)
    eng = vp.Engine(fs, B, S, nb, params=vp.default_params(keyPitch=3, gainVoice=-6.0, gainSynth=-12.0))
    outL, outR = eng.process(voice, sl, sr)
    eng.close()
    assert j["latency_samples"] == 1024 and j["kernel_launches"] > 0
    assert j["crc_outL"] == "%08x" % zlib.crc32(np.ascontiguousarray(outL).tobytes())
    assert j["crc_outR"] == "%08x" % zlib.crc32(np.ascontiguousarray(outR).tobytes())


def test_standalone_input_generator_equals_the_engines(vp):
    """tools/inputgen.cpp (what bench.py's reference arm loads instead of the product library) and vp_synth_host are one
    definition built twice: bit-identical inputs, all three flavours, two sample rates."""
    import ctypes as C
    import __graft_entry__ as ge
    lib = C.CDLL(ge.build_inputgen())
    fn = lib.vpgen_synth_host
    fn.restype = C.c_int
    fn.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    n = 20000
    for fs in (44100.0, 48000.0):
        for flavour in (0, 1, 2):
            v, l, r = vp.synth_host(fs, 3, n, flavour=flavour, first_stream=5)
            gv, gl, gr = (np.zeros((3, n), np.float32) for _ in range(3))
            assert fn(fs, flavour, 5, 3, n, n, gv.ctypes.data, gl.ctypes.data, gr.ctypes.data) == 0
            assert np.array_equal(v, gv) and np.array_equal(l, gl) and np.array_equal(r, gr)
    assert fn(48000.0, 0, 0, 0, n, n, gv.ctypes.data, None, None) != 0   # bad arguments are refused
