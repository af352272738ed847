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
