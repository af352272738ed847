"""the WAV front-end (vocoderproject_b200/csrc/vp_wavbatch.cpp over vp_wav.hpp + vp_facade.hpp).
CPU: the tool builds, parses its manifest and the WAV flavours, and fails loudly without a device. GPU: a batch of WAV
pairs of different lengths / sample formats gives exactly what the engine gives on the decoded arrays."""
import json
import os
import subprocess
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tool(vp):
    from vocoderproject_b200 import build as b
    b.build()
    return b.build_wavbatch()


def write_pcm16(path, fs, planes):
    x = np.stack(planes, axis=1)
    q = np.clip(np.rint(x * 32768.0), -32768, 32767).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(x.shape[1]); w.setsampwidth(2); w.setframerate(int(fs)); w.writeframes(q.tobytes())
    return [q[:, c].astype(np.float32) / 32768.0 for c in range(x.shape[1])]


def write_f32(path, fs, planes):
    from scipy.io import wavfile
    x = np.stack(planes, axis=1).astype(np.float32)
    wavfile.write(path, int(fs), x if x.shape[1] > 1 else x[:, 0])
    return [x[:, c].copy() for c in range(x.shape[1])]


def write_pcm24(path, fs, planes):
    x = np.stack(planes, axis=1)
    q = np.clip(np.rint(x * 8388608.0), -8388608, 8388607).astype(np.int32)
    b = np.zeros(q.shape + (3,), np.uint8)
    for k in range(3):
        b[..., k] = (q >> (8 * k)) & 255
    with wave.open(path, "wb") as w:
        w.setnchannels(x.shape[1]); w.setsampwidth(3); w.setframerate(int(fs)); w.writeframes(b.tobytes())
    return [q[:, c].astype(np.float32) / 8388608.0 for c in range(x.shape[1])]


def read_f32(path):
    from scipy.io import wavfile
    fs, x = wavfile.read(path)
    assert x.dtype == np.float32
    return fs, x


def test_wavbatch_builds_and_reports_errors(vp, tool, tmp_path):
    assert os.path.exists(tool)
    r = subprocess.run([tool, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "manifest" in r.stdout
    man = tmp_path / "m.txt"
    man.write_text("# comment\n%s %s %s\n" % (tmp_path / "missing.wav", tmp_path / "missing2.wav", tmp_path / "o.wav"))
    r = subprocess.run([tool, str(man)], capture_output=True, text=True)
    assert r.returncode == 2 and "cannot open" in r.stderr
    # mixed sample rates are refused before any device work
    t = np.zeros(4096, np.float32)
    write_pcm16(str(tmp_path / "a.wav"), 44100, [t]); write_pcm16(str(tmp_path / "b.wav"), 48000, [t, t])
    man.write_text("%s %s %s\n" % (tmp_path / "a.wav", tmp_path / "b.wav", tmp_path / "o.wav"))
    r = subprocess.run([tool, str(man)], capture_output=True, text=True)
    assert r.returncode == 2 and "sample rate" in r.stderr
    if vp.load_library().vp_device_count() == 0:
        write_pcm16(str(tmp_path / "b.wav"), 44100, [t, t])
        r = subprocess.run([tool, str(man)], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_wav_reader_writer_against_python_decoders(tmp_path):
    """vp_wav.hpp alone (no GPU): PCM16 / PCM24 / float32, mono and stereo, decode to exactly what Python's decoders give;
    the float32 rewrite is lossless, the PCM16 rewrite is the rounded signal."""
    exe = str(tmp_path / "wav_roundtrip")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "vocoderproject_b200", "csrc"),
                           os.path.join(ROOT, "tests", "adapter", "wav_roundtrip.cpp"), "-o", exe])
    rng = np.random.default_rng(7)
    for wr, nch in ((write_pcm16, 1), (write_pcm16, 2), (write_pcm24, 2), (write_f32, 1), (write_f32, 2)):
        x = [(0.8 * rng.standard_normal(3001)).clip(-0.99, 0.99).astype(np.float32) for _ in range(nch)]
        src = str(tmp_path / "in.wav")
        dec = wr(src, 22050, x)
        f32, p16 = str(tmp_path / "f.wav"), str(tmp_path / "p.wav")
        out = subprocess.run([exe, src, f32, p16], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        j = json.loads(out.stdout)
        assert (j["sample_rate"], j["channels"], j["frames"]) == (22050, nch, 3001)
        fs, y = read_f32(f32)
        y = y.reshape(3001, nch)
        for c in range(nch):
            assert np.array_equal(y[:, c], dec[c])  # decoded exactly like Python's decoder, float32 rewrite lossless
        with wave.open(p16, "rb") as w:
            q = np.frombuffer(w.readframes(3001), "<i2").reshape(3001, nch)
        for c in range(nch):
            assert np.array_equal(q[:, c], np.clip(np.rint(dec[c] * 32768.0), -32768, 32767).astype(np.int16))
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"RIFFxxxxWAVEjunk")
    assert subprocess.run([exe, str(bad), f32, p16], capture_output=True).returncode == 1


@pytest.mark.gpu
def test_wav_batch_equals_engine_on_decoded_arrays(vp, tool, tmp_path):
    fs, B, K = 44100.0, 1024, 8
    lens = [50000, 81234, 66000]
    writers = [write_pcm16, write_f32, write_pcm24]
    v, l, r = vp.synth_host(fs, 3, max(lens), flavour=0, first_stream=500)
    dec, lines = [], []
    for s, (n, wr) in enumerate(zip(lens, writers)):
        pv, ps, po = [str(tmp_path / ("%s%d.wav" % (k, s))) for k in "vso"]
        dv = wr(pv, fs, [v[s, :n]])
        ds = wr(ps, fs, [l[s, :n], r[s, :n]] if s != 1 else [l[s, :n]])   # job 1: mono side-chain feeds both channels
        dec.append((dv[0], ds[0], ds[-1]))
        lines.append("%s %s %s" % (pv, ps, po))
    man = tmp_path / "m.txt"
    man.write_text("\n".join(lines) + "\n")
    out = subprocess.run([tool, str(man), "--block", str(B), "--blocks-per-call", str(K), "--key", "3", "--gain-synth", "-20"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    j = json.loads(out.stdout.strip().splitlines()[-1])
    lat = j["latency_samples"]
    assert lat == 1024 and j["streams"] == 3 and j["latency_compensated"] is True
    m = K * B
    n = (max(lens) + lat + m - 1) // m * m
    V, L, R = (np.zeros((3, n), np.float32) for _ in range(3))
    for s, (dv, dl, dr) in enumerate(dec):
        V[s, :len(dv)], L[s, :len(dl)], R[s, :len(dr)] = dv, dl, dr
    eng = vp.Engine(fs, B, 3, n // B, params=vp.default_params(keyPitch=3, gainSynth=-20.0))
    refL, refR = eng.process(V, L, R)
    eng.close()
    for s, nlen in enumerate(lens):
        fs_o, x = read_f32(str(tmp_path / ("o%d.wav" % s)))
        assert fs_o == 44100 and x.shape == (nlen, 2)
        assert np.array_equal(x[:, 0], refL[s, lat:lat + nlen]) and np.array_equal(x[:, 1], refR[s, lat:lat + nlen])
        assert np.abs(x).max() > 0.05
