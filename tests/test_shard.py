"""N > 1 host logic on CPU: stream sharding (no data-path collective) and the
max-over-ranks / sum-over-ranks aggregation bench.py uses, world_size 2, gloo.
Each rank runs the CPU oracle on ITS shard only; the union must equal the
single-process result (streams are independent plug-in instances)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from vocoderproject_b200.shard import Group, aggregate_throughput, stream_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stream_range_partitions_exactly():
    for S in (0, 1, 7, 8, 1000, 16384):
        for G in (1, 2, 3, 4, 8):
            r = [stream_range(g, G, S) for g in range(G)]
            assert r[0][0] == 0 and r[-1][1] == S
            assert all(r[i][1] == r[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert stream_range(3, 8, 16384) == (6144, 8192)  # config 4: 2048 streams per GPU
    with pytest.raises(ValueError):
        stream_range(2, 2, 10)


def test_single_process_group_is_a_noop():
    g = Group()
    assert g.world == 1 and g.max(3.5) == 3.5 and g.sum(2.0) == 2.0
    v, t, tot = aggregate_throughput(g, 100.0, 4.0)
    assert (v, t, tot) == (25.0, 4.0, 100.0)


WORKER = r"""
import os, sys, json
import numpy as np
sys.path[:0] = [%(root)r, os.path.join(%(root)r, "tests")]
import vocoderproject_b200 as vp
from vocoderproject_b200.shard import Group, aggregate_throughput, stream_range
import oraclebind, refbind
g = Group(backend="gloo")
S, fs, B, n = 5, 44100.0, 1024, 16 * 1024
lo, hi = stream_range(g.rank, g.world, S)
v, l, r = vp.synth_host(fs, hi - lo, n, flavour=0, first_stream=lo)
outs = [oraclebind.run(fs, B, v[i], l[i], synthR=r[i], params=refbind.default_params(pitchBool=0))["outL"] for i in range(hi - lo)]
np.save(os.path.join(%(tmp)r, "shard%%d.npy" %% g.rank), np.stack(outs))
g.barrier()
secs = 1.0 + g.rank  # rank 1 is the slow one
val, t, tot = aggregate_throughput(g, (hi - lo) * n / fs, secs)
# bench.py's in-run host-link probe without a GPU: every rank fails its set-up, all agree (one collective) and return None
import bench
probe = bench.live_link_probe(g, 0, 0, 0, 1 << 20)
if g.rank == 0:
    print(json.dumps({"value": val, "t": t, "total": tot, "world": g.world, "range": [lo, hi], "probe": probe}))
g.close()
"""


def test_two_rank_gloo_sharding(tmp_path, vp, oracle):
    import json
    import refbind
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "tmp": str(tmp_path)})
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se.decode()[-2000:]
    line = json.loads(outs[0][0].decode().strip().splitlines()[-1])
    S, fs, n = 5, 44100.0, 16 * 1024
    assert line["world"] == 2 and line["range"] == [0, 2] and line["probe"] is None
    assert line["t"] == 2.0  # max over ranks
    assert abs(line["total"] - S * n / fs) < 1e-9 and abs(line["value"] - S * n / fs / 2.0) < 1e-9
    got = np.concatenate([np.load(tmp_path / "shard0.npy"), np.load(tmp_path / "shard1.npy")])
    v, l, r = vp.synth_host(fs, S, n, flavour=0, first_stream=0)
    for s in range(S):
        ref = oracle.run(fs, 1024, v[s], l[s], synthR=r[s], params=refbind.default_params(pitchBool=0))["outL"]
        assert np.array_equal(got[s], ref)
