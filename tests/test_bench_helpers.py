"""bench.py's host-side arithmetic (no GPU): the host-link bound, the committed probe logs it falls back to, and the
instruction-mix roofline read from the newest committed ncu capture."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_link_bound_is_the_slowest_of_the_three_terms():
    link = {"h2d": 50.0, "d2h": 50.0, "duplex": 80.0}
    # twice as many bytes up as down (float I/O of the chain): the upload alone bounds the step
    assert bench.link_bound(link, 100e9, 50e9) == pytest.approx(2.0)
    # equal bytes: the shared total rate does
    assert bench.link_bound(link, 60e9, 60e9) == pytest.approx(1.5)
    # a slow download direction
    assert bench.link_bound({"h2d": 100.0, "d2h": 10.0, "duplex": 200.0}, 10e9, 20e9) == pytest.approx(2.0)


def test_committed_probe_logs_cover_1_2_4_8_gpus():
    for n in (1, 2, 4, 8):
        link = bench.link_ceiling(n)
        assert link is not None and all(link[k] > 10.0 for k in ("h2d", "d2h", "duplex")), n
        assert link["duplex"] <= link["h2d"] + link["d2h"] + 1e-9   # both directions together never beat the two alone
        assert os.path.exists(link["file"])
    assert bench.link_ceiling(3) is None
    # the 8-GPU finding the end-to-end section of DESIGN.md rests on: the host serves 8 GPUs less than twice what it serves one
    assert bench.link_ceiling(8)["duplex"] < 2.0 * bench.link_ceiling(1)["duplex"]


def test_instruction_mix_capture_feeds_the_roofline():
    for wl in ("chain48", "chain44", "voc44", "pitch44"):
        w = bench.mix_capture(wl)
        assert w is not None and w["passes_captured"] >= 1, wl
        tot = w["total"]
        assert tot["thread_inst_per_sample"] > tot["fp64_ops_per_sample"] + tot["fp32_ops_per_sample"] > 0
        assert abs(sum(s["fp64_ops_per_sample"] for s in w["stages"].values()) - tot["fp64_ops_per_sample"]) < 1e-6
    peaks = {"fp64_fma_per_s": 18.3e12, "fp32_fma_per_s": 36.0e12}
    samples = 2048 * 60 * 48000
    m = bench.mix_roofline("chain48", samples, 500.0, peaks, 1965.0)
    assert m is not None and 0.3 < m["frac"] < 1.0 and m["lower_bound_ms"] == pytest.approx(m["frac"] * 500.0)
    # the FP64 term dominates the chain; a step as fast as the bound would read frac = 1
    assert bench.mix_roofline("chain48", samples, m["lower_bound_ms"], peaks, 1965.0)["frac"] == pytest.approx(1.0)
    assert bench.mix_roofline("no-such-workload", samples, 500.0, peaks, 1965.0) is None
