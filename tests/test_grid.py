"""Frame-grid bookkeeping of the engine (host logic, no GPU; vp_grid_plan in include/vp_engine.h) against the CPU oracle's
frame logs: where vocoder and pitch frames start when vocBool / pitchBool are toggled between blocks
(PluginProcessor.cpp:214-221; VocoderProcess::startSample / PitchProcess::startSample / nChunk freeze while their
process() is skipped) and however the blocks are grouped into process calls."""
import random

import numpy as np
import pytest

import refbind


def schedule_calls(vp, nb, sched, base, extra_cuts):
    """blocks [0, nb) split at the schedule's blocks and the extra cuts -> [(nBlocks, Params)], [first block of each call]"""
    cuts = sorted(set([0, nb]) | set(b for b, _ in sched if 0 < b < nb) | set(c for c in extra_cuts if 0 < c < nb))
    cur = dict(base)
    by_block = dict(sched)
    calls, firsts = [], []
    for a, b in zip(cuts, cuts[1:]):
        if a in by_block:
            cur = dict(cur, **by_block[a])
        calls.append((b - a, vp.default_params(**cur)))
        firsts.append(a)
    return calls, firsts


@pytest.mark.parametrize("fs,B", [(44100.0, 1024), (48000.0, 1024), (44100.0, 128), (48000.0, 64), (44100.0, 1000), (48000.0, 256), (96000.0, 512)])
def test_frame_grids_follow_the_reference_through_enable_toggles(vp, oracle, fs, B):
    rng = random.Random(int(fs) + B)
    secs = 1.2
    nb = int(fs * secs) // B
    voice, sl, _ = vp.synth_host(fs, 1, nb * B, flavour=0, first_stream=11, want_right=False)
    for trial in range(6):
        # toggles at random blocks (trial 0: none)
        sched, cur = [], dict(vocBool=1, pitchBool=1)
        for _ in range(0 if trial == 0 else rng.randint(2, 7)):
            b = rng.randint(1, nb - 1)
            cur = dict(vocBool=rng.randint(0, 1), pitchBool=rng.randint(0, 1))
            sched.append((b, cur))
        sched = sorted(dict(sched).items())
        # cumulative parameter sets, as the oracle's schedule wants them
        osched, acc = [], dict(vocBool=1, pitchBool=1)
        for b, d in sched:
            acc = dict(acc, **d)
            osched.append((b, refbind.default_params(**acc)))
        r = oracle.run(fs, B, voice[0], sl[0], params=refbind.default_params(), log=True, schedule=osched)
        ref_v = [x.block * B + x.startSample for x in r["voc"]]
        ref_p = [x.block * B + x.startSample for x in r["pitch"]]
        sz = vp.sizes_for(fs, B)
        for cuts in ([], [rng.randint(1, nb - 1) for _ in range(5)], list(range(1, nb))):
            calls, firsts = schedule_calls(vp, nb, sched, dict(vocBool=1, pitchBool=1), cuts)
            plan = vp.grid_plan(fs, B, calls)
            got_v, got_p = [], []
            for (n_blocks, _), first, pl in zip(calls, firsts, plan):
                assert pl.firstBlock == first
                got_v += [first * B + pl.offV + k * sz["hopV"] for k in range(pl.nFramesV)]
                got_p += [first * B + pl.offP + k * sz["hopP"] for k in range(pl.nFramesP)]
                assert all(first * B <= x < (first + n_blocks) * B for x in got_v[len(got_v) - pl.nFramesV:])
            assert got_v == ref_v, (trial, len(cuts))
            assert got_p == ref_p, (trial, len(cuts))


def test_orders_widen_the_rows_while_older_frames_are_in_flight(vp):
    """Rows are never narrower than the tuned kernels' 40 / 5 (a smaller order rides along zero padded). 40 -> 64 / 5 -> 9: the
    rows widen at once; 64 -> 48: the first calls after the change still carry order-64 frames, so the rows stay 65 wide until
    four order-48 frames have been carried."""
    fs, B = 44100.0, 128  # one vocoder frame per block
    p40, p20, p64, p48 = (vp.default_params(), vp.default_params(lpcVoice=20, lpcSynth=3), vp.default_params(lpcVoice=64, lpcSynth=9),
                          vp.default_params(lpcVoice=48, lpcSynth=7))
    plan = vp.grid_plan(fs, B, [(8, p40), (1, p20), (2, p64), (1, p48), (1, p48), (1, p48), (1, p48), (1, p48), (1, p20)])
    assert [pl.rowOrderV for pl in plan] == [40, 40, 64, 64, 64, 64, 64, 48, 48]
    assert [pl.rowOrderS for pl in plan] == [5, 5, 9, 9, 9, 9, 9, 7, 7]
    assert [pl.carriedV for pl in plan] == [0, 4, 4, 4, 4, 4, 4, 4, 4]


def test_vocoder_frames_in_flight_become_orphans_when_the_vocoder_is_switched_off(vp):
    fs, B = 48000.0, 128  # hop 139 > B: the grid slips against the blocks
    on, off = vp.default_params(), vp.default_params(vocBool=0)
    plan = vp.grid_plan(fs, B, [(20, on), (1, off), (1, off), (1, on), (3, on), (1, off)] + [(1, off)] * 6)
    assert plan[1].rowsOrphaned == 4 and plan[1].orphansLive == 4 and plan[1].vocMix == 1 and plan[1].nFramesV == 0
    assert plan[2].rowsOrphaned == 0 and 0 < plan[2].orphansLive <= 4
    assert plan[3].carriedV == 0 and plan[3].orphansLive > 0          # new grid: nothing carried on it yet
    assert plan[4].carriedV == plan[3].nFramesV
    assert plan[5].rowsOrphaned >= 1
    assert plan[-1].orphansLive == 0 and plan[-1].vocMix == 0          # 556 samples later everything has come out


def test_silence_cuts_the_pitch_frame_in_flight(vp):
    fs, B = 44100.0, 256  # one chunk per block; a frame every 3 blocks
    on, off = vp.default_params(), vp.default_params(pitchBool=0)
    plan = vp.grid_plan(fs, B, [(5, on), (1, off), (1, off), (1, on), (1, on), (1, on)])
    # after 5 blocks: frames at blocks 0 and 3; the frame of block 3 has had chunks 0 and 1 handled (blocks 3, 4)
    assert list(plan[1].carryPosP) == [-5 * 256, -2 * 256] and list(plan[1].carryChunksP) == [4, 2]
    assert plan[1].nFramesP == 0 and plan[1].pitchMix == 0   # chunk = block here: nothing of them reaches into the off block
    # nChunk froze at 2: the next chunk is a dead continuation, the one after it is nChunk == 3 -> a new frame starts there
    assert [pl.nFramesP for pl in plan[3:]] == [0, 1, 0] and plan[4].offP == 0
