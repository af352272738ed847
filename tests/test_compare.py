"""The decision comparison of the parity tests itself (tests/common.py): what counts as compared, flagged, excused."""
import types

import pytest

from common import CHECKED_FLOOR, assert_decisions, compare_decisions

VP = types.SimpleNamespace(PF_GATED=1, PF_VOICED=2, PF_HAS_MARKS=4, PF_NEAR_GATE=16, PF_NEAR_YIN=32, PF_UB=64)


def ref_row(period=200, an=(10, 210), st=(12, 215), gated=0):
    return {"gated": gated, "period": period, "periodNew": 205, "note": 5, "an": list(an), "st": list(st), "stale": 0, "beta": 0.97}


def eng_frame(period=200, an=(10, 210), st=(12, 215), flags=6):
    an, st = list(an), list(st)
    return types.SimpleNamespace(flags=flags, period=period, periodNew=205, note=5, nAn=len(an), nSt=len(st),
                                 anMarks=an + [0] * 4, stMarks=st + [0] * 4, anStale=0, beta=0.97)


def test_all_frames_counted_when_nothing_is_flagged():
    dec = compare_decisions(VP, [ref_row()] * 50, [eng_frame()] * 50)
    assert (dec.n, dec.bad, dec.flagged, dec.checked, dec.excused) == (50, 0, 0, 50, 0)
    assert_decisions(dec)


def test_a_difference_is_a_mismatch_and_names_the_frame():
    eng = [eng_frame()] * 20 + [eng_frame(period=201)] + [eng_frame()] * 20
    dec = compare_decisions(VP, [ref_row()] * 41, eng)
    assert dec.bad == 1 and dec.checked == 41 and dec.first.startswith("frame 20 ")
    with pytest.raises(AssertionError):
        assert_decisions(dec)


def test_flagged_frame_excuses_only_until_the_chains_agree_again():
    # frame 10 flagged; frames 11-12 differ (state carried from the ambiguous frame); from 13 on both agree again
    eng = [eng_frame()] * 10 + [eng_frame(flags=6 | 32)] + [eng_frame(an=(11, 211))] * 2 + [eng_frame()] * 87
    dec = compare_decisions(VP, [ref_row()] * 100, eng)
    assert (dec.flagged, dec.excused, dec.bad) == (1, 2, 0)
    assert dec.checked == 97
    # a stream that never re-synchronises is NOT silently accepted: almost nothing was compared
    eng = [eng_frame()] * 10 + [eng_frame(flags=6 | 64)] + [eng_frame(an=(11, 211))] * 89
    dec = compare_decisions(VP, [ref_row()] * 100, eng)
    assert dec.bad == 0 and dec.checked == 10 and dec.checked < CHECKED_FLOOR * dec.n
    with pytest.raises(AssertionError):
        assert_decisions(dec)


def test_mismatch_long_after_a_flag_counts():
    eng = [eng_frame(flags=6 | 16)] + [eng_frame()] * 30 + [eng_frame(st=(13, 215))] + [eng_frame()] * 30
    dec = compare_decisions(VP, [ref_row()] * 62, eng)
    assert dec.flagged == 1 and dec.bad == 1
