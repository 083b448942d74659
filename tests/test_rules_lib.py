"""Host rules inside the library (bb_rule_*) against the oracle (CPU only, no compute kernels)."""
import numpy as np
import pytest

import birda_b200 as b
from oracle import frontend as ofe
from oracle import rules as orules


def test_segment_samples_all_two_decimal_overlaps():
    # 70 of the 2000 two-decimal overlaps differ from exact arithmetic
    for rate, dur in ((48_000, 3.0), (32_000, 5.0), (44_100, 3.0), (256_000, 3.0)):
        for i in range(500):
            ovl = i / 100.0
            assert b.rules.segment_samples(dur, ovl, rate) == orules.segment_and_overlap_samples(dur, ovl, rate)
    assert b.rules.segment_samples(0.5625, 0.3, 256_000, bat_mode=True) == (144_000, 36_000)
    assert b.rules.segment_samples(float("nan"), -1.0, 48_000) == (0, 0)


def test_source_window_and_counts_random():
    rng = np.random.default_rng(11)
    rates = [8000, 11025, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000, 192000, 256000]
    for _ in range(300):
        sr, tr = int(rng.choice(rates)), int(rng.choice([32000, 48000]))
        seg = int(rng.integers(1000, 200_000)); ovl = int(rng.integers(0, seg))
        assert b.rules.source_window(seg, ovl, sr, tr) == orules.source_window(seg, ovl, sr, tr)
    for _ in range(400):
        seg = int(rng.integers(1, 5000)); ovl = int(rng.integers(0, seg)); total = int(rng.integers(0, 40_000))
        tab = orules.next_segment_table(total, seg, ovl)
        assert b.rules.segment_count(total, seg, ovl) == len(tab)
        st, tk = b.rules.segment_table(total, seg, ovl)
        assert [int(x) for x in st] == [w.start_sample for w in tab]
        assert [int(x) for x in tk] == [w.take for w in tab]


def test_segment_count_configs_and_error():
    assert b.rules.segment_count(158_760_000, 132_300, 66_150) == 2400
    assert b.rules.segment_count(921_600_000, 144_000, 36_000) == 8534
    assert b.rules.segment_count(0, 144_000, 0) == 0
    with pytest.raises(b.BirdaError) as e:
        b.rules.segment_count(1000, 100, 100)
    assert e.value.code == -2 and "must be less than segment_samples" in e.value.message


def test_chunk_times_bit_exact():
    rng = np.random.default_rng(3)
    for _ in range(2000):
        ss = int(rng.integers(0, 2**40)); sr = int(rng.choice([44100, 48000, 256000, 22050]))
        a = b.rules.chunk_times(ss, sr, 144_000, 48_000)
        o = orules.chunk_times(ss, sr, 144_000, 48_000)
        assert a[0].tobytes() == o[0].tobytes() and a[1].tobytes() == o[1].tobytes()


def test_estimate_and_effective_batch():
    for d, s, o in ((10.0, 3.0, 0.0), (10.0, 3.0, 1.0), (None, 3.0, 0.0), (10.0, 3.0, 3.0), (3600.0, 3.0, 1.5),
                    (0.0, 3.0, 0.0), (44739.0, 5.0, 0.0), (59.99, 3.0, 2.9)):
        assert b.rules.estimate_segment_count(d, s, o) == orules.estimate_segment_count(d, s, o)
    for bs, est in ((64, 2400), (64, 20), (64, 0), (64, None), (1, 5), (512, 511)):
        assert b.rules.effective_batch_size(bs, est) == orules.effective_batch_size(bs, est)


def test_date_rules():
    for m in range(1, 13):
        for d in range(1, orules.DAYS_IN_MONTH[m - 1] + 1):
            assert b.rules.date_to_week(m, d) == orules.date_to_week(m, d)
    for w in range(1, 49):
        assert b.rules.week_to_start_day(w) == orules.week_to_start_day(w)
    for doy in (1, 31, 32, 59, 60, 166, 365, 366, 400):
        assert b.rules.day_of_year_to_date(doy) == orules.day_of_year_to_date(doy)


@pytest.mark.parametrize("sr,tr", [(44_100, 48_000), (48_000, 32_000), (44_100, 32_000), (22_050, 48_000),
                                   (96_000, 48_000), (16_000, 48_000), (32_000, 48_000), (250_000, 48_000),
                                   (8_000, 48_000), (192_000, 32_000)])
def test_resampler_spec_matches_oracle(sr, tr):
    p = ofe.make_plan(sr, tr)
    n_in, n_out, n_keep, cutoff = b.rules.resampler_blocks(sr, tr)
    assert (n_in, n_out, n_keep) == (p.n_in, p.n_out, p.n_keep)
    assert abs(float(cutoff) - float(p.cutoff)) <= 2e-7
    taps = b.rules.resampler_taps(sr, tr)
    # independent f32 builds of the same windowed sinc: equal to a few ulp of the peak tap
    assert np.abs(taps - p.taps).max() <= 4e-7 * np.abs(p.taps).max()
    for n in (n_in * 3, n_in * 3 + 1, 132_300, 240_000, 5):
        assert b.rules.resampled_len(n, sr, tr) == ofe.resampled_len(n, p)


def test_unsupported_rate_is_reported():
    with pytest.raises(b.BirdaError) as e:
        b.rules.resampler_blocks(44_101, 48_000)     # gcd 1 -> a 44101-sample block
    assert e.value.code == -3
