"""Resampler restatement against the reference's own property tests
(src/audio/resample.rs:240-385), ported test by test (CPU only).  These are the only
behavioural pins the reference holds for the resampler; sample-level parity is UNPINNED."""
import math

import numpy as np
import pytest

from oracle import frontend as fe

PI = np.float32(np.pi)
TEST_RATE_HIGH, TEST_RATE_LOW, TEST_RATE_CD = 48_000, 32_000, 44_100
TEST_SIGNAL_LEN, TEST_SIGNAL_LEN_CD = 48_000, 44_100
BIRD_BAND_HZ, REFERENCE_TONE_HZ, ABOVE_NYQUIST_HZ, ALIAS_IMAGE_HZ = 6000.0, 1000.0, 20000.0, 12000.0
DOMINANCE_RATIO, MIN_TONE_POWER_FRACTION = 100.0, 0.5
FILTERED_RMS_CEILING, PRESERVED_RMS_FLOOR, RMS_TOLERANCE = 0.1, 0.6, 0.05
STEADY_STATE_MARGIN, ALIAS_POWER_FRACTION = 8, 1e-6


def sine(freq, rate, n):
    i = np.arange(n, dtype=np.float32)
    return np.sin((np.float32(2.0) * PI * np.float32(freq) * i / np.float32(rate)).astype(np.float32)).astype(np.float32)


def tone_power(samples, rate, freq):
    """Goertzel, resample.rs:188-201 (f64 here: it is the measuring instrument, not the path)."""
    n = float(len(samples))
    k = round(n * freq / rate)
    w = 2.0 * math.pi * k / n
    coeff = 2.0 * math.cos(w)
    s1 = s2 = 0.0
    for x in samples.astype(np.float64):
        s0 = coeff * s1 + x - s2
        s2, s1 = s1, s0
    return max(s1 * s1 + s2 * s2 - coeff * s1 * s2, 0.0) / n


def rms(x):
    return float(np.sqrt(np.mean(x.astype(np.float64) ** 2)))


def steady_state(x):
    m = len(x) // STEADY_STATE_MARGIN
    return x[m: len(x) - m]


def assert_tone_intact(body, rate, tone, others):
    at = tone_power(body, rate, tone)
    assert at > len(body) / 4.0 * MIN_TONE_POWER_FRACTION
    for o in others:
        assert at > tone_power(body, rate, o) * DOMINANCE_RATIO


def test_resample_preserves_tone_frequency():
    out = fe.resample(sine(REFERENCE_TONE_HZ, TEST_RATE_HIGH, TEST_SIGNAL_LEN), TEST_RATE_HIGH, TEST_RATE_LOW)
    assert_tone_intact(steady_state(out), TEST_RATE_LOW, REFERENCE_TONE_HZ, [500.0, 2000.0, 4000.0])


def test_resample_preserves_bird_band_content():
    out = fe.resample(sine(BIRD_BAND_HZ, TEST_RATE_HIGH, TEST_SIGNAL_LEN), TEST_RATE_HIGH, TEST_RATE_LOW)
    body = steady_state(out)
    assert_tone_intact(body, TEST_RATE_LOW, BIRD_BAND_HZ, [3000.0, 9000.0, 12000.0])
    assert rms(body) > PRESERVED_RMS_FLOOR


def test_resample_filters_content_above_output_nyquist():
    out = fe.resample(sine(ABOVE_NYQUIST_HZ, TEST_RATE_HIGH, TEST_SIGNAL_LEN), TEST_RATE_HIGH, TEST_RATE_LOW)
    body = steady_state(out)
    assert tone_power(body, TEST_RATE_LOW, ALIAS_IMAGE_HZ) < len(body) / 4.0 * ALIAS_POWER_FRACTION
    assert rms(body) < FILTERED_RMS_CEILING


def test_resample_from_cd_rate_filters_above_output_nyquist():
    out = fe.resample(sine(ABOVE_NYQUIST_HZ, TEST_RATE_CD, TEST_SIGNAL_LEN_CD), TEST_RATE_CD, TEST_RATE_LOW)
    assert rms(steady_state(out)) < FILTERED_RMS_CEILING


def test_resample_from_cd_rate_preserves_bird_band_content():
    out = fe.resample(sine(BIRD_BAND_HZ, TEST_RATE_CD, TEST_SIGNAL_LEN_CD), TEST_RATE_CD, TEST_RATE_LOW)
    assert_tone_intact(steady_state(out), TEST_RATE_LOW, BIRD_BAND_HZ, [3000.0, 9000.0, 12000.0])


def test_resample_preserves_amplitude():
    x = sine(REFERENCE_TONE_HZ, TEST_RATE_HIGH, TEST_SIGNAL_LEN)
    out = fe.resample(x, TEST_RATE_HIGH, TEST_RATE_LOW)
    assert abs(rms(steady_state(out)) - rms(x)) < RMS_TOLERANCE


def test_resample_same_rate_returns_input():
    x = np.array([0.1, 0.2, 0.3, 0.4, 0.5], np.float32)
    assert np.array_equal(fe.resample(x, 48000, 48000), x)


def test_resample_lengths():
    i = np.arange(48000, dtype=np.float32)
    out = fe.resample(np.sin(i * np.float32(0.001)), 48000, 32000)
    assert 20000 < out.size < 35000 and out.size == 32000
    out = fe.resample(np.sin(i[:32000] * np.float32(0.001)), 32000, 48000)
    assert 45000 < out.size < 55000


def test_block_sizes_match_reference_comment():
    # src/audio/resample.rs:313-316: 48k/32k share 16000 -> 342 fft chunks; 44.1k/32k share 100 -> 3
    p = fe.make_plan(48_000, 32_000)
    assert (p.n_in, p.n_out) == (342 * 3, 342 * 2)
    p = fe.make_plan(44_100, 32_000)
    assert (p.n_in, p.n_out) == (3 * 441, 3 * 320)
    p = fe.make_plan(44_100, 48_000)
    assert (p.n_in, p.n_out) == (1029, 1120)


@pytest.mark.parametrize("sr,tr,seg", [(44_100, 48_000, 144_000), (48_000, 32_000, 160_000),
                                        (22_050, 48_000, 144_000), (96_000, 48_000, 144_000)])
def test_resampled_len_fills_model_window(sr, tr, seg):
    # C2 -> 143360 + 640 = 144000; C3 -> 159372 + 628 = 160000
    from oracle import rules
    src_seg, _ = rules.source_window(seg, 0, sr, tr)
    p = fe.make_plan(sr, tr)
    assert fe.resampled_len(src_seg, p) == seg
    x = np.random.default_rng(1).standard_normal(src_seg).astype(np.float32)
    assert fe.resample(x, sr, tr, p).size == seg


def test_f32_path_close_to_f64_path():
    x = (np.random.default_rng(2).standard_normal(132_300) * 0.1).astype(np.float32)
    a = fe.resample(x, 44_100, 48_000, precision="f32")
    b = fe.resample(x, 44_100, 48_000, precision="f64")
    scale = max(float(np.abs(b).max()), rms(b))
    assert float(np.abs(a - b).max()) <= 3e-6 * scale
