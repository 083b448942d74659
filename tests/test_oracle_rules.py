"""Oracle host rules against the reference's own unit-test tables (CPU only).

Ports: src/audio/chunker.rs:83-124, src/output/progress.rs:172-185, src/utils/date.rs
tests, src/lib.rs:3183-3310 plus hand-derived next_segment tables."""
import numpy as np
import pytest

from oracle import rules


def test_estimate_segment_count_reference_table():
    # src/output/progress.rs:172-185
    assert rules.estimate_segment_count(10.0, 3.0, 0.0) == 4
    assert rules.estimate_segment_count(10.0, 3.0, 1.0) == 5
    assert rules.estimate_segment_count(None, 3.0, 0.0) is None
    assert rules.estimate_segment_count(10.0, 3.0, 3.0) is None
    assert rules.estimate_segment_count(10.0, 3.0, 4.0) is None


def test_chunk_audio_tables_match_next_segment_without_overlap():
    # src/audio/chunker.rs:87-93: 2 s @ 48 kHz, 1 s chunks -> starts 0.0, 1.0
    t = rules.next_segment_table(96_000, 48_000, 0)
    assert [w.start_sample for w in t] == [0, 48_000]
    assert [float(rules.chunk_times(w.start_sample, 48_000, 48_000, 48_000)[0]) for w in t] == [0.0, 1.0]
    # :107-113: 1.25 s pads the final chunk
    t = rules.next_segment_table(60_000, 48_000, 0)
    assert [(w.start_sample, w.take) for w in t] == [(0, 48_000), (48_000, 12_000)]
    # :116-120: empty input
    assert rules.next_segment_table(0, 48_000, 0) == []


def test_next_segment_overlap_ge_segment_is_error():
    # src/audio/decode.rs:156-162
    with pytest.raises(ValueError):
        rules.next_segment_table(1000, 100, 100)
    with pytest.raises(ValueError):
        rules.next_segment_table(1000, 100, 101)


def test_next_segment_tail_semantics():
    # after the last full window the retained overlap is emitted again
    t = rules.next_segment_table(1000, 400, 100)   # hop 300
    assert [(w.start_sample, w.take) for w in t] == [(0, 400), (300, 400), (600, 400), (900, 100)]
    t = rules.next_segment_table(1050, 400, 100)
    # 0,300,600 full; remaining 150 -> take 150 (advance 50) ; remaining 100 -> take 100 <= ovl -> clear
    assert [(w.start_sample, w.take) for w in t] == [(0, 400), (300, 400), (600, 400), (900, 150), (950, 100)]
    t = rules.next_segment_table(50, 400, 100)
    assert [(w.start_sample, w.take) for w in t] == [(0, 50)]


def test_next_segment_config_counts():
    # SURVEY §8 sizes
    assert len(rules.next_segment_table(2_880_000, 144_000, 0)) == 20                 # C1
    t = rules.next_segment_table(158_760_000, 132_300, 66_150)                          # C2
    assert len(t) == 2400 and t[-1].take == 66_150 and t[-1].start_sample == 2399 * 66_150
    assert len(rules.next_segment_table(172_800_000, 240_000, 0)) == 720               # C3
    assert len(rules.next_segment_table(921_600_000, 144_000, 36_000)) == 8534         # C4 1 h
    assert len(rules.next_segment_table(153_600_000, 144_000, 36_000)) == 1423         # C4 600 s


@pytest.mark.parametrize("total,seg,ovl", [(1000, 400, 100), (1050, 400, 100), (7, 3, 2), (10, 3, 0),
                                           (12345, 1000, 999), (5000, 1024, 512), (1, 5, 4)])
def test_table_matches_literal_fifo(total, seg, ovl):
    stream = np.arange(1, total + 1, dtype=np.float32)
    lit = list(rules.iter_next_segment(stream, seg, ovl))
    tab = rules.next_segment_table(total, seg, ovl)
    assert len(lit) == len(tab)
    for (samples, start), w in zip(lit, tab):
        assert start == w.start_sample
        ref = np.zeros(seg, np.float32)
        ref[: w.take] = stream[w.start_sample: w.start_sample + w.take]
        assert np.array_equal(samples, ref)


def test_f32_truncated_sample_counts():
    # f32 products truncated toward zero (processor.rs:514-521)
    assert rules.segment_and_overlap_samples(3.0, 1.3, 48_000) == (144_000, 62_399)
    assert rules.segment_and_overlap_samples(3.0, 0.53, 48_000)[1] == 25_439
    assert rules.segment_and_overlap_samples(3.0, 1.05, 48_000)[1] == 50_399
    assert rules.segment_and_overlap_samples(3.0, 2.1, 48_000)[1] == 100_799
    assert rules.segment_and_overlap_samples(3.0, 1.5, 48_000) == (144_000, 72_000)
    assert rules.segment_and_overlap_samples(5.0, 0.0, 32_000) == (160_000, 0)
    assert rules.segment_and_overlap_samples(0.5625, 2.0, 256_000, bat_mode=True) == (144_000, 36_000)


def test_source_window():
    assert rules.source_window(144_000, 72_000, 44_100, 48_000) == (132_300, 66_150)
    assert rules.source_window(160_000, 0, 48_000, 32_000) == (240_000, 0)
    assert rules.source_window(144_000, 62_399, 44_100, 48_000) == (132_300, 57_330)  # ceil(57329.08)
    assert rules.source_window(144_000, 0, 48_000, 48_000) == (144_000, 0)
    assert rules.source_window(144_000, 36_000, 256_000, 256_000) == (144_000, 36_000)


def test_chunk_times_f32():
    s, e = rules.chunk_times(66_150 * 3, 44_100, 144_000, 48_000)
    assert s.dtype == np.float32 and float(s) == 4.5 and float(e) == 7.5
    # beyond 2^24 samples the int->f32 cast rounds to nearest even (SURVEY §5 long context)
    s, _ = rules.chunk_times(2**24 + 1, 44_100, 144_000, 48_000)
    assert float(s) == float(np.float32(2**24) / np.float32(44_100))
    s, _ = rules.chunk_times(2**24 + 3, 44_100, 144_000, 48_000)
    assert float(s) == float(np.float32(2**24 + 4) / np.float32(44_100))


def test_effective_batch_and_layout():
    assert rules.effective_batch_size(64, 2400) == 64
    assert rules.effective_batch_size(64, 20) == 20
    assert rules.effective_batch_size(64, 0) == 64
    assert rules.effective_batch_size(64, None) == 64
    lay = rules.batch_layout(2400, 64)
    assert len(lay) == 38 and lay[-1] == (2368, 32, 64) and lay[0] == (0, 64, 64)
    assert rules.batch_layout(20, 8) == [(0, 8, 8), (8, 8, 8), (16, 4, 8)]


def test_date_math_reference_table():
    # src/utils/date.rs tests
    assert rules.date_to_week(1, 1) == 1
    assert rules.date_to_week(12, 31) == 48
    assert rules.date_to_week(6, 15) == 22
    assert rules.date_to_week(7, 1) == 24
    assert rules.week_to_start_day(1) == 1
    assert rules.week_to_start_day(24) == 175
    assert rules.week_to_start_day(48) == 358
    assert rules.day_of_year_to_date(1) == (1, 1)
    assert rules.day_of_year_to_date(365) == (12, 31)
    assert rules.day_of_year_to_date(166) == (6, 15)
    assert rules.day_of_year_to_date(400) == (12, 31)


def test_default_batch_sizes():
    # src/lib.rs:3183-3310
    assert rules.determine_default_batch_size("cpu", "birdnet-v24") == 8
    assert rules.determine_default_batch_size("cuda", "birdnet-v24") == 64
    assert rules.determine_default_batch_size("cuda", "bsg-finland") == 64
    assert rules.determine_default_batch_size("cuda", "birdnet-v30") == 32
    assert rules.determine_default_batch_size("cuda", "perch-v2") == 32
    assert rules.determine_default_batch_size("tensorrt", "perch-v2") == 32
    assert rules.determine_default_batch_size("other", "birdnet-v24") == 16
