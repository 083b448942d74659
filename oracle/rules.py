"""Integer / f32 host rules of the reference's front end (oracle; test infrastructure only).

Every function cites the reference source it restates (paths relative to
/root/reference).  All "f32" arithmetic is done with numpy.float32 scalars so the
rounding is the reference's, not Python's double arithmetic.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Iterator, List, Optional, Tuple

import numpy as np

f32 = np.float32

# src/constants.rs:25,28,36,44,55,178,333,525-541
DEFAULT_MIN_CONFIDENCE = 0.1
DEFAULT_OVERLAP = 0.0
DEFAULT_BATCH_SIZE = 8
MIN_BATCH_SIZE = 1
MAX_BATCH_SIZE = 512
DEFAULT_TOP_K = 5
DEFAULT_RANGE_THRESHOLD = 0.01
BAT_SAMPLE_RATE = 256_000
BAT_CHUNK_SAMPLES = 144_000
BAT_SEGMENT_DURATION = f32(BAT_CHUNK_SAMPLES) / f32(BAT_SAMPLE_RATE)
BAT_OVERLAP_SAMPLES = BAT_CHUNK_SAMPLES // 4  # src/pipeline/processor.rs:506

# src/constants.rs:313-330, :397-399
WEEKS_PER_YEAR = 48
DAYS_PER_WEEK = f32(7.6)
YEAR_START_DAY = f32(1.0)
DAYS_IN_MONTH = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]


def _trunc_f32_to_usize(x: np.float32) -> int:
    """Rust ``f32 as usize``: truncate toward zero, saturate, NaN -> 0."""
    if np.isnan(x):
        return 0
    if x <= 0:
        return 0
    if np.isinf(x):
        return 2**64 - 1
    return int(x)


def segment_and_overlap_samples(segment_duration: float, overlap: float, target_rate: int,
                                bat_mode: bool = False) -> Tuple[int, int]:
    """``(segment_samples, overlap_samples)`` at the TARGET rate.

    src/pipeline/processor.rs:502-522 — f32 products truncated toward zero; bat mode
    uses the constants 144000 / 36000 regardless of the user's overlap.
    """
    if bat_mode:
        return BAT_CHUNK_SAMPLES, BAT_OVERLAP_SAMPLES
    seg = _trunc_f32_to_usize(f32(segment_duration) * f32(target_rate))
    ovl = _trunc_f32_to_usize(f32(overlap) * f32(target_rate))
    return seg, ovl


def source_window(segment_samples: int, overlap_samples: int, source_rate: int,
                  target_rate: int) -> Tuple[int, int]:
    """Source-rate window / overlap sizes.  src/pipeline/processor.rs:61-82 (f64, ceil)."""
    if source_rate == target_rate:
        return segment_samples, overlap_samples
    src_seg = math.ceil(float(segment_samples) * float(source_rate) / float(target_rate))
    src_ovl = math.ceil(float(overlap_samples) * float(source_rate) / float(target_rate))
    return int(src_seg), int(src_ovl)


@dataclass
class RawWindow:
    """One ``RawSegment`` (src/audio/decode.rs:27-32) described by indices only."""
    start_sample: int   # samples_emitted when the window was cut
    take: int           # real samples copied (the rest up to segment is zero padding)


def next_segment_table(total_frames: int, segment_samples: int,
                       overlap_samples: int) -> List[RawWindow]:
    """All windows ``StreamingDecoder::next_segment`` yields for a stream of
    ``total_frames`` mono samples.  src/audio/decode.rs:150-202.

    Raises ValueError when ``overlap >= segment`` (``:156-162`` returns Error::Internal).
    """
    if overlap_samples >= segment_samples:
        raise ValueError(
            f"overlap_samples ({overlap_samples}) must be less than segment_samples ({segment_samples})")
    out: List[RawWindow] = []
    buffered = total_frames      # whole file counted as already decoded: the FIFO logic
    emitted = 0                  # only looks at len(buffer) once it is >= segment or EOF
    while buffered > 0:
        take = min(segment_samples, buffered)
        out.append(RawWindow(emitted, take))
        advance = take - overlap_samples if take > overlap_samples else 0  # saturating_sub
        if advance > 0:
            buffered -= advance
            emitted += advance
        else:
            buffered = 0
            emitted += take
    return out


def iter_next_segment(stream: np.ndarray, segment_samples: int,
                      overlap_samples: int) -> Iterator[Tuple[np.ndarray, int]]:
    """Literal FIFO restatement of ``next_segment`` over a mono f32 stream, yielding
    ``(samples[segment_samples], start_sample)``.  src/audio/decode.rs:150-202.
    Used to cross-check :func:`next_segment_table`."""
    if overlap_samples >= segment_samples:
        raise ValueError("overlap_samples must be less than segment_samples")
    buf = np.asarray(stream, dtype=np.float32)
    emitted = 0
    while buf.size > 0:
        take = min(segment_samples, buf.size)
        seg = np.zeros(segment_samples, dtype=np.float32)
        seg[:take] = buf[:take]
        start = emitted
        advance = max(take - overlap_samples, 0)
        if advance > 0:
            buf = buf[advance:]
            emitted += advance
        else:
            buf = buf[:0]
            emitted += take
        yield seg, start


def chunk_times(start_sample: int, source_rate: int, segment_samples: int,
                target_rate: int) -> Tuple[np.float32, np.float32]:
    """``(start_time, end_time)`` of an AudioChunk.  src/pipeline/processor.rs:89-94 — f32."""
    start = f32(start_sample) / f32(source_rate)
    dur = f32(segment_samples) / f32(target_rate)
    return start, f32(start + dur)


def estimate_segment_count(duration_secs: Optional[float], segment_duration: float,
                           overlap: float) -> Optional[int]:
    """src/output/progress.rs:80-92 — f32 step, f64 division, ceil."""
    if duration_secs is None:
        return None
    step = f32(segment_duration) - f32(overlap)
    if not (step > 0):
        return None
    v = math.ceil(float(duration_secs) / float(step))
    return max(int(v), 0)


def effective_batch_size(batch_size: int, estimated_segments: Optional[int]) -> int:
    """src/pipeline/processor.rs:525-545 — min(batch, estimate), never 0."""
    if estimated_segments is None or estimated_segments == 0:
        return batch_size
    return estimated_segments if batch_size > estimated_segments else batch_size


def batch_layout(nseg: int, batch_size: int) -> List[Tuple[int, int, int]]:
    """``(first_segment, valid_count, padded_size)`` per submitted batch.

    src/pipeline/processor.rs:132-174 (collect ``batch_size`` chunks, flush the
    remainder) and :239-260 (pad with all-zero segments up to ``target_batch_size``)."""
    out = []
    i = 0
    while i < nseg:
        valid = min(batch_size, nseg - i)
        out.append((i, valid, max(valid, batch_size)))
        i += valid
    return out


# ---- date math (range-filter query), src/utils/date.rs:21-68 ---------------------------

def date_to_week(month: int, day: int) -> int:
    doy = sum(DAYS_IN_MONTH[: month - 1]) + day
    week = int(np.floor(f32(doy - 1) / DAYS_PER_WEEK)) + 1
    return min(week, WEEKS_PER_YEAR)


def day_of_year_to_date(day_of_year: int) -> Tuple[int, int]:
    remaining = day_of_year
    for m, dim in enumerate(DAYS_IN_MONTH):
        if remaining <= dim:
            return m + 1, remaining
        remaining -= dim
    return 12, 31


def week_to_start_day(week: int) -> int:
    # ((week-1) as f32).mul_add(7.6, 1.0) as u32 — fused multiply-add, single rounding
    exact = float(f32(week - 1)) * float(DAYS_PER_WEEK) + float(YEAR_START_DAY)  # exact in f64
    return int(f32(exact))


# ---- batch-size defaults, src/lib.rs:256-288 (tests :3183-3310) -----------------------------

def determine_default_batch_size(device: str, model_type: str) -> int:
    """device in {cpu, cuda, tensorrt, other}; model_type in {birdnet-v24, birdnet-v30, perch-v2, ...}.
    src/constants.rs:58-72."""
    if device == "cpu":
        return 8
    if device == "cuda":
        return 64 if model_type in ("birdnet-v24", "bsg-finland") else 32
    if device == "tensorrt":
        return 32
    return 16
