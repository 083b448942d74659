"""ctypes loader for libbirda_b200.so.  Fails loudly when the library has not been built."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.environ.get("BIRDA_B200_LIB") or os.path.join(_HERE, "libbirda_b200.so")   # override: A/B kernel variants


class BirdaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"birda_b200 error {code}: {message}")
        self.code = code
        self.message = message


if not os.path.exists(lib_path):
    raise ImportError(
        f"{lib_path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C birda_b200/csrc`). birda_b200 has no CPU fallback.")

lib = C.CDLL(lib_path)

u8p, u32p, u64p, i32p, i64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                               C.POINTER(C.c_int32), C.POINTER(C.c_int64))
f32p = C.POINTER(C.c_float)
vp = C.c_void_p


class WavInfo(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("fmt", C.c_int32), ("frames", C.c_uint64), ("data_offset", C.c_uint64)]


class PostCfg(C.Structure):
    _fields_ = [("activation", C.c_int32), ("min_confidence", C.c_float), ("top_k", C.c_uint32),
                ("range_threshold", C.c_float), ("keep_unmatched", C.c_int32), ("rerank", C.c_int32)]


class MelSpecCfg(C.Structure):
    _fields_ = [("n_fft", C.c_uint32), ("hop", C.c_uint32), ("n_frames", C.c_uint32), ("n_mels", C.c_uint32),
                ("power", C.c_float), ("log_mode", C.c_int32), ("log_eps", C.c_float)]


class PipelineCfg(C.Structure):
    _fields_ = [("target_rate", C.c_uint32), ("segment_duration", C.c_float), ("overlap", C.c_float),
                ("batch_size", C.c_uint32), ("bat_mode", C.c_int32), ("post", PostCfg),
                ("d_mask", C.c_void_p), ("d_species_keep", C.c_void_p)]


class DetectionC(C.Structure):
    _fields_ = [("segment", C.c_uint32), ("index", C.c_uint32), ("confidence", C.c_float),
                ("start_time", C.c_float), ("end_time", C.c_float)]


class PoolResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("device", C.c_int32), ("n_detections", C.c_uint64), ("n_segments", C.c_uint64),
                ("batch_used", C.c_uint32), ("detections", C.POINTER(DetectionC)), ("error", C.c_char * 200)]


class FlacInfo(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("min_block", C.c_uint32), ("max_block", C.c_uint32), ("min_frame_bytes", C.c_uint32), ("max_frame_bytes", C.c_uint32),
                ("frames", C.c_uint64), ("first_frame_offset", C.c_uint64), ("file_bytes", C.c_uint64), ("fmt", C.c_int32)]


WATCHDOG_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_uint32)
BATCH_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64)
CLASSIFY_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32))


# name -> (restype, argtypes).  tests/test_abi.py checks this list against include/birda_b200.h.
SIGNATURES = {
    "bb_rule_segment_samples": (C.c_int32, [C.c_float, C.c_float, C.c_uint32, C.c_int32, u64p, u64p]),
    "bb_rule_source_window": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, u64p, u64p]),
    "bb_rule_segment_count": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_uint64, u64p]),
    "bb_rule_segment_table": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p, u64p]),
    "bb_rule_chunk_times": (C.c_int32, [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, f32p, f32p]),
    "bb_rule_estimate_segment_count": (C.c_int32, [C.c_double, C.c_int32, C.c_float, C.c_float, i64p]),
    "bb_rule_effective_batch_size": (C.c_uint32, [C.c_uint32, C.c_int64]),
    "bb_rule_resampler_blocks": (C.c_int32, [C.c_uint32, C.c_uint32, u32p, u32p, u32p, f32p]),
    "bb_rule_resampler_taps": (C.c_int32, [C.c_uint32, C.c_uint32, f32p, C.c_uint32]),
    "bb_rule_resampled_len": (C.c_int32, [C.c_uint64, C.c_uint32, C.c_uint32, u64p]),
    "bb_rule_date_to_week": (C.c_uint32, [C.c_uint32, C.c_uint32]),
    "bb_rule_week_to_start_day": (C.c_uint32, [C.c_uint32]),
    "bb_rule_day_of_year_to_date": (None, [C.c_uint32, u32p, u32p]),
    "bb_rule_inference_timeout_secs": (C.c_uint64, [C.c_char_p]),
    "bb_rule_scientific_name_len": (C.c_uint32, [C.c_char_p]),
    "bb_mask_build": (C.c_int32, [C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.c_char_p), C.c_uint32,
                                  C.POINTER(C.c_char_p), f32p, C.c_uint32, f32p, u32p, u32p]),
    "bb_watchdog_start": (C.c_int32, [C.c_uint64, C.c_uint32, WATCHDOG_FN, vp, C.POINTER(vp)]),
    "bb_watchdog_cancel": (None, [vp]),
    "bb_debug_inject_alloc_failure": (None, [C.c_int32]),
    "bb_version": (C.c_uint32, []),
    "bb_device_count": (C.c_int32, [i32p]),
    "bb_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(vp)]),
    "bb_ctx_create_on_stream": (C.c_int32, [C.c_int32, vp, C.POINTER(vp)]),
    "bb_ctx_destroy": (None, [vp]),
    "bb_last_error": (C.c_char_p, [vp]),
    "bb_ctx_stream": (vp, [vp]),
    "bb_sync": (C.c_int32, [vp]),
    "bb_ctx_kernel_launches": (C.c_uint64, [vp]),
    "bb_ctx_set_blocking_sync": (None, [vp, C.c_int32]),
    "bb_host_alloc": (C.c_int32, [C.c_uint64, C.POINTER(vp)]),
    "bb_host_free": (None, [vp]),
    "bb_plan_create": (C.c_int32, [vp, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(vp)]),
    "bb_plan_destroy": (None, [vp]),
    "bb_plan_source_window": (C.c_int32, [vp, u64p, u64p]),
    "bb_plan_segment_count": (C.c_int32, [vp, C.c_uint64, u64p]),
    "bb_plan_describe": (C.c_int32, [vp, C.c_char_p, C.c_uint32]),
    "bb_frontend_run": (C.c_int32, [vp, vp, C.c_uint64, C.c_int32, C.c_uint64, C.c_int32, C.c_uint32, vp, C.c_uint64,
                                    C.POINTER(vp), u64p, f32p, f32p, u64p, u64p, u64p]),
    "bb_post_run_device": (C.c_int32, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(PostCfg), vp, vp, vp, vp, vp]),
    "bb_post_run": (C.c_int32, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(PostCfg), vp, vp, u32p, f32p, u32p]),
    "bb_wav_probe": (C.c_int32, [C.c_char_p, C.POINTER(WavInfo)]),
    "bb_wav_read": (C.c_int32, [C.c_char_p, C.POINTER(WavInfo), C.c_uint64, C.c_uint64, vp]),
    "bb_wav_read_parallel": (C.c_int32, [C.c_char_p, C.POINTER(WavInfo), C.c_uint64, C.c_uint64, vp, C.c_uint32]),
    "bb_pipeline_set_read_threads": (None, [vp, C.c_uint32]),
    "bb_pipeline_create": (C.c_int32, [vp, C.POINTER(PipelineCfg), CLASSIFY_FN, vp, C.POINTER(vp)]),
    "bb_pipeline_destroy": (None, [vp]),
    "bb_pipeline_last_error": (C.c_char_p, [vp]),
    "bb_pipeline_plans_created": (C.c_uint64, [vp]),
    "bb_pipeline_set_sync_before_classify": (None, [vp, C.c_int32]),
    "bb_pipeline_set_batch_hooks": (None, [vp, BATCH_HOOK, BATCH_HOOK, vp]),
    "bb_pipeline_set_batch_timeout": (None, [vp, C.c_uint64, WATCHDOG_FN, vp]),
    "bb_pipeline_process_pcm": (C.c_int32, [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(DetectionC),
                                            C.c_uint64, u64p, u64p, u32p]),
    "bb_pipeline_process_wav": (C.c_int32, [vp, C.c_char_p, C.c_uint64, C.POINTER(DetectionC), C.c_uint64, u64p, u64p, u32p]),
    "bb_pool_create": (C.c_int32, [C.POINTER(C.c_int32), C.c_uint32, C.POINTER(PipelineCfg), CLASSIFY_FN, C.POINTER(vp), C.POINTER(vp)]),
    "bb_pool_destroy": (None, [vp]),
    "bb_pool_process_wavs": (C.c_int32, [vp, C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(PoolResult)]),
    "bb_pool_free_results": (None, [C.POINTER(PoolResult), C.c_uint32]),
    "bb_pool_kernel_launches": (C.c_uint64, [vp]),
    "bb_pool_worker_ctx": (vp, [vp, C.c_uint32]),
    "bb_pool_set_stream_ordered": (None, [vp, C.c_int32]),
    "bb_dense_run": (C.c_int32, [vp, vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_int32, vp]),
    "bb_calibrate_run": (C.c_int32, [vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp]),
    "bb_melspec_create": (C.c_int32, [vp, C.POINTER(MelSpecCfg), f32p, f32p, C.POINTER(vp)]),
    "bb_melspec_destroy": (None, [vp]),
    "bb_melspec_info": (C.c_int32, [vp, u32p, u32p, u32p]),
    "bb_melspec_run": (C.c_int32, [vp, vp, C.c_uint32, C.c_uint32, vp]),
    "bb_flac_probe": (C.c_int32, [C.c_char_p, C.POINTER(FlacInfo)]),
    "bb_flac_probe_bytes": (C.c_int32, [vp, C.c_uint64, C.POINTER(FlacInfo)]),
    "bb_flac_index": (C.c_int32, [vp, C.c_uint64, C.POINTER(FlacInfo), u64p, u64p, u32p, C.c_uint64, u64p]),
    "bb_flac_create": (C.c_int32, [vp, C.POINTER(vp)]),
    "bb_flac_destroy": (None, [vp]),
    "bb_flac_decode": (C.c_int32, [vp, vp, C.c_uint64, C.POINTER(FlacInfo), C.POINTER(vp), u64p]),
    "bb_standin_create": (C.c_int32, [C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(vp)]),
    "bb_standin_destroy": (None, [vp]),
    "bb_standin_use_stream": (None, [vp, vp, C.c_int32]),
    "bb_standin_weights": (C.c_int32, [vp, f32p, f32p]),
    "bb_standin_launches": (C.c_uint64, [vp]),
    "bb_standin_classify": (C.c_int32, [vp, vp, C.c_uint32, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint32)]),
    "bb_dev_alloc": (C.c_int32, [vp, C.c_uint64, C.POINTER(vp)]),
    "bb_dev_free": (None, [vp, vp]),
    "bb_memcpy_h2d": (C.c_int32, [vp, vp, vp, C.c_uint64]),
    "bb_memcpy_d2h": (C.c_int32, [vp, vp, vp, C.c_uint64]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here == the .so does not export the symbol
    _fn.restype = _res
    _fn.argtypes = _args


def check(code: int, ctx=None) -> None:
    if code != 0:
        msg = lib.bb_last_error(ctx)
        raise BirdaError(code, msg.decode("utf-8", "replace") if msg else "")
