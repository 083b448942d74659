//! Raw bindings to `include/birda_b200.h` and the safe wrappers birda's pipeline calls.
//! Untested source (no Rust toolchain in the build image); the C ABI is what the tests exercise.
//!
//! `Cargo.toml:76-77` sets `unsafe_code = "deny"`, so this module carries a scoped allow, as
//! `src/update/replace.rs:11` does for `libc::getuid`.
#![allow(unsafe_code)]

use std::ffi::{c_char, c_void, CStr};

use crate::error::{Error, Result};

#[repr(C)] pub struct bb_ctx { _p: [u8; 0] }
#[repr(C)] pub struct bb_plan { _p: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct bb_post_cfg {
    pub activation: i32,
    pub min_confidence: f32,
    pub top_k: u32,
    pub range_threshold: f32,
    pub keep_unmatched: i32,
    pub rerank: i32,
}

pub const BB_S16: i32 = 1;
pub const BB_S32: i32 = 2;
pub const BB_F32: i32 = 3;
pub const BB_ERR_OVERLAP_GE_SEGMENT: i32 = -2;
pub const BB_ERR_UNSUPPORTED_RATE: i32 = -3;

#[repr(C)] pub struct bb_melspec { _p: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct bb_melspec_cfg {
    pub n_fft: u32,
    pub hop: u32,
    pub n_frames: u32,
    pub n_mels: u32,
    pub power: f32,
    pub log_mode: i32,
    pub log_eps: f32,
}

extern "C" {
    pub fn bb_melspec_create(ctx: *mut bb_ctx, cfg: *const bb_melspec_cfg, window: *const f32, mel_weights: *const f32,
                             out: *mut *mut bb_melspec) -> i32;
    pub fn bb_melspec_run(ms: *mut bb_melspec, d_segments: *const f32, rows: u32, samples: u32, d_out: *mut f32) -> i32;
    pub fn bb_melspec_destroy(ms: *mut bb_melspec);
    pub fn bb_ctx_create(device: i32, out: *mut *mut bb_ctx) -> i32;
    pub fn bb_ctx_create_on_stream(device: i32, stream: *mut c_void, out: *mut *mut bb_ctx) -> i32;
    pub fn bb_ctx_destroy(ctx: *mut bb_ctx);
    pub fn bb_last_error(ctx: *const bb_ctx) -> *const c_char;
    pub fn bb_sync(ctx: *mut bb_ctx) -> i32;
    pub fn bb_host_alloc(bytes: u64, out: *mut *mut c_void) -> i32;
    pub fn bb_host_free(p: *mut c_void);
    pub fn bb_plan_create(ctx: *mut bb_ctx, src_rate: u32, channels: u32, fmt: i32, tgt_rate: u32,
                          segment_samples: u64, overlap_samples: u64, out: *mut *mut bb_plan) -> i32;
    pub fn bb_plan_destroy(plan: *mut bb_plan);
    pub fn bb_frontend_run(plan: *mut bb_plan, pcm: *const c_void, frames: u64, pcm_is_device: i32,
                           first_start_sample: u64, is_eof: i32, pad_to_batch: u32,
                           d_out_user: *mut f32, capacity_rows: u64, d_segments: *mut *mut f32,
                           start_sample: *mut u64, start_time: *mut f32, end_time: *mut f32,
                           nseg_out: *mut u64, nseg_padded: *mut u64, consumed_frames: *mut u64) -> i32;
    pub fn bb_post_run(ctx: *mut bb_ctx, d_scores: *const f32, b: u32, c: u32, valid_b: u32,
                       cfg: *const bb_post_cfg, d_mask: *const f32, d_species_keep: *const u8,
                       h_index: *mut u32, h_conf: *mut f32, h_count: *mut u32) -> i32;
}

fn check(ctx: *const bb_ctx, code: i32) -> Result<()> {
    if code == 0 { return Ok(()); }
    // SAFETY: bb_last_error returns a NUL-terminated string owned by the context (or a thread local).
    let msg = unsafe { CStr::from_ptr(bb_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(match code {
        BB_ERR_OVERLAP_GE_SEGMENT => Error::Internal { message: msg },          // decode.rs:156-162
        BB_ERR_UNSUPPORTED_RATE => Error::Resample { reason: msg },             // resample.rs:26-28
        -5 | -6 | -8 => Error::Inference { reason: msg },
        _ => Error::Internal { message: msg },
    })
}

/// One GPU.  Owned by the thread that owns the `BirdClassifier` (processor.rs:659-671).
pub struct GpuContext { raw: *mut bb_ctx }

impl GpuContext {
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        // SAFETY: out pointer is valid for the call.
        check(std::ptr::null(), unsafe { bb_ctx_create(device, &mut raw) })?;
        Ok(Self { raw })
    }
    pub fn sync(&self) -> Result<()> { check(self.raw, unsafe { bb_sync(self.raw) }) }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { bb_ctx_destroy(self.raw) } } }

/// A batch of packed segments on the device plus the host-side `AudioChunk` time stamps
/// (`src/audio/chunker.rs:5-12`): what replaces the per-segment `tx.send(Ok(chunk))`.
pub struct DeviceSegments {
    pub device_ptr: *mut f32,
    pub rows: usize,
    pub valid: usize,
    pub start_time: Vec<f32>,
    pub end_time: Vec<f32>,
}

/// Replaces `StreamingDecoder::next_segment` + `resample_chunk` + resize for one file.
pub struct GpuFrontEnd { raw: *mut bb_plan, ctx: *const bb_ctx }

impl GpuFrontEnd {
    pub fn new(ctx: &GpuContext, src_rate: u32, channels: u32, fmt: i32, tgt_rate: u32,
               segment_samples: usize, overlap_samples: usize) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(ctx.raw, unsafe {
            bb_plan_create(ctx.raw, src_rate, channels, fmt, tgt_rate, segment_samples as u64, overlap_samples as u64, &mut raw)
        })?;
        Ok(Self { raw, ctx: ctx.raw })
    }

    /// `pcm`: interleaved decoded frames in pinned host memory (see `bb_host_alloc`).
    pub fn run(&mut self, pcm: &[u8], frames: u64, first_start_sample: u64, is_eof: bool, batch: u32,
               max_rows: usize) -> Result<(DeviceSegments, u64)> {
        let mut st = vec![0f32; max_rows];
        let mut et = vec![0f32; max_rows];
        let (mut d, mut n, mut rows, mut consumed) = (std::ptr::null_mut(), 0u64, 0u64, 0u64);
        check(self.ctx, unsafe {
            bb_frontend_run(self.raw, pcm.as_ptr().cast(), frames, 0, first_start_sample, is_eof as i32, batch,
                            std::ptr::null_mut(), max_rows as u64, &mut d, std::ptr::null_mut(),
                            st.as_mut_ptr(), et.as_mut_ptr(), &mut n, &mut rows, &mut consumed)
        })?;
        st.truncate(n as usize); et.truncate(n as usize);
        Ok((DeviceSegments { device_ptr: d, rows: rows as usize, valid: n as usize, start_time: st, end_time: et }, consumed))
    }
}
impl Drop for GpuFrontEnd { fn drop(&mut self) { unsafe { bb_plan_destroy(self.raw) } } }

// ---- per-file pipeline and multi-GPU file pool (csrc/pipeline.cpp, csrc/pool.cpp) -------------------------------
#[repr(C)] pub struct bb_pipeline { _p: [u8; 0] }
#[repr(C)] pub struct bb_pool { _p: [u8; 0] }

/// `BirdClassifier::predict_batch_device` behind a C trampoline: device windows in, device scores out.
pub type bb_classify_fn = unsafe extern "C" fn(user: *mut c_void, d_segments: *const f32, batch_rows: u32, samples: u32,
                                               d_scores: *mut *const f32, classes: *mut u32) -> i32;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct bb_pipeline_cfg {
    pub target_rate: u32,
    pub segment_duration: f32,
    pub overlap: f32,
    pub batch_size: u32,
    pub bat_mode: i32,
    pub post: bb_post_cfg,
    pub d_mask: *const f32,
    pub d_species_keep: *const u8,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct bb_detection {
    pub segment: u32,
    pub index: u32,
    pub confidence: f32,
    pub start_time: f32,
    pub end_time: f32,
}

#[repr(C)]
pub struct bb_pool_result {
    pub status: i32,
    pub device: i32,
    pub n_detections: u64,
    pub n_segments: u64,
    pub batch_used: u32,
    pub detections: *mut bb_detection,
    pub error: [c_char; 200],
}

extern "C" {
    pub fn bb_pipeline_create(ctx: *mut bb_ctx, cfg: *const bb_pipeline_cfg, f: bb_classify_fn, user: *mut c_void,
                              out: *mut *mut bb_pipeline) -> i32;
    pub fn bb_pipeline_destroy(p: *mut bb_pipeline);
    pub fn bb_pipeline_last_error(p: *const bb_pipeline) -> *const c_char;
    pub fn bb_pipeline_process_wav(p: *mut bb_pipeline, path: *const c_char, piece_frames: u64, out: *mut bb_detection,
                                   capacity: u64, n_detections: *mut u64, n_segments: *mut u64, batch_used: *mut u32) -> i32;
    pub fn bb_pool_create(devices: *const i32, n_devices: u32, cfgs: *const bb_pipeline_cfg, f: bb_classify_fn,
                          users: *const *mut c_void, out: *mut *mut bb_pool) -> i32;
    pub fn bb_pool_destroy(p: *mut bb_pool);
    pub fn bb_pool_process_wavs(p: *mut bb_pool, paths: *const *const c_char, n_files: u32, results: *mut bb_pool_result) -> i32;
    pub fn bb_pool_free_results(results: *mut bb_pool_result, n: u32);
}

// ---------------------------------------------------------------------------------------------
// Round-2 additions of the C ABI (include/birda_b200.h): range-filter label projection, watchdog
// seam, FLAC ingest, pool options.  Untested source like the rest of this file (no Rust toolchain
// in the build image); the C side is exercised by tests/test_boundary.py, tests/test_flac.py and
// tests/test_gpu_pipeline.py.
pub const BB_ERR_TIMEOUT: i32 = -11;
pub const BB_S24: i32 = 4; // 3-byte packed PCM, converted as the S32 `<< 8` values symphonia presents

pub type bb_batch_hook = Option<unsafe extern "C" fn(user: *mut c_void, batch_rows: u32, valid_rows: u32, first_segment: u64)>;
pub type bb_watchdog_fn = Option<unsafe extern "C" fn(user: *mut c_void, timeout_secs: u64, batch_size: u32)>;

#[repr(C)]
pub struct bb_watchdog {
    _private: [u8; 0],
}
#[repr(C)]
pub struct bb_flac {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct bb_flac_info {
    pub sample_rate: u32,
    pub channels: u32,
    pub bits_per_sample: u32,
    pub min_block: u32,
    pub max_block: u32,
    pub min_frame_bytes: u32,
    pub max_frame_bytes: u32,
    pub frames: u64,
    pub first_frame_offset: u64,
    pub file_bytes: u64,
    pub fmt: i32,
}

extern "C" {
    // src/inference/geomodel.rs:58-87, :140-157 -> the dense [C] mask K3 reads (NaN = no geomodel entry)
    pub fn bb_mask_build(classifier_labels: *const *const c_char, n_classifier: u32,
                         geomodel_labels: *const *const c_char, n_geomodel: u32,
                         score_species: *const *const c_char, score_values: *const f32, n_scores: u32,
                         mask: *mut f32, mapped: *mut u32, unmatched: *mut u32) -> i32;
    pub fn bb_rule_scientific_name_len(label: *const c_char) -> u32;
    // src/pipeline/processor.rs:194-211
    pub fn bb_rule_inference_timeout_secs(env_value: *const c_char) -> u64;
    // src/gpu/watchdog.rs:22-66
    pub fn bb_watchdog_start(timeout_ms: u64, batch_size: u32, on_fire: bb_watchdog_fn, user: *mut c_void, out: *mut *mut bb_watchdog) -> i32;
    pub fn bb_watchdog_cancel(w: *mut bb_watchdog);
    // the per-batch seam of the library's per-file loop (processor.rs:263-277)
    pub fn bb_pipeline_set_batch_hooks(p: *mut bb_pipeline, before: bb_batch_hook, after: bb_batch_hook, user: *mut c_void);
    pub fn bb_pipeline_set_batch_timeout(p: *mut bb_pipeline, timeout_ms: u64, on_fire: bb_watchdog_fn, user: *mut c_void);
    pub fn bb_pipeline_set_read_threads(p: *mut bb_pipeline, threads: u32);
    // FLAC decoded on the GPU (src/audio/decode.rs:54-128 for that container)
    pub fn bb_flac_probe(path: *const c_char, out: *mut bb_flac_info) -> i32;
    pub fn bb_flac_create(ctx: *mut bb_ctx, out: *mut *mut bb_flac) -> i32;
    pub fn bb_flac_destroy(f: *mut bb_flac);
    pub fn bb_flac_decode(f: *mut bb_flac, file_bytes: *const c_void, n_bytes: u64, info: *const bb_flac_info,
                          d_pcm: *mut *mut c_void, frames_out: *mut u64) -> i32;
    // pool: a classifier that queues on its worker's stream (ORT user_compute_stream) needs no waits around a batch
    pub fn bb_pool_worker_ctx(p: *mut bb_pool, worker: u32) -> *mut bb_ctx;
    pub fn bb_pool_set_stream_ordered(p: *mut bb_pool, on: i32);
    pub fn bb_ctx_set_blocking_sync(ctx: *mut bb_ctx, on: i32);
    pub fn bb_plan_describe(plan: *const bb_plan, buf: *mut c_char, buf_len: u32) -> i32;
}

/// `start_inference_watchdog` (src/gpu/watchdog.rs:22-52) over the library's timer: dropping the guard cancels it.
pub struct WatchdogGuard(*mut bb_watchdog);
impl WatchdogGuard {
    pub fn start(timeout: std::time::Duration, batch_size: usize) -> Option<Self> {
        let mut w: *mut bb_watchdog = std::ptr::null_mut();
        // on_fire = None: the reference's behaviour (FATAL block on stderr, exit status 1)
        let rc = unsafe { bb_watchdog_start(timeout.as_millis() as u64, batch_size as u32, None, std::ptr::null_mut(), &mut w) };
        if rc == 0 { Some(Self(w)) } else { None }
    }
}
impl Drop for WatchdogGuard {
    fn drop(&mut self) {
        unsafe { bb_watchdog_cancel(self.0) }
    }
}
// the guard is only ever dropped, never shared: same contract as the reference's (watchdog.rs:86-91)
unsafe impl Send for WatchdogGuard {}
