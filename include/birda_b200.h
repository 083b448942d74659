/*
 * birda_b200 — C ABI of the B200-native audio front end and post-inference scoring path.
 *
 * This header is the drop-in boundary.  The reference (tphakala/birda,
 * pure Rust) has no FFI for this path today; each entry point below names the Rust
 * item (file:line under /root/reference) whose work it replaces, so that a maintainer
 * can bind it from `src/audio`, `src/pipeline` and `src/inference` through a thin
 * `extern "C"` module built in build.rs (see INTEGRATION.md and rust/).
 *
 * Conventions
 *   - every function returns int32_t: 0 = BB_OK, < 0 = BB_ERR_*; nothing throws or aborts
 *     across the boundary (the reference forbids panics: Cargo.toml:85-87);
 *   - `bb_last_error(ctx)` returns a message for the last failure on that context
 *     (ctx == NULL: the calling thread's last context-free failure);
 *   - entry points leave the calling thread's current CUDA device as they found it;
 *   - a bb_ctx is one GPU + one CUDA stream; it is NOT thread-safe (one owner thread, as
 *     BirdClassifier is only used from the main thread: src/pipeline/processor.rs:659-671).
 *     Several contexts (one per GPU) may run concurrently;
 *   - the caller owns every host pointer it passes; the library owns device memory held
 *     by a ctx / plan.  A plan-owned device tensor stays valid until the next run on that
 *     plan or bb_plan_destroy;
 *   - there is NO CPU fallback: without a CUDA device bb_ctx_create fails with
 *     BB_ERR_NO_DEVICE.  The `bb_rule_*` helpers are pure host arithmetic (the reference's
 *     integer / f32 rules) and work anywhere.
 */
#ifndef BIRDA_B200_H
#define BIRDA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BB_VERSION_MAJOR 0
#define BB_VERSION_MINOR 1

typedef struct bb_ctx  bb_ctx;   /* one per GPU                                              */
typedef struct bb_plan bb_plan;  /* one per (src_rate, channels, fmt, tgt_rate, seg, ovl)     */

typedef enum {
    BB_OK                     =  0,
    BB_ERR_INVALID_ARG        = -1,
    BB_ERR_OVERLAP_GE_SEGMENT = -2,  /* src/audio/decode.rs:156-162 -> Error::Internal        */
    BB_ERR_UNSUPPORTED_RATE   = -3,  /* -> Error::Resample { reason }                         */
    BB_ERR_UNSUPPORTED_FORMAT = -4,  /* reference silently drops: src/audio/decode.rs:407-409 */
    BB_ERR_CUDA               = -5,  /* -> Error::Inference { reason }                        */
    BB_ERR_OOM                = -6,
    BB_ERR_INTERNAL           = -7,  /* -> Error::Internal { message }                        */
    BB_ERR_NO_DEVICE          = -8,
    BB_ERR_CAPACITY           = -9,  /* caller-provided output array too small                */
    BB_ERR_IO                 = -10, /* -> Error::AudioOpen / Error::AudioDecode              */
    BB_ERR_TIMEOUT            = -11  /* a batch outlived the watchdog (src/gpu/watchdog.rs:22-52) */
} bb_status;

/* interleaved frames, little endian; the formats src/audio/decode.rs:353-411 converts.
 * BB_S24 = 3-byte packed PCM as it sits in a 24-bit WAV file.  symphonia's PCM decoder hands 24-bit PCM to
 * append_samples as AudioBufferRef::S32 holding `sample << 8` (symphonia-codec-pcm is not vendored under
 * /root/reference: stated from the crate's published behaviour), so the value converted is
 * ((s24 << 8) as f32) / 2^31 through the S32 arm, decode.rs:386-402. */
typedef enum { BB_S16 = 1, BB_S32 = 2, BB_F32 = 3, BB_S24 = 4 } bb_sample_fmt;

/* ------------------------------------------------------------------------------------------
 * Host rules — pure arithmetic, no GPU.  Bit-exact restatements of the reference's integer
 * and f32 rules; the Rust host keeps its own code for these, the C forms exist so that the
 * library, the tests and non-Rust hosts agree with it.
 * ---------------------------------------------------------------------------------------- */

/* (segment_duration * rate as f32) as usize etc.; bat_mode -> 144000 / 36000.
 * src/pipeline/processor.rs:502-522, src/constants.rs:525-541 */
int32_t bb_rule_segment_samples(float segment_duration, float overlap, uint32_t target_rate,
                                int32_t bat_mode, uint64_t* segment_samples, uint64_t* overlap_samples);

/* ceil(n * src / tgt) in f64, identity for equal rates.  src/pipeline/processor.rs:61-82 */
int32_t bb_rule_source_window(uint64_t segment_samples, uint64_t overlap_samples,
                              uint32_t src_rate, uint32_t tgt_rate,
                              uint64_t* src_segment, uint64_t* src_overlap);

/* Number of windows StreamingDecoder::next_segment yields for a stream of total_frames
 * mono samples.  src/audio/decode.rs:150-202.  BB_ERR_OVERLAP_GE_SEGMENT as :156-162. */
int32_t bb_rule_segment_count(uint64_t total_frames, uint64_t src_segment, uint64_t src_overlap,
                              uint64_t* nseg);

/* Windows [first, first+capacity) of that sequence: RawSegment.start_sample and the number of
 * real (non-padding) samples in each.  src/audio/decode.rs:175-196 */
int32_t bb_rule_segment_table(uint64_t total_frames, uint64_t src_segment, uint64_t src_overlap,
                              uint64_t first, uint64_t capacity,
                              uint64_t* start_sample, uint64_t* take, uint64_t* written);

/* AudioChunk.start_time / end_time in f32.  src/pipeline/processor.rs:89-94 */
int32_t bb_rule_chunk_times(uint64_t start_sample, uint32_t src_rate, uint64_t segment_samples,
                            uint32_t tgt_rate, float* start_time, float* end_time);

/* estimate_segment_count; *estimate = -1 encodes None.  src/output/progress.rs:80-92 */
int32_t bb_rule_estimate_segment_count(double duration_secs, int32_t has_duration,
                                       float segment_duration, float overlap, int64_t* estimate);

/* min(batch, estimate), never 0.  src/pipeline/processor.rs:525-545 */
uint32_t bb_rule_effective_batch_size(uint32_t batch_size, int64_t estimate);

/* rubato Fft::new(from, to, 1024, 1, FixedSync::Both) block sizes (src/audio/resample.rs:19-30):
 * input frames per process() call, output frames per call, spectrum bins carried over. */
int32_t bb_rule_resampler_blocks(uint32_t src_rate, uint32_t tgt_rate,
                                 uint32_t* n_in, uint32_t* n_out, uint32_t* n_keep, float* cutoff);

/* The anti-alias filter taps (f32, already / (2*n_in)) the resampler convolves with; n = n_in. */
int32_t bb_rule_resampler_taps(uint32_t src_rate, uint32_t tgt_rate, float* taps, uint32_t n);

/* Length resample() returns for src_len input samples before the resize.
 * src/audio/resample.rs:35-88 */
int32_t bb_rule_resampled_len(uint64_t src_len, uint32_t src_rate, uint32_t tgt_rate, uint64_t* out_len);

/* src/utils/date.rs:21-68 */
uint32_t bb_rule_date_to_week(uint32_t month, uint32_t day);
uint32_t bb_rule_week_to_start_day(uint32_t week);
void     bb_rule_day_of_year_to_date(uint32_t day_of_year, uint32_t* month, uint32_t* day);

/* BIRDA_INFERENCE_TIMEOUT parsing: seconds in [1, 3600], anything else (NULL, junk, out of range) -> 10.
 * src/pipeline/processor.rs:194-211 */
uint64_t bb_rule_inference_timeout_secs(const char* env_value);

/* scientific_name(): bytes of `label` that form the species key before case folding — the part before the first
 * '_' when that part contains a space, else the whole label.  src/inference/geomodel.rs:28-33 */
uint32_t bb_rule_scientific_name_len(const char* label);

/* Range-filter precompute, label side (src/inference/geomodel.rs:41-127 SpeciesMapping::build, :129-157
 * GeomodelScores::project): project geomodel scores into the classifier's label space as the dense vector K3
 * reads — mask[i] = score_of(classifier_labels[i]), NaN where that is None.  Keys are scientific names, lower-cased
 * (Unicode simple case folding of the Latin, Greek and Cyrillic blocks; everything else compares as is); the first
 * classifier label wins a key collision (its later namesakes have no entry unless their label STRING is identical);
 * mapped species the geomodel did not report read 0.0; later scores overwrite earlier ones.  score_species[j] /
 * score_values[j] are the LocationScore list of RangeFilter::predict (src/inference/classifier.rs:133-141).
 * mapped / unmatched: SpeciesMapping::mapped_count / unmatched_count (may be NULL).  Pure host code. */
int32_t bb_mask_build(const char* const* classifier_labels, uint32_t n_classifier,
                      const char* const* geomodel_labels, uint32_t n_geomodel,
                      const char* const* score_species, const float* score_values, uint32_t n_scores,
                      float* mask, uint32_t* mapped, uint32_t* unmatched);

/* ------------------------------------------------------------------------------------------
 * Watchdog (src/gpu/watchdog.rs:22-66): a timer armed around one inference batch.  If it is not cancelled
 * within timeout_ms, on_fire(user, timeout_secs, batch_size) runs on the watchdog's thread; on_fire == NULL is
 * the reference's behaviour — print its FATAL block to stderr and terminate the process with status 1
 * (watchdog.rs:31-49).  bb_watchdog_cancel is WatchdogGuard::drop: it disarms the timer and frees it.
 * ---------------------------------------------------------------------------------------- */
typedef struct bb_watchdog bb_watchdog;
typedef void (*bb_watchdog_fn)(void* user, uint64_t timeout_secs, uint32_t batch_size);
int32_t bb_watchdog_start(uint64_t timeout_ms, uint32_t batch_size, bb_watchdog_fn on_fire, void* user, bb_watchdog** out);
void    bb_watchdog_cancel(bb_watchdog*);

/* ------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------- */

uint32_t    bb_version(void);                                  /* (major << 16) | minor        */
int32_t     bb_device_count(int32_t* count);
int32_t     bb_ctx_create(int32_t device, bb_ctx** out);       /* owns a new non-blocking stream */
/* Run on a stream the caller owns (e.g. the stream ORT's CUDA EP or torch uses) */
int32_t     bb_ctx_create_on_stream(int32_t device, void* cuda_stream, bb_ctx** out);
void        bb_ctx_destroy(bb_ctx*);
const char* bb_last_error(const bb_ctx*);
void*       bb_ctx_stream(bb_ctx*);                            /* cudaStream_t                  */
int32_t     bb_sync(bb_ctx*);                                  /* cudaStreamSynchronize         */
uint64_t    bb_ctx_kernel_launches(const bb_ctx*);             /* kernels launched so far        */
/* waits of this context (bb_sync, bb_post_run) sleep on a blocking-sync event instead of spinning on the stream: for hosts
 * that run several worker threads per GPU and need their cores (bb_pool turns it on for its contexts) */
void        bb_ctx_set_blocking_sync(bb_ctx*, int32_t on);

/* Page-locked host staging (decode straight into it; H2D copies then run at PCIe rate) */
int32_t bb_host_alloc(uint64_t bytes, void** out);
void    bb_host_free(void*);

/* ------------------------------------------------------------------------------------------
 * Front end: decoded PCM -> downmix -> per-window resample -> packed model input
 * Replaces the decode thread's loop body: StreamingDecoder::next_segment
 * (src/audio/decode.rs:150-202) + append_samples' conversion/downmix (:353-411) +
 * resample_chunk (src/audio/resample.rs:97-105) + resize + time stamps
 * (src/pipeline/processor.rs:84-100) + the tensor pack inside birdnet_onnx::predict_batch*.
 * ---------------------------------------------------------------------------------------- */

/* segment_samples / overlap_samples are TARGET-rate counts exactly as processor.rs:514-521
 * computes them (use bb_rule_segment_samples).  tgt_rate == src_rate means "no resampling"
 * (also the bat path, processor.rs:464-475). */
int32_t bb_plan_create(bb_ctx*, uint32_t src_rate, uint32_t channels, bb_sample_fmt fmt,
                       uint32_t tgt_rate, uint64_t segment_samples, uint64_t overlap_samples,
                       bb_plan** out);
void    bb_plan_destroy(bb_plan*);
int32_t bb_plan_source_window(const bb_plan*, uint64_t* src_segment, uint64_t* src_overlap);
int32_t bb_plan_segment_count(const bb_plan*, uint64_t total_frames, uint64_t* nseg);
/* which kernel, transform sizes and blocking bb_frontend_run will launch for this plan (benches record it next to
 * their timings; a committed profile is only quoted when it names the same kernel) */
int32_t bb_plan_describe(const bb_plan*, char* buf, uint32_t buf_len);

/* One piece of a file in, packed segments out.
 *
 *   pcm               interleaved frames (host or device memory, see pcm_is_device)
 *   frames            frames in this piece
 *   first_start_sample RawSegment.start_sample of the first window of this piece (0 for a
 *                     whole file); only feeds the start_sample / time outputs
 *   is_eof            1: this piece ends the file -> emit the zero-padded tail windows
 *                     (decode.rs:175-196); 0: emit full windows only and report how many
 *                     frames were consumed so the caller re-presents the rest
 *   pad_to_batch      > 0: rows up to the next multiple are zero segments
 *                     (processor.rs:239-260); the tensor has nseg_padded rows
 *   d_segments        out: device pointer to [nseg_padded, segment_samples] f32, plan-owned.
 *                     Pass d_out_user != NULL to write into caller-owned device memory
 *                     (e.g. an ORT IoBinding input) of capacity_rows rows instead.
 *   start_sample, start_time, end_time   host arrays of capacity_rows entries (may be NULL)
 *   nseg_out          windows produced (valid rows); consumed_frames: frames fully retired
 *
 * Asynchronous w.r.t. the host: device work is enqueued on the ctx stream; the host arrays
 * are filled before return (they are pure host arithmetic). */
int32_t bb_frontend_run(bb_plan*, const void* pcm, uint64_t frames, int32_t pcm_is_device,
                        uint64_t first_start_sample, int32_t is_eof, uint32_t pad_to_batch,
                        float* d_out_user, uint64_t capacity_rows,
                        float** d_segments, uint64_t* start_sample, float* start_time,
                        float* end_time, uint64_t* nseg_out, uint64_t* nseg_padded,
                        uint64_t* consumed_frames);

/* ------------------------------------------------------------------------------------------
 * Post-inference: activation -> top-k + threshold -> range / species mask -> threshold
 * Replaces the tail of birdnet_onnx::Classifier::predict_batch* (configured at
 * src/inference/classifier.rs:269-273), BirdClassifier::apply_range_filter
 * (src/inference/classifier.rs:587-645 -> src/inference/geomodel_filter.rs:45-79) and the
 * second confidence test at src/pipeline/processor.rs:374.
 * ---------------------------------------------------------------------------------------- */

typedef enum { BB_ACT_NONE = 0, BB_ACT_SIGMOID = 1, BB_ACT_SOFTMAX = 2 } bb_activation;

typedef struct {
    int32_t  activation;        /* bb_activation                                              */
    float    min_confidence;    /* --min-confidence (src/constants.rs:25)                     */
    uint32_t top_k;             /* DEFAULT_TOP_K = 5 (src/constants.rs:178); 1..BB_MAX_TOP_K  */
    float    range_threshold;   /* FilterSettings.threshold, inclusive                        */
    int32_t  keep_unmatched;    /* UnmatchedPolicy::Keep                                      */
    int32_t  rerank;            /* FilterSettings.rerank                                      */
} bb_post_cfg;

#define BB_MAX_TOP_K 8

/* d_scores: device [B, C] f32 row-major (model output).  Rows [0, valid_B) are processed
 * (the rest is batch padding).  d_mask: device [C] f32 dense projection of GeomodelScores,
 * NaN = label has no geomodel entry, NULL = no range filter.  d_species_keep: device [C] u8,
 * NULL = no species list (mutually exclusive with d_mask: src/lib.rs:502-520).
 * Outputs (device, caller- or ctx-owned): index [valid_B, top_k] u32, conf [valid_B, top_k]
 * f32, count [valid_B] u32 — entries beyond count are index 0xFFFFFFFF / conf 0. */
int32_t bb_post_run_device(bb_ctx*, const float* d_scores, uint32_t B, uint32_t C, uint32_t valid_B,
                           const bb_post_cfg*, const float* d_mask, const uint8_t* d_species_keep,
                           uint32_t* d_index, float* d_conf, uint32_t* d_count);

/* Same, results copied to host arrays and the stream synchronised before return. */
int32_t bb_post_run(bb_ctx*, const float* d_scores, uint32_t B, uint32_t C, uint32_t valid_B,
                    const bb_post_cfg*, const float* d_mask, const uint8_t* d_species_keep,
                    uint32_t* h_index, float* h_conf, uint32_t* h_count);

/* ------------------------------------------------------------------------------------------
 * PCM ingest: WAV / RF64 files are already the interleaved layout the kernels
 * consume.  Replaces StreamingDecoder::open + the symphonia packet loop for PCM WAV
 * (src/audio/decode.rs:54-128, :205-245): probe once, then read frame ranges straight into a staging
 * buffer (bb_host_alloc) and hand it to bb_frontend_run.  Pure host code.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t sample_rate;
    uint32_t channels;
    uint32_t bits_per_sample;
    int32_t  fmt;            /* bb_sample_fmt, 0 = a format the reference does not convert           */
    uint64_t frames;
    uint64_t data_offset;    /* byte offset of the first sample in the file                           */
} bb_wav_info;
int32_t bb_wav_probe(const char* path, bb_wav_info* out);
int32_t bb_wav_read(const char* path, const bb_wav_info* info, uint64_t first_frame, uint64_t frames, void* dst);
/* the same read cut into slices read by `threads` threads at once (a read from the page cache is a memcpy: one thread
 * moves 6-10 GB/s, the PCIe link to the GPU takes ~55 GB/s) */
int32_t bb_wav_read_parallel(const char* path, const bb_wav_info* info, uint64_t first_frame, uint64_t frames, void* dst, uint32_t threads);

/* ------------------------------------------------------------------------------------------
 * Per-file pipeline (C++ host code in the library): the reference's process_file loop
 * (src/pipeline/processor.rs:418-796) with the front end and the post step on the GPU and the
 * classifier as a callback.  In birda the callback is BirdClassifier::predict_batch_device (ONNX
 * Runtime with IoBinding on the ctx stream); `d_segments` is written, and `*d_scores` is read, on the ctx stream
 * (bb_ctx_stream): a callback that computes on another stream must order itself against it.  It must leave
 * `*d_scores` = device [batch_rows, classes]
 * f32 valid until the next call.  Return 0 on success (anything else -> BB_ERR_INTERNAL, as
 * Error::Inference).
 * ---------------------------------------------------------------------------------------- */
typedef struct bb_pipeline bb_pipeline;
typedef int32_t (*bb_classify_fn)(void* user, const float* d_segments, uint32_t batch_rows, uint32_t samples,
                                  const float** d_scores, uint32_t* classes);
typedef struct {
    uint32_t    target_rate;        /* classifier.sample_rate()                                       */
    float       segment_duration;   /* classifier.segment_duration()                                  */
    float       overlap;            /* --overlap (src/constants.rs:28)                                */
    uint32_t    batch_size;         /* --batch-size, 1..512 (src/constants.rs:44-55)                  */
    int32_t     bat_mode;           /* src/pipeline/processor.rs:461-475                              */
    bb_post_cfg post;
    const float*   d_mask;          /* device [C] f32 or NULL                                         */
    const uint8_t* d_species_keep;  /* device [C] u8 or NULL                                          */
} bb_pipeline_cfg;
typedef struct {                    /* Detection fields the writers need (src/output/types.rs:8-23)   */
    uint32_t segment;               /* window index inside the file                                   */
    uint32_t index;                 /* class index -> label -> Detection::from_label                  */
    float    confidence;
    float    start_time;
    float    end_time;
} bb_detection;
int32_t     bb_pipeline_create(bb_ctx*, const bb_pipeline_cfg*, bb_classify_fn, void* user, bb_pipeline** out);
void        bb_pipeline_destroy(bb_pipeline*);
const char* bb_pipeline_last_error(const bb_pipeline*);
/* front-end plans are cached per (source rate, channels, format), 8 kinds, least recently used evicted; this counts
 * how many were built so far (a mixed-rate directory builds each kind once) */
uint64_t    bb_pipeline_plans_created(const bb_pipeline*);
/* Per-batch seam (src/pipeline/processor.rs:263-277): `before` is called right before the classifier callback of
 * every batch, `after` once that batch's inference and post step have completed on the device (the pipeline
 * synchronises the stream per batch while hooks or a timeout are set) — the two places where the reference
 * creates and drops its WatchdogGuard, so a Rust host arms start_inference_watchdog exactly as today. */
typedef void (*bb_batch_hook)(void* user, uint32_t batch_rows, uint32_t valid_rows, uint64_t first_segment);
void        bb_pipeline_set_batch_hooks(bb_pipeline*, bb_batch_hook before, bb_batch_hook after, void* user);
/* Or let the library keep the watchdog: every batch is bracketed by a timer of timeout_ms (0 = off).  With
 * on_fire == NULL a batch that outlives it ends the process like the reference; with a callback, on_fire runs and
 * the file fails with BB_ERR_TIMEOUT as soon as the batch returns. */
void        bb_pipeline_set_batch_timeout(bb_pipeline*, uint64_t timeout_ms, bb_watchdog_fn on_fire, void* user);
/* for classifiers that do not run on the ctx stream: synchronise it before every callback (the callback then has to
 * finish its own work before it returns).  Off by default; bb_pool turns it on. */
void        bb_pipeline_set_sync_before_classify(bb_pipeline*, int32_t on);
/* threads per file read of bb_pipeline_process_wav (bb_wav_read_parallel); default 4 */
void        bb_pipeline_set_read_threads(bb_pipeline*, uint32_t threads);
/* Whole decoded file in host memory.  Detections come back sorted (start_time asc, confidence desc:
 * processor.rs:178-187); BB_ERR_CAPACITY reports the needed count in *n_detections. */
int32_t bb_pipeline_process_pcm(bb_pipeline*, const void* pcm, uint64_t frames, uint32_t src_rate, uint32_t channels,
                                int32_t fmt, bb_detection* out, uint64_t capacity, uint64_t* n_detections,
                                uint64_t* n_segments, uint32_t* batch_used);
/* An audio file (by content: RIFF / RF64 WAVE, or FLAC).  WAV is streamed through two pinned staging buffers in pieces
 * of ~piece_frames (0 = default, ~256 MB of PCM): a reader thread fills one while the GPU works on the other, pieces
 * are cut so that only the file's last batch is padded, and the post-step results of a piece come back in one copy
 * (per batch while hooks or a batch timeout are set).  A FLAC file goes to the GPU compressed and is decoded there
 * (bb_flac_*), in one piece. */
int32_t bb_pipeline_process_wav(bb_pipeline*, const char* path, uint64_t piece_frames, bb_detection* out, uint64_t capacity,
                                uint64_t* n_detections, uint64_t* n_segments, uint32_t* batch_used);

/* ------------------------------------------------------------------------------------------
 * Directory batches across the GPUs of one box (csrc/pool.cpp): a host work queue of WAV files, longest first, one worker
 * thread + context + bb_pipeline per GPU, no collective (files are independent: src/lib.rs:694-796).
 * `cfgs[i]` / `users[i]` belong to `devices[i]` (device pointers such as d_mask live on that device); the classifier
 * callback is called from the worker thread of its device, after the context's stream has been synchronised, and must
 * finish its own device work before it returns.  Results come back per file, each list sorted the
 * reference's way; `detections` is owned by the library until bb_pool_free_results.
 * ---------------------------------------------------------------------------------------- */
typedef struct bb_pool bb_pool;
typedef struct {
    int32_t  status;                /* BB_OK or the BB_ERR_* of this file (the batch goes on: src/lib.rs:776-796)  */
    int32_t  device;                /* GPU that processed the file                                                 */
    uint64_t n_detections, n_segments;
    uint32_t batch_used;
    bb_detection* detections;
    char     error[200];
} bb_pool_result;
int32_t bb_pool_create(const int32_t* devices, uint32_t n_devices, const bb_pipeline_cfg* cfgs, bb_classify_fn fn,
                       void* const* users, bb_pool** out);
void    bb_pool_destroy(bb_pool*);
/* returns BB_OK when every file succeeded, else the status of a failed file; results[n_files] is always filled */
int32_t bb_pool_process_wavs(bb_pool*, const char* const* paths, uint32_t n_files, bb_pool_result* results);
void    bb_pool_free_results(bb_pool_result* results, uint32_t n);
/* worker i's context (its stream: bb_ctx_stream).  A classifier that queues its work on that stream (ONNX Runtime with
 * user_compute_stream, the stand-in after bb_standin_use_stream) needs no synchronisation around the callback:
 * bb_pool_set_stream_ordered(pool, 1) drops the pool's per-batch waits (default 0: the callback may use any stream and
 * must finish before it returns). */
bb_ctx* bb_pool_worker_ctx(bb_pool*, uint32_t worker);
void    bb_pool_set_stream_ordered(bb_pool*, int32_t on);
uint64_t bb_pool_kernel_launches(const bb_pool*);      /* kernels the workers' contexts launched so far (classifier's not included) */

/* ------------------------------------------------------------------------------------------
 * Tiny dense heads on the device: out[B,N] = act(x[B,K] W[K,N] + b[N]), f32.
 * The geomodel forward ([1,3] -> [1,12012] sigmoid, once per run: src/inference/classifier.rs:117-188,
 * fixture tests/fixtures/make_fixture_geomodel.py:20-28) and the bat head over embeddings
 * (src/pipeline/processor.rs:323-360) without a host hop.  b may be NULL.  Asynchronous on the ctx stream.
 * ---------------------------------------------------------------------------------------- */
int32_t bb_dense_run(bb_ctx*, const float* d_x, uint32_t B, uint32_t K, const float* d_W, const float* d_b, uint32_t N,
                     int32_t activation, float* d_out);
/* Per-class logistic (Platt) calibration of logits ahead of the post step: out[b,c] = a[c] * scores[b,c] + b[c]
 * (then bb_post_run with BB_ACT_SIGMOID).  This is the standard form of the "per-species calibration" the reference
 * applies to BSG models (src/pipeline/processor.rs:281-314 -> birdnet_onnx::BsgPostProcessor::calibrate); that crate's
 * source and its CSV format are not in the image, so the form is stated, not pinned, and the SDM adjustment (location /
 * day-of-year prior) is not restated.  d_out may equal d_scores; b may be NULL.  Asynchronous on the ctx stream. */
int32_t bb_calibrate_run(bb_ctx*, const float* d_scores, uint32_t B, uint32_t C, const float* d_a, const float* d_b, float* d_out);

/* ------------------------------------------------------------------------------------------
 * Spectrogram prefix on the device: framed STFT power + mel projection for classifiers
 * whose ONNX graph opens with an STFT / mel prefix (manifests/Perch-v2-Models.models.json:15,46; BirdNET v2.4's
 * in-graph spectrogram layers).  The truncated graph then takes [rows, n_mels, n_frames] f32 through IoBinding.
 * The reference has no host code for this step and pins none of its parameters, so window, mel weights and
 * scaling are inputs.  Frame t of a row covers samples [t*hop, t*hop + n_fft) (zero past the end of the row),
 * is multiplied by `window`, transformed by a real FFT of length n_fft; |X|^power of the bins the mel filters touch
 * is projected with `mel_weights` on the tcgen05 tensor cores (three tf32 products of an error-free hi/lo split:
 * f32 accuracy) and scaled: log_mode 0 = linear, 1 = ln(x + log_eps), 2 = 10 log10(max(x, log_eps)).
 * ---------------------------------------------------------------------------------------- */
typedef struct bb_melspec bb_melspec;
typedef struct {
    uint32_t n_fft;      /* power of two, 256..4096                                   */
    uint32_t hop;        /* frame step in samples                                      */
    uint32_t n_frames;   /* frames per row                                             */
    uint32_t n_mels;     /* multiple of 16, 16..256                                    */
    float    power;      /* 1 = magnitude, 2 = power                                   */
    int32_t  log_mode;
    float    log_eps;
} bb_melspec_cfg;
/* window: host [n_fft]; mel_weights: host [n_mels, n_fft/2 + 1] row-major.  Both are copied. */
int32_t bb_melspec_create(bb_ctx*, const bb_melspec_cfg*, const float* window, const float* mel_weights, bb_melspec** out);
void    bb_melspec_destroy(bb_melspec*);
/* bins [bin_lo, bin_lo + n_bins) carry non-zero mel weight; the GEMM runs over k_padded = n_bins rounded up to 32 */
int32_t bb_melspec_info(const bb_melspec*, uint32_t* bin_lo, uint32_t* n_bins, uint32_t* k_padded);
/* d_segments: device [rows, samples] f32 (the packed tensor of bb_frontend_run); d_out: device [rows, n_mels, n_frames].
 * Asynchronous on the ctx stream. */
int32_t bb_melspec_run(bb_melspec*, const float* d_segments, uint32_t rows, uint32_t samples, float* d_out);

/* ------------------------------------------------------------------------------------------
 * FLAC ingest (the reference decodes FLAC through symphonia on its decode thread:
 * src/audio/decode.rs:54-128, :205-245, extension list src/pipeline/coordinator.rs:181-190).  The compressed file is
 * what crosses PCIe; the host finds the frame boundaries, the GPU decodes (one thread per frame) into the interleaved
 * PCM K1/K2 consume: streams of up to 16 bits -> BB_S16 (left-justified), 17..24 -> BB_S24, 32 -> BB_S32 — the values
 * append_samples converts (symphonia presents FLAC as left-justified S32, decode.rs:386-402).  Every frame's CRC-16 is
 * verified on the device.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t sample_rate, channels, bits_per_sample;
    uint32_t min_block, max_block, min_frame_bytes, max_frame_bytes;
    uint64_t frames;               /* samples per channel from STREAMINFO (0 = unknown)   */
    uint64_t first_frame_offset;   /* where the metadata blocks end                        */
    uint64_t file_bytes;
    int32_t  fmt;                  /* bb_sample_fmt of the decoded PCM                     */
} bb_flac_info;
typedef struct bb_flac bb_flac;
int32_t bb_flac_probe(const char* path, bb_flac_info* out);
int32_t bb_flac_probe_bytes(const void* bytes, uint64_t n_bytes, bb_flac_info* out);
/* frame index (host only): offsets / first samples / block sizes as the decoder finds them; arrays may be NULL */
int32_t bb_flac_index(const void* bytes, uint64_t n_bytes, const bb_flac_info* info, uint64_t* offsets, uint64_t* first_samples,
                      uint32_t* block_sizes, uint64_t capacity, uint64_t* n_frames);
int32_t bb_flac_create(bb_ctx*, bb_flac** out);
void    bb_flac_destroy(bb_flac*);
/* the whole compressed file in host memory -> *d_pcm (device, interleaved info->fmt, valid until the next decode on this
 * object), *frames_out samples per channel.  Returns once the decode has completed. */
int32_t bb_flac_decode(bb_flac*, const void* file_bytes, uint64_t n_bytes, const bb_flac_info* info, void** d_pcm, uint64_t* frames_out);

/* ------------------------------------------------------------------------------------------
 * Stand-in classifier (benches and tests; NOT a model, nothing of the reference's is replaced by it): the I/O contract
 * of BirdClassifier::predict_batch (src/inference/classifier.rs:469-582) — [batch, samples] f32 windows on the device
 * in, [batch, classes] f32 logits on the device out — with trivial arithmetic (48 band energies times a fixed
 * pseudo-random matrix).  bb_standin_classify is a bb_classify_fn whose `user` is the bb_standin*.  By default it runs
 * on a stream of its own and finishes before it returns (what bb_pool requires); after bb_standin_use_stream(s, stream,
 * 1) it queues its two kernels on that stream and returns at once (a pipeline on its own context: bb_ctx_stream).
 * The logits stay valid until the next call.  bb_standin_weights copies out W [48, classes] and bias [classes].
 * ---------------------------------------------------------------------------------------- */
typedef struct bb_standin bb_standin;
int32_t  bb_standin_create(int32_t device, uint32_t samples, uint32_t classes, uint32_t max_batch, uint64_t seed, bb_standin** out);
void     bb_standin_destroy(bb_standin*);
void     bb_standin_use_stream(bb_standin*, void* cuda_stream, int32_t on);
int32_t  bb_standin_weights(const bb_standin*, float* W, float* bias);
uint64_t bb_standin_launches(const bb_standin*);
int32_t  bb_standin_classify(void* user, const float* d_segments, uint32_t batch, uint32_t samples, const float** d_scores, uint32_t* classes);

/* Debug hook (tests): the n-th guarded entry point called on this thread from now on fails as if a host
 * allocation had thrown (n = 1: the next one; 0 disarms).  Proves that exceptions stop at the boundary. */
void bb_debug_inject_alloc_failure(int32_t nth);

/* Device memory helpers for hosts without a CUDA binding of their own (tests, Rust shim) */
int32_t bb_dev_alloc(bb_ctx*, uint64_t bytes, void** out);
void    bb_dev_free(bb_ctx*, void*);
int32_t bb_memcpy_h2d(bb_ctx*, void* dst_dev, const void* src_host, uint64_t bytes);   /* async */
int32_t bb_memcpy_d2h(bb_ctx*, void* dst_host, const void* src_dev, uint64_t bytes);   /* async */

#ifdef __cplusplus
}
#endif
#endif /* BIRDA_B200_H */
