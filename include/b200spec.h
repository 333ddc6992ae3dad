/*
 * b200spec.h -- C ABI of libb200spec.so: the madmom-style spectral front end on B200 (sm_100a).
 *
 * The reference (alvaroortegaangulo/audio-tabs) has no FFI for this path: it reaches it through
 * madmom 0.16.1's Python Processor protocol.  Each entry point below names the madmom call it
 * replaces and the reference call site that reaches it (paths relative to /root/reference):
 *
 *   backend/app/services/grid/beats.py:71-75        RNNBeatProcessor()       (3 resolutions + diff)
 *   backend/app/services/chords/extract.py:54-55    DeepChromaProcessor()    (8192 @ fps 10, 105 bands)
 *   backend/app/services/chords/deep_chords.py:48-49,79-81                   (same; 113-band CNN chord variant)
 *   backend/app/services/theory/key.py:99-101,143-144  CNNKeyRecognitionProcessor() (8192 @ fps 5, int16 input)
 *
 * Conventions
 *   - every function returns 0 on success or a negative b200spec_status; the message is available
 *     from b200spec_last_error() (thread-local).  No C++ exception crosses this boundary.
 *   - the caller owns every device buffer and passes raw device pointers plus a CUDA stream handle
 *     (void* == cudaStream_t).  The library never allocates user-visible memory and never
 *     synchronises the stream or the device inside a compute call.
 *   - a plan owns its device-side constants (windows, twiddles, banded filterbank, projection)
 *     until b200spec_plan_destroy().  Host arrays passed to plan_create are copied.
 *   - plans are immutable after creation: shareable between host threads; calls on different
 *     streams with different workspaces are independent.
 *   - there is no CPU fallback: on a machine without an sm_100 device plan_create fails with
 *     B200SPEC_ERR_ARCH / B200SPEC_ERR_CUDA.
 */
#ifndef B200SPEC_H_
#define B200SPEC_H_

#include <stddef.h>
#include <stdint.h>

#if defined(B200SPEC_BUILD) && defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

#define B200SPEC_ABI_VERSION 7
#define B200SPEC_MAX_RES 4          /* resolutions per plan (RNNBeatProcessor uses 3) */
#define B200SPEC_MAX_DIFF_FRAMES 16 /* largest supported diff lag in frames */

typedef enum b200spec_status {
  B200SPEC_OK = 0,
  B200SPEC_ERR_ARG = -1,         /* NULL pointer, negative size, inconsistent descriptor */
  B200SPEC_ERR_UNSUPPORTED = -2, /* frame size / option the kernels do not implement */
  B200SPEC_ERR_CUDA = -3,        /* a CUDA runtime call failed (message has the CUDA error string) */
  B200SPEC_ERR_ARCH = -4         /* device is not compute capability 10.x */
} b200spec_status;

/* input sample formats: madmom Signal dtype (audio/signal.py) */
typedef enum b200spec_dtype {
  B200SPEC_F32 = 0, /* float32 samples, window used unscaled */
  B200SPEC_I16 = 1  /* int16 PCM, window pre-divided by 32767 (madmom stft.py: fft_window = window / iinfo.max) */
} b200spec_dtype;

/* madmom FramedSignal `end` modes (audio/signal.py FramedSignal.__init__) */
typedef enum b200spec_end_mode {
  B200SPEC_END_NORMAL = 0, /* num_frames = ceil(N / hop) */
  B200SPEC_END_EXTEND = 1  /* num_frames = floor(N / hop + 1) */
} b200spec_end_mode;

/*
 * One resolution of a plan = one madmom chain
 *   FramedSignalProcessor -> ShortTimeFourierTransformProcessor -> FilteredSpectrogramProcessor
 *   -> LogarithmicSpectrogramProcessor [-> SpectrogramDifferenceProcessor]
 * The filterbank is given in banded form: band j covers FFT bins
 *   [band_start[j], band_start[j] + band_len[j])  with weights  weights[band_woff[j] + i].
 * (madmom's LogarithmicFilterbank is contiguous per band; the host builds it bit-identically to
 *  madmom/audio/filters.py and passes the float32 weights through unchanged.)
 */
typedef struct b200spec_res_desc {
  int32_t frame_size;        /* FFT length: 1024, 2048, 4096 or 8192 (no Nyquist bin).  A madmom frame_size != fft_size is
                              * expressed by the caller: window zero-padded (or cut) to the FFT length and
                              * origin += frame_size / 2 - fft_size / 2, so that samples [int(n*hop) - fft/2 - origin, + fft)
                              * start where madmom's frame starts */
  double hop_size;           /* samples, may be fractional (sample_rate / fps) */
  int32_t origin;            /* integer origin as resolved by FramedSignal (0 for 'center') */
  const float *window;       /* host, frame_size floats: madmom's fft_window rounded to float32 */
  int32_t num_bands;         /* B; 0 = no filterbank (stft / magnitude only) */
  const int32_t *band_start; /* host, B */
  const int32_t *band_len;   /* host, B */
  const int32_t *band_woff;  /* host, B: offset of band j's first weight in `weights` */
  const float *weights;      /* host, sum(band_len) floats */
  int32_t log_enabled;       /* 1: out = log10(mul * y + add)  (LogarithmicSpectrogram); 0: out = y */
  float mul, add;
  int32_t diff_frames;       /* lag k >= 1 (madmom _diff_frames); 0 = no difference */
  int32_t positive_diffs;    /* 1: max(diff, 0) */
  /* optional projection of the (log-)filtered rows onto num_classes columns (chroma fold / PCP):
   *   proj[t, c] = sum_i proj_weight[i] * L[t, proj_band[i]]  for i in [proj_off[c], proj_off[c+1]) */
  int32_t num_classes;       /* 0 = none */
  const int32_t *proj_off;   /* host, num_classes + 1 */
  const int32_t *proj_band;  /* host, proj_off[num_classes] */
  const float *proj_weight;  /* host, proj_off[num_classes] */
  /* SuperFlux (madmom SpectrogramDifference diff_max_bins): the lagged row is widened by a maximum
   * filter over diff_max_bins neighbouring bands (scipy.ndimage.maximum_filter, size (1, M), 'reflect')
   * before it is subtracted.  0 or 1 = plain difference.  With M > 1 the difference is produced by
   * b200spec_diff_flux_chroma from the rows b200spec_logfilt wrote (col_diff / d_flux of logfilt must
   * be unused); see FrontEnd.run_packed for the two-call sequence. */
  int32_t diff_max_bins;
  /* librosa-style variants of the same chain (onset_strength: /root/reference/backend/app/services/
   * accompaniment/strum.py:114, analysis/content_classifier.py:48,92):
   *   power      1: the filterbank is applied to |X|^2 (librosa melspectrogram power=2.0); 0: to |X|
   *   log_scale  out = log_scale * log10(max(mul*y + add, log_floor)); madmom: 1.  power_to_db: 10 with mul 1, add 0
   *   log_floor  lower clamp of the logarithm's argument (power_to_db amin = 1e-10); <= 0: no clamp */
  int32_t power;
  float log_scale;
  float log_floor;
  /* madmom stft(circular_shift=True) with fft_size == frame_size: the halves of the windowed frame are swapped before
   * the transform, i.e. bin k of b200spec_stft's output is multiplied by (-1)^k.  Magnitude outputs do not change. */
  int32_t circular_shift;
  /* madmom stft(include_nyquist=True): b200spec_stft / b200spec_spectrogram rows hold frame_size/2 + 1 bins, the last
   * one the (real) Nyquist bin.  Resolutions with a filterbank take frame_size/2 bins: the two do not combine. */
  int32_t include_nyquist;
} b200spec_res_desc;

typedef struct b200spec_plan_desc {
  int32_t device;   /* CUDA device ordinal */
  int32_t dtype;    /* b200spec_dtype of the samples */
  int32_t channels; /* 1, or 2 = interleaved stereo down-mixed on load (madmom remix: mean, cast back) */
  int32_t num_res;  /* 1..B200SPEC_MAX_RES */
  b200spec_res_desc res[B200SPEC_MAX_RES];
} b200spec_plan_desc;

typedef struct b200spec_plan b200spec_plan;

/* where one resolution's results go inside the caller's row-major output matrices */
typedef struct b200spec_out_desc {
  float *d_out;      /* (total_frames, ld_out) float32, may be NULL when only flux/proj are wanted */
  int64_t ld_out;    /* row stride in floats */
  int32_t col_spec;  /* first column of the B (log-)filtered values, or -1 to skip */
  int32_t col_diff;  /* first column of the B difference values, or -1 to skip (np.hstack layout) */
  float *d_flux;     /* (total_frames,) row sums of the difference (spectral flux), or NULL */
  float *d_proj;     /* (total_frames, ld_proj) projection output, or NULL */
  int64_t ld_proj;
  /* optional per-clip gain g_c on the SAMPLES of clip c, applied to the band sums before the logarithm
   * (n_clips floats, device).  The magnitude path is linear, so the band sums of g_c * x are g_c times
   * those of x -- and g_c^2 times for a power spectrogram (res_desc.power = 1), which the kernel accounts
   * for.  scale[c] = 1 / (max|x_c| + 1e-9) from b200spec_clip_peak therefore gives the spectrogram of the
   * peak-normalised clip (services/audio.py:24-26 peak_normalize; madmom Signal(norm=True)) without a
   * pass that rewrites the samples.  NULL = no gain. */
  const float *d_clip_scale;
  /* optional per-clip status word (n_clips int32, device; the caller zeroes it): bit 0 is OR-ed in when a
   * non-finite value (NaN / Inf samples) reached clip c's output rows.  Clips are independent -- a bad
   * clip never changes another clip's rows -- so a batch can be triaged per clip (SURVEY.md section 5:
   * "a failed shard must not poison the batch").  NULL = not wanted. */
  int32_t *d_clip_status;
} b200spec_out_desc;

#define B200SPEC_CLIP_NONFINITE 1

int b200spec_abi_version(void);
const char *b200spec_last_error(void);

/* madmom FramedSignal.num_frames (audio/signal.py): evaluated in float64, bit-exact with numpy. */
int b200spec_num_frames(int64_t n_samples, double hop_size, int end_mode, int64_t *out);
/* madmom signal_frame(): first sample index of frame `index` = int(index*hop) - frame_size/2 - origin. */
int b200spec_frame_start(int64_t index, double hop_size, int32_t frame_size, int32_t origin, int64_t *out);

int b200spec_plan_create(const b200spec_plan_desc *desc, b200spec_plan **out);
int b200spec_plan_destroy(b200spec_plan *plan);

/* bytes of device scratch a compute call on `n_clips` clips needs (pass as d_workspace). */
size_t b200spec_workspace_bytes(int32_t n_clips);

/*
 * K1: framing + window + real FFT -> complex64 STFT.
 * Replaces madmom.audio.stft.stft() / ShortTimeFourierTransformProcessor.process().
 *   d_sig        samples of all clips, concatenated (interleaved if channels == 2)
 *   d_clip_off   n_clips+1 sample offsets (per-channel sample index) of each clip in d_sig
 *   d_frame_off  n_clips+1 row offsets of each clip's first frame in d_out
 *   total_frames d_frame_off[n_clips] (host copy, sizes the grid)
 *   d_out        (total_frames, frame_size/2) complex64 as interleaved float pairs
 *                (frame_size/2 + 1 bins per row when the resolution was planned with include_nyquist)
 */
int b200spec_stft(const b200spec_plan *plan, int32_t res, const void *d_sig, const int64_t *d_clip_off,
                  const int64_t *d_frame_off, int32_t n_clips, int64_t total_frames, float *d_out,
                  void *d_workspace, size_t workspace_bytes, void *stream);

/* K1 + |.|: magnitude spectrogram (madmom Spectrogram = np.abs(stft)); d_out is (total_frames, frame_size/2) float32. */
int b200spec_spectrogram(const b200spec_plan *plan, int32_t res, const void *d_sig, const int64_t *d_clip_off,
                         const int64_t *d_frame_off, int32_t n_clips, int64_t total_frames, float *d_out,
                         void *d_workspace, size_t workspace_bytes, void *stream);

/*
 * K1 + K2 + K3 fused: frames -> window -> FFT -> magnitude -> banded filterbank -> log10(mul*y+add)
 * -> lagged (positive) difference -> optional flux / projection, written straight into the
 * caller's stacked layout.  Replaces the whole madmom chain named above for resolution `res`.
 * Launches on `stream`: a task table, the fused kernel and -- when a difference or flux is wanted and the filtered
 * rows go to d_out (col_spec >= 0) -- a seam kernel that forms the first diff_frames difference rows of every task
 * from the rows in d_out (the fused kernel then transforms no warm-up frames; without the filtered rows in d_out
 * it does).  A clip's rows do not depend on the batch it is processed in (bitwise).
 */
int b200spec_logfilt(const b200spec_plan *plan, int32_t res, const void *d_sig, const int64_t *d_clip_off,
                     const int64_t *d_frame_off, int32_t n_clips, int64_t total_frames,
                     const b200spec_out_desc *out, void *d_workspace, size_t workspace_bytes, void *stream);

/*
 * The same chain for SEVERAL resolutions of one hop in ONE launch -- what RNNBeatProcessor's
 * ParallelProcessor + np.hstack does with three FramedSignalProcessor / STFT / filterbank / log / difference
 * branches (/root/reference/backend/app/services/grid/beats.py:73-74; madmom features/beats.py): a group of
 * threads takes a (clip, chunk of frames) task through every resolution in turn and writes all of the row's
 * columns, so the chunk's samples leave HBM once (the later resolutions read them from L2) and there is one
 * launch, one task list and one tail instead of n_res.  res[i] / outs[i] name the resolutions of the plan and
 * where each writes (the np.hstack column offsets).  Frame sizes 1024 / 2048 / 4096 sharing hop_size and
 * origin; B200SPEC_ERR_UNSUPPORTED if the tables of these resolutions do not fit one SM's shared memory
 * together (b200spec_logfilt_multi_supported tells in advance) -- the per-resolution call always works.
 */
int b200spec_logfilt_multi(const b200spec_plan *plan, int32_t n_res, const int32_t *res, const void *d_sig,
                           const int64_t *d_clip_off, const int64_t *d_frame_off, int32_t n_clips,
                           int64_t total_frames, const b200spec_out_desc *outs, void *d_workspace,
                           size_t workspace_bytes, void *stream);
/* 1 if b200spec_logfilt_multi can run these resolutions of the plan in one launch, else 0 */
int b200spec_logfilt_multi_supported(const b200spec_plan *plan, int32_t n_res, const int32_t *res);

/*
 * Per-clip peak of the (down-mixed) samples: d_peak[c] = max_n |x_c[n]| in the plan's sample format
 * (float32 units, or int16 units for B200SPEC_I16; stereo is down-mixed first exactly as the front end
 * does).  With reciprocal != 0 the result is 1 / (peak + eps) instead (32767 / (peak + eps) for
 * B200SPEC_I16, whose window is pre-divided by 32767), ready to be passed as
 * b200spec_out_desc.d_clip_scale.  Replaces peak_normalize (/root/reference/backend/app/services/audio.py:24-26)
 * and madmom Signal(norm=True) (eps = 0).  One read of the samples, no write.
 */
int b200spec_clip_peak(const b200spec_plan *plan, const void *d_sig, const int64_t *d_clip_off, int32_t n_clips,
                       float eps, int32_t reciprocal, float *d_peak, void *stream);

/*
 * Context stacking of DeepChromaProcessor (madmom audio/chroma.py: FramedSignal(frame_size=context,
 * hop_size=1) + _dcp_flatten; /root/reference/backend/app/services/chords/extract.py:54):
 *   out[t, c * B + j] = in[t + c - context/2, j]  inside the clip of row t, 0 outside
 * (T, B) -> (T, context * B), e.g. (T, 105) -> (T, 1575) for context 15.  Pure data movement.
 */
int b200spec_context_stack(const float *d_in, int64_t ld_in, int32_t num_bands, const int64_t *d_frame_off,
                           int32_t n_clips, int64_t total_frames, int32_t context, float *d_out, void *stream);

/*
 * Onset-strength envelope of (log-)filtered rows, librosa.onset.onset_strength semantics
 * (lag, max_size = 1, detrend = False, center shift):
 *   Lc = top_db >= 0 ? max(L, max_over_clip(L) - top_db) : L          (power_to_db top_db clip, per clip)
 *   env[t] = 0                                                     for t < lag + shift
 *   env[t] = agg_j max(0, Lc[t - shift, j] - Lc[t - shift - lag, j])   otherwise (t < frames of the clip)
 * aggregate: 0 = mean over the num_bands columns, 1 = median.  shift = n_fft / (2 hop) for center=True.
 * d_clip_max: n_clips floats of device scratch (receives each clip's maximum).  num_bands <= 1024.
 */
int b200spec_onset_envelope(const float *d_L, int64_t ld_L, int32_t num_bands, const int64_t *d_frame_off,
                            int32_t n_clips, int64_t total_frames, int32_t lag, float top_db, int32_t aggregate,
                            int32_t shift, float *d_clip_max, float *d_env, void *stream);

/*
 * Stand-alone stages on caller-supplied matrices (used when a madmom chain is not fusable,
 * e.g. SpectrogramDifferenceProcessor applied to an arbitrary array):
 *   magnitude   np.abs(complex64)                       madmom Spectrogram
 *   filter_log  np.dot(spec, fb) [+ log10(mul*y+add)]   madmom FilteredSpectrogram / LogarithmicSpectrogram
 *   diff        lagged difference per clip              madmom SpectrogramDifference(+Processor NaN-pad => first k rows 0)
 */
int b200spec_magnitude(const float *d_stft, int64_t n_elems, float *d_out, void *stream);
int b200spec_filter_log(const b200spec_plan *plan, int32_t res, const float *d_spec, int64_t ld_spec,
                        int64_t total_frames, int32_t apply_filter, int32_t apply_log, float *d_out,
                        int64_t ld_out, void *stream);
int b200spec_diff_flux_chroma(const b200spec_plan *plan, int32_t res, const float *d_L, int64_t ld_L,
                              const int64_t *d_frame_off, int32_t n_clips, int64_t total_frames,
                              const b200spec_out_desc *out, void *stream);

/* introspection used by tests and the host layer */
int b200spec_plan_num_res(const b200spec_plan *plan);
int b200spec_plan_num_bands(const b200spec_plan *plan, int32_t res);
/* how the fused kernel lays out the filterbank of one resolution: out[0] = bins per slab (odd),
 * out[1] = slab rounds per thread, out[2] = first used bin, out[3] = one past the last used bin,
 * out[4] = bands summed tap by tap in the band stage ("direct"), out[5] = their taps */
int b200spec_plan_filterbank_layout(const b200spec_plan *plan, int32_t res, int32_t out[6]);
/* number of kernel launches the library has issued in this process (bench.py's gpu_launches) */
int64_t b200spec_launch_count(void);

/*
 * Diagnostic, no GPU needed: the task plan a fused launch would use -- how a batch of n_clips clips with total_frames
 * frames is cut into tasks for num_sms x workers_per_sm workers (16 warps per SM for the warp kernel, 3-5 groups
 * for the others) that pull them from one counter.  out[0] = frames per task, out[1] = frames per task of the
 * last out[2] clips (their short tasks fill the last round; out[2] = 0: one size).  overhead_frames: cost of
 * starting a task, in frames (0.75 for a warp, 1 for a group); warmup_rows: extra rows a task transforms
 * (diff_frames for a flux-only call, else 0).  The model: W workers run ceil(tasks / W) rounds of
 * (frames per task + warmup_rows + overhead_frames); see DESIGN.md section 4, "Tasks".
 */
int b200spec_task_plan(int32_t num_sms, int32_t workers_per_sm, double overhead_frames, int64_t total_frames,
                       int32_t n_clips, int32_t warmup_rows, int32_t out[3]);

#ifdef __cplusplus
}
#endif
#if defined(B200SPEC_BUILD) && defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* B200SPEC_H_ */
