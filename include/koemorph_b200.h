/*
 * koemorph_b200 -- C ABI of the B200 (sm_100a) audio -> ARKit-blendshape inference path.
 *
 * The reference (atsuki-ichikawa/KoeMorph) has no FFI / plugin / operator API for this path:
 * its boundary is the PyTorch nn.Module surface.  These entry points are
 * therefore what a binding for the path would bind; each one names the reference code it
 * replaces (paths relative to the reference root).  The Python host side in
 * koemorph_b200/ (a mirror of src/model and src/features) calls them through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in _host;
 *   - every call is enqueued on `stream` (a cudaStream_t passed as void*) and returns immediately;
 *   - the caller owns every buffer; kernels borrow them for the stream-ordered duration of the call;
 *   - return value: 0 on success, a negative KOE_E_* code on misuse, a positive value is a cudaError_t;
 *     koe_last_error() returns a human-readable message for the calling thread;
 *   - there is no CPU fallback anywhere: without a CUDA device every call fails.
 */
#ifndef KOEMORPH_B200_H_
#define KOEMORPH_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KOE_OK 0
#define KOE_E_INVALID (-1)   /* bad argument (null pointer, size out of range, misaligned) */
#define KOE_E_UNSUPPORTED (-2) /* configuration outside what the kernels implement */
#define KOE_E_NODEVICE (-3)

#define KOE_N_MELS 80
#define KOE_N_FFT 1024
#define KOE_N_BLENDSHAPES 52
#define KOE_N_MOUTH 28
#define KOE_N_EXPR 24
#define KOE_D_MODEL 256
#define KOE_N_HEADS 8
#define KOE_NO_EDGE (-1000000) /* lo_rel_hops / hi_rel_hops: no window edge on that side */
#define KOE_MAX_EDGE 2         /* ceil((n_fft/2) / hop) for hop >= 256 */

const char* koe_last_error(void);
int koe_version(void);
/* sizeof of the public argument structs, for bindings that mirror them (ctypes, cgo, JNI): 0 koe_frontend_config,
 * 1 koe_logmel_args, 2 koe_core_weights, 3 koe_stream_args, 4 koe_forward_args; -1 for any other index */
int koe_sizeof_struct(int which);
/* number of kernels launched by this library in this process since load / since the last reset */
int64_t koe_launch_count(void);
void koe_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Log-mel frontend.
 * Replaces: SimplifiedDualStreamModel.extract_mel_features, src/model/simplified_dual_stream_model.py:166-229
 * (librosa.feature.melspectrogram(n_fft=1024, hop, n_mels=80, fmin, fmax) + power_to_db(ref=max) + (x+80)/80)
 * and the librosa calls of MelSlidingWindowExtractor, src/features/mel_sliding_window.py:224-230,280-295.
 * ------------------------------------------------------------------------------------------ */
typedef struct koe_frontend koe_frontend_t; /* opaque: Hann window, FFT twiddles, sparse Slaney filterbank on one device */

int koe_frontend_create(int device, int sample_rate, int n_fft, int n_mels, float fmin, float fmax,
                        koe_frontend_t** out);

/*
 * General form: also the torchaudio flavour of the reference's MelSpectrogramExtractor (src/features/stft.py:23-142:
 * T.MelSpectrogram(n_fft=512, hop=sr/fps, mel_scale "htk", norm None, normalized=True, pad_mode "reflect") followed by
 * log(mel + eps)).  n_fft is 1024 or 512 (a 512-point frame is transformed as the middle of a zero-extended 1024-point
 * one); window_normalized divides the power by sum(window^2); log_mode selects what koe_logmel_power* store.
 */
#define KOE_MEL_SLANEY 0
#define KOE_MEL_HTK 1
#define KOE_MEL_NORM_NONE 0
#define KOE_MEL_NORM_SLANEY 1
#define KOE_LOG_DB 0     /* 10*log10(max(p, 1e-10)) */
#define KOE_LOG_LN_EPS 1 /* ln(p + log_eps) */
typedef struct {
  int32_t device, sample_rate, n_fft, n_mels;
  float fmin, fmax;
  int32_t mel_scale, mel_norm;
  int32_t window_normalized;
  int32_t log_mode;
  float log_eps;
  int32_t win_length; /* 0 = n_fft; else the periodic Hann of win_length <= n_fft points, centred in the n_fft frame
                         (librosa / torch.stft semantics of win_length < n_fft) */
} koe_frontend_config;
int koe_frontend_create_ex(const koe_frontend_config* cfg, koe_frontend_t** out);
int koe_frontend_destroy(koe_frontend_t* fe);
/* 1 when this frontend's filterbank has the structure of the path's default bank (sr 16000, 80 mels, 80..8000 Hz,
 * simplified_dual_stream_model.py:188-199) and runs the unrolled filterbank phase; 0: generic looped kernel */
int koe_frontend_uses_unrolled_bank(const koe_frontend_t* fe);
/* copy the dense (n_mels x (1+n_fft/2)) float32 filterbank to host memory (for parity tests) */
int koe_frontend_filterbank_host(const koe_frontend_t* fe, float* fb_host);

/*
 * Mel power, in dB, of n_frames frames of every clip.  Output row j is frame g = frame_offset + j*frame_step:
 * the 1024-sample periodic-Hann frame centred on sample g*hop (librosa center=True).  Samples outside
 * [0, n_samples) read as zero (librosa pad_mode="constant"); in addition, when lo_rel_hops != KOE_NO_EDGE
 * samples before (g + lo_rel_hops)*hop read as zero, and when hi_rel_hops != KOE_NO_EDGE samples at or
 * after (g + hi_rel_hops)*hop read as zero.  The edge variants are the first / last frames of the sliding
 * windows of SequentialDualStreamModel.forward (src/model/sequential_dual_stream_model.py:101-120):
 * a window starting at frame i sees frame i with lo_rel_hops=0 and frame i+W with hi_rel_hops=0.
 *   audio      [n_clips][audio_stride] float32
 *   power      [n_clips][n_frames][80] float32   10*log10(max(sum_k fb[m][k] * |X_g[k]|^2, 1e-10)): the first term of
 *                                                librosa.power_to_db, so that a consumer only subtracts its reference
 *   frame_max  [n_clips][n_frames]     float32   (max_m power[.,j,m]); may be NULL
 */
int koe_logmel_power(const koe_frontend_t* fe, const float* audio, int64_t audio_stride, int n_clips,
                     int n_samples, int hop, int n_frames, int frame_offset, int frame_step, int lo_rel_hops,
                     int hi_rel_hops, float* power, float* frame_max, void* stream);

/*
 * General form of koe_logmel_power for the streaming paths: frame j is centred on sample
 * sample_offset + (frame_offset + j*frame_step)*hop (the window-edge masks move with it); pad_mode 1 reflects the
 * clip about its first / last sample instead of zero padding (numpy "reflect", the default of
 * MelSlidingWindowExtractor, src/features/mel_sliding_window.py:178,291); the output blocks of consecutive clips are
 * power_clip_stride / frame_max_clip_stride elements apart, so rows can land directly in per-stream ring buffers.
 */
typedef struct {
  const float* audio;
  int64_t audio_stride;
  int32_t n_clips, n_samples, hop, n_frames;
  int32_t frame_offset, frame_step, sample_offset;
  int32_t lo_rel_hops, hi_rel_hops;
  int32_t pad_mode;
  float* power;
  int64_t power_clip_stride;
  float* frame_max;
  int64_t frame_max_clip_stride;
  /* optional second destination: when power_b != NULL the odd frames (2j + 1, the second frame of every pair the kernel
   * transforms together) are written to power_b[clip][j] / frame_max_b[clip][j] instead of row 2j + 1 of `power`.  The
   * streaming step uses it with n_frames = 2: the new plain frame goes into the stream's ring, the next hop's
   * "window ends here" frame (same launch, same transform) into its own row. */
  float* power_b;
  int64_t power_b_clip_stride;
  float* frame_max_b;
  int64_t frame_max_b_clip_stride;
} koe_logmel_args;
int koe_logmel_power_ex(const koe_frontend_t* fe, const koe_logmel_args* args, void* stream);

/*
 * 16-bit PCM -> float32 audio, sample / 32768: the normalisation libsndfile applies when the reference's loaders read a
 * WAV file as float (sf.read(path, dtype="float32"), src/data/io.py:71, src/data/sequential_dataset.py:99).  Lets a host
 * application hand the samples over as stored on disk, halving the PCIe bytes of the host-buffer path; the result is
 * bit-identical to converting on the host.  Both buffers 16-byte aligned, n_samples counts samples over all clips.
 */
int koe_pcm16_to_float(const int16_t* pcm, int64_t n_samples, float* audio, void* stream);

/*
 * Rest of power_to_db on koe_logmel_power's output (ref = max over the clip's n_frames frames, top_db=80), then (x+80)/80.
 * Writes the long-term features [n_clips][n_frames][80] and the short-term detail = last three frames
 * [n_clips][3][80] (zero rows when n_frames < 3: simplified_dual_stream_model.py:206-212).
 * db_only != 0 skips the (x+80)/80 rescale (MelSlidingWindowExtractor semantics, mel_sliding_window.py:295).
 */
int koe_logmel_normalise(const float* power, const float* frame_max, int n_clips, int n_frames, int db_only,
                         float* long_term, float* short_term, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dual-stream attention core.
 * Replaces: DualStreamCrossAttention.forward, src/model/dual_stream_attention.py:162-280, and the
 * 264 -> 256 compression of OpenSMILEeGeMAPSExtractor.get_concatenated_features,
 * src/features/opensmile_extractor.py:583-604.
 * The host folds batch-invariant products once per weight update (koemorph_b200/model/folding.py);
 * all matrices below are row-major float32 and stored TRANSPOSED ([in][out]) for coalesced streaming.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t k_mel;     /* mel_sequence_length + 3: 259 (30 fps) or 515 (60 fps) */
  int32_t k_mel_pad; /* rows allocated in wc_t, multiple of 16, zero filled */
  int32_t emo_in;    /* 264 (compression folded in) or 256 (DualStreamCrossAttention standalone) */
  int32_t emo_in_pad;
  float b2;          /* blendshape_decoder.3.bias */
  float ln_eps;      /* 1e-5 */
  const float* wc_t;   /* [k_mel_pad][256]  mel_channel_encoder.weight^T                         (:211) */
  const float* bc;     /* [256] */
  const float* ln_g;   /* [256] mel_norm.weight                                                  (:212) */
  const float* ln_b;   /* [256] */
  const float* qk_t;   /* [256][256]: column h*28+q = ((Wq q_q + bq)_h / sqrt(32)) Wk_h; cols 224..255 zero (:225-230) */
  const float* wv_t;   /* [256][256] mel_attention.in_proj_weight[512:768]^T */
  const float* bv;     /* [256] */
  const float* wa_t;   /* [256][128] (decoder.0 . mel_output_proj . out_proj)^T                  (:231,248) */
  const float* ba;     /* [128] */
  const float* w2;     /* [128] blendshape_decoder.3.weight */
  const float* coef;   /* [52] 0.5*(softmax(mel_weights/T) + softmax(emotion_weights/T))         (:252-267) */
  const int32_t* mouth_idx; /* [28] */
  const int32_t* expr_idx;  /* [24] */
  const float* we1_t;  /* [emo_in_pad][256] (emotion_encoder . compression)^T                    (:216) */
  const float* be1;    /* [256] */
  const float* eln_g;  /* [256] emotion_norm */
  const float* eln_b;  /* [256] */
  const float* we2_t;  /* [256][128] (decoder.0 . emotion_output_proj . out_proj_e . Wv_e)^T     (:234-240,248) */
  const float* be2;    /* [128] */
  /* tcgen05 path (precision 2): the bf16 weight matrices pre-tiled into the UMMA no-swizzle K-major core-matrix
   * layout, as 16 KiB pipeline-stage images in consumption order (see csrc/dual_stream_tc.cu); NULL if absent */
  const void* tc_bf16;
  int32_t tc_stages;   /* number of 16 KiB stage images behind tc_bf16 */
  int32_t reserved_;
  /* tcgen05 path: the affine parts around the LayerNorm are folded into the pre-tiled weights (bc rides in the first GEMM
   * as weight row k_mel against a constant-one operand column; mel_norm.weight scales the columns of qk / wv; its bias
   * drops out of the scores (softmax-invariant) and moves into this value bias): tc_bv = wv . mel_norm.bias + bv, [256] */
  const float* tc_bv;
} koe_core_weights;

/*
 * Emotion stream: one key per clip, so softmax == 1 and all 24 expression queries give the same value
 * (dual_stream_attention.py:234-240).  emo_in [n_clips][w->emo_in] -> expr_sigmoid [n_clips].
 */
int koe_emotion_stream(const koe_core_weights* w, const float* emo_in, int n_clips, float* expr_sigmoid,
                       void* stream);

/*
 * y[r] = x[r] . w_t + b for n_rows rows: the stand-alone form of the 264 -> 256 eGeMAPS compression
 * (get_concatenated_features, src/features/opensmile_extractor.py:583-604), which SimplifiedDualStreamModel
 * .extract_emotion_features returns.  w_t is [n_in][n_out] (the nn.Linear weight transposed), row-major float32.
 */
int koe_affine_rows(const float* x, int n_rows, int n_in, const float* w_t, const float* b, int n_out, float* y,
                    void* stream);

/*
 * Windows of mel power -> blendshape frames.  Output frame i of clip b uses the window whose first frame
 * is g0 = i*stride_frames and that holds frames_per_window (T) frames; its frame k comes from
 *   power[1 + 2*k]       row (b, i)        lo-edge variant k      if k <  n_edge
 *   power[2 + 2*(T-1-k)] row (b, i)        hi-edge variant T-1-k  if k >= T - n_edge
 *   power[0]             row (b, g0 + k)   plain frames           otherwise.
 * i.e. power[0] is [n_clips][n_frames][80] over the clip's global frames, and every edge buffer is
 * [n_clips][n_out][80], one row per window (koe_logmel_power with frame_offset/frame_step = the window grid).
 * dB reference = max over the window's frames (librosa ref=np.max), -80 dB clamp, (x+80)/80,
 * long-term = first min(T, mel_seq) frames (zero padded to mel_seq), short-term = last 3 frames.
 *   frame_max[j]  same leading shape as power[j] without the 80
 *   expr_sigmoid  [n_clips] from koe_emotion_stream
 *   out           [n_clips][n_out][52]  final blendshapes (before temporal smoothing)
 *   sigmoid_out   [n_clips][n_out][52]  decoder output before the stream-weight fusion, or NULL
 *   attn_out      [n_clips][n_out][28][80] head-averaged mel attention weights, or NULL
 * precision: 0 = fp32 CUDA-core FMA; 2 = bf16 operands on tcgen05 with fp32 accumulation (1 is reserved and rejected).
 */
int koe_dual_stream_windows(const koe_core_weights* w, const float* const* power, const float* const* frame_max,
                            int n_edge, int n_clips, int n_frames, int n_out, int stride_frames,
                            int frames_per_window, const float* expr_sigmoid, float* out, float* sigmoid_out,
                            float* attn_out, int precision, void* stream);

/*
 * Same core on already-normalised features: the DualStreamCrossAttention.forward signature
 * (dual_stream_attention.py:162-168).  mel_long [n_clips][n_long][80] (zero padded / truncated to mel_seq),
 * mel_short [n_clips][3][80]; outputs as above with n_out = 1.
 */
int koe_dual_stream_features(const koe_core_weights* w, const float* mel_long, int n_long, const float* mel_short,
                             int n_clips, const float* expr_sigmoid, float* out, float* sigmoid_out,
                             float* attn_out, int precision, void* stream);

/*
 * Streaming step (rt.py-style sliding window, stride one hop, all streams in lockstep): one window per stream whose
 * frames live in per-stream rings of ring_frames mel rows.  Frame k of the window is
 *   k == 0                      lo-edge row:  power_lo_ring[s][(ring_base) % ring_frames]
 *   0 < k < frames_per_window-1 plain row:    power_ring  [s][(ring_base + k) % ring_frames]
 *   k == frames_per_window-1    hi-edge row:  power_hi    [s]
 * (30 fps geometry: hop >= n_fft/2, one edge frame per side).  Outputs as koe_dual_stream_windows with n_out = 1.
 */
int koe_dual_stream_ring(const koe_core_weights* w, const float* power_ring, const float* fmax_ring,
                         const float* power_lo_ring, const float* fmax_lo_ring, const float* power_hi,
                         const float* fmax_hi, int n_streams, int ring_frames, int ring_base, int frames_per_window,
                         const float* expr_sigmoid, float* out, float* sigmoid_out, float* attn_out, int precision,
                         void* stream);

/*
 * General form for hop < n_fft/2 (60 fps: hop 266, two edge frames per window side): power[0] / frame_max[0] the plain
 * ring, power[1 + 2m] the ring of lo-edge variant m (frame g with everything before (g - m) * hop zeroed, in the slot of
 * global frame g), power[2 + 2m] the per-stream row of hi-edge variant m (frame T-1-m of the window that ends now).
 */
int koe_dual_stream_ring_edges(const koe_core_weights* w, const float* const* power, const float* const* frame_max,
                               int n_edge, int n_streams, int ring_frames, int ring_base, int frames_per_window,
                               const float* expr_sigmoid, float* out, float* sigmoid_out, float* attn_out, int precision,
                               void* stream);

/*
 * One hop of every stream as ONE call (the real-time loop of scripts/rt.py:465-519 over
 * SimplifiedDualStreamModel.process_audio_frame_realtime, src/model/simplified_dual_stream_model.py:452-500, with the
 * sliding window of SequentialDualStreamModel.forward at stride 1): appends hop_audio to each stream's audio tail,
 * computes the new frames of the step (plain, window-start and window-end variants, SURVEY.md section 8 note E: three
 * at 30 fps, five at 60 fps where hop < n_fft/2) into the streams' rings, and -- once window_frames hops have been pushed --
 * runs koe_dual_stream_ring on the window that ends at the newest sample and smooths the result (koe_ema_scan with
 * n_out = 1).  Queues five kernels (seven at 60 fps); replaces the
 * Python driver's six separate calls (about 150 us of host time per hop).  All buffers are caller-owned device memory
 * and persist between calls; `step` counts the hops pushed before this one (0, 1, 2, ...) and selects the ping-pong tail
 * buffer (reads tail[step & 1], writes tail[(step + 1) & 1]) and the ring slot.  *emitted = 1 when `out` holds a frame.
 */
typedef struct {
  const koe_frontend_t* frontend;
  const koe_core_weights* weights;
  int32_t n_streams, hop, window_frames, half_fft; /* half_fft = n_fft / 2 (512) */
  const float* hop_audio;                          /* [n_streams][hop]: the next hop of every stream */
  float* tail[2];                                  /* [n_streams][half_fft + n_edge * hop] each, n_edge = ceil(half_fft / hop);
                                                      zero before the first hop */
  float* ring_f; float* fmax_f;                    /* [n_streams][window_frames][80], [n_streams][window_frames] */
  float* ring_r; float* fmax_r;                    /* same shapes: the window-start variants */
  float* row_l; float* fmax_l;                     /* [n_streams][80], [n_streams]: the window-end variant */
  /* hop < n_fft/2 (60 fps): the second edge frame of either side -- frame g with everything before (g - 1) * hop zeroed
   * (ring) and the last-but-one frame of the window that ends now (row); NULL when hop >= n_fft/2 */
  float* ring_r2; float* fmax_r2;
  float* row_l2; float* fmax_l2;
  const float* expr_sigmoid;                       /* [n_streams], from koe_emotion_stream */
  float* out;                                      /* [n_streams][52] */
  float* ema_state;                                /* [n_streams][52], or NULL: no temporal smoothing */
  float alpha;                                     /* sigmoid(smoothing_alpha) */
  int32_t has_state;                               /* 0 for the first emitted frame of a stream set */
  int32_t precision;
  int64_t step;
} koe_stream_args;
int koe_stream_push(const koe_stream_args* args, int* emitted, void* stream);

/*
 * The whole forward of SequentialDualStreamModel.forward (src/model/sequential_dual_stream_model.py:63-167) -- or, with
 * n_out = 1 and n_edge = 0, of SimplifiedDualStreamModel.forward (src/model/simplified_dual_stream_model.py:370-415) --
 * as ONE call: koe_logmel_power on the clip's global frames, the 2 * n_edge edge-variant launches on the window grid,
 * the emotion stream, koe_dual_stream_windows and (smooth != 0, n_out > 1) koe_ema_scan, queued back to back on `stream`.
 * Because the call knows which kernel precedes which, every kernel is chained to the one before it with programmatic
 * dependent launch (it takes SMs as that kernel's CTAs leave them and orders itself with griddepcontrol.wait); the public
 * single-kernel entries above keep plain stream order.  The expression entries of `out` depend on the emotion stream only
 * and the mouth entries on the mel stream only, so here the emotion kernel writes its 24 entries of every output row
 * itself (and expr_sigmoid, as koe_emotion_stream would) and the core leaves them alone: neither waits for the other's
 * results.  With one window per clip (n_out = 1, n_edge = 0, precision 2, no attention output) and more clips than SMs,
 * the core does not wait for the whole frontend either: it starts every round of windows but its last on a release /
 * acquire flag that the frontend counts up once those clips are stored (two words per device and stream, owned by the
 * library, zero between calls).  Results are identical to the kernels run one after the other.
 * Window i of a clip starts at global frame i * stride_frames and holds
 * frames_per_window frames; n_frames >= (n_out - 1) * stride_frames + frames_per_window global frames are computed.
 * Workspace (caller-owned): power[0] / frame_max[0] [n_clips][n_frames][80] / [n_clips][n_frames]; for m < n_edge
 * power[1 + 2m], power[2 + 2m] [n_clips][n_out][80] (+ frame_max); expr_sigmoid [n_clips].
 */
typedef struct {
  const koe_frontend_t* frontend;
  const koe_core_weights* weights;
  const float* audio;      /* [n_clips][audio_stride] */
  int64_t audio_stride;
  int32_t n_clips, n_samples, hop;
  int32_t n_frames, frames_per_window, stride_frames, n_out, n_edge;
  const float* egemaps;    /* [n_clips][weights->emo_in] */
  float* power[1 + 2 * KOE_MAX_EDGE];
  float* frame_max[1 + 2 * KOE_MAX_EDGE];
  float* expr_sigmoid;
  float* out;              /* [n_clips][n_out][52] */
  float* sigmoid_out;      /* or NULL */
  float* attn_out;         /* or NULL */
  float alpha;             /* sigmoid(smoothing_alpha) */
  int32_t smooth;          /* != 0: EMA over the n_out frames of every clip (first frame passes through) */
  int32_t precision;
} koe_forward_args;
int koe_forward_windows(const koe_forward_args* args, void* stream);

/*
 * Learnable-alpha exponential smoothing along the frame axis, in place
 * (apply_temporal_smoothing, src/model/simplified_dual_stream_model.py:341-368):
 *   y_0 = x_0 (or alpha*x_0 + (1-alpha)*state when state != NULL and has_state != 0); y_t = alpha*x_t + (1-alpha)*y_{t-1}.
 * frames [n_clips][n_out][52]; state [n_clips][52] receives y_{n_out-1} when non-NULL.
 */
int koe_ema_scan(float* frames, int n_clips, int n_out, float alpha, float* state, int has_state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KOEMORPH_B200_H_ */
