// One streaming step as ONE native call: the hop-aligned sliding window of scripts/rt.py:465-519 /
// SimplifiedDualStreamModel.process_audio_frame_realtime (src/model/simplified_dual_stream_model.py:452-500) for many
// lock-step streams.  The Python driver used to issue the step as a tensor concatenation, three frontend calls, the core
// call and the smoothing call: six kernels but ~150 us of interpreter / ctypes / allocator time per hop, a third of the
// latency of 4096 streams and nearly all of it for a handful.  Here the five launches are queued back to back from C++.
//
// State per stream (all caller-owned device memory, see koe_stream_args): the last n_fft/2 + hop samples (two buffers,
// ping-pong: step n reads tail[n & 1] and writes tail[(n + 1) & 1]), rings of W plain (F) and "window starts here" (R)
// mel rows with their per-frame maxima, one "window ends here" row (L), the smoothing state.  SURVEY.md section 8
// note E: frame n is centred on sample n * hop; R[n] is the same frame with everything before its centre zeroed; L[n + 1]
// is centred on (n + 1) * hop with everything from there on zeroed.
#include <map>
#include <mutex>
#include <utility>

#include "core_params.cuh"

namespace koe {

// new_tail[s] = [old_tail[s][hop:], hop_audio[s]]: one CTA per stream row (strided over the grid), no index division
__global__ void __launch_bounds__(256) stream_tail_shift_kernel(const float* __restrict__ old_tail,
                                                               const float* __restrict__ hop_audio, int n_streams,
                                                               int tail_len, int hop, float* __restrict__ new_tail) {
  const int keep = tail_len - hop;
  for (int s = blockIdx.x; s < n_streams; s += gridDim.x) {
    const float* o = old_tail + (long long)s * tail_len + hop;
    const float* x = hop_audio + (long long)s * hop;
    float* t = new_tail + (long long)s * tail_len;
    for (int k = threadIdx.x; k < keep; k += blockDim.x) t[k] = o[k];
    for (int k = threadIdx.x; k < hop; k += blockDim.x) t[keep + k] = __ldcs(x + k);
  }
}

}  // namespace koe

using namespace koe;

extern "C" int koe_stream_push(const koe_stream_args* a, int* emitted, void* stream_v) {
  KOE_REQUIRE(a != nullptr && emitted != nullptr, "koe_stream_push: NULL argument");
  KOE_REQUIRE(a->frontend != nullptr && a->weights != nullptr, "koe_stream_push: NULL frontend / weights");
  KOE_REQUIRE(a->n_streams >= 0 && a->hop > 0 && a->window_frames > 0 && a->half_fft > 0 && a->step >= 0,
              "koe_stream_push: bad geometry");
  // frames within n_fft/2 of a window edge see zeros beyond it: ceil(half / hop) frames per side (1 at 30 fps, 2 at 60 fps)
  const int n_edge = (a->half_fft + a->hop - 1) / a->hop;
  KOE_REQUIRE(n_edge >= 1 && n_edge <= KOE_MAX_EDGE, "koe_stream_push: hop %d needs %d edge frames per window side (max %d)",
              a->hop, n_edge, KOE_MAX_EDGE);
  KOE_REQUIRE(a->hop_audio && a->tail[0] && a->tail[1] && a->ring_f && a->fmax_f && a->ring_r && a->fmax_r && a->row_l &&
                  a->fmax_l && a->expr_sigmoid && a->out,
              "koe_stream_push: NULL buffer");
  KOE_REQUIRE(n_edge == 1 || (a->ring_r2 && a->fmax_r2 && a->row_l2 && a->fmax_l2),
              "koe_stream_push: hop < n_fft/2 needs the second edge buffers (ring_r2, row_l2)");
  *emitted = 0;
  if (a->n_streams == 0) return KOE_OK;
  cudaStream_t stream = (cudaStream_t)stream_v;
  const int S = a->n_streams, hop = a->hop, W = a->window_frames, half = a->half_fft, tail_len = half + n_edge * hop;
  const int64_t n = a->step;
  const float* old_tail = a->tail[n & 1];
  float* tail = a->tail[(n + 1) & 1];
  {
    int dev = 0, sms = 0;
    KOE_CUDA(cudaGetDevice(&dev));
    KOE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = std::min(S, sms * 8);
    stream_tail_shift_kernel<<<grid, 256, 0, stream>>>(old_tail, a->hop_audio, S, tail_len, hop, tail);
    count_launch();
    KOE_CUDA(cudaGetLastError());
  }
  // The tail ends at sample (n + 1) * hop.  The newest frame whose samples are all there is global frame g = n - lag,
  // lag = n_edge - 1, centred `half` samples into the tail; the frames centred one hop further each are the window-end
  // variants of the window that ends now: beyond the end of the tail the clip reads as zeros, which is exactly their mask.
  const int lag = n_edge - 1;
  const int64_t g = n - lag;
  koe_logmel_args f = {};
  f.audio = tail;
  f.audio_stride = tail_len;
  f.n_clips = S;
  f.n_samples = tail_len;
  f.hop = hop;
  f.frame_offset = 0;
  f.frame_step = 1;
  f.pad_mode = 0;
  f.sample_offset = half;
  f.lo_rel_hops = KOE_NO_EDGE, f.hi_rel_hops = KOE_NO_EDGE;
  f.power_clip_stride = (int64_t)W * KOE_N_MELS, f.frame_max_clip_stride = W;
  f.power_b_clip_stride = KOE_N_MELS, f.frame_max_b_clip_stride = 1;
  const int slot = (int)(((g % W) + W) % W);
  // plain frame g (-> ring slot) and the frame one hop later as ONE pair of one launch: at 30 fps that second frame is the
  // window-end variant L (centred on the tail's end), at 60 fps the last-but-one frame of the window (row_l2)
  f.n_frames = 2;
  f.power = a->ring_f + (size_t)slot * KOE_N_MELS, f.frame_max = a->fmax_f + slot;
  f.power_b = n_edge == 1 ? a->row_l : a->row_l2, f.frame_max_b = n_edge == 1 ? a->fmax_l : a->fmax_l2;
  if (int rc = koe_logmel_power_ex(a->frontend, &f, stream)) return rc;
  f.power_b = nullptr, f.frame_max_b = nullptr;
  if (n_edge == 2) {  // the frame centred on the tail's end: the window's last frame
    f.n_frames = 1;
    f.frame_offset = 2;
    f.power = a->row_l, f.power_clip_stride = KOE_N_MELS;
    f.frame_max = a->fmax_l, f.frame_max_clip_stride = 1;
    if (int rc = koe_logmel_power_ex(a->frontend, &f, stream)) return rc;
    f.frame_offset = 0;
    f.power_clip_stride = (int64_t)W * KOE_N_MELS, f.frame_max_clip_stride = W;
  }
  // window-start variants of frame g: nothing before its centre (first frame of the window that starts at g), and at
  // 60 fps nothing before the centre of frame g - 1 (second frame of the window that starts at g - 1)
  f.n_frames = 1;
  for (int m = 0; m < n_edge; ++m) {
    f.lo_rel_hops = -m;
    f.power = (m == 0 ? a->ring_r : a->ring_r2) + (size_t)slot * KOE_N_MELS;
    f.frame_max = (m == 0 ? a->fmax_r : a->fmax_r2) + slot;
    if (int rc = koe_logmel_power_ex(a->frontend, &f, stream)) return rc;
  }
  if (n + 1 < W) return KOE_OK;  // the context is not full yet
  const int base = (int)((n + 1 - W) % W);  // ring slot of the first frame of the window that ends now
  const float* power[1 + 2 * KOE_MAX_EDGE] = {a->ring_f, a->ring_r, a->row_l, a->ring_r2, a->row_l2};
  const float* fmax[1 + 2 * KOE_MAX_EDGE] = {a->fmax_f, a->fmax_r, a->fmax_l, a->fmax_r2, a->fmax_l2};
  if (int rc = koe_dual_stream_ring_edges(a->weights, power, fmax, n_edge, S, W, base, W + 1, a->expr_sigmoid, a->out,
                                          nullptr, nullptr, a->precision, stream))
    return rc;
  if (a->ema_state != nullptr)
    if (int rc = koe_ema_scan(a->out, S, 1, a->alpha, a->ema_state, a->has_state, stream)) return rc;
  *emitted = 1;
  return KOE_OK;
}

// ---- early release of the core ------------------------------------------------------------------------------------------
// With one window per clip the core's CTA b handles clips b, b + G, b + 2G, ... (G = its grid).  The frontend stores the
// clips in index order, so all but the last round of windows can be read long before the frontend's slowest CTA has
// finished -- and the core's CTAs become resident exactly then, as the frontend's first CTAs leave.  The frontend's
// consumer warps count up a flag when the first E = G * floor((n - 1) / G) clips are stored; the core acquires it instead
// of waiting for the whole frontend, and executes griddepcontrol.wait only before its last round (step 199.1 -> 191.9 us
// at most, measured with a probe that did not wait at all).  The flag and the core's exit counter live in two words per
// (device, stream), zero between forwards: the last core CTA out clears them.
namespace {
struct EarlyFlags {
  static constexpr int kPool = 64;   // word pairs per device, allocated (and zeroed) at the first use outside a capture
  std::mutex mu;
  std::map<std::pair<int, void*>, unsigned*> words;
  std::map<int, std::pair<unsigned*, int>> pool;   // device -> (block, pairs handed out)
  unsigned* get(void* stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    auto it = words.find({dev, stream});
    if (it != words.end()) return it->second;
    auto pl = pool.find(dev);
    if (pl == pool.end()) {
      // (no allocation inside a stream capture: a forward captured before any eager one takes the plain chain)
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing((cudaStream_t)stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return nullptr;
      unsigned* block = nullptr;
      if (cudaMalloc(&block, kPool * 2 * sizeof(unsigned)) != cudaSuccess) return nullptr;
      if (cudaMemset(block, 0, kPool * 2 * sizeof(unsigned)) != cudaSuccess) {
        cudaFree(block);
        return nullptr;
      }
      pl = pool.emplace(dev, std::make_pair(block, 0)).first;
    }
    if (pl->second.second >= kPool) return nullptr;   // more streams than pairs: those forwards take the plain chain
    unsigned* w = pl->second.first + 2 * pl->second.second++;
    words[{dev, stream}] = w;
    return w;
  }
};
EarlyFlags g_early;
}  // namespace

// ---- the batch forward as one native call (see koe_forward_windows in the header) ------------------------------------
extern "C" int koe_forward_windows(const koe_forward_args* a, void* stream) {
  KOE_REQUIRE(a != nullptr && a->frontend != nullptr && a->weights != nullptr, "koe_forward_windows: NULL argument");
  KOE_REQUIRE(a->audio && a->egemaps && a->power[0] && a->frame_max[0] && a->expr_sigmoid && a->out,
              "koe_forward_windows: NULL buffer");
  KOE_REQUIRE(a->n_clips >= 0 && a->n_samples >= 0 && a->hop > 0 && a->n_out >= 1 && a->stride_frames >= 1 &&
                  a->frames_per_window >= 1 && a->n_edge >= 0 && a->n_edge <= KOE_MAX_EDGE,
              "koe_forward_windows: bad geometry");
  KOE_REQUIRE((long long)(a->n_out - 1) * a->stride_frames + a->frames_per_window <= a->n_frames,
              "koe_forward_windows: windows run past n_frames");
  if (a->n_clips == 0) return KOE_OK;
  koe_logmel_args f = {};
  f.audio = a->audio;
  f.audio_stride = a->audio_stride;
  f.n_clips = a->n_clips;
  f.n_samples = a->n_samples;
  f.hop = a->hop;
  f.pad_mode = 0;
  f.sample_offset = 0;
  // every global frame once
  f.n_frames = a->n_frames;
  f.frame_offset = 0, f.frame_step = 1;
  f.lo_rel_hops = KOE_NO_EDGE, f.hi_rel_hops = KOE_NO_EDGE;
  f.power = a->power[0], f.power_clip_stride = (int64_t)a->n_frames * KOE_N_MELS;
  f.frame_max = a->frame_max[0], f.frame_max_clip_stride = a->n_frames;
  // early release of the core (above): one window per clip, tensor-core core, no attention output, more than one round
  static const bool early_off = getenv("KOE_NO_EARLY_CORE") != nullptr;  // experiment switch
  unsigned* early_flag = nullptr;
  unsigned early_target = 0;
  int early_clips = 0;
  if (!early_off && a->n_out == 1 && a->n_edge == 0 && a->precision == 2 && a->attn_out == nullptr &&
      (a->weights->k_mel == 259 || a->weights->k_mel == 515)) {
    const int G = dual_stream_tc_grid(a->n_clips);
    if (G > 0 && a->n_clips > G) {
      early_clips = G * ((a->n_clips - 1) / G);
      early_flag = g_early.get(stream);
    }
  }
  if (int rc = launch_logmel(a->frontend, &f, stream, /*follows_frontend_launch=*/false, early_clips, early_flag, &early_target))
    return rc;
  if (early_target == 0) early_flag = nullptr;
  bool early_applied = false;
  // the frames within n_fft/2 of a window edge, once per window (SURVEY.md section 8 note E)
  for (int m = 0; m < a->n_edge; ++m) {
    KOE_REQUIRE(a->power[1 + 2 * m] && a->power[2 + 2 * m] && a->frame_max[1 + 2 * m] && a->frame_max[2 + 2 * m],
                "koe_forward_windows: NULL edge buffer");
    f.n_frames = a->n_out;
    f.frame_step = a->stride_frames;
    f.power_clip_stride = (int64_t)a->n_out * KOE_N_MELS, f.frame_max_clip_stride = a->n_out;
    // (these launches follow a frontend launch of this forward and touch none of its buffers: they do not wait for it)
    f.frame_offset = m, f.lo_rel_hops = -m, f.hi_rel_hops = KOE_NO_EDGE;
    f.power = a->power[1 + 2 * m], f.frame_max = a->frame_max[1 + 2 * m];
    if (int rc = launch_logmel(a->frontend, &f, stream, /*follows_frontend_launch=*/true, 0, nullptr, nullptr)) return rc;
    f.frame_offset = a->frames_per_window - 1 - m, f.lo_rel_hops = KOE_NO_EDGE, f.hi_rel_hops = m;
    f.power = a->power[2 + 2 * m], f.frame_max = a->frame_max[2 + 2 * m];
    if (int rc = launch_logmel(a->frontend, &f, stream, /*follows_frontend_launch=*/true, 0, nullptr, nullptr)) return rc;
  }
  // The mouth entries of every output row come from the mel stream (the core) and the expression entries from the emotion
  // stream alone, so the two kernels do not depend on each other's results: the emotion kernel writes its entries of `out`
  // itself and the core skips them.  The kernel queued last is a frontend launch; the emotion stream (which reads none of
  // its outputs) runs on the SMs the frontend's first CTAs leave, the core sets itself up behind both.
  if (int rc = launch_emotion_stream(a->weights, a->egemaps, a->n_clips, a->expr_sigmoid, a->out, a->sigmoid_out, a->n_out,
                                     stream, /*after_frontend=*/true)) {
    if (early_flag != nullptr) cudaMemsetAsync(early_flag, 0, 2 * sizeof(unsigned), (cudaStream_t)stream);
    return rc;
  }
  if (int rc = launch_dual_stream_windows(a->weights, a->power, a->frame_max, a->n_edge, a->n_clips, a->n_frames, a->n_out,
                                          a->stride_frames, a->frames_per_window, nullptr, a->out, a->sigmoid_out,
                                          a->attn_out, a->precision, stream, /*expr_by_emotion_kernel=*/true, early_flag,
                                          early_target, early_clips, &early_applied)) {
    if (early_flag != nullptr) cudaMemsetAsync(early_flag, 0, 2 * sizeof(unsigned), (cudaStream_t)stream);  // frontend counted, nobody clears
    return rc;
  }
  // (the frontend counted but the core took the plain kernel: the flag must not survive this forward)
  if (early_flag != nullptr && !early_applied) KOE_CUDA(cudaMemsetAsync(early_flag, 0, 2 * sizeof(unsigned), (cudaStream_t)stream));
  if (a->smooth && a->n_out > 1)
    if (int rc = koe_ema_scan(a->out, a->n_clips, a->n_out, a->alpha, nullptr, 0, stream)) return rc;
  return KOE_OK;
}
