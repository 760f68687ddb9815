// Launch parameters shared by the CUDA-core and the tcgen05 implementations of the dual-stream core.
#pragma once
#include "common.cuh"

namespace koe {

struct CoreParams {
  koe_core_weights w;
  const float* power[1 + 2 * KOE_MAX_EDGE];
  const float* fmax[1 + 2 * KOE_MAX_EDGE];
  int n_edge, n_clips, n_frames, n_out, stride_frames, frames_per_window, mel_seq;
  const float* expr_sigmoid;
  float* out;
  float* sigmoid_out;
  float* attn_out;
  // prenormalised mode (koe_dual_stream_features): rows come from mel_long / mel_short as they are
  const float* mel_long;
  const float* mel_short;
  int n_long;
  // ring mode (koe_dual_stream_ring): plain and lo-edge rows are slots of per-stream rings, the hi-edge row is per stream
  int ring_frames, ring_base;
  // early release (csrc/session.cu): items [0, early_items) may be read once *early_flag >= early_target (acquire); the
  // rest after griddepcontrol.wait.  early_flag == NULL: everything after griddepcontrol.wait, at the start of the kernel.
  // early_flag[1] counts the CTAs that have finished; the last one clears both words for the next forward.
  unsigned* early_flag;
  unsigned early_target;
  int early_items;
  long long* dbg;  // optional phase timestamps of CTA 0 (bring-up / profiling only), NULL in production
};

// frame k of window wi of clip b: which buffer and which row (see koe_dual_stream_windows in the header)
__device__ __forceinline__ int window_variant(const CoreParams& p, int k) {
  if (k < p.n_edge) return 1 + 2 * k;
  if (k >= p.frames_per_window - p.n_edge) return 2 + 2 * (p.frames_per_window - 1 - k);
  return 0;
}
__device__ __forceinline__ long long window_row(const CoreParams& p, int variant, int b, int wi, int k) {
  if (p.ring_frames > 0)  // plain and lo-edge rows are ring slots of the global frame; a hi-edge row (variants 2, 4) is per stream
    return variant > 0 && (variant & 1) == 0 ? (long long)b
                                             : (long long)b * p.ring_frames + (p.ring_base + k) % p.ring_frames;
  return variant == 0 ? (long long)b * p.n_frames + (long long)wi * p.stride_frames + k
                      : (long long)b * p.n_out + wi;
}

// emotion stream launch shared by the public entry and the fused forward (csrc/session.cu, which has it write the
// expression entries of `out`); the core launch that leaves those entries alone (expr_sigmoid == NULL); see dual_stream.cu
int launch_emotion_stream(const koe_core_weights* w, const float* emo_in, int n_clips, float* expr_sigmoid, float* out,
                          float* sigmoid_out, int n_out, void* stream, bool after_frontend);
int launch_dual_stream_windows(const koe_core_weights* w, const float* const* power, const float* const* frame_max, int n_edge,
                               int n_clips, int n_frames, int n_out, int stride_frames, int frames_per_window,
                               const float* expr_sigmoid, float* out, float* sigmoid_out, float* attn_out, int precision,
                               void* stream, bool expr_by_emotion_kernel, unsigned* early_flag = nullptr,
                               unsigned early_target = 0, int early_items = 0, bool* early_applied = nullptr);
// how many SMs' worth of CTAs the tensor-core core launches for n_items windows (the early-release split follows it)
int dual_stream_tc_grid(int n_items);

}  // namespace koe
