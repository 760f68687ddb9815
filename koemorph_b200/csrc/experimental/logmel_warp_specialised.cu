// EXPERIMENT -- NOT BUILT, NOT SHIPPED.  Warp-specialised variant of logmel.cu kept for the record (round 1):
// 13 producer warps (FFT of one frame pair each, audio staged by cp.async.bulk + mbarrier, pairs handed out by ticket,
// ping-pong spectrum slots, plane-wise 32x32 transpose) + 3 consumer warps (filterbank with lane = frame, immediate
// weights), rows stored by the producers one pair later.  Parity: identical results to logmel.cu (scripts/k1_check.py).
// Measured on B200, 512 clips x 257 frames: 237-265 us against 222 us for the lock-step kernel in ../logmel.cu.
// Why it lost (per-warp clock64 instrumentation + ncu source counters): with 128 registers per
// thread an SM holds 16 warps, every warp advances at ~1 instruction per 8-10 cycles (FMA-pipe contention in the
// butterflies, shared-memory latency elsewhere), so three warps that do not transform cost 3/16 of the FFT throughput
// while the consumers' single-warp dependency chains (4-5k cycles per batch) put the hand-over on the critical path.
// It needs melbank_default_3runs.inc (generated with WARPS = 3 and baked weights) in place of melbank_default.inc.
//
// Log-mel frontend for sm_100a: framing + periodic Hann + 1024-point FFT + |X|^2 + sparse Slaney
// filterbank, fused in one kernel; then a small dB-normalisation kernel.
//
// Replaces the per-clip librosa loop of SimplifiedDualStreamModel.extract_mel_features
// (reference src/model/simplified_dual_stream_model.py:184-229).
//
// Kernel design (see DESIGN.md section "K1"): one persistent, warp-specialised CTA of 16 warps per SM.
//   * PRODUCER warps (12, three per SM sub-partition): each transforms TWO real frames of one clip at once as one
//     complex 1024-point FFT (z = a + i b), decomposed 32 x 32: radix-2 DIT FFT of 32 points in registers, twiddle,
//     32x32 transpose through a padded shared-memory tile, second 32-point FFT in registers.  Every complex value is
//     one 64-bit register pair and every butterfly is written with the packed fp32x2 instructions of sm_100
//     (FADD2 / FMUL2 / FFMA2, whose operands take per-half negation and a half swap, so "times -i" is free and a
//     butterfly with a general twiddle is 3 instructions: x = a + w b as two FFMA2, y = 2a - x as one).
//     After the second pass lane l holds Z[l + 32 r]; the conjugate-symmetric partner Z[1024 - k] lives in lane
//     (32 - l) & 31, so the two real spectra are separated with 32 warp shuffles; the pair (|A_k|^2, |B_k|^2) comes
//     out of one FMUL2 + one FFMA2 and is parked in the warp's tile.
//     The 1557 contiguous samples of the NEXT pair are fetched by one TMA bulk copy (cp.async.bulk, completion on an
//     mbarrier) into a per-warp staging buffer while the current pair is transformed: no load instruction, no
//     register, and no DRAM latency on the producers' critical path.
//   * CONSUMER warps (4, one per sub-partition) apply the Slaney filterbank with lane = frame (24 frames per batch):
//     the four warps split the bins, every weight is a constant-bank operand of straight-line code unrolled from
//     the bank's structure (melbank_default.inc), every spectrum read is a conflict-free LDS.128 and the control
//     flow is warp-uniform; then they convert to dB and write coalesced rows of 80 values plus the per-frame max.
//   * producers and consumers meet only on two mbarriers (spectra full / spectra consumed): nobody waits at a
//     CTA-wide barrier, so the FMA pipe, the shared-memory pipe and the copy engine overlap across warps.
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace koe {

constexpr int kFrameLen = 1024;
constexpr int kBins = 513;
constexpr int kProducers = 13;          // FFT warps: one frame pair each per batch
constexpr int kConsumers = 3;           // filterbank / store warps
constexpr int kWarps = kProducers + kConsumers;
constexpr int kThreads = kWarps * 32;
constexpr int kBatchFrames = 2 * kProducers;  // 26 <= 32: one lane per frame in the filterbank phase
constexpr int kRow = 33;                // floats per row of the transposition tile (padded: conflict-free both ways)
constexpr int kSlotFloats = 32 * kRow;  // 1056 floats: one plane of the 32x32 transpose, then the two power spectra
constexpr int kSecondFrame = 516;       // offset of the second spectrum, == 4 (mod 32)
constexpr int kStageFloats = 1576;      // TMA staging: 3 + 533 + 1024 samples rounded up, sized so that ...
constexpr int kProdStride = 2 * kSlotFloats + kStageFloats;  // (two slots, ping-pong) ... the per-producer stride is == 8 (mod 32) floats: with
                                        // lane = frame the bases of 8 consecutive frames are 16 bytes apart mod 128
                                        // -> the consumers' LDS.128 are conflict-free
constexpr int kMaxTmaSpan = kStageFloats - 4;  // samples from the first of frame A to the last of frame B
constexpr int kMaxBins = 512;           // spectrum bins that carry filterbank weight (506 for 80..8000 Hz)
constexpr int kMaxGroups = 96;          // groups of consecutive bins feeding the same pair of adjacent filters
constexpr int kTileStride = 81;         // mel staging row stride (floats), odd: conflict-free across frames
static_assert(kProdStride % 32 == 8 && kSecondFrame % 32 == 4 && kSlotFloats * 4 % 16 == 0 &&
                  kSecondFrame + kBins <= kSlotFloats, "bank skew / slot size");

struct FrontendTables {
  const float* hann;     // [1024]
  const float2* tw;      // [32][32] W_1024^(k1*n2)
  // Slaney filterbank, bin-major: a spectrum bin feeds at most two ADJACENT filters (fl, fl + 1)
  const float2* binw;    // [n_bins] 0.25 * (weight into filter fl, weight into filter fl + 1)
  const int4* groups;    // [n_groups] {first bin k, first entry of binw, number of bins, fl}: fl rises by one per group
  const int* runs;       // [kConsumers + 1] group range of every consumer warp (generic bank)
  int n_bins, n_groups;
};

struct LogmelParams {
  const float* audio;
  int64_t audio_stride;
  int n_clips, n_samples, hop, n_frames;
  int lo_rel, hi_rel;
  int frame_offset, frame_step;  // output row j is the frame centred on sample_offset + (frame_offset + j*frame_step)*hop
  int sample_offset;
  int pad_mode;                  // 0: samples outside the clip are zero; 1: reflected (numpy "reflect")
  float* power;
  float* frame_max;
  long long power_clip_stride, fmax_clip_stride;  // elements between consecutive clips' output blocks
};

__host__ __device__ constexpr int bitrev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// cos(2*pi*t/32) for t = 0..15 as literals so the unrolled butterflies use immediates
__device__ __forceinline__ float cos32(int t) {
  switch (t) {
    case 0: return 1.0f;
    case 1: return 0.98078528040323043f;
    case 2: return 0.92387953251128674f;
    case 3: return 0.83146961230254524f;
    case 4: return 0.70710678118654752f;
    case 5: return 0.55557023301960218f;
    case 6: return 0.38268343236508978f;
    case 7: return 0.19509032201612825f;
    case 8: return 0.0f;
    case 9: return -0.19509032201612825f;
    case 10: return -0.38268343236508978f;
    case 11: return -0.55557023301960218f;
    case 12: return -0.70710678118654752f;
    case 13: return -0.83146961230254524f;
    case 14: return -0.92387953251128674f;
    default: return -0.98078528040323043f;
  }
}

// ---- complex arithmetic on (re, im) register pairs with the packed fp32x2 pipe ----------------------------------
__device__ __forceinline__ float2 bcast(float s) { return make_float2(s, s); }
// a + w*b and a - w*b for a compile-time twiddle w = W_32^t = cos(2 pi t/32) - i sin(2 pi t/32)
__device__ __forceinline__ void butterfly(int t, float2& a, float2& b) {
  if (t == 0) {
    const float2 x = __fadd2_rn(a, b);
    b = __fadd2_rn(a, make_float2(-b.x, -b.y));
    a = x;
  } else if (t == 8) {  // w = -i: w*b = (b.y, -b.x)
    const float2 x = __fadd2_rn(a, make_float2(b.y, -b.x));
    b = __fadd2_rn(a, make_float2(-b.y, b.x));
    a = x;
  } else {
    const float wr = cos32(t);
    const float ws = cos32(t > 8 ? t - 8 : 8 - t);  // sin(2 pi t/32)
    // w*b = (wr b.x + ws b.y, wr b.y - ws b.x)
    float2 x = __ffma2_rn(b, bcast(wr), a);
    x = __ffma2_rn(make_float2(b.y, -b.x), bcast(ws), x);
    b = __ffma2_rn(a, bcast(2.0f), make_float2(-x.x, -x.y));
    a = x;
  }
}

// In-register radix-2 decimation-in-time FFT of 32 complex values (forward, e^{-i...}).
// On entry element i holds x[bitrev5(i)]; on return element k holds X[k].
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
#pragma unroll
  for (int h = 1; h <= 16; h <<= 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if ((i & h) == 0) butterfly((i & (h - 1)) * (16 / h), v[i], v[i + h]);
    }
  }
}

// Where a frame pair lives: clip, first frame (< 0: no pair), first sample of both frames, valid sample ranges.
struct PairInfo {
  int clip, frame;
  int fa_lo, fb_lo;            // first sample of frame A / frame B (may be negative)
  int lo_a, hi_a, lo_b, hi_b;  // samples outside [lo, hi) read as zero (or are reflected, pad_mode 1)
  bool interior, has_b;
  bool tma;                    // the pair's samples come through the staging buffer
  int stage_off;               // index of frame A's first sample in the staging buffer (0..3)
  unsigned tma_bytes;
  const float* tma_src;        // 16-byte aligned
};

// pair number `pic` of clip `b` (frames 2 pic, 2 pic + 1); b >= n_clips: no pair
__device__ __forceinline__ PairInfo describe_pair(const LogmelParams& p, int b, int pic) {
  PairInfo pi;
  pi.clip = 0;
  pi.frame = -1;
  pi.interior = false;
  pi.tma = false;
  pi.has_b = false;
  if (b >= p.n_clips) return pi;
  const int ga = 2 * pic;
  pi.clip = b;
  pi.frame = ga;
  pi.has_b = ga + 1 < p.n_frames;
  const int fa = p.frame_offset + ga * p.frame_step, fb = fa + p.frame_step;  // frame indices in hops
  pi.lo_a = 0, pi.hi_a = p.n_samples, pi.lo_b = 0, pi.hi_b = p.n_samples;
  if (p.lo_rel != KOE_NO_EDGE) {
    pi.lo_a = max(pi.lo_a, p.sample_offset + (fa + p.lo_rel) * p.hop);
    pi.lo_b = max(pi.lo_b, p.sample_offset + (fb + p.lo_rel) * p.hop);
  }
  if (p.hi_rel != KOE_NO_EDGE) {
    pi.hi_a = min(pi.hi_a, p.sample_offset + (fa + p.hi_rel) * p.hop);
    pi.hi_b = min(pi.hi_b, p.sample_offset + (fb + p.hi_rel) * p.hop);
  }
  if (!pi.has_b) pi.hi_b = pi.lo_b;  // empty range: second frame reads as silence
  pi.fa_lo = p.sample_offset + fa * p.hop - kFrameLen / 2;
  pi.fb_lo = p.sample_offset + fb * p.hop - kFrameLen / 2;
  pi.interior = pi.fa_lo >= pi.lo_a && pi.fa_lo + kFrameLen <= pi.hi_a && pi.fb_lo >= pi.lo_b &&
                pi.fb_lo + kFrameLen <= pi.hi_b;
  // one bulk copy covers [first sample of A, last sample of B], widened to 16-byte boundaries; it must stay inside the buffer
  const int span = pi.fb_lo - pi.fa_lo + kFrameLen;
  if (pi.interior && span <= kMaxTmaSpan) {
    const float* first = p.audio + (long long)b * p.audio_stride + pi.fa_lo;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(first);
    const uintptr_t base = addr & ~(uintptr_t)15;
    pi.stage_off = (int)((addr - base) >> 2);
    pi.tma_bytes = (unsigned)(((pi.stage_off + span) * 4 + 15) & ~15);
    pi.tma_src = reinterpret_cast<const float*>(base);
    const uintptr_t buf_end = reinterpret_cast<uintptr_t>(p.audio + (long long)(p.n_clips - 1) * p.audio_stride + p.n_samples);
    pi.tma = base >= reinterpret_cast<uintptr_t>(p.audio) && base + pi.tma_bytes <= buf_end;
  }
  return pi;
}

// A warp's position in the pair sequence, advanced without divisions: every batch moves it by the same number of pairs.
struct PairCursor {
  int clip, pic;
  __device__ __forceinline__ void init(unsigned pair, unsigned ppc) {
    clip = (int)(pair / ppc);
    pic = (int)(pair - (unsigned)clip * ppc);
  }
  __device__ __forceinline__ void advance(int dclip, int dpic, int ppc) {
    clip += dclip;
    pic += dpic;
    while (pic >= ppc) {  // (one turn, unless a clip has fewer pairs than the step)
      pic -= ppc;
      ++clip;
    }
  }
};

// interior frames (all but the first / last of a clip): no masking.  v[bitrev5(n1)] = (a[32 n1 + lane], b[32 n1 + lane])
__device__ __forceinline__ void load_interior(const LogmelParams& p, const PairInfo& pi, int lane, float2 (&v)[32]) {
  const float* clip = p.audio + (long long)pi.clip * p.audio_stride;
  const float* __restrict__ pa = clip + pi.fa_lo + lane;
  const float* __restrict__ pb = clip + pi.fb_lo + lane;
#pragma unroll
  for (int n1 = 0; n1 < 32; ++n1) v[bitrev5(n1)] = make_float2(__ldg(pa + 32 * n1), __ldg(pb + 32 * n1));
}

// frames that touch a clip / window edge (~2 pairs per clip): masked or reflected samples, staged through the warp's
// staging buffer sixteen rows at a time as tile[(n1 - n1_begin) * 32 + lane] = (a, b); out of line and not unrolled to
// keep the hot loop small
__device__ __noinline__ void load_edge(const float* __restrict__ clip, int n_samples, int pad_mode, int fa_lo, int fb_lo,
                                       int lo_a, int hi_a, int lo_b, int hi_b, float2* tile, int n1_begin) {
  const int lane = threadIdx.x & 31;
  const int last = n_samples - 1;
#pragma unroll 1
  for (int n1 = n1_begin; n1 < n1_begin + 16; ++n1) {
    int sa = fa_lo + lane + 32 * n1, sb = fb_lo + lane + 32 * n1;
    bool oka, okb;
    if (pad_mode == 1) {  // numpy "reflect" padding about the first / last sample (MelSlidingWindowExtractor default)
      sa = sa < 0 ? -sa : (sa > last ? 2 * last - sa : sa);
      sb = sb < 0 ? -sb : (sb > last ? 2 * last - sb : sb);
      oka = sa >= 0 && sa <= last;
      okb = hi_b > lo_b && sb >= 0 && sb <= last;
    } else {
      oka = sa >= lo_a && sa < hi_a;
      okb = sb >= lo_b && sb < hi_b;
    }
    tile[(n1 - n1_begin) * 32 + lane] = make_float2(oka ? __ldg(clip + sa) : 0.0f, okb ? __ldg(clip + sb) : 0.0f);
  }
}

__device__ __forceinline__ float db_from_power(float p) {
  // 10 log10(max(p, amin)) = (10 log10 2) * log2(.): the argument is a normal number, MUFU.LG2 is within 2 ulp
  return 3.0102999566398120f * __log2f(fmaxf(p, kAmin));
}
// order-preserving map float -> int32, so that a warp maximum is one REDUX
__device__ __forceinline__ int float_order(float f) {
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// ---- filterbank phase, default bank: straight-line code from the compile-time tables --------------------------------
#include "melbank_default.inc"

// Bins [b0, b1) of one spectrum (lane = frame): every bin feeds the falling half of filter g and the rising half of filter
// g + 1 with immediate weights; two accumulator pairs alternate so that the FFMA chains are half as long.  Returns the
// rising half of the first filter after the run; `row` receives the complete sums of the filters whose falling half lies
// in the run (the first of them lacks the rising half that the previous run returns: see mel_fix_boundaries).
template <int W>
__device__ __forceinline__ float mel_run_default(const float* __restrict__ spec, float* __restrict__ row) {
  constexpr int b0 = kDefRunBinDev[W], b1 = kDefRunBinDev[W + 1];
  float lo0 = 0.0f, lo1 = 0.0f, hi0 = 0.0f, hi1 = 0.0f, carry = 0.0f;
#pragma unroll
  for (int k4 = (b0 & ~3); k4 < b1; k4 += 4) {
    const float4 x4 = *reinterpret_cast<const float4*>(spec + k4);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k4 + j;
      if (k >= b0 && k < b1) {
        if (k > b0 && kDefBinGroupDev[k] != kDefBinGroupDev[k - 1]) {  // next interval: retire filter g, carry g + 1's rise
          row[kDefBinGroupDev[k - 1]] = (lo0 + lo1) + carry;
          carry = hi0 + hi1;
          lo0 = lo1 = hi0 = hi1 = 0.0f;
        }
        const float wl = kDefBinWDev[2 * (k - kDefFirstBin)], wh = kDefBinWDev[2 * (k - kDefFirstBin) + 1];
        if (k & 1) {
          lo1 = fmaf(wl, xs[j], lo1);
          hi1 = fmaf(wh, xs[j], hi1);
        } else {
          lo0 = fmaf(wl, xs[j], lo0);
          hi0 = fmaf(wh, xs[j], hi0);
        }
      }
    }
  }
  row[kDefBinGroupDev[b1 - 1]] = (lo0 + lo1) + carry;
  return hi0 + hi1;
}

__device__ __forceinline__ float mel_phase_default(int consumer, const float* spec, float* row) {
  static_assert(kConsumers == 3, "melbank_default.inc is generated for three filterbank warps");
  switch (consumer) {
    case 0: return mel_run_default<0>(spec, row);
    case 1: return mel_run_default<1>(spec, row);
    default: return mel_run_default<2>(spec, row);
  }
}

// generic bank (any other sample rate / band edges): warp-uniform loops over the group table
__device__ __forceinline__ float mel_phase_generic(int g, int gend, const int4* s_groups, const float2* s_binw,
                                                   const float* spec, float* row) {
  float carry = 0.0f;
  for (; g < gend; ++g) {
    const int4 gi = s_groups[g];
    const float* x = spec + gi.x;
    const float2* w = s_binw + gi.y;
    float lo = carry, hi = 0.0f;
    for (int i = 0; i < gi.z; ++i) {
      const float xv = x[i];
      const float2 wv = w[i];
      lo = fmaf(wv.x, xv, lo);
      hi = fmaf(wv.y, xv, hi);
    }
    row[gi.w] = lo;
    carry = hi;
  }
  return carry;
}

// ---- mbarrier / TMA wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (error to the host) instead of hanging the GPU.  A waiting warp must not spin:
// the SM sub-partition arbitrates by warp id, so a spinning high-id warp would starve the very warps it waits for.
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_test(bar, parity)) return;
  for (uint32_t spin = 0; !mbar_test(bar, parity); ++spin) {
    __nanosleep(20);
    if (spin > (1u << 20)) __trap();
  }
}
// Monotonic progress counter in shared memory (release / acquire at CTA scope).  The producers take pairs by ticket and
// may skip whole batches, so they cannot follow the phase parity of an mbarrier (a wait two phases late would alias).
__device__ __forceinline__ void publish(unsigned* counter, unsigned value) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_addr(counter)), "r"(value) : "memory");
}
__device__ __forceinline__ void await_at_least(const unsigned* counter, unsigned value) {
  for (uint32_t spin = 0;; ++spin) {
    unsigned seen;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(smem_addr(counter)) : "memory");
    if (seen >= value) return;
    __nanosleep(20);
    if (spin > (1u << 20)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers * 32) : "memory"); }

constexpr int kSmemBarBytes = 256;  // full[2], -, stored[2], staged[kProducers], ticket counter

// One frame pair, producer side after the windowed samples are in v: transform_first = FFT, twiddle; transform_second =
// transpose, FFT, separation, which leaves 4 |A_k|^2 at slot[k] and 4 |B_k|^2 at slot[kSecondFrame + k], k = 0..512.
__device__ __forceinline__ void transform_first(float2 (&v)[32], const float2* s_tw, int lane) {
  fft32(v);  // v[k1] = Y[k1] of column n2 = lane
#pragma unroll
  for (int k1 = 1; k1 < 32; ++k1) {
    const float2 w = s_tw[k1 * 32 + lane];  // W_1024^(k1 * n2)
    const float2 z = v[k1];
    float2 r = __fmul2_rn(z, bcast(w.x));
    v[k1] = __ffma2_rn(make_float2(-z.y, z.x), bcast(w.y), r);
  }
}

__device__ __forceinline__ void transform_second(float2 (&v)[32], float* slot, int lane) {
  // 32x32 transpose through the padded slot, real plane then imaginary plane (the slot is sized for one plane so that two
  // slots and the staging buffer of 13 producers fit in shared memory)
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) slot[k1 * kRow + lane] = v[k1].x;
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) v[bitrev5(n2)].x = slot[lane * kRow + n2];  // (.y still holds the old imaginary parts)
  __syncwarp();
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) slot[k1 * kRow + lane] = v[k1].y;
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) v[bitrev5(n2)].y = slot[lane * kRow + n2];
  __syncwarp();
  fft32(v);  // v[k2] = Z[lane + 32 * k2]

  // separate the two real spectra: partner of k = lane + 32 r is 1024 - k = ((32-lane)&31) + 32 r'.
  // Lanes 1..31: partner register r' = 31 - r; lane 0: r' = 32 - r, which is the value it fetched (from itself) one step
  // earlier, and r = 0 is its own partner.  Four steps at a time, so that only 8 shuffle results are live at once.
  const int src = (32 - lane) & 31;
  float* pa = slot;
  float* pb = slot + kSecondFrame;
  float2 prev = v[0];
#pragma unroll
  for (int r0 = 0; r0 < 16; r0 += 4) {
    float2 q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q[j].x = __shfl_sync(kFullMask, v[31 - r0 - j].x, src);
      q[j].y = __shfl_sync(kFullMask, v[31 - r0 - j].y, src);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + j;
      const float2 z = v[r];
      const float2 c = lane == 0 ? prev : q[j];
      prev = q[j];
      // 2A = z + conj(c), 2iB = z - conj(c): (4|A|^2, 4|B|^2) = u*u + w*w, u = (z.x + c.x, z.x - c.x), w = (z.y - c.y, z.y + c.y)
      const float2 u = __fadd2_rn(bcast(z.x), make_float2(c.x, -c.x));
      const float2 w = __fadd2_rn(bcast(z.y), make_float2(-c.y, c.y));
      const float2 pw = __ffma2_rn(w, w, __fmul2_rn(u, u));
      pa[lane + 32 * r] = pw.x;
      pb[lane + 32 * r] = pw.y;
    }
    asm volatile("" ::: "memory");  // keep the compiler from hoisting the next group's shuffles (register pressure)
  }
  if (lane == 0) {  // Nyquist bin 512 = register 16, self-paired: A = z.x, B = z.y (x4 like the others)
    pa[512] = 4.0f * v[16].x * v[16].x;
    pb[512] = 4.0f * v[16].y * v[16].y;
  }
}

// dB rows + per-frame max of one frame pair, from the filterbank tile (rows 2 * position, 2 * position + 1)
__device__ __forceinline__ void store_pair(const LogmelParams& p, const PairInfo& pi, const float* rows, int lane) {
  float* dst = p.power + (long long)pi.clip * p.power_clip_stride + (long long)pi.frame * KOE_N_MELS;
  // the 160 values of the two rows (row B follows row A in the clip's block), five per lane: j = lane + 32 q;
  // j < 80 -> frame A filter j, else frame B filter j - 80, which sits kTileStride - 80 = 1 float further in the tile
  float db[5];
#pragma unroll
  for (int qd = 0; qd < 5; ++qd) {
    const int j = lane + 32 * qd;
    const bool second = qd > 2 || (qd == 2 && lane >= KOE_N_MELS - 64);
    db[qd] = db_from_power(rows[j + (second ? kTileStride - KOE_N_MELS : 0)]);  // stored in dB: the consumer of the
  }                                                                              // buffer only subtracts its reference
  float mx_a = fmaxf(db[0], db[1]), mx_b = fmaxf(db[3], db[4]);
  if (lane < KOE_N_MELS - 64) mx_a = fmaxf(mx_a, db[2]); else mx_b = fmaxf(mx_b, db[2]);
  dst[lane] = db[0];
  dst[lane + 32] = db[1];
  if (pi.has_b || lane < KOE_N_MELS - 64) dst[lane + 64] = db[2];
  if (pi.has_b) {
    dst[lane + 96] = db[3];
    dst[lane + 128] = db[4];
  }
  if (p.frame_max != nullptr) {
    const int ia = __reduce_max_sync(kFullMask, float_order(mx_a));
    const int ib = __reduce_max_sync(kFullMask, float_order(mx_b));
    float* fm = p.frame_max + (long long)pi.clip * p.fmax_clip_stride + pi.frame;
    if (lane == 0) fm[0] = order_float(ia);
    if (lane == 1 && pi.has_b) fm[1] = order_float(ib);
  }
}

template <bool kDefaultBank>
__global__ void __launch_bounds__(kThreads, 1)
logmel_power_kernel(FrontendTables tab, LogmelParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw);                    // full[2], -, stored[2], staged[kProducers]
  unsigned* s_ticket = reinterpret_cast<unsigned*>(smem_raw + kSmemBarBytes - 8);  // next pair to hand out
  unsigned* s_consumed = s_ticket + 1;  // batches whose spectra the consumers have turned into filterbank sums
  float* s_hann = reinterpret_cast<float*>(smem_raw + kSmemBarBytes);         // 1024
  float2* s_tw = reinterpret_cast<float2*>(s_hann + kFrameLen);               // 1024 float2
  float2* s_binw = s_tw + 1024;                                               // kMaxBins float2 (generic bank)
  int4* s_groups = reinterpret_cast<int4*>(s_binw + kMaxBins);                // kMaxGroups       (generic bank)
  int* s_runs = reinterpret_cast<int*>(s_groups + kMaxGroups);                // kConsumers + 1 (+ pad to 8)
  unsigned* s_pending = reinterpret_cast<unsigned*>(s_runs + 8);              // kProducers x 16: tickets whose rows are not stored yet
  float* s_bpart = reinterpret_cast<float*>(s_pending + kProducers * 16);                      // 2 x kConsumers x 32: rising half handed to the next run
  float* s_mel = s_bpart + 2 * kConsumers * 32;                                   // 2 x 26 x 81 filterbank sums of batch i in tile i & 1
  float* s_prod = s_mel + 2 * kBatchFrames * kTileStride;                     // kProducers x (slot 0 | slot 1 | staging), 16-byte aligned

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // batch i lives in slot i & 1 and signals on full[i & 1] / empty[i & 1]; those barriers are in phase i >> 1
  const uint32_t bar_full = smem_addr(s_bar), bar_stored = smem_addr(s_bar + 4);

  for (int i = tid; i < kFrameLen; i += kThreads) s_hann[i] = tab.hann[i];
  for (int i = tid; i < 1024; i += kThreads) s_tw[i] = tab.tw[i];
  if (!kDefaultBank) {
    for (int i = tid; i < tab.n_bins; i += kThreads) s_binw[i] = tab.binw[i];
    for (int i = tid; i < tab.n_groups; i += kThreads) s_groups[i] = tab.groups[i];
    if (tid <= kConsumers) s_runs[tid] = tab.runs[tid];
  }
  if (tid == 0) {
    *s_ticket = 0;
    *s_consumed = 0;
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, kProducers);
      mbar_init(bar_stored + 8 * s, kProducers);
    }
    for (int w = 0; w < kProducers; ++w) mbar_init(smem_addr(s_bar + 6 + w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const unsigned ppc = (unsigned)(p.n_frames + 1) >> 1;  // frame pairs per clip
  const unsigned total_pairs = (unsigned)p.n_clips * ppc;
  const unsigned n_batches = (total_pairs + kProducers - 1) / kProducers;
  // every warp of the CTA runs the same number of batches: blockIdx.x, blockIdx.x + gridDim.x, ...
  const unsigned my_batches = blockIdx.x < n_batches ? (n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp < kProducers) {
    // ================================================================== producer: FFT of one frame pair per batch
    // Pairs are handed out by ticket (ticket t = position t % 13 of this CTA's batch t / 13), not per warp: the SM
    // sub-partitions arbitrate by warp id, so some producers run persistently faster than others, and with a fixed
    // assignment every batch would wait for the slowest one.  A warp owns its staging buffer; the two spectrum slots
    // belong to the POSITION, so consecutive batches of one position may be written by different warps.
    float* stage = s_prod + warp * kProdStride + 2 * kSlotFloats;
    const uint32_t bar_staged = smem_addr(s_bar + 6 + warp), stage_addr = smem_addr(stage);
    unsigned staged_phase = 0;
    const unsigned n_tickets = my_batches * kProducers;
    auto take_ticket = [&]() -> unsigned {
      unsigned t = 0;
      if (lane == 0) t = atomicAdd(s_ticket, 1u);
      return __shfl_sync(kFullMask, t, 0);
    };
    auto describe_ticket = [&](unsigned t) -> PairInfo {
      if (t >= n_tickets) return describe_pair(p, p.n_clips, 0);
      const unsigned b = t / kProducers, j = t - b * kProducers;
      PairCursor pos;
      pos.init((blockIdx.x + b * gridDim.x) * kProducers + j, ppc);
      return describe_pair(p, pos.clip, pos.pic);
    };
    auto request = [&](const PairInfo& pi) {
      if (pi.tma && lane == 0) {
        mbar_expect_tx(bar_staged, pi.tma_bytes);
        bulk_g2s(stage_addr, pi.tma_src, pi.tma_bytes, bar_staged);
      }
    };
    // Nothing but the next ticket stays live across the transform (it needs every register, and with ~222 KB of the
    // SM's 228 KB configured as shared memory a spill would miss L1 every time): the pair is described again when needed.
    // The dB rows of a pair are written by the warp that transformed it, normally one pair later (by then the consumers
    // have turned its spectra into filterbank sums): flush_rows waits until that batch is consumed, stores the two rows
    // and arrives on `stored`, which the consumers await before they overwrite the tile two batches later.  A batch is
    // consumed only after ALL its pairs have arrived, so a warp may flush a pair only when every ticket it still holds
    // belongs to a later batch; until then the ticket waits in a small per-warp queue.
    unsigned* pending = s_pending + warp * 16;
    unsigned q_head = 0, q_tail = 0;
    auto flush_rows = [&](unsigned t) {
      const unsigned b = t / kProducers, j = t - b * kProducers;
      const PairInfo pi = describe_ticket(t);
      await_at_least(s_consumed, b + 1);  // (also without a pair: the arrival below must land in this batch's phase)
      if (pi.frame >= 0) store_pair(p, pi, s_mel + ((b & 1) * kBatchFrames + 2 * j) * kTileStride, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_stored + 8 * (b & 1));
    };
    auto flush_older_than = [&](unsigned batch_limit) {  // every queued ticket of a batch < batch_limit
      while (q_head != q_tail) {
        const unsigned t = pending[q_head & 15];
        if (t / kProducers >= batch_limit) break;
        flush_rows(t);
        ++q_head;
      }
    };
    unsigned ticket = take_ticket();
    request(describe_ticket(ticket));
    float2 v[32];
    while (ticket < n_tickets) {
      const unsigned batch = ticket / kProducers, position = ticket - batch * kProducers;
      float* slots = s_prod + position * kProdStride;
      float* slot = slots + (batch & 1) * kSlotFloats;
      const PairInfo cur = describe_ticket(ticket);
      const bool have_pair = cur.frame >= 0;
      if (have_pair) {
        if (cur.tma) {
          mbar_wait(bar_staged, staged_phase & 1);
          ++staged_phase;
          const float* sa = stage + cur.stage_off + lane;
          const float* sb = sa + (cur.fb_lo - cur.fa_lo);
#pragma unroll
          for (int n1 = 0; n1 < 32; ++n1) v[bitrev5(n1)] = make_float2(sa[32 * n1], sb[32 * n1]);
        } else if (cur.interior) {
          load_interior(p, cur, lane, v);
        } else {
          // edge frames: masked loads, staged through this warp's (idle: no bulk copy was requested) staging buffer
          float2* st2 = reinterpret_cast<float2*>(stage);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            load_edge(p.audio + (long long)cur.clip * p.audio_stride, p.n_samples, p.pad_mode, cur.fa_lo, cur.fb_lo,
                      cur.lo_a, cur.hi_a, cur.lo_b, cur.hi_b, st2, 16 * half);
            __syncwarp();
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) v[bitrev5(16 * half + n1)] = st2[n1 * 32 + lane];
            __syncwarp();
          }
        }
        // periodic Hann window; every staged sample has now reached a register (the products depend on the loads) ...
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) v[bitrev5(n1)] = __fmul2_rn(v[bitrev5(n1)], bcast(s_hann[32 * n1 + lane]));
      }
      __syncwarp();
      // ... in every lane, so the copy engine may refill the staging buffer with the next pair
      const unsigned next_ticket = take_ticket();
      request(describe_ticket(next_ticket));
      if (have_pair) transform_first(v, s_tw, lane);
      flush_older_than(batch);  // rows of earlier pairs (normally the previous one), while the FMA pipe drains
      // the slot still holds the spectra of batch - 2 until the consumers have read them (and an arrival must not run
      // ahead into that batch's phase of the `full` barrier either)
      if (batch >= 2) await_at_least(s_consumed, batch - 1);
      if (have_pair) transform_second(v, slot, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * (batch & 1));
      if (lane == 0) pending[q_tail & 15] = ticket;
      ++q_tail;
      __syncwarp();
      ticket = next_ticket;
    }
    flush_older_than(0xffffffffu);
  } else {
    // ================================================================== consumer: filterbank, dB, store
    const int cons = warp - kProducers;
    const float* spec0 = s_prod + (lane >> 1) * kProdStride + (lane & 1) * kSecondFrame;
    for (unsigned it = 0; it < my_batches; ++it) {
      mbar_wait(bar_full + 8 * (it & 1), (it >> 1) & 1);
      if (it >= 2) mbar_wait(bar_stored + 8 * (it & 1), ((it >> 1) - 1) & 1);  // the rows of batch it - 2 have left the tile
      float* row = s_mel + ((it & 1) * kBatchFrames + lane) * kTileStride;
      if (lane < kBatchFrames) {
        const float* spec = spec0 + (it & 1) * kSlotFloats;
        const float rise = kDefaultBank ? mel_phase_default(cons, spec, row)
                                        : mel_phase_generic(s_runs[cons], s_runs[cons + 1], s_groups, s_binw, spec, row);
        s_bpart[((it & 1) * kConsumers + cons) * 32 + lane] = rise;
      }
      consumers_sync();  // the rising halves are visible
      if (lane < kBatchFrames && cons > 0) {
        const int m = kDefaultBank ? (cons == 1 ? kDefBinGroupDev[kDefRunBinDev[1]] : kDefBinGroupDev[kDefRunBinDev[2]])
                                   : s_groups[s_runs[cons]].w;
        row[m] += s_bpart[((it & 1) * kConsumers + cons - 1) * 32 + lane];  // first filter of the run: rising half from the previous warp
      }
      consumers_sync();
      // spectra consumed (the producers may overwrite the slot) and filterbank sums in place (they may store the rows)
      if (warp == kProducers && lane == 0) publish(s_consumed, it + 1);
    }
  }
}

constexpr size_t kLogmelSmem = kSmemBarBytes + sizeof(float) * kFrameLen + sizeof(float2) * 1024 + sizeof(float2) * kMaxBins +
                               sizeof(int4) * kMaxGroups + sizeof(int) * (8 + kProducers * 16) +
                               sizeof(float) * (2 * kConsumers * 32 + 2 * kBatchFrames * kTileStride + kProducers * kProdStride);
static_assert(kLogmelSmem <= 232448, "shared memory budget of one SM");

// ---- dB normalisation: ref = clip max, clamp, rescale; emits long-term and last-3 short-term features
__global__ void logmel_normalise_kernel(const float* __restrict__ power, const float* __restrict__ frame_max,
                                        int n_frames, int db_only, float* __restrict__ long_term,
                                        float* __restrict__ short_term) {
  __shared__ float s_red[32];
  __shared__ float s_ref_db;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* fm = frame_max + (long long)b * n_frames;
  float mx = -INFINITY;
  for (int g = tid; g < n_frames; g += blockDim.x) mx = fmaxf(mx, fm[g]);
  mx = warp_max(mx);
  if ((tid & 31) == 0) s_red[tid >> 5] = mx;
  __syncthreads();
  if (tid < 32) {
    float v = tid < (blockDim.x >> 5) ? s_red[tid] : -INFINITY;
    v = warp_max(v);
    if (tid == 0) s_ref_db = v;
  }
  __syncthreads();
  const float ref_db = s_ref_db;
  const float4* src = reinterpret_cast<const float4*>(power + (long long)b * n_frames * KOE_N_MELS);
  float4* dst = reinterpret_cast<float4*>(long_term + (long long)b * n_frames * KOE_N_MELS);
  const int n4 = n_frames * (KOE_N_MELS / 4);
  const bool rescale = db_only == 0;
  for (int i = tid; i < n4; i += blockDim.x) {
    float4 v = src[i];
    v.x = normalise_db(v.x, ref_db, rescale);
    v.y = normalise_db(v.y, ref_db, rescale);
    v.z = normalise_db(v.z, ref_db, rescale);
    v.w = normalise_db(v.w, ref_db, rescale);
    dst[i] = v;
  }
  if (short_term != nullptr) {
    // last three frames; clips shorter than 3 frames: rows [0, n_frames) then zeros (reference :206-212)
    float* st = short_term + (long long)b * 3 * KOE_N_MELS;
    for (int i = tid; i < 3 * KOE_N_MELS; i += blockDim.x) {
      const int row = i / KOE_N_MELS, m = i % KOE_N_MELS;
      const int g = n_frames >= 3 ? n_frames - 3 + row : row;
      float v = 0.0f;
      if (g < n_frames)
        v = normalise_db(power[((long long)b * n_frames + g) * KOE_N_MELS + m], ref_db, rescale);
      st[i] = v;
    }
  }
}

// ---- host side: Slaney filterbank (librosa.filters.mel restated, float64 then float32) -----------
static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static std::vector<float> slaney_filterbank(int sr, int n_fft, int n_mels, double fmin, double fmax) {
  const int n_bins = 1 + n_fft / 2;
  std::vector<double> mel_f(n_mels + 2);
  const double m0 = hz_to_mel(fmin), m1 = hz_to_mel(fmax);
  for (int i = 0; i < n_mels + 2; ++i) {
    // numpy.linspace: start + i * step, last point pinned to stop
    const double step = (m1 - m0) / (n_mels + 1);
    mel_f[i] = mel_to_hz(i == n_mels + 1 ? m1 : m0 + i * step);
  }
  std::vector<float> fb((size_t)n_mels * n_bins, 0.0f);
  for (int m = 0; m < n_mels; ++m) {
    const double d0 = mel_f[m + 1] - mel_f[m], d1 = mel_f[m + 2] - mel_f[m + 1];
    const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
    for (int k = 0; k < n_bins; ++k) {
      const double f = (double)k * sr / n_fft;
      const double lower = -(mel_f[m] - f) / d0, upper = (mel_f[m + 2] - f) / d1;
      const double w = std::fmax(0.0, std::fmin(lower, upper));
      // librosa stores float32 weights, then multiplies the float32 array by the float64 norm
      const float w32 = (float)w;
      fb[(size_t)m * n_bins + k] = (float)((double)w32 * enorm);
    }
  }
  return fb;
}

}  // namespace koe

using namespace koe;

struct koe_frontend {
  int device = 0, sample_rate = 0, n_fft = 0, n_mels = 0;
  float fmin = 0, fmax = 0;
  float* d_hann = nullptr;
  float2* d_tw = nullptr;
  float2* d_binw = nullptr;
  int* d_tables = nullptr;  // groups[kMaxGroups] (int4) | runs[kConsumers + 1 -> 32]
  int n_bins = 0, n_groups = 0;
  bool default_bank = false;  // structure == melbank_default.inc: the unrolled filterbank phase applies
  int num_sms = 0, occupancy = 0;
  std::vector<float> fb_host;
};

extern "C" int koe_frontend_create(int device, int sample_rate, int n_fft, int n_mels, float fmin, float fmax,
                                   koe_frontend_t** out) {
  KOE_REQUIRE(out != nullptr, "koe_frontend_create: out is NULL");
  if (n_fft != KOE_N_FFT || n_mels != KOE_N_MELS)
    return fail(KOE_E_UNSUPPORTED, "koe_frontend_create: only n_fft=1024, n_mels=80 are implemented (got %d, %d)",
                n_fft, n_mels);
  KOE_REQUIRE(sample_rate > 0 && fmin >= 0 && fmax > fmin && fmax <= sample_rate / 2.0f,
              "koe_frontend_create: bad sample_rate/fmin/fmax");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return fail(KOE_E_NODEVICE, "koe_frontend_create: no CUDA device (this library has no CPU path)");
  KOE_REQUIRE(device >= 0 && device < n_dev, "koe_frontend_create: device %d out of range", device);
  int prev = 0;
  KOE_CUDA(cudaGetDevice(&prev));
  KOE_CUDA(cudaSetDevice(device));

  auto* fe = new koe_frontend();
  fe->device = device;
  fe->sample_rate = sample_rate;
  fe->n_fft = n_fft;
  fe->n_mels = n_mels;
  fe->fmin = fmin;
  fe->fmax = fmax;
  fe->fb_host = slaney_filterbank(sample_rate, n_fft, n_mels, fmin, fmax);

  // bin-major sparse filterbank: every weighted bin feeds one filter or two adjacent ones (fl, fl + 1); consecutive bins
  // with the same fl form a group (the interval between two filter centres), and fl rises by one from group to group
  std::vector<float2> binw;
  std::vector<int> tables(4 * kMaxGroups + 32, 0);
  std::vector<unsigned char> bin_group(kBins, 255);
  int n_groups = 0;
  auto unsupported = [&](const char* why, int k) {
    delete fe;
    cudaSetDevice(prev);
    return fail(KOE_E_UNSUPPORTED, "koe_frontend_create: filterbank is not a bank of adjacent triangles (%s at bin %d)", why, k);
  };
  {
    int prev_k = -1, prev_fl = -1;
    for (int k = 0; k < kBins; ++k) {
      int first = -1, count = 0, last = -1;
      for (int m = 0; m < n_mels; ++m)
        if (fe->fb_host[(size_t)m * kBins + k] > 1e-12f) {  // (a band edge that falls on a bin leaves ~1e-17 there)
          if (first < 0) first = m;
          last = m;
          ++count;
        }
      if (count == 0) continue;
      if (count > 2 || last - first > 1) return unsupported("more than two / non-adjacent filters", k);
      if ((int)binw.size() >= kMaxBins) return unsupported("too many weighted bins", k);
      if (prev_k >= 0 && k != prev_k + 1) return unsupported("gap in the weighted bins", k);
      if (first != prev_fl) {
        if (first != prev_fl + 1) return unsupported("filters skipped", k);
        if (n_groups >= kMaxGroups) return unsupported("too many groups", k);
        tables[4 * n_groups + 0] = k;
        tables[4 * n_groups + 1] = (int)binw.size();
        tables[4 * n_groups + 2] = 0;
        tables[4 * n_groups + 3] = first;
        ++n_groups;
      }
      ++tables[4 * (n_groups - 1) + 2];
      bin_group[k] = (unsigned char)first;
      // 0.25: the kernel leaves 4 |X|^2 in the spectrum (exact power-of-two scaling)
      binw.push_back(make_float2(0.25f * fe->fb_host[(size_t)first * kBins + k],
                                 count == 2 ? 0.25f * fe->fb_host[(size_t)last * kBins + k] : 0.0f));
      prev_k = k;
      prev_fl = first;
    }
    if (n_groups != n_mels) return unsupported("a filter without a falling half", kBins);
  }
  fe->n_bins = (int)binw.size();
  fe->n_groups = n_groups;
  // runs: contiguous group ranges, one per consumer warp, balanced on bins + a per-group overhead (generic kernel only)
  {
    int* runs = tables.data() + 4 * kMaxGroups;
    auto cost = [&](int g) { return 3 + tables[4 * g + 2]; };
    long long total = 0;
    for (int g = 0; g < n_groups; ++g) total += cost(g);
    long long acc = 0;
    int r = 1;
    runs[0] = 0;
    for (int g = 0; g < n_groups && r < kConsumers; ++g) {
      acc += cost(g);
      if (acc * kConsumers >= total * r) runs[r++] = g + 1;
    }
    for (; r <= kConsumers; ++r) runs[r] = n_groups;
  }
  // the unrolled filterbank phase applies when the bank has the structure of melbank_default.inc and its weights
  // (computed above from the run-time arguments) equal the baked-in ones to within one float32 ulp
  fe->default_bank = fe->n_bins == kDefNumBins && tables[0] == kDefFirstBin &&
                     std::equal(bin_group.begin(), bin_group.end(), kDefBinGroup);
  for (int i = 0; fe->default_bank && i < kDefNumBins; ++i) {
    const float got[2] = {binw[i].x, binw[i].y}, want[2] = {kDefBinW[2 * i], kDefBinW[2 * i + 1]};
    for (int h = 0; h < 2; ++h)
      if (std::fabs(got[h] - want[h]) > 1.2e-7f * std::fabs(want[h])) fe->default_bank = false;
  }
  std::vector<float> hann(kFrameLen);
  for (int n = 0; n < kFrameLen; ++n) hann[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * n / kFrameLen));
  std::vector<float2> tw(1024);
  for (int k1 = 0; k1 < 32; ++k1)
    for (int n2 = 0; n2 < 32; ++n2) {
      const double a = -2.0 * M_PI * (double)(k1 * n2) / 1024.0;
      tw[k1 * 32 + n2] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
  cudaError_t e = cudaSuccess;
  auto up = [&](void** dptr, const void* src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(dptr, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice);
  };
  up((void**)&fe->d_hann, hann.data(), hann.size() * sizeof(float));
  up((void**)&fe->d_tw, tw.data(), tw.size() * sizeof(float2));
  binw.resize(kMaxBins, make_float2(0.f, 0.f));
  up((void**)&fe->d_binw, binw.data(), binw.size() * sizeof(float2));
  up((void**)&fe->d_tables, tables.data(), tables.size() * sizeof(int));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(logmel_power_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLogmelSmem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(logmel_power_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLogmelSmem);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess) {
    fe->num_sms = prop.multiProcessorCount;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fe->occupancy, logmel_power_kernel<true>, kThreads, kLogmelSmem);
  }
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    koe_frontend_destroy(fe);
    return fail((int)e, "koe_frontend_create: %s", cudaGetErrorString(e));
  }
  if (fe->occupancy < 1) fe->occupancy = 1;
  *out = fe;
  return KOE_OK;
}

extern "C" int koe_frontend_destroy(koe_frontend_t* fe) {
  if (fe == nullptr) return KOE_OK;
  cudaFree(fe->d_hann);
  cudaFree(fe->d_tw);
  cudaFree(fe->d_binw);
  cudaFree(fe->d_tables);
  delete fe;
  return KOE_OK;
}

extern "C" int koe_frontend_uses_unrolled_bank(const koe_frontend_t* fe) { return fe != nullptr && fe->default_bank ? 1 : 0; }

extern "C" int koe_frontend_filterbank_host(const koe_frontend_t* fe, float* fb_host) {
  KOE_REQUIRE(fe != nullptr && fb_host != nullptr, "koe_frontend_filterbank_host: NULL argument");
  std::copy(fe->fb_host.begin(), fe->fb_host.end(), fb_host);
  return KOE_OK;
}

extern "C" int koe_logmel_power_ex(const koe_frontend_t* fe, const koe_logmel_args* a, void* stream) {
  KOE_REQUIRE(fe != nullptr && a != nullptr && a->audio != nullptr && a->power != nullptr,
              "koe_logmel_power: NULL argument");
  KOE_REQUIRE(a->n_clips >= 0 && a->n_samples >= 0 && a->n_frames >= 0, "koe_logmel_power: negative size");
  KOE_REQUIRE(a->hop > 0 && a->audio_stride >= a->n_samples, "koe_logmel_power: bad hop/stride");
  KOE_REQUIRE(a->frame_offset >= 0 && a->frame_step >= 1 && a->sample_offset >= 0,
              "koe_logmel_power: bad frame_offset/frame_step/sample_offset");
  KOE_REQUIRE((long long)a->sample_offset +
                      ((long long)a->frame_offset + (long long)(a->n_frames + 1) * a->frame_step + KOE_MAX_EDGE + 1) *
                          a->hop < (1ll << 31) && a->n_samples < (1 << 30),
              "koe_logmel_power: clip too long for 32-bit sample indices");
  KOE_REQUIRE(a->pad_mode == 0 || a->pad_mode == 1, "koe_logmel_power: pad_mode must be 0 (constant) or 1 (reflect)");
  KOE_REQUIRE(a->pad_mode == 0 || (a->lo_rel_hops == KOE_NO_EDGE && a->hi_rel_hops == KOE_NO_EDGE &&
                                   a->n_samples > KOE_N_FFT / 2),
              "koe_logmel_power: reflect padding needs n_samples > n_fft/2 and no window edges");
  KOE_REQUIRE(a->power_clip_stride >= (int64_t)a->n_frames * KOE_N_MELS && a->power_clip_stride % 4 == 0,
              "koe_logmel_power: bad power_clip_stride");
  KOE_REQUIRE(a->frame_max == nullptr || a->frame_max_clip_stride >= a->n_frames,
              "koe_logmel_power: bad frame_max_clip_stride");
  if (a->n_clips == 0 || a->n_frames == 0) return KOE_OK;
  FrontendTables tab;
  tab.hann = fe->d_hann;
  tab.tw = fe->d_tw;
  tab.binw = fe->d_binw;
  tab.groups = reinterpret_cast<const int4*>(fe->d_tables);
  tab.runs = fe->d_tables + 4 * kMaxGroups;
  tab.n_bins = fe->n_bins;
  tab.n_groups = fe->n_groups;
  LogmelParams p;
  p.audio = a->audio;
  p.audio_stride = a->audio_stride;
  p.n_clips = a->n_clips;
  p.n_samples = a->n_samples;
  p.hop = a->hop;
  p.n_frames = a->n_frames;
  p.lo_rel = a->lo_rel_hops;
  p.hi_rel = a->hi_rel_hops;
  p.frame_offset = a->frame_offset;
  p.frame_step = a->frame_step;
  p.sample_offset = a->sample_offset;
  p.pad_mode = a->pad_mode;
  p.power = a->power;
  p.frame_max = a->frame_max;
  p.power_clip_stride = a->power_clip_stride;
  p.fmax_clip_stride = a->frame_max_clip_stride;
  const long long ppc = (a->n_frames + 1) / 2;
  KOE_REQUIRE((long long)a->n_clips * ppc < (1ll << 31) - kProducers * 65536ll,
              "koe_logmel_power: more than 2^31 frame pairs in one call");
  const long long n_blocks = ((long long)a->n_clips * ppc + kProducers - 1) / kProducers;
  const long long max_grid = (long long)fe->num_sms * fe->occupancy;
  const int grid = (int)std::min(n_blocks, max_grid);
  if (fe->default_bank)
    logmel_power_kernel<true><<<grid, kThreads, kLogmelSmem, (cudaStream_t)stream>>>(tab, p);
  else
    logmel_power_kernel<false><<<grid, kThreads, kLogmelSmem, (cudaStream_t)stream>>>(tab, p);
  count_launch();
  KOE_CUDA(cudaGetLastError());
  return KOE_OK;
}

extern "C" int koe_logmel_power(const koe_frontend_t* fe, const float* audio, int64_t audio_stride, int n_clips,
                                int n_samples, int hop, int n_frames, int frame_offset, int frame_step,
                                int lo_rel_hops, int hi_rel_hops, float* power, float* frame_max, void* stream) {
  koe_logmel_args a;
  a.audio = audio;
  a.audio_stride = audio_stride;
  a.n_clips = n_clips;
  a.n_samples = n_samples;
  a.hop = hop;
  a.n_frames = n_frames;
  a.frame_offset = frame_offset;
  a.frame_step = frame_step;
  a.sample_offset = 0;
  a.lo_rel_hops = lo_rel_hops;
  a.hi_rel_hops = hi_rel_hops;
  a.pad_mode = 0;
  a.power = power;
  a.power_clip_stride = (int64_t)n_frames * KOE_N_MELS;
  a.frame_max = frame_max;
  a.frame_max_clip_stride = n_frames;
  return koe_logmel_power_ex(fe, &a, stream);
}

extern "C" int koe_logmel_normalise(const float* power, const float* frame_max, int n_clips, int n_frames,
                                    int db_only, float* long_term, float* short_term, void* stream) {
  KOE_REQUIRE(power != nullptr && frame_max != nullptr && long_term != nullptr,
              "koe_logmel_normalise: NULL argument");
  KOE_REQUIRE(n_clips >= 0 && n_frames >= 0, "koe_logmel_normalise: negative size");
  KOE_REQUIRE(((reinterpret_cast<uintptr_t>(power) | reinterpret_cast<uintptr_t>(long_term)) & 15) == 0,
              "koe_logmel_normalise: buffers must be 16-byte aligned");
  if (n_clips == 0) return KOE_OK;
  logmel_normalise_kernel<<<n_clips, 256, 0, (cudaStream_t)stream>>>(power, frame_max, n_frames, db_only, long_term,
                                                                      short_term);
  count_launch();
  KOE_CUDA(cudaGetLastError());
  return KOE_OK;
}
