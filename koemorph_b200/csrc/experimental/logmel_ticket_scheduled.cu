// EXPERIMENT -- NOT BUILT, NOT SHIPPED.  Barrier-free variant of logmel.cu kept for the record (round 1): 16 identical
// warps; frame pairs, filterbank tasks and store tasks handed out by tickets; dependencies as release / acquire counters
// with "helping" waits.  Parity: identical results to logmel.cu (scripts/k1_check.py).
// Measured on B200, 512 clips x 257 frames: 268 us against 222-226 us for the lock-step kernel in ../logmel.cu.
// Why it lost (ncu source counters, profiles/ncu_k1_r01_experiment_ticket_scheduled_summary.txt): the kernel is ~5400
// SASS instructions (87 KB) of straight-line code; once the warps drift apart each one streams its own instructions and
// "no instruction" becomes the second largest stall (19 %), where the lock-step kernel keeps all 16 warps within the same
// few cache lines of the unrolled FFT.  It needs melbank_default_16runs.inc (WARPS = 16 and baked weights).
//
// Log-mel frontend for sm_100a: framing + periodic Hann + 1024-point FFT + |X|^2 + sparse Slaney
// filterbank, fused in one kernel; then a small dB-normalisation kernel.
//
// Replaces the per-clip librosa loop of SimplifiedDualStreamModel.extract_mel_features
// (reference src/model/simplified_dual_stream_model.py:184-229).
//
// Kernel design (see DESIGN.md section "K1"): one persistent CTA of 16 identical warps per SM (128 registers per thread
// allow no more), no CTA-wide barrier in the main loop.
//   * FFT: a warp transforms TWO real frames of one clip at once as one complex 1024-point FFT (z = a + i b), decomposed
//     32 x 32: radix-2 DIT FFT of 32 points in registers, twiddle, 32x32 transpose through a padded shared-memory slot
//     (real plane, then imaginary plane), second 32-point FFT in registers.  Every complex value is one 64-bit register
//     pair and every butterfly is written with the packed fp32x2 instructions of sm_100 (FADD2 / FMUL2 / FFMA2, whose
//     operands take per-half negation and a half swap, so "times -i" is free and a butterfly with a general twiddle is 3
//     instructions: x = a + w b as two FFMA2, y = 2a - x as one).  After the second pass lane l holds Z[l + 32 r]; the
//     conjugate-symmetric partner Z[1024 - k] lives in lane (32 - l) & 31, so the two real spectra are separated with 32
//     warp shuffles; the pair (|A_k|^2, |B_k|^2) comes out of one FMUL2 + one FFMA2 and is parked in the slot.
//   * Filterbank: lane = frame.  Sixteen frame pairs form a batch (32 frames = 32 lanes); the Slaney filterbank of a batch
//     is cut into 16 tasks (runs of bins), each straight-line code unrolled from the bank's compile-time tables
//     (melbank_default.inc: immediate weights, conflict-free LDS.128 of four bins, warp-uniform control flow).
//   * Rows: 16 store tasks per batch convert two rows of 80 sums to dB and write them, with the per-frame maximum.
//   * Scheduling: frame pairs, filterbank tasks and store tasks are handed out by tickets (shared-memory counters) to
//     whichever warp is free -- the SM sub-partitions arbitrate by warp id, so warps run at persistently different
//     speeds and any fixed assignment would wait for the slowest.  Dependencies (batch complete -> filterbank; filterbank
//     complete -> rows, slot reuse two batches later; rows stored -> tile reuse) are monotonic counters with
//     release / acquire semantics, and every wait "helps": it executes pending filterbank / store tasks instead of idling,
//     which also makes the scheme deadlock-free.  The audio of a warp's next pair is loaded into registers before it
//     turns to filterbank / store tasks, so DRAM latency hides behind them.
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace koe {

constexpr int kFrameLen = 1024;
constexpr int kBins = 513;
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kBatchPairs = 16;         // frame pairs per batch: 32 frames = one lane each in the filterbank tasks
constexpr int kRow = 33;                // floats per row of the transposition slot (padded: conflict-free both ways)
constexpr int kSlotFloats = 32 * kRow;  // 1056 floats: one plane of the 32x32 transpose, then the two power spectra
constexpr int kSecondFrame = 516;       // offset of the second spectrum, == 4 (mod 32)
constexpr int kPosStride = 2 * kSlotFloats + 8;  // two slots (batch parity) per batch position; == 8 (mod 32) floats: with
                                        // lane = frame the bases of 8 consecutive frames are 16 bytes apart mod 128
                                        // -> the filterbank's LDS.128 are conflict-free
constexpr int kMaxBins = 512;           // spectrum bins that carry filterbank weight (506 for 80..8000 Hz)
constexpr int kMaxGroups = 96;          // groups of consecutive bins feeding the same pair of adjacent filters
constexpr int kTileStride = 81;         // filterbank tile row stride (floats), odd: conflict-free across frames
constexpr int kTileFloats = 2 * kBatchPairs * kTileStride;  // one tile: 32 rows
static_assert(kPosStride % 32 == 8 && kSecondFrame % 32 == 4 && kSecondFrame + kBins <= kSlotFloats, "bank skew / slot size");

struct FrontendTables {
  const float* hann;     // [1024]
  const float2* tw;      // [32][32] W_1024^(k1*n2)
  // Slaney filterbank, bin-major: a spectrum bin feeds at most two ADJACENT filters (fl, fl + 1)
  const float2* binw;    // [n_bins] 0.25 * (weight into filter fl, weight into filter fl + 1)
  const int4* groups;    // [n_groups] {first bin k, first entry of binw, number of bins, fl}: fl rises by one per group
  const int* runs;       // [kBatchPairs + 1] group range of every filterbank task (generic bank)
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
};

// pair number `pic` of clip `b` (frames 2 pic, 2 pic + 1); b >= n_clips: no pair
__device__ __forceinline__ PairInfo describe_pair(const LogmelParams& p, int b, int pic) {
  PairInfo pi;
  pi.clip = 0;
  pi.frame = -1;
  pi.interior = false;
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

// frames that touch a clip / window edge (~2 pairs per clip): masked or reflected samples, staged through the pair's
// spectrum slot sixteen rows at a time as tile[(n1 - n1_begin) * 32 + lane] = (a, b); out of line and not unrolled to
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

// ---- filterbank tasks, default bank: straight-line code from the compile-time tables -------------------------------
#include "melbank_default.inc"

// Bins [b0, b1) of run W for one spectrum (lane = frame): every bin feeds the falling half of filter g (-> tlo[g]) and
// the rising half of filter g + 1 (-> thi[g + 1]) with immediate weights; two accumulator pairs alternate so that the
// FFMA chains are half as long.  Runs are cut between groups, so every (tile, filter) cell has exactly one writer.
template <int W>
__device__ __forceinline__ void mel_run_default(const float* __restrict__ spec, float* __restrict__ tlo,
                                                float* __restrict__ thi) {
  constexpr int b0 = kDefRunBinDev[W], b1 = kDefRunBinDev[W + 1];
  if (b0 >= b1) return;
  float lo0 = 0.0f, lo1 = 0.0f, hi0 = 0.0f, hi1 = 0.0f;
#pragma unroll
  for (int k4 = (b0 & ~3); k4 < b1; k4 += 4) {
    const float4 x4 = *reinterpret_cast<const float4*>(spec + k4);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k4 + j;
      if (k >= b0 && k < b1) {
        if (k > b0 && kDefBinGroupDev[k] != kDefBinGroupDev[k - 1]) {  // next interval: retire the two halves
          tlo[kDefBinGroupDev[k - 1]] = lo0 + lo1;
          thi[kDefBinGroupDev[k - 1] + 1] = hi0 + hi1;
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
  tlo[kDefBinGroupDev[b1 - 1]] = lo0 + lo1;
  thi[kDefBinGroupDev[b1 - 1] + 1] = hi0 + hi1;  // the rising half of "filter 80" lands in the spare column of the tile
}

__device__ __noinline__ void mel_task_default(int run, const float* spec, float* tlo, float* thi) {
  static_assert(kBatchPairs == 16, "melbank_default.inc is generated for sixteen filterbank tasks");
  switch (run) {
    case 0: mel_run_default<0>(spec, tlo, thi); break;
    case 1: mel_run_default<1>(spec, tlo, thi); break;
    case 2: mel_run_default<2>(spec, tlo, thi); break;
    case 3: mel_run_default<3>(spec, tlo, thi); break;
    case 4: mel_run_default<4>(spec, tlo, thi); break;
    case 5: mel_run_default<5>(spec, tlo, thi); break;
    case 6: mel_run_default<6>(spec, tlo, thi); break;
    case 7: mel_run_default<7>(spec, tlo, thi); break;
    case 8: mel_run_default<8>(spec, tlo, thi); break;
    case 9: mel_run_default<9>(spec, tlo, thi); break;
    case 10: mel_run_default<10>(spec, tlo, thi); break;
    case 11: mel_run_default<11>(spec, tlo, thi); break;
    case 12: mel_run_default<12>(spec, tlo, thi); break;
    case 13: mel_run_default<13>(spec, tlo, thi); break;
    case 14: mel_run_default<14>(spec, tlo, thi); break;
    default: mel_run_default<15>(spec, tlo, thi); break;
  }
}

// generic bank (any other sample rate / band edges): warp-uniform loops over the group table
__device__ __noinline__ void mel_task_generic(int g, int gend, const int4* s_groups, const float2* s_binw,
                                              const float* spec, float* tlo, float* thi) {
  for (; g < gend; ++g) {
    const int4 gi = s_groups[g];
    const float* x = spec + gi.x;
    const float2* w = s_binw + gi.y;
    float lo = 0.0f, hi = 0.0f;
    for (int i = 0; i < gi.z; ++i) {
      const float xv = x[i];
      const float2 wv = w[i];
      lo = fmaf(wv.x, xv, lo);
      hi = fmaf(wv.y, xv, hi);
    }
    tlo[gi.w] = lo;
    thi[gi.w + 1] = hi;
  }
}

// ---- progress counters in shared memory (monotonic, release / acquire at CTA scope) -----------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned load_acquire(const unsigned* counter) {
  unsigned seen;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(smem_addr(counter)) : "memory");
  return seen;
}
__device__ __forceinline__ void add_release(unsigned* counter) {  // this thread's (and, after __syncwarp, its warp's) writes first
  asm volatile("red.release.cta.shared.add.u32 [%0], 1;" ::"r"(smem_addr(counter)) : "memory");
}
// counters are per batch parity and advance by kBatchPairs per batch: batch b is through when its counter reaches this
__device__ __forceinline__ unsigned batch_goal(unsigned b) { return kBatchPairs * ((b >> 1) + 1); }

struct Counters {
  unsigned fft_ticket, mel_ticket, store_ticket, pad_;
  unsigned arrived[2];   // frame pairs whose spectra are in their slot
  unsigned mel_done[2];  // filterbank tasks finished
  unsigned stored[2];    // store tasks finished
};
constexpr int kSmemCounterBytes = 256;

// One frame pair after the windowed samples are in v: transform_first = FFT, twiddle; transform_second =
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
  // 32x32 transpose through the padded slot, real plane then imaginary plane (the slot is sized for one plane: two slots
  // per batch position leave the rest of shared memory to L1, which serves the 48 % overlap of consecutive frames)
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

// dB rows + per-frame max of one frame pair from the filterbank tile (rows 2 * position, 2 * position + 1 of the two halves)
__device__ __forceinline__ void store_pair(const LogmelParams& p, const PairInfo& pi, const float* tlo, const float* thi,
                                           int lane) {
  float* dst = p.power + (long long)pi.clip * p.power_clip_stride + (long long)pi.frame * KOE_N_MELS;
  // the 160 values of the two rows (row B follows row A in the clip's block), five per lane: j = lane + 32 q;
  // j < 80 -> frame A filter j, else frame B filter j - 80, which sits kTileStride - 80 = 1 float further in the tile
  float db[5];
#pragma unroll
  for (int qd = 0; qd < 5; ++qd) {
    const int j = lane + 32 * qd;
    const bool second = qd > 2 || (qd == 2 && lane >= KOE_N_MELS - 64);
    const int t = j + (second ? kTileStride - KOE_N_MELS : 0);
    db[qd] = db_from_power(tlo[t] + thi[t]);  // stored in dB: the consumer of the buffer only subtracts its reference
  }
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
  Counters* ctr = reinterpret_cast<Counters*>(smem_raw);
  float* s_hann = reinterpret_cast<float*>(smem_raw + kSmemCounterBytes);     // 1024
  float2* s_tw = reinterpret_cast<float2*>(s_hann + kFrameLen);               // 1024 float2
  float2* s_binw = s_tw + 1024;                                               // kMaxBins float2 (generic bank)
  int4* s_groups = reinterpret_cast<int4*>(s_binw + kMaxBins);                // kMaxGroups       (generic bank)
  int* s_runs = reinterpret_cast<int*>(s_groups + kMaxGroups);                // kBatchPairs + 1 (+ pad to 32)
  float* s_tlo = reinterpret_cast<float*>(s_runs + 32);                       // [2 parities][32 rows][81]: falling halves
  float* s_thi = s_tlo + 2 * kTileFloats;                                     // [2 parities][32 rows][81]: rising halves
  float* s_slots = s_thi + 2 * kTileFloats;                                   // kBatchPairs x (slot parity 0 | slot parity 1)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < kFrameLen; i += kThreads) s_hann[i] = tab.hann[i];
  for (int i = tid; i < 1024; i += kThreads) s_tw[i] = tab.tw[i];
  if (!kDefaultBank) {
    for (int i = tid; i < tab.n_bins; i += kThreads) s_binw[i] = tab.binw[i];
    for (int i = tid; i < tab.n_groups; i += kThreads) s_groups[i] = tab.groups[i];
    if (tid <= kBatchPairs) s_runs[tid] = tab.runs[tid];
  }
  if (tid < 2 * 2 * kBatchPairs) s_thi[tid * kTileStride] = 0.0f;  // filter 0 has no rising half from a lower interval
  if (tid < (int)(sizeof(Counters) / sizeof(unsigned))) reinterpret_cast<unsigned*>(ctr)[tid] = 0;
  __syncthreads();

  const unsigned ppc = (unsigned)(p.n_frames + 1) >> 1;  // frame pairs per clip
  const unsigned total_pairs = (unsigned)p.n_clips * ppc;
  const unsigned n_batches = (total_pairs + kBatchPairs - 1) / kBatchPairs;
  // this CTA's batches: blockIdx.x, blockIdx.x + gridDim.x, ...; its tickets number them 0, 1, ...
  const unsigned my_batches = blockIdx.x < n_batches ? (n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const unsigned n_tickets = my_batches * kBatchPairs;  // of each kind: frame pairs, filterbank tasks, store tasks

  auto describe_ticket = [&](unsigned t) -> PairInfo {  // pair at position t % 16 of this CTA's batch t / 16
    if (t >= n_tickets) return describe_pair(p, p.n_clips, 0);
    const unsigned b = t / kBatchPairs, j = t - b * kBatchPairs;
    PairCursor pos;
    pos.init((blockIdx.x + b * gridDim.x) * kBatchPairs + j, ppc);
    return describe_pair(p, pos.clip, pos.pic);
  };

  // ---- filterbank / store tasks: taken by compare-and-swap only when their inputs are ready, so taking never blocks
  auto try_mel_task = [&]() -> bool {
    unsigned t = 0, ok = 0;
    if (lane == 0) {
      t = load_acquire(&ctr->mel_ticket);
      if (t < n_tickets) {
        const unsigned b = t / kBatchPairs;
        // the batch's 32 spectra are in their slots, and the rows of the batch that used this tile before are stored
        ok = load_acquire(&ctr->arrived[b & 1]) >= batch_goal(b) &&
             (b < 2 || load_acquire(&ctr->stored[b & 1]) >= batch_goal(b - 2));
        if (ok) ok = atomicCAS(&ctr->mel_ticket, t, t + 1) == t;
      }
    }
    ok = __shfl_sync(kFullMask, ok, 0);
    if (!ok) return false;
    t = __shfl_sync(kFullMask, t, 0);
    const unsigned b = t / kBatchPairs, run = t - b * kBatchPairs;
    const float* spec = s_slots + (lane >> 1) * kPosStride + (b & 1) * kSlotFloats + (lane & 1) * kSecondFrame;
    float* tlo = s_tlo + ((b & 1) * 2 * kBatchPairs + lane) * kTileStride;
    float* thi = s_thi + ((b & 1) * 2 * kBatchPairs + lane) * kTileStride;
    if (kDefaultBank)
      mel_task_default((int)run, spec, tlo, thi);
    else
      mel_task_generic(s_runs[run], s_runs[run + 1], s_groups, s_binw, spec, tlo, thi);
    __syncwarp();
    if (lane == 0) add_release(&ctr->mel_done[b & 1]);
    return true;
  };
  auto try_store_task = [&]() -> bool {
    unsigned t = 0, ok = 0;
    if (lane == 0) {
      t = load_acquire(&ctr->store_ticket);
      if (t < n_tickets) {
        const unsigned b = t / kBatchPairs;
        ok = load_acquire(&ctr->mel_done[b & 1]) >= batch_goal(b);
        if (ok) ok = atomicCAS(&ctr->store_ticket, t, t + 1) == t;
      }
    }
    ok = __shfl_sync(kFullMask, ok, 0);
    if (!ok) return false;
    t = __shfl_sync(kFullMask, t, 0);
    const unsigned b = t / kBatchPairs, j = t - b * kBatchPairs;
    const PairInfo pi = describe_ticket(t);
    if (pi.frame >= 0) {
      const int row = ((b & 1) * 2 * kBatchPairs + 2 * j) * kTileStride;
      store_pair(p, pi, s_tlo + row, s_thi + row, lane);
    }
    __syncwarp();
    if (lane == 0) add_release(&ctr->stored[b & 1]);
    return true;
  };
  // wait until the filterbank of batch b is complete (its spectra may be overwritten) -- working, not idling
  auto help_until_mel_done = [&](unsigned b) {
    for (unsigned spin = 0; load_acquire(&ctr->mel_done[b & 1]) < batch_goal(b);) {
      if (try_store_task() || try_mel_task()) continue;
      __nanosleep(20);
      if (++spin > (1u << 22)) __trap();  // a protocol bug traps (error to the host) instead of hanging the GPU
    }
  };

  // ---- main loop: one frame pair per ticket
  unsigned ticket = 0;
  if (lane == 0) ticket = atomicAdd(&ctr->fft_ticket, 1u);
  ticket = __shfl_sync(kFullMask, ticket, 0);
  float2 v[32];
  {
    const PairInfo first = describe_ticket(ticket);
    if (first.interior) load_interior(p, first, lane, v);
  }
  while (ticket < n_tickets) {
    const unsigned batch = ticket / kBatchPairs, position = ticket - batch * kBatchPairs;
    float* slot = s_slots + position * kPosStride + (batch & 1) * kSlotFloats;
    const PairInfo cur = describe_ticket(ticket);
    const bool have_pair = cur.frame >= 0;
    bool slot_free = batch < 2;
    if (have_pair) {
      if (!cur.interior) {
        // edge frames: masked loads, staged through the pair's own slot sixteen rows at a time
        if (!slot_free) help_until_mel_done(batch - 2);
        slot_free = true;
        float2* st2 = reinterpret_cast<float2*>(slot);
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
      // periodic Hann window
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) v[bitrev5(n1)] = __fmul2_rn(v[bitrev5(n1)], bcast(s_hann[32 * n1 + lane]));
      transform_first(v, s_tw, lane);
    }
    // the slot still holds the spectra of batch - 2 until their filterbank is complete (and an arrival must not overtake
    // that batch's arrivals in the counter either)
    if (!slot_free) help_until_mel_done(batch - 2);
    if (have_pair) transform_second(v, slot, lane);
    __syncwarp();
    if (lane == 0) add_release(&ctr->arrived[batch & 1]);
    // next pair: its audio goes to registers now and is in flight while this warp serves filterbank / store tasks
    if (lane == 0) ticket = atomicAdd(&ctr->fft_ticket, 1u);
    ticket = __shfl_sync(kFullMask, ticket, 0);
    {
      const PairInfo nxt = describe_ticket(ticket);
      if (nxt.interior) load_interior(p, nxt, lane, v);
    }
#pragma unroll 1
    for (int n = 0; n < 2; ++n)
      if (!try_store_task()) break;
#pragma unroll 1
    for (int n = 0; n < 2; ++n)
      if (!try_mel_task()) break;
  }
  // ---- no pairs left: drain the filterbank and store tasks
  for (unsigned spin = 0; load_acquire(&ctr->store_ticket) < n_tickets;) {
    if (try_store_task() || try_mel_task()) continue;
    __nanosleep(20);
    if (++spin > (1u << 22)) __trap();
  }
}

constexpr size_t kLogmelSmem = kSmemCounterBytes + sizeof(float) * kFrameLen + sizeof(float2) * 1024 +
                               sizeof(float2) * kMaxBins + sizeof(int4) * kMaxGroups + sizeof(int) * 32 +
                               sizeof(float) * (4 * kTileFloats + kBatchPairs * kPosStride);
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
  int* d_tables = nullptr;  // groups[kMaxGroups] (int4) | runs[kBatchPairs + 1 -> 32]
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
  // runs: contiguous group ranges, one per filterbank task, balanced on bins + a per-group overhead (generic kernel only)
  {
    int* runs = tables.data() + 4 * kMaxGroups;
    auto cost = [&](int g) { return 3 + tables[4 * g + 2]; };
    long long total = 0;
    for (int g = 0; g < n_groups; ++g) total += cost(g);
    long long acc = 0;
    int r = 1;
    runs[0] = 0;
    for (int g = 0; g < n_groups && r < kBatchPairs; ++g) {
      acc += cost(g);
      if (acc * kBatchPairs >= total * r) runs[r++] = g + 1;
    }
    for (; r <= kBatchPairs; ++r) runs[r] = n_groups;
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
  KOE_REQUIRE((long long)a->n_clips * ppc < (1ll << 31) - kBatchPairs * 65536ll,
              "koe_logmel_power: more than 2^31 frame pairs in one call");
  const long long n_blocks = ((long long)a->n_clips * ppc + kBatchPairs - 1) / kBatchPairs;
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
