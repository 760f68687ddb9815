// Log-mel frontend for sm_100a: framing + periodic Hann + 1024-point FFT + |X|^2 + sparse Slaney
// filterbank, fused in one kernel; then a small dB-normalisation kernel.
//
// Replaces the per-clip librosa loop of SimplifiedDualStreamModel.extract_mel_features
// (reference src/model/simplified_dual_stream_model.py:184-229).
//
// Kernel design (see DESIGN.md section "K1"):
//   * one warp transforms TWO real frames of the same clip at once as one complex 1024-point FFT
//     (z = a + i b), decomposed 32 x 32: radix-32 FFT in registers, twiddle, 32x32 transpose through a
//     padded shared-memory tile, radix-32 FFT in registers.  After the second pass lane l holds
//     X[l + 32 r]; the conjugate-symmetric partner X[1024 - k] lives in lane (32 - l) & 31, so the two
//     real spectra are separated with 32 warp shuffles and no further shared-memory round trip.
//   * a CTA (8 warps) therefore produces 16 power spectra per iteration, parked in shared memory with
//     bank-skewed strides; the Slaney filterbank (<= 2 filters per bin, 992 non-zeros) is then applied
//     by all 8 warps with lanes = 16 frames x 2 adjacent filters (weights broadcast, spectra conflict-free).
//   * results leave through a shared staging tile as coalesced float4 rows of 80 mel powers.
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace koe {

constexpr int kFrameLen = 1024;
constexpr int kBins = 513;
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kSlots = 2 * kWarps;      // frames per CTA iteration
constexpr int kXbufStride = 1058;       // floats per warp buffer: >= 32*33, == 2 (mod 32)
constexpr int kSecondFrame = 513;       // offset of the warp's second spectrum, == 1 (mod 32)
constexpr int kMaxBins = 512;           // spectrum bins that carry filterbank weight (507 for 80..8000 Hz)
constexpr int kRuns = 2 * kWarps;       // the weighted bins are cut into 16 equal runs, two per warp
constexpr int kTileStride = 81;         // mel accumulator row stride (floats), odd: conflict-free across slots

struct FrontendTables {
  const float* hann;     // [1024]
  const float2* tw;      // [32][32] W_1024^(k1*n2)
  // Slaney filterbank, bin-major: a spectrum bin feeds at most two ADJACENT filters (fl, fl + 1)
  const float2* binw;    // [n_bins] (weight into filter fl, weight into filter fl + 1)
  const int* binkf;      // [n_bins] bin index k | fl << 16
  const int* runs;       // [kRuns + 1] run boundaries into the bin list
  int n_bins;
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

// cos / sin of 2*pi*t/32 for t = 0..15 as literals so the unrolled butterflies use immediates
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
// multiply (tr + i ti) by W_32^t = cos(2 pi t / 32) - i sin(2 pi t / 32); t is a compile-time value after unrolling
__device__ __forceinline__ void mul_w32(int t, float tr, float ti, float& orr, float& oi) {
  if (t == 0) {
    orr = tr;
    oi = ti;
  } else if (t == 8) {  // -i
    orr = ti;
    oi = -tr;
  } else if (t == 4) {  // (1 - i)/sqrt2
    const float c = 0.70710678118654752f;
    orr = c * (tr + ti);
    oi = c * (ti - tr);
  } else if (t == 12) {  // (-1 - i)/sqrt2
    const float c = 0.70710678118654752f;
    orr = c * (ti - tr);
    oi = -c * (tr + ti);
  } else {
    const float wr = cos32(t);
    const float ws = cos32(t > 8 ? t - 8 : 8 - t);  // sin(2 pi t/32) = cos(2 pi (t-8)/32), cos even
    orr = tr * wr + ti * ws;
    oi = ti * wr - tr * ws;
  }
}

// In-register radix-2 decimation-in-frequency FFT of 32 complex values (forward, e^{-i...}).
// On return element i holds X[bitrev5(i)].
__device__ __forceinline__ void fft32(float (&re)[32], float (&im)[32]) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if ((i & s) == 0) {
        const int j = i + s;
        const int t = (i & (s - 1)) * (16 / s);
        const float ar = re[i], ai = im[i], br = re[j], bi = im[j];
        re[i] = ar + br;
        im[i] = ai + bi;
        mul_w32(t, ar - br, ai - bi, re[j], im[j]);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 2)
logmel_power_kernel(FrontendTables tab, LogmelParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_hann = reinterpret_cast<float*>(smem_raw);              // 1024
  float2* s_tw = reinterpret_cast<float2*>(s_hann + kFrameLen);    // 1024 float2
  float2* s_binw = s_tw + 1024;                                    // kMaxBins float2
  int* s_binkf = reinterpret_cast<int*>(s_binw + kMaxBins);        // kMaxBins
  int* s_runs = s_binkf + kMaxBins;                                // kRuns + 1 (+ pad to 24)
  float* s_xbuf = reinterpret_cast<float*>(s_runs + 24);           // kWarps * kXbufStride
  float* s_tile = s_xbuf + kWarps * kXbufStride;                   // kSlots * kTileStride mel accumulators

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < kFrameLen; i += kThreads) s_hann[i] = tab.hann[i];
  for (int i = tid; i < 1024; i += kThreads) s_tw[i] = tab.tw[i];
  for (int i = tid; i < tab.n_bins; i += kThreads) {
    s_binw[i] = tab.binw[i];
    s_binkf[i] = tab.binkf[i];
  }
  if (tid <= kRuns) s_runs[tid] = tab.runs[tid];
  __syncthreads();

  const int ppc = (p.n_frames + 1) >> 1;  // frame pairs per clip
  const long long total_pairs = (long long)p.n_clips * ppc;
  const long long n_blocks = (total_pairs + kWarps - 1) / kWarps;
  float* xb = s_xbuf + warp * kXbufStride;

  for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    for (int i = tid; i < kSlots * kTileStride; i += kThreads) s_tile[i] = 0.0f;  // consumed after two barriers
    // ------------------------------------------------------------------ FFT phase (per warp)
    const long long pair = blk * kWarps + warp;
    if (pair < total_pairs) {
      const int b = (int)(pair / ppc);
      const int ga = 2 * (int)(pair % ppc);
      const bool has_b = ga + 1 < p.n_frames;
      const float* clip = p.audio + (long long)b * p.audio_stride;
      const int fa = p.frame_offset + ga * p.frame_step, fb = fa + p.frame_step;  // frame indices in hops
      int lo_a = 0, hi_a = p.n_samples, lo_b = 0, hi_b = p.n_samples;
      if (p.lo_rel != KOE_NO_EDGE) {
        lo_a = max(lo_a, p.sample_offset + (fa + p.lo_rel) * p.hop);
        lo_b = max(lo_b, p.sample_offset + (fb + p.lo_rel) * p.hop);
      }
      if (p.hi_rel != KOE_NO_EDGE) {
        hi_a = min(hi_a, p.sample_offset + (fa + p.hi_rel) * p.hop);
        hi_b = min(hi_b, p.sample_offset + (fb + p.hi_rel) * p.hop);
      }
      if (!has_b) hi_b = lo_b;  // empty range: second frame reads as silence
      const int fa_lo = p.sample_offset + fa * p.hop - kFrameLen / 2;  // first sample of each frame
      const int fb_lo = p.sample_offset + fb * p.hop - kFrameLen / 2;
      const int sa0 = fa_lo + lane, sb0 = fb_lo + lane;

      float re[32], im[32];
      if (fa_lo >= lo_a && fa_lo + kFrameLen <= hi_a && fb_lo >= lo_b && fb_lo + kFrameLen <= hi_b) {
        // interior frames (all but the first / last of a clip): no masking, one base pointer per frame
        const float* __restrict__ pa = clip + sa0;
        const float* __restrict__ pb = clip + sb0;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
          const float w = s_hann[32 * n1 + lane];
          re[n1] = __ldg(pa + 32 * n1) * w;
          im[n1] = __ldg(pb + 32 * n1) * w;
        }
      } else if (p.pad_mode == 1) {
        // numpy "reflect" padding about the first / last sample (MelSlidingWindowExtractor default)
        const int last = p.n_samples - 1;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
          int sa = sa0 + 32 * n1, sb = sb0 + 32 * n1;
          sa = sa < 0 ? -sa : (sa > last ? 2 * last - sa : sa);
          sb = sb < 0 ? -sb : (sb > last ? 2 * last - sb : sb);
          const float w = s_hann[32 * n1 + lane];
          const float va = (sa >= 0 && sa <= last) ? __ldg(clip + sa) : 0.0f;
          const float vb = (has_b && sb >= 0 && sb <= last) ? __ldg(clip + sb) : 0.0f;
          re[n1] = va * w;
          im[n1] = vb * w;
        }
      } else {
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
          const int sa = sa0 + 32 * n1, sb = sb0 + 32 * n1;
          const float w = s_hann[32 * n1 + lane];
          const float va = (sa >= lo_a && sa < hi_a) ? __ldg(clip + sa) : 0.0f;
          const float vb = (sb >= lo_b && sb < hi_b) ? __ldg(clip + sb) : 0.0f;
          re[n1] = va * w;
          im[n1] = vb * w;
        }
      }
      fft32(re, im);  // element i = Y[k1 = bitrev5(i)] for column n2 = lane
#pragma unroll
      for (int i = 1; i < 32; ++i) {
        const float2 w = s_tw[bitrev5(i) * 32 + lane];  // W_1024^(k1 * n2)
        const float tr = re[i], ti = im[i];
        re[i] = tr * w.x - ti * w.y;
        im[i] = tr * w.y + ti * w.x;
      }
      // 32x32 transpose, real plane then imaginary plane, through the warp's padded tile
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 32; ++i) xb[bitrev5(i) * 33 + lane] = re[i];
      __syncwarp();
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) re[n2] = xb[lane * 33 + n2];
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 32; ++i) xb[bitrev5(i) * 33 + lane] = im[i];
      __syncwarp();
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) im[n2] = xb[lane * 33 + n2];
      __syncwarp();
      fft32(re, im);  // element i = Z[lane + 32 * bitrev5(i)]

      // separate the two real spectra: partner of k = lane + 32 r is 1024 - k = ((32-lane)&31) + 32 r'
      const int src = (32 - lane) & 31;
      float pr[16], pi[16];
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        pr[s] = __shfl_sync(kFullMask, re[bitrev5(16 + s)], src);
        pi[s] = __shfl_sync(kFullMask, im[bitrev5(16 + s)], src);
      }
      float* pa = xb;
      float* pb = xb + kSecondFrame;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float zr = re[bitrev5(r)], zi = im[bitrev5(r)];
        // lanes 1..31: partner register r' = 31 - r (slot 15 - r); lane 0: r' = 32 - r (slot 16 - r), r = 0 is its own partner
        float qr = pr[15 - r], qi = pi[15 - r];
        if (lane == 0) {
          qr = (r == 0) ? zr : pr[(16 - r) & 15];
          qi = (r == 0) ? zi : pi[(16 - r) & 15];
        }
        const float ar = zr + qr, ai = zi - qi, br = zi + qi, bi = zr - qr;
        pa[lane + 32 * r] = 0.25f * (ar * ar + ai * ai);
        pb[lane + 32 * r] = 0.25f * (br * br + bi * bi);
      }
      if (lane == 0) {  // Nyquist bin 512 = register r = 16, self-paired
        const float zr = re[bitrev5(16)], zi = im[bitrev5(16)];
        pa[512] = zr * zr;
        pb[512] = zi * zi;
      }
    }
    __syncthreads();

    // ------------------------------------------------------------------ mel phase (whole CTA)
    // lane = (frame slot, run): walk the run's bins once, feeding the two adjacent filters each bin touches;
    // runs are cut on interval boundaries, so a filter receives at most two partial sums (its rising half from one
    // run, its falling half from the same or the next) and the float atomics are order-free: 0 + a + b == 0 + b + a.
    {
      const int slot = lane & 15, run = 2 * warp + (lane >> 4);
      const float* spec = s_xbuf + (slot >> 1) * kXbufStride + (slot & 1) * kSecondFrame;
      float* trow = s_tile + slot * kTileStride;
      int i = s_runs[run];
      const int iend = s_runs[run + 1];
      int fl = i < iend ? (s_binkf[i] >> 16) : 0;
      float acc_lo = 0.0f, acc_hi = 0.0f;
      for (; i < iend; ++i) {
        const int kf = s_binkf[i];
        const int f = kf >> 16;
        if (f != fl) {
          atomicAdd(trow + fl, acc_lo);
          if (f == fl + 1) {
            acc_lo = acc_hi;
          } else {
            if (fl + 1 < KOE_N_MELS) atomicAdd(trow + fl + 1, acc_hi);
            acc_lo = 0.0f;
          }
          acc_hi = 0.0f;
          fl = f;
        }
        const float2 w = s_binw[i];
        const float x = spec[kf & 0xffff];
        acc_lo = fmaf(w.x, x, acc_lo);
        acc_hi = fmaf(w.y, x, acc_hi);
      }
      atomicAdd(trow + fl, acc_lo);
      if (fl + 1 < KOE_N_MELS) atomicAdd(trow + fl + 1, acc_hi);
    }
    __syncthreads();

    // ------------------------------------------------------------------ store phase: warp w owns slots 2w, 2w+1
    {
      const long long pr_ = blk * kWarps + warp;
      if (pr_ < total_pairs) {
        const int b = (int)(pr_ / ppc);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int g = 2 * (int)(pr_ % ppc) + h;
          if (g < p.n_frames) {
            const float* trow = s_tile + (2 * warp + h) * kTileStride;
            float* dst = p.power + (long long)b * p.power_clip_stride + (long long)g * KOE_N_MELS;
            // stored in dB: 10 log10(max(power, amin)); the consumer only subtracts its reference and clamps
            const float v0 = power_db(trow[lane]), v1 = power_db(trow[lane + 32]);
            const float v2 = lane < 16 ? power_db(trow[lane + 64]) : -INFINITY;
            dst[lane] = v0;
            dst[lane + 32] = v1;
            if (lane < 16) dst[lane + 64] = v2;
            const float mx = warp_max(fmaxf(fmaxf(v0, v1), v2));
            if (lane == 0 && p.frame_max != nullptr) p.frame_max[(long long)b * p.fmax_clip_stride + g] = mx;
          }
        }
      }
    }
    __syncthreads();
  }
}

constexpr size_t kLogmelSmem = sizeof(float) * kFrameLen + sizeof(float2) * 1024 + sizeof(float2) * kMaxBins +
                               sizeof(int) * (kMaxBins + 24) +
                               sizeof(float) * (kWarps * kXbufStride + kSlots * kTileStride);

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
  int* d_tables = nullptr;  // binkf[kMaxBins] | runs[kRuns + 1]
  int n_bins = 0;
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

  // bin-major sparse filterbank: every weighted bin feeds one filter or two adjacent ones
  std::vector<float2> binw;
  std::vector<int> tables(kMaxBins + kRuns + 1, 0);
  for (int k = 0; k < kBins; ++k) {
    int first = -1, count = 0, last = -1;
    for (int m = 0; m < n_mels; ++m)
      if (fe->fb_host[(size_t)m * kBins + k] > 0.0f) {
        if (first < 0) first = m;
        last = m;
        ++count;
      }
    if (count == 0) continue;
    if (count > 2 || last - first > 1 || (int)binw.size() >= kMaxBins) {
      delete fe;
      cudaSetDevice(prev);
      return fail(KOE_E_UNSUPPORTED, "koe_frontend_create: filterbank is not a bank of adjacent triangles at bin %d", k);
    }
    tables[binw.size()] = k | (first << 16);
    binw.push_back(make_float2(fe->fb_host[(size_t)first * kBins + k],
                               count == 2 ? fe->fb_host[(size_t)last * kBins + k] : 0.0f));
  }
  fe->n_bins = (int)binw.size();
  // run boundaries on filter-interval boundaries (where fl changes), as balanced as that allows: then a
  // filter's rising half lies in one run and its falling half in the same or the next run -> <= 2 partial sums
  {
    int r = 1;
    tables[kMaxBins] = 0;
    for (int i = 1; i < fe->n_bins && r < kRuns; ++i) {
      const bool boundary = (tables[i] >> 16) != (tables[i - 1] >> 16);
      if (boundary && (long long)i * kRuns >= (long long)fe->n_bins * r) tables[kMaxBins + r++] = i;
    }
    for (; r <= kRuns; ++r) tables[kMaxBins + r] = fe->n_bins;
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
    e = cudaFuncSetAttribute(logmel_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLogmelSmem);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess) {
    fe->num_sms = prop.multiProcessorCount;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fe->occupancy, logmel_power_kernel, kThreads, kLogmelSmem);
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
  tab.binkf = fe->d_tables;
  tab.runs = fe->d_tables + kMaxBins;
  tab.n_bins = fe->n_bins;
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
  const long long n_blocks = ((long long)a->n_clips * ppc + kWarps - 1) / kWarps;
  const long long max_grid = (long long)fe->num_sms * fe->occupancy;
  const int grid = (int)std::min(n_blocks, max_grid);
  logmel_power_kernel<<<grid, kThreads, kLogmelSmem, (cudaStream_t)stream>>>(tab, p);
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
