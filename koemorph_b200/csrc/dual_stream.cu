// Dual-stream attention core on CUDA cores (precision 0: exact fp32 FMA), emotion stream and EMA scan.
//
// Replaces DualStreamCrossAttention.forward (reference src/model/dual_stream_attention.py:162-280),
// the eGeMAPS compression (src/features/opensmile_extractor.py:583-604) and apply_temporal_smoothing
// (src/model/simplified_dual_stream_model.py:341-368).
//
// One persistent CTA per SM walks over (clip, window) work items.  Per window, everything between the
// mel-power rows in HBM and the 52 output coefficients stays in shared memory / registers:
//   P0  window dB reference (max over the window's frame maxima), normalise rows -> A block
//   G1  z = Xn[80 x K] Wc^T + bc, LayerNorm          (K = 259 / 515, streamed in 64-row blocks)
//   G3  V = enc Wv^T + bv                            [80 x 256]
//   S   scores^T = enc Qk^T                          [80 x 224]; Qk folds (Wq q + bq)/sqrt(32) and Wk,
//                                                    the key bias adds a per-row constant that softmax drops
//   softmax over the 80 mel-channel tokens, P V per head, O -> h = relu(O Wa^T + ba), sigmoid(h w2 + b2)
//   fusion: out[idx] = clamp(coef[idx] * y), coef = 0.5 (softmax(mel_w/T) + softmax(emo_w/T)).
#include <algorithm>

#include "common.cuh"
#include "core_params.cuh"

namespace koe {

constexpr int kTok = KOE_N_MELS;         // 80 tokens (mel channels) per window
constexpr int kD = KOE_D_MODEL;          // 256
constexpr int kHQ = KOE_N_HEADS * KOE_N_MOUTH;  // 224 score rows
constexpr int kCoreThreads = 256;
constexpr int kKC = 16;                  // weight rows per cp.async stage
constexpr int kABlockRows = 64;          // rows of Xn staged per G1 block
constexpr int kEncStride = 81;           // encT[n][m] row stride (odd: conflict-free transposed stores)
constexpr int kVStride = kD;             // V[m][n]
constexpr int kSStride = 84;             // S[hq][t] row stride (float4-aligned)
constexpr int kOStride = 36;             // OT[k][q] row stride (float4-aligned)

// acc[TM][TN] += A[k][ty*TM + i] * B[k][tx + 32 j] for k < K.
// A lives in shared memory ([K][lda]); B is streamed from global ([ceil16(K)][32*TN], zero padded) through a
// two-stage cp.async ring.  Ends with a __syncthreads().
template <int TM, int TN, bool A_VEC2>
__device__ __forceinline__ void gemm_acc(const float* __restrict__ As, int lda, const float* __restrict__ Bg, int K,
                                         float* __restrict__ stage, float (&acc)[TM][TN], int tid) {
  constexpr int NC = 32 * TN;
  constexpr int F4 = kKC * NC / 4;
  const int tx = tid & 31, ty = tid >> 5;
  const int nchunks = (K + kKC - 1) / kKC;
  auto issue = [&](int c) {
    const float4* src = reinterpret_cast<const float4*>(Bg + (size_t)c * kKC * NC);
    float4* dst = reinterpret_cast<float4*>(stage + (c & 1) * kKC * NC);
#pragma unroll
    for (int i = 0; i < F4 / kCoreThreads; ++i) cp_async16(dst + tid + i * kCoreThreads, src + tid + i * kCoreThreads);
    cp_async_commit();
  };
  // packed fp32x2 accumulators: acc2[i][j] = (acc[i][2j], acc[i][2j+1]); one FFMA2 (broadcast a, pair of b) does the work
  // of two FFMAs in one issue slot -- same rounding per element, so the results are bit-identical to the scalar form
  static_assert(TN % 2 == 0, "TN must be even");
  float2 acc2[TM][TN / 2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN / 2; ++j) acc2[i][j] = make_float2(acc[i][2 * j], acc[i][2 * j + 1]);
  issue(0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      issue(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* bs = stage + (c & 1) * kKC * NC + tx;
    const float* as = As + (size_t)c * kKC * lda + ty * TM;
    const int kmax = min(kKC, K - c * kKC);
#pragma unroll 4
    for (int kk = 0; kk < kmax; ++kk) {
      float a[TM];
      float2 b[TN / 2];
      if constexpr (A_VEC2) {
#pragma unroll
        for (int i = 0; i < TM / 2; ++i) {
          const float2 v = *reinterpret_cast<const float2*>(as + kk * lda + 2 * i);
          a[2 * i] = v.x;
          a[2 * i + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = as[kk * lda + i];
      }
#pragma unroll
      for (int j = 0; j < TN / 2; ++j) b[j] = make_float2(bs[kk * NC + 64 * j], bs[kk * NC + 64 * j + 32]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) acc2[i][j] = __ffma2_rn(make_float2(a[i], a[i]), b[j], acc2[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN / 2; ++j) {
      acc[i][2 * j] = acc2[i][j].x;
      acc[i][2 * j + 1] = acc2[i][j].y;
    }
}

constexpr size_t kCoreSmemFloats = (size_t)kABlockRows * kTok      // A block
                                   + (size_t)kD * kEncStride       // encT, later S/P
                                   + (size_t)kTok * kVStride       // V, later OT
                                   + 2 * kKC * kD                  // weight stages
                                   + 64;                           // reductions
constexpr size_t kCoreSmem = kCoreSmemFloats * sizeof(float);

__global__ void __launch_bounds__(kCoreThreads, 1) dual_stream_fp32_kernel(CoreParams p) {
  extern __shared__ __align__(16) float smem[];
  float* s_a = smem;                               // [64][80]
  float* s_enc = s_a + kABlockRows * kTok;         // [256][81]  (S: [224][84])
  float* s_v = s_enc + kD * kEncStride;            // [80][256]  (OT: [256][36])
  float* s_stage = s_v + kTok * kVStride;          // [2][16][256]
  float* s_red = s_stage + 2 * kKC * kD;           // [64]

  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const koe_core_weights& W = p.w;
  const int T = p.frames_per_window;
  const int n_items = p.n_clips * p.n_out;

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / p.n_out, wi = item % p.n_out;
    const bool prenorm = p.mel_long != nullptr;

    // ---- P0: dB reference of this window -------------------------------------------------------
    float ref_db = 0.0f;
    if (!prenorm) {
      float mx = -INFINITY;
      for (int k = tid; k < T; k += kCoreThreads) {
        const int v = window_variant(p, k);
        mx = fmaxf(mx, p.fmax[v][window_row(p, v, b, wi, k)]);
      }
      mx = warp_max(mx);
      if (tx == 0) s_red[ty] = mx;
      __syncthreads();
      if (tid < 32) {
        float v = tid < 8 ? s_red[tid] : -INFINITY;
        v = warp_max(v);
        if (tid == 0) s_red[8] = v;
      }
      __syncthreads();
      ref_db = s_red[8];
    }

    // ---- G1: z = Xn Wc^T, K streamed in 64-row blocks -------------------------------------------
    float acc[10][8];
#pragma unroll
    for (int i = 0; i < 10; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    for (int kb = 0; kb < W.k_mel; kb += kABlockRows) {
      const int rows = min(kABlockRows, W.k_mel - kb);
      for (int idx = tid; idx < rows * (kTok / 4); idx += kCoreThreads) {
        const int r = idx / (kTok / 4), q = idx % (kTok / 4);
        const int t = kb + r;
        int k;  // window frame feeding row t, or -1 for zero padding
        if (t < p.mel_seq) {
          k = t < T ? t : -1;
        } else {
          const int s = t - p.mel_seq;
          k = T >= 3 ? T - 3 + s : (s < T ? s : -1);
        }
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (prenorm) {
          if (t >= p.mel_seq)
            v = __ldg(reinterpret_cast<const float4*>(p.mel_short + ((size_t)b * 3 + (t - p.mel_seq)) * kTok) + q);
          else if (t < p.n_long)
            v = __ldg(reinterpret_cast<const float4*>(p.mel_long + ((size_t)b * p.n_frames + t) * kTok) + q);
        } else if (k >= 0) {
          const int var = window_variant(p, k);
          const float4 pw =
              __ldg(reinterpret_cast<const float4*>(p.power[var] + window_row(p, var, b, wi, k) * kTok) + q);
          v.x = normalise_db(pw.x, ref_db, true);
          v.y = normalise_db(pw.y, ref_db, true);
          v.z = normalise_db(pw.z, ref_db, true);
          v.w = normalise_db(pw.w, ref_db, true);
        }
        *reinterpret_cast<float4*>(s_a + r * kTok + 4 * q) = v;
      }
      __syncthreads();
      gemm_acc<10, 8, true>(s_a, kTok, W.wc_t + (size_t)kb * kD, rows, s_stage, acc, tid);
    }

    // ---- bias + LayerNorm over the 256 features of each token (one warp owns 10 whole tokens) ----
    {
      float bcv[8], gv[8], bv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bcv[j] = __ldg(W.bc + tx + 32 * j);
        gv[j] = __ldg(W.ln_g + tx + 32 * j);
        bv[j] = __ldg(W.ln_b + tx + 32 * j);
      }
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] += bcv[j];
          s += acc[i][j];
        }
        const float mean = warp_sum(s) * (1.0f / kD);
        float q = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = acc[i][j] - mean;
          q = fmaf(d, d, q);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + W.ln_eps);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          s_enc[(tx + 32 * j) * kEncStride + ty * 10 + i] = (acc[i][j] - mean) * rstd * gv[j] + bv[j];
      }
    }
    __syncthreads();

    // ---- G3: V = enc Wv^T + bv -> s_v[m][n] -----------------------------------------------------
#pragma unroll
    for (int i = 0; i < 10; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    gemm_acc<10, 8, false>(s_enc, kEncStride, W.wv_t, kD, s_stage, acc, tid);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float bias = __ldg(W.bv + tx + 32 * j);
#pragma unroll
      for (int i = 0; i < 10; ++i) s_v[(ty * 10 + i) * kVStride + tx + 32 * j] = acc[i][j] + bias;
    }

    // ---- S^T = enc Qk^T (columns hq = h*28 + q; 224..255 are zero padding) ----------------------
#pragma unroll
    for (int i = 0; i < 10; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    gemm_acc<10, 8, false>(s_enc, kEncStride, W.qk_t, kD, s_stage, acc, tid);
    // every warp is past its last read of s_enc (gemm_acc ends with a barrier): reuse it for S[hq][t]
    float* s_s = s_enc;
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
      for (int i = 0; i < 10; ++i) s_s[(tx + 32 * j) * kSStride + ty * 10 + i] = acc[i][j];
    __syncthreads();

    // ---- softmax over the 80 tokens of each of the 224 rows (warp per row) ----------------------
    for (int r = ty; r < kHQ; r += 8) {
      float* row = s_s + r * kSStride;
      const float v0 = row[tx], v1 = row[tx + 32], v2 = tx < 16 ? row[tx + 64] : -INFINITY;
      const float m = warp_max(fmaxf(fmaxf(v0, v1), v2));
      const float e0 = expf(v0 - m), e1 = expf(v1 - m), e2 = tx < 16 ? expf(v2 - m) : 0.0f;
      const float inv = 1.0f / warp_sum(e0 + e1 + e2);
      row[tx] = e0 * inv;
      row[tx + 32] = e1 * inv;
      if (tx < 16) row[tx + 64] = e2 * inv;
    }
    __syncthreads();

    if (p.attn_out != nullptr) {  // head-averaged weights (nn.MultiheadAttention need_weights=True)
      float* dst = p.attn_out + (size_t)item * KOE_N_MOUTH * kTok;
      for (int idx = tid; idx < KOE_N_MOUTH * kTok; idx += kCoreThreads) {
        const int q = idx / kTok, t = idx % kTok;
        float s = 0.0f;
#pragma unroll
        for (int h = 0; h < KOE_N_HEADS; ++h) s += s_s[(h * KOE_N_MOUTH + q) * kSStride + t];
        dst[idx] = s * (1.0f / KOE_N_HEADS);
      }
    }

    // ---- O[h][q][d] = sum_t P[h*28+q][t] V[t][32 h + d]: warp = head, lane = d --------------------
    float o[KOE_N_MOUTH];
#pragma unroll
    for (int q = 0; q < KOE_N_MOUTH; ++q) o[q] = 0.0f;
    {
      const float* prow = s_s + (ty * KOE_N_MOUTH) * kSStride;
      const float* vcol = s_v + 32 * ty + tx;
      for (int t = 0; t < kTok; t += 4) {
        const float v0 = vcol[(t + 0) * kVStride], v1 = vcol[(t + 1) * kVStride];
        const float v2 = vcol[(t + 2) * kVStride], v3 = vcol[(t + 3) * kVStride];
#pragma unroll
        for (int q = 0; q < KOE_N_MOUTH; ++q) {
          const float4 pq = *reinterpret_cast<const float4*>(prow + q * kSStride + t);
          o[q] = fmaf(pq.x, v0, o[q]);
          o[q] = fmaf(pq.y, v1, o[q]);
          o[q] = fmaf(pq.z, v2, o[q]);
          o[q] = fmaf(pq.w, v3, o[q]);
        }
      }
    }
    __syncthreads();  // all reads of s_v done -> reuse as OT[k = 32 h + d][q]
    float* s_o = s_v;
#pragma unroll
    for (int q = 0; q < KOE_N_MOUTH; ++q) s_o[(32 * ty + tx) * kOStride + q] = o[q];
#pragma unroll
    for (int q = KOE_N_MOUTH; q < 32; ++q) s_o[(32 * ty + tx) * kOStride + q] = 0.0f;
    __syncthreads();

    // ---- decoder: h = relu(O Wa^T + ba) [28 x 128], logit = h . w2 + b2 --------------------------
    float hacc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) hacc[i][j] = 0.0f;
    gemm_acc<4, 4, true>(s_o, kOStride, W.wa_t, kD, s_stage, hacc, tid);
    {
      float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float ba = __ldg(W.ba + tx + 32 * j), w2 = __ldg(W.w2 + tx + 32 * j);
#pragma unroll
        for (int i = 0; i < 4; ++i) part[i] = fmaf(fmaxf(hacc[i][j] + ba, 0.0f), w2, part[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) part[i] = warp_sum(part[i]);
      if (tx < 4) {
        const int q = ty * 4 + tx;
        if (q < KOE_N_MOUTH) {
          const float logit = (tx == 0 ? part[0] : tx == 1 ? part[1] : tx == 2 ? part[2] : part[3]) + W.b2;
          const float y = 1.0f / (1.0f + expf(-logit));
          const int idx = __ldg(W.mouth_idx + q);
          const size_t o_off = (size_t)item * KOE_N_BLENDSHAPES + idx;
          p.out[o_off] = fminf(fmaxf(__ldg(W.coef + idx) * y, 0.0f), 1.0f);
          if (p.sigmoid_out != nullptr) p.sigmoid_out[o_off] = y;
        }
      }
    }
    if (tid < KOE_N_EXPR) {
      const float y = __ldg(p.expr_sigmoid + b);
      const int idx = __ldg(W.expr_idx + tid);
      const size_t o_off = (size_t)item * KOE_N_BLENDSHAPES + idx;
      p.out[o_off] = fminf(fmaxf(__ldg(W.coef + idx) * y, 0.0f), 1.0f);
      if (p.sigmoid_out != nullptr) p.sigmoid_out[o_off] = y;
    }
    __syncthreads();
  }
}

// ---- emotion stream: NC clips per CTA (16 for large batches, 4 when that leaves most SMs idle), weights streamed
// through shared memory in 64-row cp.async chunks
constexpr int kEmoClipsMax = 16;
constexpr int kEmoInMax = 272;
constexpr int kEmoChunkRows = 64;
constexpr int kEmoBufFloats = kEmoChunkRows * kD;   // one chunk buffer: 64 rows of up to 256 floats (64 KB)
constexpr size_t kEmoSmem = sizeof(float) * (2 * kEmoBufFloats + kEmoInMax * kEmoClipsMax + kD * kEmoClipsMax + 16 * 8 + 32);

template <int kEmoClips>
__global__ void __launch_bounds__(256) emotion_stream_kernel(koe_core_weights W, const float* __restrict__ emo_in,
                                                             int n_clips, float* __restrict__ expr_sigmoid) {
  extern __shared__ __align__(16) float esm[];
  float* s_buf = esm;                                   // [2][64][<=256]
  float* s_x = s_buf + 2 * kEmoBufFloats;               // [emo_in][16]  (clip-contiguous: broadcast float4 loads)
  float* s_z = s_x + kEmoInMax * kEmoClips;             // [256][16]
  float* s_red = s_z + kD * kEmoClips;                  // [16][8]
  float* s_stat = s_red + 16 * 8;                       // [16][2]
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  pdl_launch_dependents();  // the core kernel may set itself up (barriers, TMEM, constants) while this one runs
  const int c0 = blockIdx.x * kEmoClips;
  const int nc = min(kEmoClips, n_clips - c0);
  const int n1 = (W.emo_in + kEmoChunkRows - 1) / kEmoChunkRows;   // chunks of we1_t [emo_in][256]
  const int n2 = kD / kEmoChunkRows;                               // chunks of we2_t [256][128]

  auto issue = [&](int c) {  // chunk c of the concatenated chunk sequence -> buffer c & 1
    const float* src;
    int floats;
    if (c < n1) {
      const int r0 = c * kEmoChunkRows;
      src = W.we1_t + (size_t)r0 * kD;
      floats = min(kEmoChunkRows, W.emo_in - r0) * kD;
    } else {
      src = W.we2_t + (size_t)(c - n1) * kEmoChunkRows * 128;
      floats = kEmoChunkRows * 128;
    }
    float* dst = s_buf + (c & 1) * kEmoBufFloats;
    for (int i = tid; i < floats / 4; i += 256) cp_async16(dst + 4 * i, src + 4 * i);
    cp_async_commit();
  };
  issue(0);
  issue(1);
  for (int i = tid; i < kEmoClips * W.emo_in; i += 256) {
    const int c = i / W.emo_in, k = i % W.emo_in;
    s_x[k * kEmoClips + c] = c < nc ? emo_in[(size_t)(c0 + c) * W.emo_in + k] : 0.0f;
  }

  // ---- z[c][n] = we1_t[:, n] . x[c] + be1[n]   (thread = feature n, 16 clips in registers)
  float z[kEmoClips];
  {
    const float b = __ldg(W.be1 + tid);
#pragma unroll
    for (int c = 0; c < kEmoClips; ++c) z[c] = b;
  }
  int chunk = 0;
  for (; chunk < n1; ++chunk) {
    cp_async_wait<1>();
    __syncthreads();
    const float* wb = s_buf + (chunk & 1) * kEmoBufFloats + tid;
    const int r0 = chunk * kEmoChunkRows, rows = min(kEmoChunkRows, W.emo_in - r0);
    // eight rows per step: the 8 weight loads and 8 broadcast input loads are issued together (two warps per scheduler
    // cannot hide a shared-memory round trip per row: `short_scoreboard` was 37 % of the samples)
    const float4* xr0 = reinterpret_cast<const float4*>(s_x + r0 * kEmoClips);
    int k = 0;
    for (; k + 8 <= rows; k += 8) {
      float w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = wb[(k + u) * kD];
#pragma unroll
      for (int q = 0; q < kEmoClips / 4; ++q) {
        float4 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = xr0[(k + u) * (kEmoClips / 4) + q];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          z[4 * q + 0] = fmaf(w[u], x[u].x, z[4 * q + 0]);
          z[4 * q + 1] = fmaf(w[u], x[u].y, z[4 * q + 1]);
          z[4 * q + 2] = fmaf(w[u], x[u].z, z[4 * q + 2]);
          z[4 * q + 3] = fmaf(w[u], x[u].w, z[4 * q + 3]);
        }
      }
    }
    for (; k < rows; ++k) {
      const float w = wb[k * kD];
#pragma unroll
      for (int q = 0; q < kEmoClips / 4; ++q) {
        const float4 x = xr0[k * (kEmoClips / 4) + q];
        z[4 * q + 0] = fmaf(w, x.x, z[4 * q + 0]);
        z[4 * q + 1] = fmaf(w, x.y, z[4 * q + 1]);
        z[4 * q + 2] = fmaf(w, x.z, z[4 * q + 2]);
        z[4 * q + 3] = fmaf(w, x.w, z[4 * q + 3]);
      }
    }
    __syncthreads();
    if (chunk + 2 < n1 + n2) issue(chunk + 2); else cp_async_commit();
  }
  // ---- LayerNorm per clip over the 256 threads (two passes)
#pragma unroll
  for (int c = 0; c < kEmoClips; ++c) {
    const float v = warp_sum(z[c]);
    if (tx == 0) s_red[c * 8 + ty] = v;
  }
  __syncthreads();
  if (tid < kEmoClips) {
    float v = 0.0f;
    for (int w = 0; w < 8; ++w) v += s_red[tid * 8 + w];
    s_stat[tid * 2] = v * (1.0f / kD);
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < kEmoClips; ++c) {
    const float d = z[c] - s_stat[c * 2];
    const float v = warp_sum(d * d);
    if (tx == 0) s_red[c * 8 + ty] = v;
  }
  __syncthreads();
  if (tid < kEmoClips) {
    float v = 0.0f;
    for (int w = 0; w < 8; ++w) v += s_red[tid * 8 + w];
    s_stat[tid * 2 + 1] = rsqrtf(v * (1.0f / kD) + W.ln_eps);
  }
  __syncthreads();
  {
    const float g = __ldg(W.eln_g + tid), be = __ldg(W.eln_b + tid);
#pragma unroll
    for (int c = 0; c < kEmoClips; ++c)
      s_z[tid * kEmoClips + c] = (z[c] - s_stat[c * 2]) * s_stat[c * 2 + 1] * g + be;
  }
  // ---- h[c][j] = relu(we2_t[:, j] . zn[c] + be2[j]); thread = (j = tid & 127, clip half = tid >> 7)
  const int j = tid & 127, half = tid >> 7;
  float h[kEmoClips / 2];
  {
    const float b = __ldg(W.be2 + j);
#pragma unroll
    for (int c = 0; c < kEmoClips / 2; ++c) h[c] = b;
  }
  for (; chunk < n1 + n2; ++chunk) {
    cp_async_wait<1>();
    __syncthreads();   // also orders the s_z writes above before the first read
    const float* wb = s_buf + (chunk & 1) * kEmoBufFloats + j;
    const int r0 = (chunk - n1) * kEmoChunkRows;
#pragma unroll 8
    for (int k = 0; k < kEmoChunkRows; ++k) {
      const float w = wb[k * 128];
      const float* zr = s_z + (r0 + k) * kEmoClips + half * (kEmoClips / 2);
      if constexpr (kEmoClips / 2 >= 4) {
#pragma unroll
        for (int q = 0; q < kEmoClips / 8; ++q) {
          const float4 x = reinterpret_cast<const float4*>(zr)[q];
          h[4 * q + 0] = fmaf(w, x.x, h[4 * q + 0]);
          h[4 * q + 1] = fmaf(w, x.y, h[4 * q + 1]);
          h[4 * q + 2] = fmaf(w, x.z, h[4 * q + 2]);
          h[4 * q + 3] = fmaf(w, x.w, h[4 * q + 3]);
        }
      } else {
        static_assert(kEmoClips / 2 == 2, "clips per half: 2 or a multiple of 4");
        const float2 x = *reinterpret_cast<const float2*>(zr);
        h[0] = fmaf(w, x.x, h[0]);
        h[1] = fmaf(w, x.y, h[1]);
      }
    }
    __syncthreads();
    if (chunk + 2 < n1 + n2) issue(chunk + 2); else cp_async_commit();
  }
  {
    const float w2 = __ldg(W.w2 + j);
#pragma unroll
    for (int c = 0; c < kEmoClips / 2; ++c) {
      const float v = warp_sum(fmaxf(h[c], 0.0f) * w2);
      if (tx == 0) s_red[(half * (kEmoClips / 2) + c) * 8 + (ty & 3)] = v;
    }
  }
  __syncthreads();
  if (tid < nc) {
    const float logit = s_red[tid * 8] + s_red[tid * 8 + 1] + s_red[tid * 8 + 2] + s_red[tid * 8 + 3] + W.b2;
    expr_sigmoid[c0 + tid] = 1.0f / (1.0f + expf(-logit));
  }
  // this kernel reads nothing the frontend writes, so it may run beside the frontend's last CTAs; it still must not
  // complete before the frontend has: the core kernel waits on THIS kernel only
  pdl_wait();
}

// ---- EMA scan: y_t = alpha x_t + (1 - alpha) y_{t-1}, warp-parallel over 32 frames per step --------
constexpr int kEmaTile = 32;
constexpr int kEmaStride = KOE_N_BLENDSHAPES + 1;

// one frame per clip (the streaming step): y = alpha x + (1 - alpha) state, or x itself for the first frame -- the same
// roundings as the scan kernel below for n_out == 1
__global__ void __launch_bounds__(256) ema_step_kernel(float* __restrict__ frames, int n, float alpha,
                                                       float* __restrict__ state, int has_state) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float x = frames[i];
  const float y = (has_state && state != nullptr) ? fmaf(1.0f - alpha, state[i], alpha * x) : x;
  frames[i] = y;
  if (state != nullptr) state[i] = y;
}

__global__ void __launch_bounds__(256) ema_scan_kernel(float* __restrict__ frames, int n_out, float alpha,
                                                       float* __restrict__ state, int has_state) {
  __shared__ float s_t[kEmaTile * kEmaStride];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* x = frames + (size_t)b * n_out * KOE_N_BLENDSHAPES;
  const float a = 1.0f - alpha;
  float apow[5];  // a^1, a^2, a^4, a^8, a^16
  apow[0] = a;
#pragma unroll
  for (int i = 1; i < 5; ++i) apow[i] = apow[i - 1] * apow[i - 1];
  const float a_lane = powf(a, (float)(lane + 1));
  // warp w owns coefficients j = w, w + 8, ...; carry[j-slot] is y_{t0 - 1}
  float carry[7];
#pragma unroll
  for (int s = 0; s < 7; ++s) {
    const int j = warp + 8 * s;
    carry[s] = (has_state && state != nullptr && j < KOE_N_BLENDSHAPES) ? state[(size_t)b * KOE_N_BLENDSHAPES + j] : 0.0f;
  }
  for (int t0 = 0; t0 < n_out; t0 += kEmaTile) {
    const int nt = min(kEmaTile, n_out - t0);
    for (int i = tid; i < nt * KOE_N_BLENDSHAPES; i += 256)
      s_t[(i / KOE_N_BLENDSHAPES) * kEmaStride + i % KOE_N_BLENDSHAPES] = x[(size_t)t0 * KOE_N_BLENDSHAPES + i];
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const int j = warp + 8 * s;
      if (j < KOE_N_BLENDSHAPES) {
        float c = 0.0f;
        if (lane < nt) {
          const float v = s_t[lane * kEmaStride + j];
          // first frame of a fresh sequence passes through (reference :357-359)
          c = (t0 + lane == 0 && !has_state) ? v : alpha * v;
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const float up = __shfl_up_sync(kFullMask, c, 1 << i);
          if (lane >= (1 << i)) c = fmaf(apow[i], up, c);
        }
        const float y = fmaf(a_lane, carry[s], c);
        if (lane < nt) s_t[lane * kEmaStride + j] = y;
        carry[s] = __shfl_sync(kFullMask, y, nt - 1);
      }
    }
    __syncthreads();
    for (int i = tid; i < nt * KOE_N_BLENDSHAPES; i += 256)
      x[(size_t)t0 * KOE_N_BLENDSHAPES + i] = s_t[(i / KOE_N_BLENDSHAPES) * kEmaStride + i % KOE_N_BLENDSHAPES];
    __syncthreads();
  }
  if (state != nullptr && lane == 0) {
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const int j = warp + 8 * s;
      if (j < KOE_N_BLENDSHAPES) state[(size_t)b * KOE_N_BLENDSHAPES + j] = carry[s];
    }
  }
}

int launch_dual_stream_tc(const CoreParams& p, int precision, cudaStream_t stream);  // dual_stream_tc.cu

}  // namespace koe

using namespace koe;

static int validate_weights(const koe_core_weights* w) {
  KOE_REQUIRE(w != nullptr, "koe_core_weights is NULL");
  KOE_REQUIRE(w->wc_t && w->bc && w->ln_g && w->ln_b && w->qk_t && w->wv_t && w->bv && w->wa_t && w->ba && w->w2 &&
                  w->coef && w->mouth_idx && w->expr_idx && w->we1_t && w->be1 && w->eln_g && w->eln_b && w->we2_t &&
                  w->be2,
              "koe_core_weights has a NULL field");
  KOE_REQUIRE(w->k_mel >= 4 && w->k_mel <= 4096 && w->k_mel_pad >= w->k_mel && w->k_mel_pad % 16 == 0,
              "koe_core_weights: bad k_mel/k_mel_pad");
  KOE_REQUIRE(w->emo_in > 0 && w->emo_in <= kEmoInMax && w->emo_in_pad >= w->emo_in, "koe_core_weights: bad emo_in");
  return KOE_OK;
}

// `after_frontend`: the caller guarantees that the kernel queued just before this one on `stream` is the log-mel frontend,
// which writes nothing this kernel reads -- only then may it start beside that kernel's last CTAs (programmatic dependent
// launch).  From the public entry the producer of emo_in may be the previous kernel, so that launch keeps full stream order.
int koe::launch_emotion_stream(const koe_core_weights* w, const float* emo_in, int n_clips, float* expr_sigmoid,
                               void* stream, bool after_frontend) {
  if (int rc = validate_weights(w)) return rc;
  KOE_REQUIRE(n_clips >= 0, "koe_emotion_stream: negative size");
  if (n_clips == 0) return KOE_OK;  // nothing to do: empty buffers may be NULL
  KOE_REQUIRE(emo_in != nullptr && expr_sigmoid != nullptr, "koe_emotion_stream: NULL argument");
  static bool configured[64] = {false};
  int dev = 0;
  KOE_CUDA(cudaGetDevice(&dev));
  KOE_REQUIRE(dev >= 0 && dev < 64, "device index too large");
  if (!configured[dev]) {
    KOE_CUDA(cudaFuncSetAttribute(emotion_stream_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmoSmem));
    KOE_CUDA(cudaFuncSetAttribute(emotion_stream_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmoSmem));
    configured[dev] = true;
  }
  // every CTA streams all the weights (0.4 MB, L2 resident): few clips per CTA while that keeps the grid within ~4 waves
  const bool small = n_clips <= 4 * 600;
  const dim3 grid(small ? (n_clips + 3) / 4 : (n_clips + 15) / 16);
  if (after_frontend) {
    if (small)
      KOE_CUDA(launch_after_primary_starts(emotion_stream_kernel<4>, grid, dim3(256), kEmoSmem, (cudaStream_t)stream, *w,
                                           emo_in, n_clips, expr_sigmoid));
    else
      KOE_CUDA(launch_after_primary_starts(emotion_stream_kernel<16>, grid, dim3(256), kEmoSmem, (cudaStream_t)stream, *w,
                                           emo_in, n_clips, expr_sigmoid));
  } else if (small) {
    emotion_stream_kernel<4><<<grid, 256, kEmoSmem, (cudaStream_t)stream>>>(*w, emo_in, n_clips, expr_sigmoid);
  } else {
    emotion_stream_kernel<16><<<grid, 256, kEmoSmem, (cudaStream_t)stream>>>(*w, emo_in, n_clips, expr_sigmoid);
  }
  count_launch();
  KOE_CUDA(cudaGetLastError());
  return KOE_OK;
}

extern "C" int koe_emotion_stream(const koe_core_weights* w, const float* emo_in, int n_clips, float* expr_sigmoid,
                                  void* stream) {
  return koe::launch_emotion_stream(w, emo_in, n_clips, expr_sigmoid, stream, /*after_frontend=*/false);
}

// ---- y = x W^T + b for a handful of rows: the 264 -> 256 eGeMAPS compression as a stand-alone call
// (OpenSMILEeGeMAPSExtractor.get_concatenated_features, src/features/opensmile_extractor.py:583-604; the batch path folds
// this layer into the emotion stream instead)
namespace koe {
__global__ void __launch_bounds__(256) affine_rows_kernel(const float* __restrict__ x, int n_in,
                                                          const float* __restrict__ w_t, const float* __restrict__ b,
                                                          int n_out, float* __restrict__ y) {
  extern __shared__ float s_x[];
  const int r = blockIdx.x;
  for (int i = threadIdx.x; i < n_in; i += blockDim.x) s_x[i] = x[(long long)r * n_in + i];
  __syncthreads();
  for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
    float acc = 0.0f;
    for (int i = 0; i < n_in; ++i) acc = fmaf(s_x[i], w_t[(long long)i * n_out + o], acc);
    y[(long long)r * n_out + o] = acc + b[o];
  }
}
}  // namespace koe

extern "C" int koe_affine_rows(const float* x, int n_rows, int n_in, const float* w_t, const float* b, int n_out, float* y,
                               void* stream) {
  KOE_REQUIRE(n_rows >= 0 && n_in > 0 && n_in <= 8192 && n_out > 0, "koe_affine_rows: bad sizes");
  if (n_rows == 0) return KOE_OK;
  KOE_REQUIRE(x && w_t && b && y, "koe_affine_rows: NULL argument");
  koe::affine_rows_kernel<<<n_rows, 256, n_in * sizeof(float), (cudaStream_t)stream>>>(x, n_in, w_t, b, n_out, y);
  count_launch();
  KOE_CUDA(cudaGetLastError());
  return KOE_OK;
}

static long long* g_tc_debug = nullptr;
// bring-up / profiling hook (not in the public header; scripts/tc_timeline.py): a device buffer of 128 + 2 * gridDim
// 64-bit slots that the tensor-core kernel fills with phase timestamps of CTA 0 (clock64) and every CTA's start / end
// (globaltimer); NULL (the default) turns the stamps off
extern "C" void koe_debug_set_tc_timestamps(long long* device_buffer) { g_tc_debug = device_buffer; }

static int launch_core(const CoreParams& p_in, int precision, cudaStream_t stream) {
  CoreParams p = p_in;
  p.dbg = g_tc_debug;
  if (precision == 0) {
    static int num_sms[64] = {0};
    int dev = 0;
    KOE_CUDA(cudaGetDevice(&dev));
    KOE_REQUIRE(dev >= 0 && dev < 64, "device index too large");
    if (num_sms[dev] == 0) {
      KOE_CUDA(cudaFuncSetAttribute(dual_stream_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kCoreSmem));
      int n = 0;
      KOE_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
      num_sms[dev] = n;
    }
    const int grid = std::min(p.n_clips * p.n_out, num_sms[dev]);
    dual_stream_fp32_kernel<<<grid, kCoreThreads, kCoreSmem, stream>>>(p);
    count_launch();
    KOE_CUDA(cudaGetLastError());
    return KOE_OK;
  }
  if (precision == 2) return launch_dual_stream_tc(p, precision, stream);
  return fail(KOE_E_INVALID, "precision must be 0 (fp32) or 2 (bf16 operands, tcgen05), got %d", precision);
}

extern "C" int koe_dual_stream_windows(const koe_core_weights* w, const float* const* power,
                                       const float* const* frame_max, int n_edge, int n_clips, int n_frames,
                                       int n_out, int stride_frames, int frames_per_window,
                                       const float* expr_sigmoid, float* out, float* sigmoid_out, float* attn_out,
                                       int precision, void* stream) {
  if (int rc = validate_weights(w)) return rc;
  if (n_clips == 0 || n_out == 0) return KOE_OK;  // nothing to do: empty buffers may be NULL
  KOE_REQUIRE(power != nullptr && frame_max != nullptr && expr_sigmoid != nullptr && out != nullptr,
              "koe_dual_stream_windows: NULL argument");
  KOE_REQUIRE(n_edge >= 0 && n_edge <= KOE_MAX_EDGE, "koe_dual_stream_windows: n_edge out of range");
  KOE_REQUIRE(n_clips >= 0 && n_out >= 0 && stride_frames >= 1 && frames_per_window >= 1,
              "koe_dual_stream_windows: bad sizes");
  KOE_REQUIRE(frames_per_window > 2 * n_edge, "koe_dual_stream_windows: window shorter than its edge frames");
  KOE_REQUIRE(n_out == 0 || (long long)(n_out - 1) * stride_frames + frames_per_window <= n_frames,
              "koe_dual_stream_windows: windows run past n_frames");
  KOE_REQUIRE((long long)n_clips * n_out < (1ll << 31), "koe_dual_stream_windows: too many windows");
  if (n_clips == 0 || n_out == 0) return KOE_OK;
  CoreParams p{};
  p.w = *w;
  for (int j = 0; j < 1 + 2 * KOE_MAX_EDGE; ++j) {
    const bool used = j < 1 + 2 * n_edge;
    p.power[j] = used ? power[j] : nullptr;
    p.fmax[j] = used ? frame_max[j] : nullptr;
    KOE_REQUIRE(!used || (p.power[j] != nullptr && p.fmax[j] != nullptr), "koe_dual_stream_windows: NULL power[%d]", j);
    KOE_REQUIRE(!used || (reinterpret_cast<uintptr_t>(p.power[j]) & 15) == 0,
                "koe_dual_stream_windows: power[%d] must be 16-byte aligned", j);
  }
  p.n_edge = n_edge;
  p.n_clips = n_clips;
  p.n_frames = n_frames;
  p.n_out = n_out;
  p.stride_frames = stride_frames;
  p.frames_per_window = frames_per_window;
  p.mel_seq = w->k_mel - 3;
  p.expr_sigmoid = expr_sigmoid;
  p.out = out;
  p.sigmoid_out = sigmoid_out;
  p.attn_out = attn_out;
  return launch_core(p, precision, (cudaStream_t)stream);
}

extern "C" int koe_dual_stream_ring_edges(const koe_core_weights* w, const float* const* power,
                                          const float* const* frame_max, int n_edge, int n_streams, int ring_frames,
                                          int ring_base, int frames_per_window, const float* expr_sigmoid, float* out,
                                          float* sigmoid_out, float* attn_out, int precision, void* stream) {
  if (int rc = validate_weights(w)) return rc;
  KOE_REQUIRE(power != nullptr && frame_max != nullptr && expr_sigmoid != nullptr && out != nullptr,
              "koe_dual_stream_ring: NULL argument");
  KOE_REQUIRE(n_edge >= 1 && n_edge <= KOE_MAX_EDGE, "koe_dual_stream_ring: n_edge out of range");
  KOE_REQUIRE(n_streams >= 0 && ring_base >= 0 && frames_per_window > 2 * n_edge && ring_frames >= frames_per_window - 1,
              "koe_dual_stream_ring: bad sizes (the ring must hold every plain frame of a window)");
  if (n_streams == 0) return KOE_OK;
  CoreParams p{};
  p.w = *w;
  for (int j = 0; j < 1 + 2 * n_edge; ++j) {
    KOE_REQUIRE(power[j] != nullptr && frame_max[j] != nullptr, "koe_dual_stream_ring: NULL buffer %d", j);
    KOE_REQUIRE((reinterpret_cast<uintptr_t>(power[j]) & 15) == 0, "koe_dual_stream_ring: mel rows must be 16-byte aligned");
    p.power[j] = power[j], p.fmax[j] = frame_max[j];
  }
  p.n_edge = n_edge;
  p.n_clips = n_streams;
  p.n_frames = ring_frames;
  p.n_out = 1;
  p.stride_frames = 1;
  p.frames_per_window = frames_per_window;
  p.mel_seq = w->k_mel - 3;
  p.expr_sigmoid = expr_sigmoid;
  p.out = out;
  p.sigmoid_out = sigmoid_out;
  p.attn_out = attn_out;
  p.ring_frames = ring_frames;
  p.ring_base = ring_base % ring_frames;
  return launch_core(p, precision, (cudaStream_t)stream);
}

extern "C" int koe_dual_stream_ring(const koe_core_weights* w, const float* power_ring, const float* fmax_ring,
                                    const float* power_lo_ring, const float* fmax_lo_ring, const float* power_hi,
                                    const float* fmax_hi, int n_streams, int ring_frames, int ring_base,
                                    int frames_per_window, const float* expr_sigmoid, float* out, float* sigmoid_out,
                                    float* attn_out, int precision, void* stream) {
  const float* power[3] = {power_ring, power_lo_ring, power_hi};
  const float* fmax[3] = {fmax_ring, fmax_lo_ring, fmax_hi};
  return koe_dual_stream_ring_edges(w, power, fmax, 1, n_streams, ring_frames, ring_base, frames_per_window, expr_sigmoid,
                                    out, sigmoid_out, attn_out, precision, stream);
}

extern "C" int koe_dual_stream_features(const koe_core_weights* w, const float* mel_long, int n_long,
                                        const float* mel_short, int n_clips, const float* expr_sigmoid, float* out,
                                        float* sigmoid_out, float* attn_out, int precision, void* stream) {
  if (int rc = validate_weights(w)) return rc;
  KOE_REQUIRE(n_clips >= 0 && n_long >= 0, "koe_dual_stream_features: bad sizes");
  if (n_clips == 0) return KOE_OK;  // nothing to do: empty buffers may be NULL
  KOE_REQUIRE(mel_long != nullptr && mel_short != nullptr && expr_sigmoid != nullptr && out != nullptr,
              "koe_dual_stream_features: NULL argument");
  KOE_REQUIRE(((reinterpret_cast<uintptr_t>(mel_long) | reinterpret_cast<uintptr_t>(mel_short)) & 15) == 0,
              "koe_dual_stream_features: features must be 16-byte aligned");
  if (n_clips == 0) return KOE_OK;
  CoreParams p{};
  p.w = *w;
  p.n_clips = n_clips;
  p.n_out = 1;
  p.stride_frames = 1;
  p.frames_per_window = 1;
  p.mel_seq = w->k_mel - 3;
  p.expr_sigmoid = expr_sigmoid;
  p.out = out;
  p.sigmoid_out = sigmoid_out;
  p.attn_out = attn_out;
  p.mel_long = mel_long;
  p.mel_short = mel_short;
  p.n_long = n_long < p.mel_seq ? n_long : p.mel_seq;
  // rows of a longer mel_long are still addressed with its true length
  p.n_frames = n_long;
  return launch_core(p, precision, (cudaStream_t)stream);
}

extern "C" int koe_ema_scan(float* frames, int n_clips, int n_out, float alpha, float* state, int has_state,
                            void* stream) {
  KOE_REQUIRE(frames != nullptr && n_clips >= 0 && n_out >= 0, "koe_ema_scan: bad argument");
  KOE_REQUIRE(alpha >= 0.0f && alpha <= 1.0f, "koe_ema_scan: alpha must be in [0, 1]");
  KOE_REQUIRE(!has_state || state != nullptr, "koe_ema_scan: has_state set but state is NULL");
  if (n_clips == 0 || n_out == 0) return KOE_OK;
  if (n_out == 1)
    ema_step_kernel<<<(n_clips * KOE_N_BLENDSHAPES + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        frames, n_clips * KOE_N_BLENDSHAPES, alpha, state, has_state);
  else
    ema_scan_kernel<<<n_clips, 256, 0, (cudaStream_t)stream>>>(frames, n_out, alpha, state, has_state);
  count_launch();
  KOE_CUDA(cudaGetLastError());
  return KOE_OK;
}
