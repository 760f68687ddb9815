// Microbenchmark: issue rate of FFMA / FFMA2 / FADD2 / FMUL2 on sm_100a (packed fp32x2 vs scalar), alone and
// mixed with shared-memory loads, to decide how the FFT butterflies of logmel.cu should be packed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_rate fp32x2_rate.cu && ./fp32x2_rate
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
  u64 d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 d;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fadd1(float a, float b) {
  float d;
  asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

constexpr int kChains = 8;
constexpr int kInner = 64;

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float seed) {
  __shared__ float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float s[kChains];
  u64 p[kChains];
  float b = seed + 1.0f, c = seed * 0.5f;
  u64 pb, pc;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c));
#pragma unroll
  for (int k = 0; k < kChains; ++k) {
    s[k] = seed * (k + threadIdx.x);
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[k]) : "f"(s[k]), "f"(s[k] + 1.0f));
  }
  int idx = threadIdx.x & 31;
  int ix[kChains];
  const int r0 = __float_as_int(seed) | 5, r1 = (int)(seed * 1e6f) | 3;
#pragma unroll
  for (int k = 0; k < kChains; ++k) ix[k] = threadIdx.x + k;
  float lacc = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < kInner / kChains; ++j) {
#pragma unroll
      for (int k = 0; k < kChains; ++k) {
        if (MODE == 0) s[k] = ffma1(s[k], b, c);                 // FFMA 3-reg
        if (MODE == 1) p[k] = ffma2(p[k], pb, pc);               // FFMA2
        if (MODE == 2) p[k] = fadd2(p[k], pb);                   // FADD2
        if (MODE == 3) p[k] = fmul2(p[k], pb);                   // FMUL2
        if (MODE == 4) s[k] = fadd1(s[k], b);                    // FADD
        if (MODE == 5) { s[k] = ffma1(s[k], b, c); p[k] = ffma2(p[k], pb, pc); }   // 1 FFMA + 1 FFMA2
        if (MODE == 6) { p[k] = ffma2(p[k], pb, pc); lacc += sm[(idx + 32 * k + j) & 1023]; }  // FFMA2 + LDS + FADD
        if (MODE == 7) { s[k] = ffma1(s[k], b, c); lacc += sm[(idx + 32 * k + j) & 1023]; }    // FFMA + LDS + FADD
        if (MODE == 8) { p[k] = ffma2(p[k], pb, pc); ix[k] = (ix[k] ^ r0) + r1; }                     // FFMA2 + 2 ALU (LOP3, IADD3)
        if (MODE == 9) { s[k] = ffma1(s[k], b, c); ix[k] = (ix[k] ^ r0) + r1; }                       // FFMA + 2 ALU
        if (MODE == 10) { p[k] = ffma2(p[k], pb, pc); ix[k] = ix[k] ^ (ix[k] >> 3); }                  // FFMA2 + 1 ALU (LOP3 w/ shift = SHF+LOP3?)
        if (MODE == 11) { p[k] = ffma2(p[k], pb, pc); ix[k] = __vimax3_s32(ix[k], r0, r1) + 1; }       // FFMA2 + 1-2 ALU
        if (MODE == 12) { ix[k] = (ix[k] ^ r0) + r1; }                                                 // 2 ALU only
      }
    }
  }
  float r = lacc + idx;
#pragma unroll
  for (int k = 0; k < kChains; ++k) r += ix[k];
#pragma unroll
  for (int k = 0; k < kChains; ++k) {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[k]));
    r += s[k] + lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, int sms, float* d_out, double ops_per_inner) {
  const int iters = 2000, grid = sms * 8;
  bench<MODE><<<grid, 256>>>(d_out, 10, 1e-3f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<MODE><<<grid, 256>>>(d_out, iters, 1e-3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = (double)grid * 8 * iters * kInner;   // "primary" instructions per warp
  const double per_clk_smsp = warp_instr / (ms * 1e-3) / (sms * 4) / 1.965e9;
  printf("%-28s %8.3f ms  primary warp-instr/clk/SMSP (at 1965 MHz) = %.3f   fp32 lane-ops/clk/SM = %.1f\n", name, ms,
         per_clk_smsp, per_clk_smsp * 4 * 32 * ops_per_inner);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
  float* d_out;
  cudaMalloc(&d_out, sizeof(float) * prop.multiProcessorCount * 8 * 256);
  const int sms = prop.multiProcessorCount;
  run<0>("FFMA", sms, d_out, 1);
  run<1>("FFMA2", sms, d_out, 2);
  run<2>("FADD2", sms, d_out, 2);
  run<3>("FMUL2", sms, d_out, 2);
  run<4>("FADD", sms, d_out, 1);
  run<5>("FFMA + FFMA2", sms, d_out, 3);
  run<6>("FFMA2 + LDS + FADD", sms, d_out, 2);
  run<7>("FFMA + LDS + FADD", sms, d_out, 1);
  run<8>("FFMA2 + LOP3 + IADD3", sms, d_out, 2);
  run<9>("FFMA + LOP3 + IADD3", sms, d_out, 1);
  run<10>("FFMA2 + SHF/LOP3", sms, d_out, 2);
  run<11>("FFMA2 + VIMNMX3 + IADD", sms, d_out, 2);
  run<12>("LOP3 + IADD3 only", sms, d_out, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
