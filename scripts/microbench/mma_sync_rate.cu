// Issue rate of the legacy warp-level tensor path (mma.sync, SASS HMMA) on sm_100a: tf32 m16n8k8 and bf16 m16n8k16,
// fp32 accumulation, 4 independent accumulator chains per warp.  Decides whether an error-compensated (3-term tf32 split)
// fp32-accurate core is worth building on mma.sync.     nvcc -arch=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void rate_kernel(float* out, int iters) {
  float d[4][4];
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f900000u, 0x3fa00000u, 0x3fb00000u}, b[2] = {0x3f800000u, 0x3f880000u};
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  float s = 0.f;
  for (int c = 0; c < 4; ++c)
    for (int i = 0; i < 4; ++i) s += d[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * prop.multiProcessorCount * 1024);
  const int iters = 20000;
  for (int kind = 0; kind < 2; ++kind)
    for (int warps : {4, 8, 16, 32}) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0), cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (kind == 0) rate_kernel<0><<<prop.multiProcessorCount, warps * 32>>>(out, iters);
        else rate_kernel<1><<<prop.multiProcessorCount, warps * 32>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flop_per_mma = kind == 0 ? 2.0 * 16 * 8 * 8 : 2.0 * 16 * 8 * 16;
      const double total = (double)prop.multiProcessorCount * warps * iters * 4 * flop_per_mma;
      printf("%s  %2d warps/SM: %.1f TFLOP/s, %.0f FLOP/clk/SM at %.0f MHz\n", kind == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16",
             warps, total / (ms * 1e-3) / 1e12, total / (ms * 1e-3) / prop.multiProcessorCount / (prop.clockRate * 1e3),
             prop.clockRate / 1e3);
    }
  return 0;
}
