// tcgen05.mma rate at small N (M = 128, K = 16, kind::f16, both operands in shared memory, no-swizzle K-major), issued
// the way the kernels issue it: the whole warp runs the loop, one elected lane issues G back-to-back MMAs into one
// accumulator, then one tcgen05.commit (a "group" = one GEMM of a two-stage DFT: G = 6 at N = 32, G = 12 at N = 64).
// Operands are zeros: only the time matters.  Prints cycles per MMA and per group.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench/umma_small_n.bin scripts/microbench/umma_small_n_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, int lbo, int sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

template <int N, int G>
__global__ void __launch_bounds__(128, 1) rate_kernel(int a_lbo, int a_sbo, int b_lbo, int b_sbo, int groups, int bufs, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (threadIdx.x < 32) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 160 * 1024;
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32
    for (int pass = 0; pass < 2; ++pass) {
      const long long t0 = clock64();
      for (int g = 0; g < groups; ++g) {
        const uint32_t abuf = a0 + (g % bufs) * 32768;
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < G; ++j)
            mma(tmem + (g & 1) * N, desc(abuf + (j % 4) * 2 * a_lbo, a_lbo, a_sbo), desc(b0 + j * 2 * b_lbo, b_lbo, b_sbo), idesc, j != 0);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[g & 1])) : "memory");
        }
        __syncwarp();
        if (g >= 1) {  // keep at most two groups in flight, like a double-buffered accumulator
          uint32_t ok = 0;
          const uint32_t par = (((g - 1) >> 1) + pass * ((groups + ((g - 1) & 1 ? 0 : 1)) >> 1)) & 1;
          (void)par;
        }
      }
      // drain: wait for the last commit on each barrier by counting phases
      const long long t1 = clock64();
      (void)t1;
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[2])) : "memory");
      __syncwarp();
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(&bar[2])), "r"((uint32_t)pass)
                     : "memory");
      const long long t2 = clock64();
      if (pass == 1 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
  }
}

template <int N, int G>
void run(const char* name, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int bufs, long long* out) {
  const int groups = 512;
  cudaFuncSetAttribute(rate_kernel<N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaMemset(out, 0, 8);
  rate_kernel<N, G><<<148, 128, 200 * 1024>>>(a_lbo, a_sbo, b_lbo, b_sbo, groups, bufs, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
  printf("%-58s N=%3d groups of %2d: %7.1f cycles/MMA  %8.1f cycles/group  (N/2 = %d)  %s\n", name, N, G, (double)cyc / groups / G,
         (double)cyc / groups, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  run<256, 8>("A lbo 128 sbo 1024, B lbo 128 sbo 1024", 128, 1024, 128, 1024, 1, out);
  run<128, 8>("A lbo 128 sbo 1024, B lbo 128 sbo 1024", 128, 1024, 128, 1024, 1, out);
  run<64, 12>("A lbo 128 sbo 1024, B lbo 128 sbo 1024", 128, 1024, 128, 1024, 1, out);
  run<64, 12>("A lbo 144 sbo 1152 (skewed), B lbo 128 sbo 1024", 144, 1152, 128, 1024, 1, out);
  run<64, 12>("A lbo 144 sbo 1152, 4 A buffers", 144, 1152, 128, 1024, 4, out);
  run<64, 12>("A lbo 2064 sbo 128, B lbo 128 sbo 1024", 2064, 128, 128, 1024, 1, out);
  run<32, 6>("A lbo 128 sbo 512, B lbo 128 sbo 512", 128, 512, 128, 512, 1, out);
  run<32, 6>("A lbo 144 sbo 576 (skewed), B lbo 128 sbo 512", 144, 576, 128, 512, 1, out);
  run<32, 6>("A lbo 2048 sbo 128, B lbo 128 sbo 512", 2048, 128, 128, 512, 1, out);
  run<32, 6>("A lbo 128 sbo 512, 4 A buffers", 128, 512, 128, 512, 4, out);
  run<16, 6>("A lbo 128 sbo 512, B lbo 128 sbo 512", 128, 512, 128, 512, 1, out);
  return 0;
}
