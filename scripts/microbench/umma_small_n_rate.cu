// tcgen05.mma issue/execute rate per shared-memory operand layout (sm_100a): cycles per M=128, K=16 bf16 MMA when both
// operands come from shared memory (SS form), for the canonical no-swizzle K-major core-matrix layout the dual-stream
// core uses and for the 128-byte-swizzled K-major layout, at N = 80 / 128 / 256.  Operand contents are zeros: only the
// time matters.  One CTA per SM, one issuing thread, R back-to-back MMAs into the same accumulator, one commit.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate scripts/microbench/umma_operand_layout_rate.cu && /tmp/umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Case {
  int n;                 // MMA N
  int layout;            // descriptor bits 61..63: 0 none, 2 swizzle-128B
  int a_lbo, a_sbo;      // bytes
  int b_lbo, b_sbo;
  int a_step, b_step;    // bytes added to the operand start per MMA (K advance), cycling over 4 positions
  int commit_every;      // 0: one commit at the end; k: a tcgen05.commit to a second mbarrier after every k MMAs (a ring's "slot free")
  int fence_every;       // k: tcgen05.fence::after_thread_sync + an mbarrier try_wait on a completed barrier before every k MMAs
};

__global__ void __launch_bounds__(128, 1) rate_kernel(Case c, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1048575;" ::"r"(smem_u32(&bar2)));  // never completes
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar3)) : "memory");  // phase 0 complete
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
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 100 * 1024;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto desc = [&](uint32_t addr, int lbo, int sbo) {
      return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
             ((uint64_t)c.layout << 61);
    };
    for (int pass = 0; pass < 2; ++pass) {  // pass 0 warms the instruction cache
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (c.fence_every && r % c.fence_every == 0) {
          uint32_t ok = 0;
          while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok)
                         : "r"(smem_u32(&bar3)), "r"(0u)
                         : "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint64_t a = desc(a0 + (r & 3) * c.a_step, c.a_lbo, c.a_sbo);
        const uint64_t b = desc(b0 + (r & 3) * c.b_step, c.b_lbo, c.b_sbo);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
            "l"(a), "l"(b), "r"(idesc), "r"(r)
            : "memory");
        if (c.commit_every && (r + 1) % c.commit_every == 0)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(&bar)), "r"((uint32_t)pass)
                     : "memory");
      const long long t1 = clock64();
      if (pass == 1 && blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Named { const char* name; Case c; };
  const Named cases[] = {
      {"none  compact   N=256  A lbo 2048 sbo 128  B lbo 4096 sbo 128", {256, 0, 2048, 128, 4096, 128, 4096, 8192, 0, 0}},
      {"none  compact   N=128  A lbo 2048 sbo 128  B lbo 2048 sbo 128", {128, 0, 2048, 128, 2048, 128, 4096, 4096, 0, 0}},
      {"none  compact   N=64   A lbo 2048 sbo 128  B lbo 1024 sbo 128", {64, 0, 2048, 128, 1024, 128, 4096, 2048, 0, 0}},
      {"none  compact   N=48   A lbo 2048 sbo 128  B lbo 768  sbo 128", {48, 0, 2048, 128, 768, 128, 4096, 1536, 0, 0}},
      {"none  compact   N=32   A lbo 2048 sbo 128  B lbo 512  sbo 128", {32, 0, 2048, 128, 512, 128, 4096, 1024, 0, 0}},
      {"none  compact   N=16   A lbo 2048 sbo 128  B lbo 256  sbo 128", {16, 0, 2048, 128, 256, 128, 4096, 512, 0, 0}},
      {"none  skewed    N=64   A lbo 2064 sbo 128  B lbo 1024 sbo 128", {64, 0, 2064, 128, 1024, 128, 4128, 2048, 0, 0}},
      {"none  skewed    N=32   A lbo 2064 sbo 128  B lbo 512  sbo 128", {32, 0, 2064, 128, 512, 128, 4128, 1024, 0, 0}},
      {"none  N=64   commit every 12                                  ", {64, 0, 2064, 128, 1024, 128, 4128, 2048, 12, 0}},
      {"none  N=32   commit every 6                                   ", {32, 0, 2048, 128, 512, 128, 4096, 1024, 6, 0}},
  };
  const int reps = 2048;
  for (const Named& nc : cases) {
    for (int grid : {148}) {
      cudaMemset(out, 0, 8);
      rate_kernel<<<grid, 128, 200 * 1024>>>(nc.c, reps, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
      const double per = (double)cyc / reps;
      printf("%s grid %3d: %7.1f cycles/MMA  (floor %3d, %4.0f flop/clk/SM)  %s\n", nc.name, grid, per, nc.c.n / 2,
             2.0 * 128 * nc.c.n * 16 / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
