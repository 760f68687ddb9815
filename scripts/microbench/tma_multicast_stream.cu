// Does a 2-CTA cluster that shares a weight stream by TMA multicast get its stages faster than two CTAs that each load
// everything (the tcgen05 core streams 474 KB of weights per window and is bound by that)?  148 CTAs stream the same
// 512 KB buffer through a ring of 16 KB stages, `passes` times; consumers only wait for a stage and release it.
//   mode 0: every CTA loads every stage itself (unicast), CTAs started `skew` cycles apart to desynchronise them
//   mode 1: clusters of 2; CTA rank r issues the stages with (s & 1) == r, multicast to both CTAs; a slot is refilled
//           when BOTH CTAs have released it (remote mbarrier arrive)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench/tma_multicast.bin scripts/microbench/tma_multicast_stream.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

constexpr int kStage = 16384, kSlots = 4, kStages = 32;  // 512 KB per pass

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(64, 1) stream_kernel(const unsigned char* w, int passes, int skew, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[2 * kSlots];
  const uint32_t full = smem_u32(bars), empty = full + 8 * kSlots;
  const int tid = threadIdx.x;
  uint32_t rank = 0;
  if (MODE == 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (tid == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(full + 8 * i, 1);
      mbar_init(empty + 8 * i, MODE == 1 ? 2 : 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (MODE == 1) cg::this_cluster().sync();
  if (MODE == 0 && skew > 0) {  // desynchronise the CTAs
    const long long until = clock64() + (long long)(blockIdx.x % 16) * skew;
    while (clock64() < until) {}
  }
  const long long t0 = clock64();
  const int total = passes * kStages;
  if (tid == 0) {  // producer
    uint32_t slot = 0, phase = 0;
    for (int s = 0; s < total; ++s) {
      if (MODE == 0 || (uint32_t)(s & 1) == rank) {
        mbar_wait(empty + 8 * slot, phase ^ 1);
        if (MODE == 0) {
          mbar_expect_tx(full + 8 * slot, kStage);
          bulk_g2s(smem_u32(smem) + slot * kStage, w + (size_t)(s % kStages) * kStage, kStage, full + 8 * slot);
        } else {
          // both CTAs expect the bytes on their own barrier; the issuing CTA sends to both
          bulk_g2s_mc(smem_u32(smem) + slot * kStage, w + (size_t)(s % kStages) * kStage, kStage, full + 8 * slot, (uint16_t)3);
        }
      }
      if (++slot == kSlots) slot = 0, phase ^= 1;
    }
  } else if (tid == 32) {  // consumer
    uint32_t slot = 0, phase = 0;
    for (int s = 0; s < total; ++s) {
      if (MODE == 1) mbar_expect_tx(full + 8 * slot, kStage);  // arms this CTA's barrier for the multicast bytes
      mbar_wait(full + 8 * slot, phase);
      if (MODE == 0) {
        mbar_arrive_local(empty + 8 * slot);
      } else {  // release the slot in the CTA that will refill it: stage s + kSlots is issued by rank ((s + kSlots) & 1) == (s & 1)
        mbar_arrive_remote(empty + 8 * slot, (uint32_t)(s & 1));
      }
      if (++slot == kSlots) slot = 0, phase ^= 1;
    }
  }
  __syncthreads();
  if (MODE == 1) cg::this_cluster().sync();
  if (tid == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  unsigned char* w;
  long long* out;
  cudaMalloc(&w, kStages * kStage);
  cudaMemset(w, 1, kStages * kStage);
  cudaMalloc(&out, 148 * 8);
  const int passes = 64;
  auto report = [&](const char* name) {
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; double sum = 0;
    for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
    const double bytes = (double)passes * kStages * kStage;
    printf("%-44s mean %.0f cycles per 512 KB pass (%.1f B/clk/SM), slowest CTA %.0f  %s\n", name, sum / 148 / passes,
           bytes / (sum / 148), (double)mx / passes, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  cudaFuncSetAttribute(stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kStage);
  cudaFuncSetAttribute(stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kStage);
  for (int rep = 0; rep < 2; ++rep) {
    stream_kernel<0><<<148, 64, kSlots * kStage>>>(w, passes, 0, out);
    report("unicast, CTAs in step");
    stream_kernel<0><<<148, 64, kSlots * kStage>>>(w, passes, 700, out);
    report("unicast, CTAs 0..10.5 k cycles apart");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = kSlots * kStage;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, stream_kernel<1>, (const unsigned char*)w, passes, 0, out);
    report("clusters of 2, each CTA issues half, multicast");
  }
  return 0;
}
