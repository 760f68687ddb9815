// tcgen05 / TMEM implementation of the dual-stream core (precision 1 = tf32, 2 = bf16).
#include "common.cuh"

namespace koe {
struct CoreParams;
int launch_dual_stream_tc(const CoreParams& p, int precision, cudaStream_t stream) {
  (void)p;
  (void)stream;
  return fail(KOE_E_UNSUPPORTED, "koe_dual_stream_windows: precision %d (tensor-core path) is not built yet", precision);
}
}  // namespace koe
