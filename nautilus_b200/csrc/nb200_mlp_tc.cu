// tcgen05 (kind::tf32) emulator forward pass -- placeholder until the
// tensor-core kernel lands; fails loudly instead of falling back.
#include "nb200_common.cuh"
namespace nb200 {
int launch_mlp_tf32(const int32_t*, const int32_t*, const double*, int, int,
                    const double*, const uint8_t*, int64_t, double*, uint8_t*,
                    cudaStream_t) {
  return fail("nautilus_b200: %s", "NB200_MLP_TF32 is not built yet");
}
}  // namespace nb200
