// assemble_box.cu -- tiled write-once assembly for structured Hex8 blocks (placeholder until
// the tiled kernel lands: reports "unsupported" so AUTO mode uses the general kernel).
#include "internal.h"
namespace nalu_b200 {
bool box_supported(const Graph &, const nalu_b200_kernel_desc *, int) { return false; }
int launch_assemble_box(const Graph &, const nalu_b200_kernel_desc *, int, double *, double *, int, cudaStream_t)
{
  set_error("tiled kernel not built");
  return NALU_B200_ERR_ARG;
}
} // namespace nalu_b200
