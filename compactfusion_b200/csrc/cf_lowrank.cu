// LOW_RANK projector (placeholder until the kernels land; see include/compactb200.h).
#include "cf_common.cuh"
namespace cf {
size_t lowrank_workspace_bytes(int64_t, int64_t, int) { return 256; }
}
extern "C" {
int cf_lowrank_project(const void*, const void*, const float*, void*, void*, float*, int64_t, int64_t, int, int, void*,
                       size_t, cf_stream_t) {
  cf::set_error("cf_lowrank_project: not implemented yet");
  return CF_ERR_UNSUPPORTED;
}
int cf_lowrank_reconstruct(const void*, const void*, const void*, void*, int64_t, int64_t, int, cf_stream_t) {
  cf::set_error("cf_lowrank_reconstruct: not implemented yet");
  return CF_ERR_UNSUPPORTED;
}
}
