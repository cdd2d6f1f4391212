// me_tiled.cu -- tuned full-search kernel (placeholder until the first parity run).
#include "me_device.cuh"
namespace me {
struct TiledPlan { int unused; };
bool tiled_supported(const Geom &, size_t, size_t, const void *, const void *) { return false; }
cudaError_t tiled_plan_create(TiledPlan **plan, const Geom &, int) { *plan = nullptr; return cudaErrorNotSupported; }
void tiled_plan_destroy(TiledPlan *) {}
cudaError_t launch_tiled(TiledPlan *, const Geom &, const Frames &, int, const Out &, cudaStream_t, const char **) {
  return cudaErrorNotSupported;
}
}  // namespace me
