// BINNED particle store: particles kept binned by (species, cell), re-binned inside the push kernel every step.
#pragma once
#include "jic_device.cuh"
#include "jic_host.cuh"

namespace jic {

template <typename R>
struct BinnedStore {
  int create(Engine& e, const DevParams<R>&, const jic_params&, int) { return e.fail(JIC_ERR_UNSUPPORTED, "BINNED engine not built yet"); }
  void destroy() {}
  int start(Engine& e, const DevParams<R>&, const R*, const R*, R*, cudaStream_t) { return e.fail(JIC_ERR_UNSUPPORTED, "BINNED engine not built yet"); }
  int after_fields(Engine&, const DevParams<R>&, cudaStream_t) { return JIC_OK; }
  int step(Engine& e, const DevParams<R>&, const R*, R*, cudaStream_t) { return e.fail(JIC_ERR_UNSUPPORTED, "BINNED engine not built yet"); }
  int export_particles(Engine& e, const DevParams<R>&, R*, R*, uint8_t*, cudaStream_t) { return e.fail(JIC_ERR_UNSUPPORTED, "BINNED engine not built yet"); }
  int kinetic(Engine& e, const DevParams<R>&, double*, cudaStream_t) { return e.fail(JIC_ERR_UNSUPPORTED, "BINNED engine not built yet"); }
  long long extra_launches_per_step() const { return 0; }
};

}  // namespace jic
