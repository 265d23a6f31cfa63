// DMoL register kernels for __nv_bfloat16 parameters (one translation unit per element type: parallel build).
#include "dmol_dispatch.cuh"

namespace blvm_host {
template int dmol_dispatch_tp<__nv_bfloat16>(const blvm::DmolArgs&, bool, int64_t, cudaStream_t);
template int sample_dispatch_tp<__nv_bfloat16>(const blvm::SampleArgs&, int64_t, cudaStream_t);
}  // namespace blvm_host
