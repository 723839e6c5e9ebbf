// batch_stable.cu — placeholder, replaced by the EllStable + LMI kernel.
#include "common.cuh"
namespace ellcuda {
int launch_batch_stable(const BatchArgs&, int, cudaStream_t, int, long long*) {
    set_error("EllStable/LMI batched kernel not built yet");
    return ELLCUDA_E_ARG;
}
}  // namespace ellcuda
