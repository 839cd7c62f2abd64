// placeholder, replaced by the tcgen05 implicit-GEMM kernel
#include "common.cuh"
namespace b2 {
int conv3d_tcgen05_launch(const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t) {
    set_error("conv3d(tcgen05): not built yet");
    return B2_ERR_UNSUPPORTED;
}
}
