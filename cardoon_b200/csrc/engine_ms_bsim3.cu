// Newton-engine kernels instantiated for the model set CDN_MS_BSIM3
// (one translation unit per set keeps register allocation and compile time
// of the light circuits independent of the heavy models).
#include "engine_launch.cuh"

cudaError_t cdn_launch_ms_bsim3(const cdn_engine_args& A, const cdn_launch_cfg& c, cudaStream_t st) {
    return launch_engine<CDN_MS_BSIM3>(A, c, st);
}
