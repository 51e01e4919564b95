// Boundary/interior transient kernel (schur.cuh): model set CDN_MS_DIODE
// (one translation unit per instantiation: they compile in parallel).
#include "engine_launch.cuh"
#include "schur.cuh"

cudaError_t cdn_launch_diode_schur(const cdn_schur_args& A, long smem, cudaStream_t st) {
    return launch_schur<CDN_MS_DIODE>(A, smem, st);
}
