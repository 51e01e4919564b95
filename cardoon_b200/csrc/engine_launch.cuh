// Launch glue shared by the per-model-set translation units
// (engine_ms_*.cu): picks the kernel instantiation for a team kind.
#pragma once
#include "engine.cuh"

enum { CDN_TEAM_TILE = 0, CDN_TEAM_BLOCK = 1, CDN_TEAM_GRID = 2 };

struct cdn_launch_cfg {
    int team;        // CDN_TEAM_*
    int T;           // lanes per instance (tile team)
    int tpb;         // threads per block
    int blocks;      // grid size (0: derive)
    long smem;       // dynamic shared memory (tile team without global state)
    int sm_count;
};

#define CDN_MS_DIODE (CDN_MS(CDN_MODEL_DIODE))
#define CDN_MS_BSIM3 (CDN_MS(CDN_MODEL_BSIM3))
#define CDN_MS_ALL (CDN_MS(CDN_MODEL_DIODE) | CDN_MS(CDN_MODEL_BJT) | CDN_MS(CDN_MODEL_SVBJT) | \
                    CDN_MS(CDN_MODEL_BSIM3) | CDN_MS(CDN_MODEL_EKV) | CDN_MS(CDN_MODEL_MEMR) | \
                    CDN_MS(CDN_MODEL_MESFETC))

template <class K>
static cudaError_t set_smem(K kernel, long smem) {
    if (smem > 48 * 1024)
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return cudaSuccess;
}

template <int MS, int T>
static cudaError_t launch_tile(const cdn_engine_args& A, const cdn_launch_cfg& c, cudaStream_t st) {
    cudaError_t e = set_smem(engine_tile_kernel<MS, T>, c.smem);
    if (e != cudaSuccess) return e;
    engine_tile_kernel<MS, T><<<c.blocks, c.tpb, c.smem, st>>>(A);
    return cudaGetLastError();
}

template <int MS>
static cudaError_t launch_engine(const cdn_engine_args& A, cdn_launch_cfg c, cudaStream_t st) {
    switch (c.team) {
    case CDN_TEAM_TILE:
        switch (c.T) {
        case 4: return launch_tile<MS, 4>(A, c, st);
        case 8: return launch_tile<MS, 8>(A, c, st);
        case 16: return launch_tile<MS, 16>(A, c, st);
        case 32: return launch_tile<MS, 32>(A, c, st);
        default: return cudaErrorInvalidValue;
        }
    case CDN_TEAM_BLOCK: {
        // light model sets run 1024 threads per CTA (64 registers each); the
        // heavy models need their 255 registers
        if (MS == CDN_MS_DIODE && c.tpb == 1024) {
            cudaError_t e = set_smem(engine_block_kernel<MS, (MS == CDN_MS_DIODE ? 1024 : 256)>, c.smem);
            if (e != cudaSuccess) return e;
            engine_block_kernel<MS, (MS == CDN_MS_DIODE ? 1024 : 256)><<<c.blocks, 1024, c.smem, st>>>(A);
            return cudaGetLastError();
        }
        cudaError_t e = set_smem(engine_block_kernel<MS, 256>, c.smem);
        if (e != cudaSuccess) return e;
        engine_block_kernel<MS, 256><<<c.blocks, 256, c.smem, st>>>(A);
        return cudaGetLastError();
    }
    case CDN_TEAM_GRID: {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, engine_grid_kernel<MS, 256>, 256, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        int blocks = c.sm_count * per_sm;
        if (c.blocks > 0 && c.blocks < blocks) blocks = c.blocks;
        void* args[] = {(void*)&A};
        return cudaLaunchCooperativeKernel((void*)engine_grid_kernel<MS, 256>, dim3(blocks),
                                           dim3(256), args, 0, st);
    }
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t cdn_launch_ms_diode(const cdn_engine_args& A, const cdn_launch_cfg& c, cudaStream_t st);
cudaError_t cdn_launch_ms_bsim3(const cdn_engine_args& A, const cdn_launch_cfg& c, cudaStream_t st);
cudaError_t cdn_launch_ms_all(const cdn_engine_args& A, const cdn_launch_cfg& c, cudaStream_t st);
