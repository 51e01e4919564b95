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
                    CDN_MS(CDN_MODEL_MESFETC) | CDN_MS(CDN_MODEL_DIODE_T) | \
                    CDN_MS(CDN_MODEL_BJT_T) | CDN_MS(CDN_MODEL_SVBJT_T) | \
                    CDN_MS(CDN_MODEL_MESFETC_T) | CDN_MS(CDN_MODEL_SVDIODE) | \
                    CDN_MS(CDN_MODEL_SVDIODE_T) | CDN_MS(CDN_MODEL_RES_T) | \
                    CDN_MS(CDN_MODEL_EKV_T))

template <class K>
static cudaError_t set_smem(K kernel, long smem) {
    if (smem > 48 * 1024)
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return cudaSuccess;
}

// One translation unit per (model set, kernel kind): the big kernels compile in
// parallel and a light circuit never pays for the heavy models.
enum { CDN_KIND_TILE8 = 0, CDN_KIND_TILE16, CDN_KIND_TILE32, CDN_KIND_BLOCK, CDN_KIND_GRID,
       CDN_NKIND };
enum { CDN_SET_DIODE = 0, CDN_SET_BSIM3, CDN_SET_ALL, CDN_NSET };

template <int MS, int T>
static cudaError_t launch_tile(const cdn_engine_args& A, const cdn_launch_cfg& c, cudaStream_t st) {
    cudaError_t e = set_smem(engine_tile_kernel<MS, T>, c.smem);
    if (e != cudaSuccess) return e;
    engine_tile_kernel<MS, T><<<c.blocks, c.tpb, c.smem, st>>>(A);
    return cudaGetLastError();
}

#ifndef CDN_BLOCK_TPB_LIGHT
#define CDN_BLOCK_TPB_LIGHT 1024
#endif
template <int MS, int KIND>
static cudaError_t launch_kind(const cdn_engine_args& A, const cdn_launch_cfg& c, cudaStream_t st) {
    // (if constexpr: a translation unit instantiates exactly one kernel)
    if constexpr (KIND == CDN_KIND_TILE8) {
        return launch_tile<MS, 8>(A, c, st);
    } else if constexpr (KIND == CDN_KIND_TILE16) {
        return launch_tile<MS, 16>(A, c, st);
    } else if constexpr (KIND == CDN_KIND_TILE32) {
        return launch_tile<MS, 32>(A, c, st);
    } else if constexpr (KIND == CDN_KIND_BLOCK) {
        // light model sets run 1024 threads per CTA (64 registers each); the
        // heavy models need their 255 registers
        constexpr int TPB = (MS == CDN_MS_DIODE) ? CDN_BLOCK_TPB_LIGHT : 256;
        cudaError_t e = set_smem(engine_block_kernel<MS, TPB>, c.smem);
        if (e != cudaSuccess) return e;
        engine_block_kernel<MS, TPB><<<c.blocks, TPB, c.smem, st>>>(A);
        return cudaGetLastError();
    } else {
        // cooperative grid
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
}

typedef cudaError_t (*cdn_launch_fn)(const cdn_engine_args&, const cdn_launch_cfg&, cudaStream_t);
#define CDN_DEFINE_LAUNCH(name, MS, KIND)                                                    \
    cudaError_t name(const cdn_engine_args& A, const cdn_launch_cfg& c, cudaStream_t st) {   \
        return launch_kind<MS, KIND>(A, c, st);                                              \
    }
#define CDN_DECLARE_LAUNCH(name) \
    cudaError_t name(const cdn_engine_args& A, const cdn_launch_cfg& c, cudaStream_t st);
CDN_DECLARE_LAUNCH(cdn_launch_diode_tile8)
CDN_DECLARE_LAUNCH(cdn_launch_diode_tile16)
CDN_DECLARE_LAUNCH(cdn_launch_diode_tile32)
CDN_DECLARE_LAUNCH(cdn_launch_diode_block)
CDN_DECLARE_LAUNCH(cdn_launch_diode_grid)
CDN_DECLARE_LAUNCH(cdn_launch_bsim3_tile8)
CDN_DECLARE_LAUNCH(cdn_launch_bsim3_tile16)
CDN_DECLARE_LAUNCH(cdn_launch_bsim3_tile32)
CDN_DECLARE_LAUNCH(cdn_launch_bsim3_block)
CDN_DECLARE_LAUNCH(cdn_launch_bsim3_grid)
CDN_DECLARE_LAUNCH(cdn_launch_all_tile8)
CDN_DECLARE_LAUNCH(cdn_launch_all_tile16)
CDN_DECLARE_LAUNCH(cdn_launch_all_tile32)
CDN_DECLARE_LAUNCH(cdn_launch_all_block)
CDN_DECLARE_LAUNCH(cdn_launch_all_grid)
