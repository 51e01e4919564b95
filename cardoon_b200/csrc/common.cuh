// Shared declarations of libcardoon_b200.so (context, error reporting).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <cstdio>

struct cdn_ctx {
    int device;
    int sm_count;
    std::string err;
};

#include "cardoon_b200.h"

#define CDN_CUDA(ctx, call)                                                    \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);   \
            return CDN_ERR_CUDA;                                               \
        }                                                                      \
    } while (0)

// A context is bound to one GPU: every launching entry point selects it, so
// that several contexts (one per GPU) can live in one process.
#define CDN_USE_DEVICE(ctx)                                                    \
    do {                                                                       \
        if ((ctx) && (ctx)->device >= 0) {                                     \
            int cur_ = -1;                                                     \
            if (cudaGetDevice(&cur_) != cudaSuccess || cur_ != (ctx)->device)  \
                CDN_CUDA(ctx, cudaSetDevice((ctx)->device));                   \
        }                                                                      \
    } while (0)

static inline int cdn_fail(cdn_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}
