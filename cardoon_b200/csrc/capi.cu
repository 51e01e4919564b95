// Context management and introspection entry points of libcardoon_b200.so.
#include "common.cuh"
#include "param_index.h"

extern "C" int cdn_version(void) { return 1; }

extern "C" int cdn_init(int device, cdn_ctx** out) {
    if (!out) return CDN_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return CDN_ERR_CUDA;
    if (device < 0 || device >= count) return CDN_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return CDN_ERR_CUDA;
    cdn_ctx* c = new cdn_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return CDN_ERR_CUDA; }
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return CDN_OK;
}

// context without a GPU: host-side symbolic analysis only (cdn_plan_create,
// cdn_plan_info, cdn_plan_array); every launch through it fails
extern "C" int cdn_init_host(cdn_ctx** out) {
    if (!out) return CDN_ERR_ARG;
    cdn_ctx* c = new cdn_ctx();
    c->device = -1;
    c->sm_count = 0;
    *out = c;
    return CDN_OK;
}

extern "C" int cdn_free(cdn_ctx* ctx) {
    delete ctx;
    return CDN_OK;
}

extern "C" const char* cdn_last_error(cdn_ctx* ctx) {
    return ctx ? ctx->err.c_str() : "no context";
}

extern "C" int cdn_sm_count(cdn_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

// number of packed parameters a model's kernel reads (memr: fixed head only)
extern "C" int cdn_model_nparams(int model_id) {
    switch (model_id) {
    case CDN_MODEL_DIODE: return NP_DIODE;
    case CDN_MODEL_BJT: case CDN_MODEL_SVBJT: return NP_BJT;
    case CDN_MODEL_BSIM3: return NP_BSIM3;
    case CDN_MODEL_EKV: return NP_EKV;
    case CDN_MODEL_MESFETC: return NP_MESFETC;
    case CDN_MODEL_MEMR: return NP_MEMR;
    case CDN_MODEL_DIODE_T: return NP_DIODE_T;
    case CDN_MODEL_BJT_T: case CDN_MODEL_SVBJT_T: return NP_BJT_T;
    case CDN_MODEL_MESFETC_T: return NP_MESFETC_T;
    case CDN_MODEL_SVDIODE: return NP_SVDIODE;
    case CDN_MODEL_SVDIODE_T: return NP_SVDIODE_T;
    case CDN_MODEL_RES_T: return NP_RES_T;
    case CDN_MODEL_EKV_T: return NP_EKV_T;
    default: return -1;
    }
}
