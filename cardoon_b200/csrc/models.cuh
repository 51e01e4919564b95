// Model dispatch: evaluate one device instance of any model.
// `S` is an output sink with put(slot, Dual<P>) receiving out = [i..., q...]
// in the reference's order (cppaddev.create_tape :127: concatenate(i, q)).
#pragma once
#include "models/diode.cuh"
#include "models/bjt.cuh"
#include "models/mosext.cuh"
#include "models/bsim3.cuh"
#include "models/ekv.cuh"
#include "models/exprvm.cuh"

#define CDN_MAX_XIN 5
#define CDN_MAX_OUT 12
#define CDN_MAX_OUT_BUF 11          // base-model outputs of a `_t` device
#include "models/thermal.cuh"

template <int P, class S>
CDN_HD void mos_eval(int model, const PV& p0, unsigned flags, const Dual<P>* v, S& out) {
    Dual<P> iv[3], qv[3];
    PV p = p0;
    if (model == CDN_MODEL_BSIM3) {
        bsim3_intrinsic(p, flags, v, iv, qv);
        mosext_apply(p, P_BSIM3_m, flags, p[P_BSIM3_tf], v, iv, qv);
    } else {
        ekv_intrinsic(p, flags, v, iv, qv);
        mosext_apply(p, P_EKV_m, flags, p[P_EKV_tf], v, iv, qv);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) out.put(k, iv[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) out.put(3 + k, qv[k]);
}

// number of control voltages each model family is differentiated against
__host__ __device__ inline int model_lanes(int model) {
    switch (model) {
    case CDN_MODEL_DIODE: return 1;
    case CDN_MODEL_DIODE_T: case CDN_MODEL_SVDIODE_T: case CDN_MODEL_RES_T: return 2;
    case CDN_MODEL_SVDIODE: return 1;
    case CDN_MODEL_EKV_T: return 4;
    case CDN_MODEL_BJT_T: case CDN_MODEL_SVBJT_T: case CDN_MODEL_MESFETC_T: return 5;
    case CDN_MODEL_BSIM3: case CDN_MODEL_EKV: return 3;
    case CDN_MODEL_MEMR: return 2;
    default: return 4;
    }
}
