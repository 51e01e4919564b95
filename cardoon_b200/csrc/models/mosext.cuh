// Extrinsic MOSFET additions: drain/source area + sidewall junctions and the
// parallel multiplier.  Reference: devices/mosExt.py eval_cqs :281-324
// (junction voltage is -v*tf; contributions are *subtracted* from idb/isb and
// qd/qs; everything times m).  `base` is the index of the 'm' entry of the
// model's parameter list; flag bits are the first 8 of the model's flags.
#pragma once
#include "junction.cuh"

#define F_MOSEXT_AD 1u
#define F_MOSEXT_CJ_D 2u
#define F_MOSEXT_PD 4u
#define F_MOSEXT_CJSW_D 8u
#define F_MOSEXT_ASRC 16u
#define F_MOSEXT_CJ_S 32u
#define F_MOSEXT_PS 64u
#define F_MOSEXT_CJSW_S 128u

// `A` is the record accessor (see bsim3.cuh); the marks sit outside the flag
// tests so that the ring-staged accessor's bookkeeping stays compile-time.
template <int P, class A>
CDN_HD void mosext_apply(A& p, int base, unsigned flags, double tf,
                         const Dual<P>* v, Dual<P>* iv, Dual<P>* qv) {
    p.mark(base);
    const double m = val(p[base]);
    const auto dj = p.sub(base + 1), djsw = p.sub(base + 1 + CDN_NJ);
    const auto sj = p.sub(base + 1 + 2 * CDN_NJ), sjsw = p.sub(base + 1 + 3 * CDN_NJ);
    const Dual<P> vdj = -v[0] * tf, vsj = -v[2] * tf;
    if (flags & F_MOSEXT_AD) {
        iv[1] = iv[1] - junction_id(dj, vdj) * tf;
        if (flags & F_MOSEXT_CJ_D) qv[0] = qv[0] - junction_qd(dj, vdj) * tf;
    }
    p.mark(base + 1 + CDN_NJ);
    if (flags & F_MOSEXT_PD) {
        iv[1] = iv[1] - junction_id(djsw, vdj) * tf;
        if (flags & F_MOSEXT_CJSW_D) qv[0] = qv[0] - junction_qd(djsw, vdj) * tf;
    }
    p.mark(base + 1 + 2 * CDN_NJ);
    if (flags & F_MOSEXT_ASRC) {
        iv[2] = iv[2] - junction_id(sj, vsj) * tf;
        if (flags & F_MOSEXT_CJ_S) qv[2] = qv[2] - junction_qd(sj, vsj) * tf;
    }
    p.mark(base + 1 + 3 * CDN_NJ);
    if (flags & F_MOSEXT_PS) {
        iv[2] = iv[2] - junction_id(sjsw, vsj) * tf;
        if (flags & F_MOSEXT_CJSW_S) qv[2] = qv[2] - junction_qd(sjsw, vsj) * tf;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { iv[k] = iv[k] * m; qv[k] = qv[k] * m; }
}
