// Batched nonlinear device evaluation (value + full Jacobian): one thread
// per device instance, structure-of-arrays parameters and port voltages.
//
// Replaces the reference's per-element `elem.eval_and_deriv(xin)` /
// `elem.eval(xin)` loop (analyses/nodalSP.py:858-885, devices/cppaddev.py
// :134-170).  Layouts (n = instances, all column-major over instances so
// that a warp reads 32 consecutive doubles):
//   params[k*ldp + pidx[i]]   k-th packed parameter of instance i
//   xin[c*n + i]              control voltage c
//   out[r*n + i]              output r of [currents..., charges...]
//   jac[(r*nxin + c)*n + i]   d out_r / d xin_c  (row-major per instance)
#include "common.cuh"
#include "models.cuh"

template <int P>
struct GlobalSink {
    double* out;
    double* jac;
    long ld;
    long i;
    int nxin;
    CDN_HD void put(int slot, const Dual<P>& d) const {
        out[(long)slot * ld + i] = d.v;
        if (P > 0 && jac) {
#pragma unroll
            for (int c = 0; c < P; ++c)
                if (c < nxin) jac[((long)slot * nxin + c) * ld + i] = d.d[c];
        }
    }
};

template <int P, class S>
CDN_HD void eval_any(int model, const PV& p, unsigned flags, const Dual<P>* v, S& sink) {
    switch (model) {
    case CDN_MODEL_DIODE: diode_eval(p, flags, v, sink); break;
    case CDN_MODEL_BJT: bjt_eval<P, false>(p, flags, v, sink); break;
    case CDN_MODEL_SVBJT: bjt_eval<P, true>(p, flags, v, sink); break;
    case CDN_MODEL_BSIM3: case CDN_MODEL_EKV: mos_eval(model, p, flags, v, sink); break;
    case CDN_MODEL_MESFETC: mesfetc_eval(p, flags, v, sink); break;
    case CDN_MODEL_MEMR: memr_eval(p, flags, v, sink); break;
    default: break;
    }
}

// MINB = resident CTAs per SM the register allocation is limited for
template <int MODEL, int P, int MINB = 1>
__global__ void __launch_bounds__(128, MINB)
dev_eval_kernel(long n, unsigned flags, const double* __restrict__ params, long ldp,
                const int* __restrict__ pidx, const double* __restrict__ xin,
                double* __restrict__ out, double* __restrict__ jac, int nxin, int npre) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PV p;
    p.base = params + (pidx ? (long)pidx[i] : i);
    p.ld = ldp;
    // the record is read on demand all through the model code: start moving
    // its first `npre` rows towards L1 now (one 256-byte segment per warp and
    // row), so that those reads are cache hits instead of HBM round trips
    for (int k = 0; k < npre; ++k)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p.base + (long)k * ldp));
    Dual<P> v[CDN_MAX_XIN];
#pragma unroll
    for (int c = 0; c < CDN_MAX_XIN; ++c)
        v[c] = (c < nxin) ? make_var<P>(xin[(long)c * n + i], c) : Dual<P>(0.0);
    GlobalSink<P> sink;
    sink.out = out; sink.jac = jac; sink.ld = n; sink.i = i; sink.nxin = nxin;
    eval_any<P>(MODEL, p, flags, v, sink);
}

static int g_minblocks = 0;      // tuning: 0 = default register allocation
static int g_npre = -1;          // tuning: record rows prefetched to L1 (-1 = default)

extern "C" int cdn_set_tuning(int key, int value) {
    if (key == 0) { g_minblocks = value; return CDN_OK; }
    if (key == 1) { g_npre = value; return CDN_OK; }
    return CDN_ERR_ARG;
}

template <int MODEL, int P>
static cudaError_t launch_one(long n, unsigned flags, const double* params, long ldp,
                              const int* pidx, const double* xin, double* out, double* jac,
                              int nxin, cudaStream_t st) {
    const int tpb = 128;
    const long blocks = (n + tpb - 1) / tpb;
    // BSIM3 with Jacobian: 228 registers unconstrained (8 warps/SM).  Limiting
    // the allocation to 128 registers (16 warps/SM, ~200 B of spills) measured
    // +50 % throughput on B200: the kernel is latency-bound, not pipe-bound.
    const int mb = (MODEL == CDN_MODEL_BSIM3 && P == 3) ? (g_minblocks ? g_minblocks : 4) : 1;
    const int npre = (g_npre >= 0) ? g_npre : 0;
#define CDN_LAUNCH_MB(M) dev_eval_kernel<MODEL, P, M><<<(unsigned)blocks, tpb, 0, st>>>( \
        n, flags, params, ldp, pidx, xin, out, jac, nxin, npre)
    if (MODEL == CDN_MODEL_BSIM3 && P == 3) {
        switch (mb) {
        case 2: CDN_LAUNCH_MB(2); break;
        case 3: CDN_LAUNCH_MB(3); break;
        case 5: CDN_LAUNCH_MB(5); break;
        case 6: CDN_LAUNCH_MB(6); break;
        default: CDN_LAUNCH_MB(4); break;
        }
    } else {
        CDN_LAUNCH_MB(1);
    }
#undef CDN_LAUNCH_MB
    return cudaGetLastError();
}

// lanes each model family is instantiated for (0 = values only)
#define LAUNCH(M, L) launch_one<M, L>(n, flags, params, ldp, pidx, xin, out, jac, nxin, st)

static cudaError_t launch_model(int model, int lanes, long n, unsigned flags,
                                const double* params, long ldp, const int* pidx,
                                const double* xin, double* out, double* jac, int nxin,
                                cudaStream_t st) {
    switch (model) {
    case CDN_MODEL_DIODE:
        return lanes ? LAUNCH(CDN_MODEL_DIODE, 1) : LAUNCH(CDN_MODEL_DIODE, 0);
    case CDN_MODEL_BJT:
        switch (lanes) {
        case 0: return LAUNCH(CDN_MODEL_BJT, 0);
        case 2: return LAUNCH(CDN_MODEL_BJT, 2);
        case 3: return LAUNCH(CDN_MODEL_BJT, 3);
        default: return LAUNCH(CDN_MODEL_BJT, 4);
        }
    case CDN_MODEL_SVBJT:
        switch (lanes) {
        case 0: return LAUNCH(CDN_MODEL_SVBJT, 0);
        case 2: return LAUNCH(CDN_MODEL_SVBJT, 2);
        case 3: return LAUNCH(CDN_MODEL_SVBJT, 3);
        default: return LAUNCH(CDN_MODEL_SVBJT, 4);
        }
    case CDN_MODEL_BSIM3:
        return lanes ? LAUNCH(CDN_MODEL_BSIM3, 3) : LAUNCH(CDN_MODEL_BSIM3, 0);
    case CDN_MODEL_EKV:
        return lanes ? LAUNCH(CDN_MODEL_EKV, 3) : LAUNCH(CDN_MODEL_EKV, 0);
    case CDN_MODEL_MESFETC:
        return lanes ? LAUNCH(CDN_MODEL_MESFETC, 4) : LAUNCH(CDN_MODEL_MESFETC, 0);
    case CDN_MODEL_MEMR:
        return lanes ? LAUNCH(CDN_MODEL_MEMR, 2) : LAUNCH(CDN_MODEL_MEMR, 0);
    default:
        return cudaErrorInvalidValue;
    }
}

extern "C" int cdn_dev_eval(cdn_ctx* ctx, int model_id, unsigned flags, long n, int nxin,
                            const double* params, long ldp, const int* pidx,
                            const double* xin, double* out, double* jac, void* stream) {
    if (!ctx) return CDN_ERR_ARG;
    if (n <= 0) return CDN_OK;
    if (nxin < 1 || nxin > CDN_MAX_XIN) return cdn_fail(ctx, CDN_ERR_ARG, "cdn_dev_eval: bad nxin");
    cudaStream_t st = (cudaStream_t)stream;
    const int lanes = jac ? nxin : 0;
    if (model_id < 0 || model_id > CDN_MODEL_MEMR)
        return cdn_fail(ctx, CDN_ERR_ARG, "cdn_dev_eval: unknown model id");
    const cudaError_t e = launch_model(model_id, lanes, n, flags, params, ldp, pidx, xin,
                                       out, jac, nxin, st);
    if (e != cudaSuccess) return cdn_fail(ctx, CDN_ERR_CUDA, std::string("cdn_dev_eval launch: ") + cudaGetErrorString(e));
    return CDN_OK;
}
