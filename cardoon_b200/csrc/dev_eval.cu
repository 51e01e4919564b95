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
    case CDN_MODEL_DIODE_T: diode_t_eval(p, flags, v, sink); break;
    case CDN_MODEL_BJT_T: bjt_t_eval<P, false>(p, flags, v, sink); break;
    case CDN_MODEL_SVBJT_T: bjt_t_eval<P, true>(p, flags, v, sink); break;
    case CDN_MODEL_MESFETC_T: mesfetc_t_eval(p, flags, v, sink); break;
    case CDN_MODEL_SVDIODE: svdiode_eval(p, flags, v, sink); break;
    case CDN_MODEL_SVDIODE_T: svdiode_t_eval(p, flags, v, sink); break;
    case CDN_MODEL_RES_T: res_t_eval(p, flags, v, sink); break;
    case CDN_MODEL_EKV_T: ekv_t_eval(p, flags, v, sink); break;
    default: break;
    }
}

// Ring-staged record accessor of the BSIM3 evaluation kernel.  The record
// (78 intrinsic rows, sorted by first use) is streamed through a shared-memory
// ring of R rows with per-thread 8-byte cp.async copies: every thread copies
// and reads only its own column, so no barrier is needed, and the copies of
// the rows two code sections ahead are in flight while the current section
// computes -- the HBM latency of the on-demand __ldg of each row (45 % of the
// stall samples of the on-demand kernel) is taken off the dependent chain
// without holding the rows in registers.  mark(first) is called by the model
// code at section boundaries: rows below `first` are dead and their slots are
// refilled.  All bookkeeping is compile-time after inlining (row numbers are
// literals); a row outside the resident window falls back to __ldg, so
// correctness does not depend on the placement of the marks.
template <int R, int NROWS, int TPB, int DEPTH, bool SYNC = false>
struct PVRing {
    const double* base;
    long ld;
    double* sm;            // this thread's column: row k at sm[(k % R) * TPB]
    int lo, ok, hi;        // rows [lo, ok) resident, [ok, hi) in flight
    int h1;                // `hi` before the previous mark's copies
    int jbase;             // first row of the trailing junction records
    unsigned jskip;        // bit j set: junction record j is not read (flags)
    __device__ __forceinline__ void issue(int upto) {
        const int h0 = hi;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int k = h0 + j;
            // (k is a literal after unrolling: the junction test is one
            // warp-uniform predicate per row)
            if (k < upto && k < NROWS
                && !(k >= jbase && ((jskip >> ((k - jbase) / CDN_NJ)) & 1u))) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + (k % R) * TPB);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst),
                             "l"(base + (long)k * ld) : "memory");
            }
        }
        // one commit group per mark (an empty group completes at once)
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (upto > hi) hi = (upto < NROWS) ? upto : NROWS;
    }
    // fill the ring and wait for it (kernel start)
    __device__ __forceinline__ void start() {
        lo = 0; hi = 0;
        issue(R);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        ok = h1 = hi;
    }
    // Rows below `first` are dead: refill their slots.  The copies issued by
    // the last DEPTH marks may still be in flight afterwards, so a row has
    // DEPTH code sections to arrive.
    __device__ __forceinline__ void mark(int first) {
        // SYNC: the warps of the CTA walk the model code section by section
        // together, so that they share the instruction cache lines
        if (SYNC) __syncthreads();
        const int before = hi;
        issue(first + R);
        if (DEPTH == 2) {          // all but the groups of this and the previous mark
            asm volatile("cp.async.wait_group 2;" ::: "memory");
            ok = h1;
        } else {                   // all but this mark's group
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            ok = before;
        }
        h1 = before;
        lo = first;
    }
    __device__ __forceinline__ double operator[](int k) const {
        if (k >= lo && k < ok) return sm[(k % R) * TPB];
        return __ldg(base + (long)k * ld);
    }
    // view of a sub-record (junction) that follows the ring's state
    struct Sub {
        const PVRing* r;
        int off;
        __device__ __forceinline__ double operator[](int k) const { return (*r)[off + k]; }
    };
    __device__ __forceinline__ Sub sub(int off) const { Sub v; v.r = this; v.off = off; return v; }
};

// BSIM3 value + Jacobian with the record streamed through the ring.
// TPB threads per CTA, MINB CTAs per SM (register limit), R ring rows.
template <int TPB, int MINB, int R, bool SYNC>
__global__ void __launch_bounds__(TPB, MINB)
bsim3_ring_kernel(long n, unsigned flags, const double* __restrict__ params, long ldp,
                  const int* __restrict__ pidx, const double* __restrict__ xin,
                  double* __restrict__ out, double* __restrict__ jac) {
    extern __shared__ double ring[];                  // [R][TPB]
    long i = (long)blockIdx.x * TPB + threadIdx.x;
    const bool live = i < n;
    if (!live) {
        if (!SYNC) return;
        i = n - 1;                 // keeps the barriers; nothing is stored
    }
    PVRing<R, NP_BSIM3, TPB, (R >= 48) ? 2 : 1, SYNC> pr;
    pr.base = params + (pidx ? (long)pidx[i] : i);
    pr.ld = ldp;
    pr.sm = ring + threadIdx.x;
    // junction records whose flag is off are never read: leave them in HBM
    pr.jbase = P_BSIM3_m + 1;
    pr.jskip = ((flags & F_MOSEXT_AD) ? 0u : 1u) | ((flags & F_MOSEXT_PD) ? 0u : 2u)
               | ((flags & F_MOSEXT_ASRC) ? 0u : 4u) | ((flags & F_MOSEXT_PS) ? 0u : 8u);
    pr.start();
    Dual<3> v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = make_var<3>(xin[(long)c * n + i], c);
    Dual<3> iv[3], qv[3];
    const double tf = pr[P_BSIM3_tf];
    bsim3_intrinsic(pr, flags, v, iv, qv);
    mosext_apply(pr, P_BSIM3_m, flags, tf, v, iv, qv);
    if (!live) return;
    GlobalSink<3> sink;
    sink.out = out; sink.jac = jac; sink.ld = n; sink.i = i; sink.nxin = 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) sink.put(k, iv[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) sink.put(3 + k, qv[k]);
}

template <int TPB, int MINB, int R, bool SYNC>
static void launch_ring(long n, unsigned flags, const double* params, long ldp,
                        const int* pidx, const double* xin, double* out, double* jac,
                        cudaStream_t st) {
    static bool configured = false;
    const int smem = R * TPB * (int)sizeof(double);
    if (!configured) {
        cudaFuncSetAttribute(bsim3_ring_kernel<TPB, MINB, R, SYNC>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    const long blocks = (n + TPB - 1) / TPB;
    bsim3_ring_kernel<TPB, MINB, R, SYNC><<<(unsigned)blocks, TPB, smem, st>>>(
        n, flags, params, ldp, pidx, xin, out, jac);
}

// ---- split records: model-card rows shared, instance rows per device ----------------
// Devices that share a `.model` card differ in a few record rows only (BSIM3:
// the ~25 rows that depend on w and l).  The caller splits the record once
// (cdn_dev_eval_split): `card[k]` holds row k of the rows all instances agree
// on, `inst[vrow[k] * ld + i]` the others.  The map row -> instance row travels
// in the kernel parameters (constant bank): at every access the row number is
// a literal, so the choice is one uniform-datapath test and both loads are
// predicated; card rows are warp-wide broadcasts served by L1.  HBM traffic per
// evaluation falls from ~1.1 KB to ~0.4 KB and the kernel is bound by the FP64
// pipe instead.
struct cdn_rowmap {
    short vrow[NP_BSIM3 + 1];      // -1: card row; else row of `inst`
};

// Accessor of the split kernel: both kinds of rows are read from shared
// memory -- the card rows as warp-wide broadcasts, the instance rows from the
// thread's own column, staged by one burst of cp.async at kernel start -- so
// that no global-memory latency sits on the model's dependent chain.
template <int TPB>
struct PVSplit {
    const double* smc;             // card rows [nrows] (shared)
    const double* smi;             // this thread's column: instance row v at smi[v * TPB]
    const cdn_rowmap* map;         // kernel parameter
    int off;
    __device__ __forceinline__ double operator[](int k) const {
        const int v = map->vrow[off + k];
        return (v >= 0) ? smi[v * TPB] : smc[off + k];
    }
    __device__ __forceinline__ PVSplit sub(int o) const { PVSplit r = *this; r.off = off + o; return r; }
    __device__ __forceinline__ void mark(int) {}
};

template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
bsim3_split_kernel(long n, unsigned flags, const double* __restrict__ card,
                   const double* __restrict__ inst, long ld, int nvar,
                   const __grid_constant__ cdn_rowmap map,
                   const double* __restrict__ xin, double* __restrict__ out,
                   double* __restrict__ jac) {
    extern __shared__ double sm[];                    // card[NP_BSIM3 + 1] | inst[nvar][TPB]
    long i = (long)blockIdx.x * TPB + threadIdx.x;
    const bool live = i < n;
    if (!live) i = n - 1;                             // (keeps the barrier; stores nothing)
    double* smi = sm + NP_BSIM3 + 1 + threadIdx.x;
    for (int v = 0; v < nvar; ++v) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(smi + v * TPB);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst),
                     "l"(inst + (long)v * ld + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int k = threadIdx.x; k < NP_BSIM3; k += TPB) sm[k] = __ldg(card + k);
    Dual<3> v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = make_var<3>(xin[(long)c * n + i], c);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    PVSplit<TPB> p;
    p.smc = sm; p.smi = smi; p.map = &map; p.off = 0;
    Dual<3> iv[3], qv[3];
    const double tf = p[P_BSIM3_tf];
    bsim3_intrinsic(p, flags, v, iv, qv);
    mosext_apply(p, P_BSIM3_m, flags, tf, v, iv, qv);
    if (!live) return;
    GlobalSink<3> sink;
    sink.out = out; sink.jac = jac; sink.ld = n; sink.i = i; sink.nxin = 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) sink.put(k, iv[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) sink.put(3 + k, qv[k]);
}

template <int TPB, int MINB>
static cudaError_t launch_split(long n, unsigned flags, const double* card, const double* inst,
                                long ld, int nvar, const cdn_rowmap& map, const double* xin,
                                double* out, double* jac, cudaStream_t st) {
    const int smem = (NP_BSIM3 + 1 + nvar * TPB) * (int)sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(bsim3_split_kernel<TPB, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem > 48 * 1024 ? smem : 48 * 1024);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)((n + TPB - 1) / TPB);
    bsim3_split_kernel<TPB, MINB><<<blocks, TPB, smem, st>>>(n, flags, card, inst, ld, nvar, map,
                                                              xin, out, jac);
    return cudaGetLastError();
}

// tuning (cdn_set_tuning key 6): CTA shape of the split kernel.  Measured on
// B200, 2 x 2M mosbsim3 evaluations (G evals/s): 128 thr x 2/SM 3.18, x 3 3.52,
// x 4 3.38; 256 x 2 4.07 (default), 256 x 3 3.25; 384 x 1 3.55; 512 x 1 3.74;
// 640 x 1 3.49; 768 x 1 3.13 -- 16 warps per SM at 128 registers, in CTAs
// large enough for their warps to share instruction fetches.
static int g_split_minb = 5;

extern "C" int cdn_dev_eval_split(cdn_ctx* ctx, int model_id, unsigned flags, long n,
                                  const double* card, int nrows, const short* vrow,
                                  const double* inst, long ld, const double* xin, double* out,
                                  double* jac, void* stream) {
    if (!ctx) return CDN_ERR_ARG;
    if (n <= 0) return CDN_OK;
    CDN_USE_DEVICE(ctx);
    if (model_id != CDN_MODEL_BSIM3)
        return cdn_fail(ctx, CDN_ERR_ARG, "cdn_dev_eval_split: only mosbsim3 / bsim3_i records are split");
    if (nrows != NP_BSIM3 || !card || !vrow || !xin || !out || !jac)
        return cdn_fail(ctx, CDN_ERR_ARG, "cdn_dev_eval_split: bad arguments");
    cdn_rowmap map;
    int nvar = 0;
    for (int k = 0; k < NP_BSIM3; ++k) {
        map.vrow[k] = vrow[k];
        if (vrow[k] >= nvar) nvar = vrow[k] + 1;
    }
    map.vrow[NP_BSIM3] = -1;
    if (nvar && !inst) return cdn_fail(ctx, CDN_ERR_ARG, "cdn_dev_eval_split: instance rows missing");
    if (nvar > 64) return cdn_fail(ctx, CDN_ERR_ARG, "cdn_dev_eval_split: more than 64 instance rows; use cdn_dev_eval");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    switch (g_split_minb) {
    case 2: e = launch_split<128, 2>(n, flags, card, inst, ld, nvar, map, xin, out, jac, st); break;
    case 4: e = launch_split<128, 4>(n, flags, card, inst, ld, nvar, map, xin, out, jac, st); break;
    case 5: e = launch_split<256, 2>(n, flags, card, inst, ld, nvar, map, xin, out, jac, st); break;
    case 6: e = launch_split<512, 1>(n, flags, card, inst, ld, nvar, map, xin, out, jac, st); break;
    case 7: e = launch_split<384, 1>(n, flags, card, inst, ld, nvar, map, xin, out, jac, st); break;
    case 8: e = launch_split<640, 1>(n, flags, card, inst, ld, nvar, map, xin, out, jac, st); break;
    case 9: e = launch_split<768, 1>(n, flags, card, inst, ld, nvar, map, xin, out, jac, st); break;
    case 10: e = launch_split<256, 3>(n, flags, card, inst, ld, nvar, map, xin, out, jac, st); break;
    case 3: e = launch_split<128, 3>(n, flags, card, inst, ld, nvar, map, xin, out, jac, st); break;
    default: e = launch_split<256, 2>(n, flags, card, inst, ld, nvar, map, xin, out, jac, st); break;
    }
    if (e != cudaSuccess) return cdn_fail(ctx, CDN_ERR_CUDA, std::string("cdn_dev_eval_split launch: ") + cudaGetErrorString(e));
    return CDN_OK;
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

static int g_npre = -1;          // tuning: record rows prefetched to L1 (-1 = default)
static int g_ring = 1;           // tuning: BSIM3 record through the shared-memory ring

extern "C" int cdn_tune_tile_tpb;     // engine.cu
extern "C" int cdn_tune_ac_blocks;    // ac.cu
extern "C" int cdn_tune_ac_tpb;

extern "C" int cdn_set_tuning(int key, int value) {
    if (key == 0) return CDN_OK;          // (retired: register-limit variants, see git history)
    if (key == 1) { g_npre = value; return CDN_OK; }
    if (key == 2) { g_ring = value; return CDN_OK; }
    if (key == 3) { cdn_tune_tile_tpb = value; return CDN_OK; }
    if (key == 4) { cdn_tune_ac_blocks = value; return CDN_OK; }
    if (key == 5) { cdn_tune_ac_tpb = value; return CDN_OK; }
    if (key == 6) { g_split_minb = value; return CDN_OK; }
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
    const int npre = (g_npre >= 0) ? g_npre : 0;
#define CDN_LAUNCH_MB(M) dev_eval_kernel<MODEL, P, M><<<(unsigned)blocks, tpb, 0, st>>>( \
        n, flags, params, ldp, pidx, xin, out, jac, nxin, npre)
    if constexpr (MODEL == CDN_MODEL_BSIM3 && P == 3) {
      if (g_ring && nxin == 3 && jac) {
#define CDN_RING(T, M, R, S) launch_ring<T, M, R, S>(n, flags, params, ldp, pidx, xin, out, jac, st)
        // tuning variants (bench.py --ring): measured on B200 3.44 / 3.48 /
        // 3.24 / 2.96 G evals/s -- more warps cost spills, a CTA-wide
        // lockstep costs more at the CTA ramps than it saves in fetches
        switch (g_ring) {
        case 2: CDN_RING(128, 3, 32, false); break;
        case 3: CDN_RING(128, 4, 48, false); break;
        case 4: CDN_RING(384, 1, 48, true); break;
        default: CDN_RING(128, 3, 48, false); break;
        }
#undef CDN_RING
      } else {
        CDN_LAUNCH_MB(4);
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
    case CDN_MODEL_DIODE_T:
        return lanes ? LAUNCH(CDN_MODEL_DIODE_T, 2) : LAUNCH(CDN_MODEL_DIODE_T, 0);
    case CDN_MODEL_BJT_T:
        return lanes ? LAUNCH(CDN_MODEL_BJT_T, 5) : LAUNCH(CDN_MODEL_BJT_T, 0);
    case CDN_MODEL_SVBJT_T:
        return lanes ? LAUNCH(CDN_MODEL_SVBJT_T, 5) : LAUNCH(CDN_MODEL_SVBJT_T, 0);
    case CDN_MODEL_MESFETC_T:
        return lanes ? LAUNCH(CDN_MODEL_MESFETC_T, 5) : LAUNCH(CDN_MODEL_MESFETC_T, 0);
    case CDN_MODEL_SVDIODE:
        return lanes ? LAUNCH(CDN_MODEL_SVDIODE, 1) : LAUNCH(CDN_MODEL_SVDIODE, 0);
    case CDN_MODEL_SVDIODE_T:
        return lanes ? LAUNCH(CDN_MODEL_SVDIODE_T, 2) : LAUNCH(CDN_MODEL_SVDIODE_T, 0);
    case CDN_MODEL_RES_T:
        return lanes ? LAUNCH(CDN_MODEL_RES_T, 2) : LAUNCH(CDN_MODEL_RES_T, 0);
    case CDN_MODEL_EKV_T:
        return lanes ? LAUNCH(CDN_MODEL_EKV_T, 4) : LAUNCH(CDN_MODEL_EKV_T, 0);
    default:
        return cudaErrorInvalidValue;
    }
}

extern "C" int cdn_dev_eval(cdn_ctx* ctx, int model_id, unsigned flags, long n, int nxin,
                            const double* params, long ldp, const int* pidx,
                            const double* xin, double* out, double* jac, void* stream) {
    if (!ctx) return CDN_ERR_ARG;
    if (n <= 0) return CDN_OK;
    CDN_USE_DEVICE(ctx);
    if (nxin < 1 || nxin > CDN_MAX_XIN) return cdn_fail(ctx, CDN_ERR_ARG, "cdn_dev_eval: bad nxin");
    cudaStream_t st = (cudaStream_t)stream;
    const int lanes = jac ? nxin : 0;
    if (model_id < 0 || model_id > CDN_MODEL_EKV_T)
        return cdn_fail(ctx, CDN_ERR_ARG, "cdn_dev_eval: unknown model id");
    const cudaError_t e = launch_model(model_id, lanes, n, flags, params, ldp, pidx, xin,
                                       out, jac, nxin, st);
    if (e != cudaSuccess) return cdn_fail(ctx, CDN_ERR_CUDA, std::string("cdn_dev_eval launch: ") + cudaGetErrorString(e));
    return CDN_OK;
}
