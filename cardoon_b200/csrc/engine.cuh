// Device-resident Newton engine: device evaluation, gather assembly, LU,
// triangular solves, damped-Newton control, integration and the fixed-step
// time loop, all inside one kernel launch, written once over a "team" of
// cooperating threads:
//
//   TileTeam<T>  T (<= 32) lanes of a warp per circuit instance, state in
//                shared memory: batches of small circuits (Monte Carlo)
//   BlockTeam    one CTA per circuit instance, state in global memory (L2)
//   GridTeam     the whole cooperative grid on one circuit (large sparse)
//
// Reference functions restated here (cechrist/cardoon, file:line):
//   set_xin / set_i / set_Jac            analyses/nodal.py:44-80
//   get_i, get_i_Jac                     nodal.py:761-821, 1046-1133;
//                                        nodalSP.py:471-523, 804-887
//   factor_and_solve, get_chord_deltax   nodal.py:606-659, nodalSP.py:281-354
//   fsolve_Newton                        analyses/fsolve.py:59-128
//   update_q, accept, get_rhs, set_IC    nodalSP.py:666-802
//   BEuler / Trapezoidal                 analyses/integration.py:13-97
//   time loop incl. chord predictor      analyses/tran.py:171-206
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "plan.h"
#include "models.cuh"

namespace cg = cooperative_groups;

// Plan arrays are read with generic loads: the tile kernels keep a copy of the
// (small) plan of a dense circuit in shared memory, the other teams read the
// global copy.
#define CDN_LD(p) (*(p))

// ---- run description ---------------------------------------------------------------
// operation codes CDN_OP_*: include/cardoon_b200.h
struct cdn_opts {
    double abstol, reltol, maxdelta;
    double a0;        // integration coefficient (0 in DC)
    double gmin;      // gmin-stepping conductance (multiplies the Gones pattern)
    double gmin_ext;  // gmin on the diagonal of the first n_ext unknowns (nodal.py:469-489)
    double lambda;    // source-stepping factor (1 = full sources)
    double h;         // time step (delayed control ports)
    int maxiter;
    int errfunc;
    int use_q;        // 0: DC (currents only); 1: transient (charges too)
    int im;           // 0 backward Euler, 1 trapezoidal
    int n_ext;        // external terminals = first unknowns (0: no such helper)
};

struct cdn_engine_args {
    cdn_plan_dev plan;          // by value: kernel parameters live in the constant bank, so
                                // the plan's scalars and array pointers are uniform loads that
                                // never have to be re-read after a barrier
    cdn_opts o;
    int op;
    int B;                      // circuit instances in the batch
    double* x;                  // [B][n] in/out
    const double* sv_in;        // [n] source vector (shared by the batch)
    double* ws;                 // [B][ws_total] global state (NULL: shared memory)
    long ws_stride;
    // transient
    int nsamples, first, init_ic, nsrc, nsave;
    int helpers;                // 1: fall back to gmin / source stepping in-kernel
    int smem_lu;                // block team: 1 = LU values + solve work vector in shared
                                // memory, 2 = also a staging area for compact programs
    const int* src_rows;        // [nsrc] rows with a time-dependent source
    const double* src_val;      // [nsamples][nsrc] summed source value per row
    const double* src_dc;       // [n] constant part of the source vector
    const int* save_rows;       // [nsave]
    double* wave;               // [B][nsamples][nsave]
    int* iters;                 // [B][nsamples] (TRAN) or [B] (NEWTON)
    double* res;                // same shape as iters
    int* status;                // [B]: 0 ok, k > 0 first failed sample, -1 Newton failed
    double* red;                // GridTeam reduction scratch
    int plan_smem;              // tile kernels: bytes reserved for a shared-memory copy
                                // of the plan's gather lists (0: read the global copy)
    long long* prof;            // optional [16]: cycles of instance 0 per phase (device
                                // sweep, assembly, factor, solve, update, accept, chord)
    int* info;                  // optional [B][2]: highest helper used, pivot-trouble flag
};

// status bits or'ed into the negative range are avoided: singular pivots are
// reported as non-convergence (the NaN/inf step never passes the test).

// ---- per-instance state ----------------------------------------------------------
struct Ws {
    double *x, *dx, *ivec, *sv, *q, *qprev, *dq, *y, *tmp, *xb, *lu, *out, *jac, *hist;
    double* fact;   // small dense plans: factors kept for the chord step (sd8_store)
    unsigned sd_plan;  // shared-memory address of the CTA's Sd8Plan (0: generic path)
    unsigned sd_base;  // shared-memory address of this instance's state
    int* piv;       // [n] pivot rows (dense LU); piv[n] = pivot-trouble flag of the static-
                    // pivot sparse LU; piv[n + 1] = head of the history rings
    unsigned* stage;   // block team: shared-memory staging area for compact solve programs
    long long* dprof;  // optional fine-grained timers of the staged sweeps (slots 8..15)
};

// doubles of factor storage of a small dense plan (rows, reciprocal pivots, steps)
// (layout: plan.h)

__host__ __device__ inline long ws_doubles(int n, long nlu, long out_size, long jac_size,
                                           long hist_size = 0, long fact_size = 0) {
    return 10L * n + nlu + out_size + jac_size + hist_size + fact_size + (n + 1) / 2 + 2;
}
__host__ __device__ inline long ws_doubles(const cdn_plan_dev& P) {
    return ws_doubles(P.n, P.nlu, P.dev_out_size, P.dev_jac_size, P.hist_size,
                      P.sd ? CDN_SD8_FACT : 0);
}

__device__ inline Ws ws_carve(double* base, int n, long nlu, long out_size, long jac_size,
                              long hist_size = 0, long fact_size = 0) {
    Ws w;
    w.x = base; w.dx = w.x + n; w.ivec = w.dx + n; w.sv = w.ivec + n; w.q = w.sv + n;
    w.qprev = w.q + n; w.dq = w.qprev + n; w.y = w.dq + n; w.tmp = w.y + n; w.xb = w.tmp + n; w.lu = w.xb + n;
    w.out = w.lu + nlu; w.jac = w.out + out_size;
    w.hist = w.jac + jac_size;
    w.fact = w.hist + hist_size;
    w.piv = (int*)(w.fact + fact_size);
    w.stage = nullptr;
    w.dprof = nullptr;
    w.sd_plan = 0u;
    w.sd_base = 0u;
    return w;
}
__device__ inline Ws ws_carve(double* base, const cdn_plan_dev& P) {
    return ws_carve(base, P.n, P.nlu, P.dev_out_size, P.dev_jac_size, P.hist_size,
                    P.sd ? CDN_SD8_FACT : 0);
}

// ---- teams ---------------------------------------------------------------------------
template <int T>
struct TileTeam {
    cg::thread_block_tile<T> tile;
    static constexpr bool kFlat = true;      // flattened device list
    __device__ TileTeam() : tile(cg::tiled_partition<T>(cg::this_thread_block())) {}
    __device__ int rank() const { return tile.thread_rank(); }
    __device__ int size() const { return T; }
    __device__ void sync() const { tile.sync(); }
    __device__ double max(double v) const {
#pragma unroll
        for (int o = T / 2; o > 0; o >>= 1) v = fmax(v, tile.shfl_xor(v, o));
        return v;
    }
    __device__ double sum(double v) const {
#pragma unroll
        for (int o = T / 2; o > 0; o >>= 1) v += tile.shfl_xor(v, o);
        return v;
    }
    __device__ bool all(bool p) const { return tile.all(p); }
    // True if p holds for any tile of the CTA.  Called by every thread of the
    // CTA the same number of times: the tiles of a CTA iterate in lockstep so
    // that their warps run the (large) device-evaluation code at the same time
    // and share instruction-cache lines.
    __device__ bool cta_any(bool p) const { return __syncthreads_or(p) != 0; }
    // largest v, smallest index on ties; NaN never wins
    __device__ void argmax(double& v, int& i) const {
#pragma unroll
        for (int o = T / 2; o > 0; o >>= 1) {
            const double v2 = tile.shfl_xor(v, o);
            const int i2 = tile.shfl_xor(i, o);
            if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
        }
    }
};

struct BlockTeam {
    double* sred;      // shared scratch [64]
    int* ired;         // shared scratch [33]
    static constexpr bool kFlat = false;
    __device__ int rank() const { return threadIdx.x; }
    __device__ int size() const { return blockDim.x; }
    __device__ void sync() const { __syncthreads(); }
    __device__ double max(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if ((threadIdx.x & 31) == 0) sred[w] = v;
        __syncthreads();
        double r = sred[0];
        for (int k = 1; k < nw; ++k) r = fmax(r, sred[k]);
        __syncthreads();
        return r;
    }
    __device__ double sum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if ((threadIdx.x & 31) == 0) sred[w] = v;
        __syncthreads();
        double r = sred[0];
        for (int k = 1; k < nw; ++k) r += sred[k];
        __syncthreads();
        return r;
    }
    __device__ bool all(bool p) const { return __syncthreads_and(p) != 0; }
    __device__ bool cta_any(bool p) const { return p; }       // p is team-uniform
    __device__ void argmax(double& v, int& i) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
            if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
        }
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if ((threadIdx.x & 31) == 0) { sred[w] = v; ired[w] = i; }
        __syncthreads();
        v = sred[0]; i = ired[0];
        for (int k = 1; k < nw; ++k)
            if (sred[k] > v || (sred[k] == v && ired[k] < i)) { v = sred[k]; i = ired[k]; }
        __syncthreads();
    }
};

struct GridTeam {
    cg::grid_group grid;
    double* sred;      // shared scratch [64]
    double* gred;      // global scratch [2][gridDim.x] (+ flags)
    int phase;
    static constexpr bool kFlat = false;
    __device__ GridTeam() : grid(cg::this_grid()), phase(0) {}
    __device__ int rank() const { return blockIdx.x * blockDim.x + threadIdx.x; }
    __device__ int size() const { return gridDim.x * blockDim.x; }
    __device__ void sync() const { grid.sync(); }
    __device__ double max(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if ((threadIdx.x & 31) == 0) sred[w] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double r = sred[0];
            for (int k = 1; k < nw; ++k) r = fmax(r, sred[k]);
            gred[phase * gridDim.x + blockIdx.x] = r;
        }
        grid.sync();
        double r = -INFINITY;
        for (int k = threadIdx.x & 31; k < (int)gridDim.x; k += 32)
            r = fmax(r, __ldcg(gred + phase * gridDim.x + k));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
        phase ^= 1;
        return r;
    }
    __device__ double sum(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if ((threadIdx.x & 31) == 0) sred[w] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double r = sred[0];
            for (int k = 1; k < nw; ++k) r += sred[k];
            gred[phase * gridDim.x + blockIdx.x] = r;
        }
        grid.sync();
        double r = 0.0;
        for (int k = threadIdx.x & 31; k < (int)gridDim.x; k += 32)
            r += __ldcg(gred + phase * gridDim.x + k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        phase ^= 1;
        return r;
    }
    __device__ bool all(bool p) { return max(p ? 0.0 : 1.0) == 0.0; }
    __device__ bool cta_any(bool p) const { return p; }       // p is team-uniform
    __device__ void argmax(double& v, int& i) {
        // only used by the dense LU, which is never run grid-wide
        v = max(v);
    }
};

// ---- device evaluation ------------------------------------------------------------
template <int P>
struct WsSink {
    double* out;
    double* jac;
    int ld, i, nxin;
    int ndel;            // the last ndel control ports are delayed:
    double coef[2];      // d v(t - tau) / d v(t) of each (nodalSP.py:872-875)
    int ncs;             // outputs >= ncs are charges: their Jacobian rows are stored
    double qs;           // times a0, as the reference stamps them (qJac = a0 * Jac[ncs:],
                         // nodal.py:1117, nodalSP.py:878); the values stay unscaled
    CDN_HD void put(int slot, const Dual<P>& d) const {
        out[slot * ld + i] = d.v;
        if (P > 0) {
            const double s = (slot >= ncs) ? qs : 1.0;
#pragma unroll
            for (int c = 0; c < P; ++c)
                if (c < nxin) {
                    double t = d.d[c];
                    if (ndel > 0 && c >= nxin - ndel) t *= coef[c - (nxin - ndel)];
                    jac[(slot * nxin + c) * ld + i] = t * s;
                }
        }
    }
};

// v(t - td) by linear interpolation over the stored history of a port voltage
// (reference nodal.py:82-108 with a fixed step h).  ring[pos(back)] is the
// value `back` accepted samples ago (back = 0: last accepted sample); vp is the
// value at the time point being solved.  Returns the interpolated value and
// sets *coeff to its derivative with respect to vp.
__device__ __forceinline__ double delay_interp(double td, double vp, double h,
                                               const double* ring, int R, int head,
                                               double* coeff) {
    int p0 = head;
    if (td <= h) {
        const double c = (h - td) / h;
        *coeff = c;
        return ring[p0] + (vp - ring[p0]) * c;
    }
    *coeff = 0.0;
    double timeacc = h;
    for (int k = 0; k < R - 1; ++k) {
        timeacc += h;
        int p1 = p0 - 1;
        if (p1 < 0) p1 += R;
        if (td < timeacc) return ring[p1] + (ring[p0] - ring[p1]) * (timeacc - td) / h;
        p0 = p1;
    }
    return ring[p0];
}

template <int P>
__device__ __forceinline__ void gather_xin(const cdn_group_dev& G, int i, const double* x,
                                           Dual<P>* v) {
#pragma unroll
    for (int c = 0; c < CDN_MAX_XIN; ++c) {
        double xv = 0.0;
        if (c < G.nxin) {
            const int a = CDN_LD(G.vpos + c * G.n + i), b = CDN_LD(G.vneg + c * G.n + i);
            if (a >= 0) xv += x[a];
            if (b >= 0) xv -= x[b];
        }
        v[c] = make_var<P>(xv, c);
    }
}

template <int MODEL, int P>
__device__ __forceinline__ void eval_model(const cdn_group_dev& G, const PV& p, int i,
                                           const double* x, const Ws& w, int nhist, double h,
                                           int head, double qs) {
    Dual<P> v[CDN_MAX_XIN];
    gather_xin<P>(G, i, x, v);
    WsSink<P> sink;
    sink.out = w.out + G.out_off; sink.jac = w.jac + G.jac_off;
    sink.ld = G.n; sink.i = i; sink.nxin = G.nxin;
    sink.ncs = G.ncs; sink.qs = qs;
    sink.ndel = 0;
    if ((MODEL == CDN_MODEL_MESFETC || MODEL == CDN_MODEL_MESFETC_T) && nhist > 0 && G.ndelay > 0) {
        // delayed copies of control voltages (nodalSP.py:860-875)
        sink.ndel = G.ndelay;
#pragma unroll
        for (int d = 0; d < 2; ++d)
            if (d < G.ndelay) {
                const int c = G.nxin - G.ndelay + d;
                const double* ring = w.hist + G.hist_off + ((long)d * G.n + i) * nhist;
                v[c].v = delay_interp(CDN_LD(G.delay + d * G.n + i), v[c].v, h, ring, nhist, head,
                                      &sink.coef[d]);
            }
    }
    if (MODEL == CDN_MODEL_DIODE) diode_eval(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_BJT) bjt_eval<P, false>(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_SVBJT) bjt_eval<P, true>(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_BSIM3 || MODEL == CDN_MODEL_EKV) mos_eval(MODEL, p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_MESFETC) mesfetc_eval(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_MEMR) memr_eval(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_DIODE_T) diode_t_eval(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_BJT_T) bjt_t_eval<P, false>(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_SVBJT_T) bjt_t_eval<P, true>(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_MESFETC_T) mesfetc_t_eval(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_SVDIODE) svdiode_eval(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_SVDIODE_T) svdiode_t_eval(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_RES_T) res_t_eval(p, G.flags, v, sink);
    else if (MODEL == CDN_MODEL_EKV_T) ekv_t_eval(p, G.flags, v, sink);
}

#define CDN_MS(m) (1 << (m))
#define CDN_HAS(MS, m) (((MS) >> (m)) & 1)

// one device: instance i of group G of circuit instance b
template <int MS, bool DERIV>
__device__ __noinline__ void eval_one(const cdn_group_dev& G, int b, int i, const double* x,
                                      const Ws& w, int nhist, double h, int head, double qs) {
    long col = (G.per_inst ? (long)b * G.n : 0L) + i;
    if (G.pidx) col = CDN_LD(G.pidx + col);
    PV p;
    p.base = G.params + col;
    p.ld = G.ldp;
    switch (G.model) {
    case CDN_MODEL_DIODE:
        if (CDN_HAS(MS, CDN_MODEL_DIODE)) eval_model<CDN_MODEL_DIODE, DERIV ? 1 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_BJT:
        if (CDN_HAS(MS, CDN_MODEL_BJT)) eval_model<CDN_MODEL_BJT, DERIV ? 4 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_SVBJT:
        if (CDN_HAS(MS, CDN_MODEL_SVBJT)) eval_model<CDN_MODEL_SVBJT, DERIV ? 4 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_BSIM3:
        if (CDN_HAS(MS, CDN_MODEL_BSIM3)) eval_model<CDN_MODEL_BSIM3, DERIV ? 3 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_EKV:
        if (CDN_HAS(MS, CDN_MODEL_EKV)) eval_model<CDN_MODEL_EKV, DERIV ? 3 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_MEMR:
        if (CDN_HAS(MS, CDN_MODEL_MEMR)) eval_model<CDN_MODEL_MEMR, DERIV ? 2 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_DIODE_T:
        if (CDN_HAS(MS, CDN_MODEL_DIODE_T))
            eval_model<CDN_MODEL_DIODE_T, DERIV ? 2 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_BJT_T:
        if (CDN_HAS(MS, CDN_MODEL_BJT_T))
            eval_model<CDN_MODEL_BJT_T, DERIV ? 5 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_SVBJT_T:
        if (CDN_HAS(MS, CDN_MODEL_SVBJT_T))
            eval_model<CDN_MODEL_SVBJT_T, DERIV ? 5 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_MESFETC_T:
        if (CDN_HAS(MS, CDN_MODEL_MESFETC_T))
            eval_model<CDN_MODEL_MESFETC_T, DERIV ? 5 : 0>(G, p, i, x, w, nhist, h, head, qs);
        break;
    case CDN_MODEL_SVDIODE:
        if (CDN_HAS(MS, CDN_MODEL_SVDIODE))
            eval_model<CDN_MODEL_SVDIODE, DERIV ? 1 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_SVDIODE_T:
        if (CDN_HAS(MS, CDN_MODEL_SVDIODE_T))
            eval_model<CDN_MODEL_SVDIODE_T, DERIV ? 2 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_RES_T:
        if (CDN_HAS(MS, CDN_MODEL_RES_T))
            eval_model<CDN_MODEL_RES_T, DERIV ? 2 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_EKV_T:
        if (CDN_HAS(MS, CDN_MODEL_EKV_T))
            eval_model<CDN_MODEL_EKV_T, DERIV ? 4 : 0>(G, p, i, x, w, 0, h, 0, qs);
        break;
    case CDN_MODEL_MESFETC:
        if (CDN_HAS(MS, CDN_MODEL_MESFETC))
            eval_model<CDN_MODEL_MESFETC, DERIV ? 4 : 0>(G, p, i, x, w, nhist, h, head, qs);
        break;
    default: break;
    }
}

// `delayed`: evaluate the delayed control ports at t - tau (Newton, get_i);
// update_q evaluates them at the present voltages (nodalSP.py:719-733)
template <int MS, bool DERIV, class Team>
__device__ __forceinline__ void eval_devices(Team& tm, const cdn_plan_dev& P, int b,
                                             const double* x, const Ws& w, bool delayed = false,
                                             double h = 0.0, double qs = 1.0) {
    const int nhist = delayed ? P.nhist : 0;
    const int head = nhist ? w.piv[P.n + 1] : 0;
    if (Team::kFlat) {
        for (int d = tm.rank(); d < P.ndev; d += tm.size())
            eval_one<MS, DERIV>(P.groups[CDN_LD(P.dev_g + d)], b, CDN_LD(P.dev_i + d), x, w, nhist,
                                h, head, qs);
    } else {
        for (int g = 0; g < P.ngroups; ++g) {
            const cdn_group_dev& G = P.groups[g];
            for (int i = tm.rank(); i < G.n; i += tm.size())
                eval_one<MS, DERIV>(G, b, i, x, w, nhist, h, head, qs);
        }
    }
    tm.sync();
}

// Port-voltage history of the delayed control ports (nodalSP.py:686-696,
// 741-758): a ring of P.nhist accepted samples per port, head in piv[n + 1].
template <class Team>
__device__ __forceinline__ void hist_store(Team& tm, const cdn_plan_dev& P, const Ws& w,
                                           bool fill) {
    if (P.nhist <= 0) return;
    const int R = P.nhist;
    int head = fill ? 0 : w.piv[P.n + 1] + 1;
    if (head >= R) head = 0;
    for (int g = 0; g < P.ngroups; ++g) {
        const cdn_group_dev& G = P.groups[g];
        if (G.ndelay <= 0) continue;
        for (int t = tm.rank(); t < G.ndelay * G.n; t += tm.size()) {
            const int d = t / G.n, i = t - d * G.n;
            const int c = G.nxin - G.ndelay + d;
            const int a = CDN_LD(G.vpos + c * G.n + i), bb = CDN_LD(G.vneg + c * G.n + i);
            double v = 0.0;
            if (a >= 0) v += w.x[a];
            if (bb >= 0) v -= w.x[bb];
            double* ring = w.hist + G.hist_off + ((long)d * G.n + i) * R;
            if (fill) for (int k = 0; k < R; ++k) ring[k] = v;
            else ring[head] = v;
        }
    }
    tm.sync();
    if (tm.rank() == 0) w.piv[P.n + 1] = head;
    tm.sync();
}

// ---- assembly (gather; no atomics) ------------------------------------------------------
// coefficient of a packed gather entry: +-1 for a current row, +-a0 for a
// charge row (DC plans hold no charge entries)
__device__ __forceinline__ double pk_coef(unsigned pk, double a0) {
    const double c = (pk & (1u << 30)) ? a0 : 1.0;
    return (pk & (1u << 29)) ? -c : c;
}

// iVec = G'x + Ti i(x) + a0 Tq q(x) (+ gmin Gones x);  y = lambda*sv - iVec
// Rows / slots with more than CDN_HEAVY gather terms (hub nodes such as a
// supply rail) are summed by the whole team, everything else by one thread.
template <class Team>
__device__ __forceinline__ double row_sum(const cdn_plan_dev& P, const Ws& w, const cdn_opts& o,
                                          double a0, int r, int first, int stride) {
    double acc = 0.0;
    const int l0 = CDN_LD(P.r_lin_ptr + r), l1 = CDN_LD(P.r_lin_ptr + r + 1);
    if (stride == 1) {
        // groups of four entries with independent loads (short rows are the
        // rule: the index -> coefficient -> x chain of one entry is three
        // dependent loads, and four of them overlap); lanes past the end of
        // the row re-read its first entry with a zero coefficient
        double b1 = 0.0, b2 = 0.0, b3 = 0.0;
        for (int k = l0; k < l1; k += 4) {
            const int k1 = (k + 1 < l1) ? k + 1 : l0, k2 = (k + 2 < l1) ? k + 2 : l0;
            const int k3 = (k + 3 < l1) ? k + 3 : l0;
            double g0 = CDN_LD(P.r_lin_g + k), g1 = CDN_LD(P.r_lin_g + k1);
            double g2 = CDN_LD(P.r_lin_g + k2), g3 = CDN_LD(P.r_lin_g + k3);
            if (o.use_q) {
                g0 += a0 * CDN_LD(P.r_lin_c + k); g1 += a0 * CDN_LD(P.r_lin_c + k1);
                g2 += a0 * CDN_LD(P.r_lin_c + k2); g3 += a0 * CDN_LD(P.r_lin_c + k3);
            }
            if (o.gmin != 0.0) {
                g0 += o.gmin * CDN_LD(P.r_lin_gones + k); g1 += o.gmin * CDN_LD(P.r_lin_gones + k1);
                g2 += o.gmin * CDN_LD(P.r_lin_gones + k2); g3 += o.gmin * CDN_LD(P.r_lin_gones + k3);
            }
            const double x0 = w.x[CDN_LD(P.r_lin_col + k)], x1 = w.x[CDN_LD(P.r_lin_col + k1)];
            const double x2 = w.x[CDN_LD(P.r_lin_col + k2)], x3 = w.x[CDN_LD(P.r_lin_col + k3)];
            acc = fma(g0, x0, acc);
            b1 = fma((k + 1 < l1) ? g1 : 0.0, x1, b1);
            b2 = fma((k + 2 < l1) ? g2 : 0.0, x2, b2);
            b3 = fma((k + 3 < l1) ? g3 : 0.0, x3, b3);
        }
        acc += (b1 + b2) + b3;
    } else {
        for (int k = l0 + first; k < l1; k += stride) {
            double g = CDN_LD(P.r_lin_g + k);
            if (o.use_q) g += a0 * CDN_LD(P.r_lin_c + k);
            if (o.gmin != 0.0) g += o.gmin * CDN_LD(P.r_lin_gones + k);
            acc += g * w.x[CDN_LD(P.r_lin_col + k)];
        }
    }
    // device outputs: branch-free, the sign and the a0 factor of charge rows
    // are decoded from the packed entry (DC plans hold no charge entries)
    const int c0 = CDN_LD(P.r_con_ptr + r), c1 = CDN_LD(P.r_con_ptr + r + 1);
    int k = c0 + first;
    if (stride == 1) {
        double a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (; k + 3 < c1; k += 4) {
            const unsigned p0 = CDN_LD(P.r_con_pk + k), p1 = CDN_LD(P.r_con_pk + k + 1);
            const unsigned p2 = CDN_LD(P.r_con_pk + k + 2), p3 = CDN_LD(P.r_con_pk + k + 3);
            const double o0 = w.out[p0 & 0x1fffffffu], o1 = w.out[p1 & 0x1fffffffu];
            const double o2 = w.out[p2 & 0x1fffffffu], o3 = w.out[p3 & 0x1fffffffu];
            acc = fma(pk_coef(p0, a0), o0, acc);
            a1 = fma(pk_coef(p1, a0), o1, a1);
            a2 = fma(pk_coef(p2, a0), o2, a2);
            a3 = fma(pk_coef(p3, a0), o3, a3);
        }
        acc += (a1 + a2) + a3;
    }
    for (; k < c1; k += stride) {
        const unsigned pk = CDN_LD(P.r_con_pk + k);
        acc = fma(pk_coef(pk, a0), w.out[pk & 0x1fffffffu], acc);
    }
    return acc;
}

__device__ __forceinline__ bool row_heavy(const cdn_plan_dev& P, int r) {
    return (CDN_LD(P.r_lin_ptr + r + 1) - CDN_LD(P.r_lin_ptr + r)) +
           (CDN_LD(P.r_con_ptr + r + 1) - CDN_LD(P.r_con_ptr + r)) > P.heavy;
}

// Returns whether the rows this thread wrote are finite (dense plans: input of
// the non-finite test of newton()).
template <class Team>
__device__ __forceinline__ bool residual(Team& tm, const cdn_plan_dev& P, const Ws& w,
                                         const cdn_opts& o) {
    const double a0 = o.use_q ? o.a0 : 0.0;
    bool fin = true;
    for (int r = tm.rank(); r < P.n; r += tm.size()) {
        if (P.n_r_heavy && row_heavy(P, r)) continue;
        double acc = row_sum<Team>(P, w, o, a0, r, 0, 1);
        if (o.gmin_ext != 0.0 && r < o.n_ext) acc += o.gmin_ext * w.x[r];   // nodal.py:478
        w.ivec[r] = acc;
        const double y = o.lambda * w.sv[r] - acc;
        w.y[r] = y;
        fin = fin && isfinite(y);
    }
    for (int h = 0; h < P.n_r_heavy; ++h) {
        const int r = CDN_LD(P.r_heavy + h);
        double acc = tm.sum(row_sum<Team>(P, w, o, a0, r, tm.rank(), tm.size()));
        if (tm.rank() == 0) {
            if (o.gmin_ext != 0.0 && r < o.n_ext) acc += o.gmin_ext * w.x[r];
            w.ivec[r] = acc;
            w.y[r] = o.lambda * w.sv[r] - acc;
        }
    }
    return fin;
}

// Jacobian slots: G' + Ti J Tcv + a0 Tq Jq Tcv (+ gmin Gones).  The charge rows of
// the device Jacobians arrive already multiplied by a0 (WsSink), so every
// contribution is +-1 times a stored value.
__device__ __forceinline__ double pk_sign(unsigned pk) { return (pk & (1u << 29)) ? -1.0 : 1.0; }

template <class Team>
__device__ __forceinline__ double slot_sum(const cdn_plan_dev& P, const Ws& w, const cdn_opts& o,
                                           double a0, int e, int first, int stride) {
    double v = 0.0;
    const int c0 = CDN_LD(P.e_con_ptr + e), c1 = CDN_LD(P.e_con_ptr + e + 1);
    int k = c0 + first;
    if (stride == 1) {
        // four independent accumulators: the loads of a group overlap
        double v1 = 0.0, v2 = 0.0, v3 = 0.0;
        for (; k + 3 < c1; k += 4) {
            const unsigned p0 = CDN_LD(P.e_con_pk + k), p1 = CDN_LD(P.e_con_pk + k + 1);
            const unsigned p2 = CDN_LD(P.e_con_pk + k + 2), p3 = CDN_LD(P.e_con_pk + k + 3);
            const double j0 = w.jac[p0 & 0x1fffffffu], j1 = w.jac[p1 & 0x1fffffffu];
            const double j2 = w.jac[p2 & 0x1fffffffu], j3 = w.jac[p3 & 0x1fffffffu];
            v = fma(pk_sign(p0), j0, v);
            v1 = fma(pk_sign(p1), j1, v1);
            v2 = fma(pk_sign(p2), j2, v2);
            v3 = fma(pk_sign(p3), j3, v3);
        }
        v += (v1 + v2) + v3;
    }
    for (; k < c1; k += stride) {
        const unsigned pk = CDN_LD(P.e_con_pk + k);
        v = fma(pk_sign(pk), w.jac[pk & 0x1fffffffu], v);
    }
    return v;
}

template <class Team>
__device__ __forceinline__ bool assemble(Team& tm, const cdn_plan_dev& P, const Ws& w,
                                         const cdn_opts& o) {
    const double a0 = o.use_q ? o.a0 : 0.0;
    bool fin = true;
    for (int q = tm.rank(); q < P.nlu; q += tm.size()) {
        const int e = P.dense ? CDN_LD(P.e_bal + q) : q;
        if (P.n_e_heavy && CDN_LD(P.e_con_ptr + e + 1) - CDN_LD(P.e_con_ptr + e) > P.heavy) continue;
        double v = CDN_LD(P.e_g + e);
        if (o.use_q) v += a0 * CDN_LD(P.e_c + e);
        if (o.gmin != 0.0) v += o.gmin * CDN_LD(P.e_gones + e);
        // gmin from ground to the external nodes (nodal.py:479): dense slot
        // e = i * n + j is on the diagonal iff e is a multiple of n + 1
        if (o.gmin_ext != 0.0 && P.dense && e % (P.n + 1) == 0 && e / (P.n + 1) < o.n_ext)
            v += o.gmin_ext;
        v += slot_sum<Team>(P, w, o, a0, e, 0, 1);
        w.lu[e] = v;
        fin = fin && isfinite(v);
    }
    for (int h = 0; h < P.n_e_heavy; ++h) {
        const int e = CDN_LD(P.e_heavy + h);
        const double t = tm.sum(slot_sum<Team>(P, w, o, a0, e, tm.rank(), tm.size()));
        if (tm.rank() == 0) {
            double v = CDN_LD(P.e_g + e);
            if (o.use_q) v += a0 * CDN_LD(P.e_c + e);
            if (o.gmin != 0.0) v += o.gmin * CDN_LD(P.e_gones + e);
            w.lu[e] = v + t;
        }
    }
    return fin;
}

// q = C x + Tq q(x)   (device outputs must be current)
template <class Team>
__device__ __forceinline__ double charge_sum(const cdn_plan_dev& P, const Ws& w, int r, int first,
                                             int stride) {
    double acc = 0.0;
    const int l0 = CDN_LD(P.r_lin_ptr + r), l1 = CDN_LD(P.r_lin_ptr + r + 1);
    for (int k = l0 + first; k < l1; k += stride)
        acc += CDN_LD(P.r_lin_c + k) * w.x[CDN_LD(P.r_lin_col + k)];
    const int c0 = CDN_LD(P.r_con_ptr + r), c1 = CDN_LD(P.r_con_ptr + r + 1);
    for (int k = c0 + first; k < c1; k += stride) {
        const unsigned pk = CDN_LD(P.r_con_pk + k);
        if (pk & (1u << 30)) {
            const double v = w.out[pk & 0x1fffffffu];
            acc += (pk & (1u << 29)) ? -v : v;
        }
    }
    return acc;
}

template <class Team>
__device__ __forceinline__ void charge(Team& tm, const cdn_plan_dev& P, const Ws& w) {
    for (int r = tm.rank(); r < P.n; r += tm.size()) {
        if (P.n_r_heavy && row_heavy(P, r)) continue;
        w.q[r] = charge_sum<Team>(P, w, r, 0, 1);
    }
    for (int h = 0; h < P.n_r_heavy; ++h) {
        const int r = CDN_LD(P.r_heavy + h);
        const double acc = tm.sum(charge_sum<Team>(P, w, r, tm.rank(), tm.size()));
        if (tm.rank() == 0) w.q[r] = acc;
    }
}

// ---- dense LU with partial pivoting (LAPACK getrf/getrs semantics) ----------------------
template <class Team>
__device__ __forceinline__ void dense_factor(Team& tm, double* a, int* piv, int n) {
    for (int k = 0; k < n; ++k) {
        double best = -1.0;
        int bi = n;
        for (int i = k + tm.rank(); i < n; i += tm.size()) {
            const double v = fabs(a[i * n + k]);
            if (v > best) { best = v; bi = i; }
        }
        tm.argmax(best, bi);
        if (bi >= n) bi = k;                      // column of NaNs: keep the row
        if (tm.rank() == 0) piv[k] = bi;
        if (bi != k)
            for (int j = tm.rank(); j < n; j += tm.size()) {
                const double t = a[k * n + j];
                a[k * n + j] = a[bi * n + j];
                a[bi * n + j] = t;
            }
        tm.sync();
        // (dgetf2: a zero pivot is recorded, the column is left unscaled and
        // the elimination goes on)
        const double pk = a[k * n + k];
        const double pinv = (pk != 0.0) ? 1.0 / pk : 1.0;
        for (int i = k + 1 + tm.rank(); i < n; i += tm.size()) a[i * n + k] *= pinv;
        tm.sync();
        const int m = n - k - 1;
        for (int idx = tm.rank(); idx < m * m; idx += tm.size()) {
            const int i = k + 1 + idx / m, j = k + 1 + idx % m;
            a[i * n + j] -= a[i * n + k] * a[k * n + j];
        }
        tm.sync();
    }
}

// b <- A^{-1} b
template <class Team>
__device__ __forceinline__ void dense_solve(Team& tm, const double* a, const int* piv,
                                            double* b, int n) {
    if (tm.rank() == 0)
        for (int k = 0; k < n; ++k) {
            const int p = piv[k];
            if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; }
        }
    tm.sync();
    for (int k = 0; k < n - 1; ++k) {
        const double bk = b[k];
        for (int i = k + 1 + tm.rank(); i < n; i += tm.size()) b[i] -= a[i * n + k] * bk;
        tm.sync();
    }
    for (int k = n - 1; k >= 0; --k) {
        if (tm.rank() == 0) b[k] /= a[k * n + k];
        tm.sync();
        const double bk = b[k];
        for (int i = tm.rank(); i < k; i += tm.size()) b[i] -= a[i * n + k] * bk;
        tm.sync();
    }
}

// ---- sparse LU: static pivot order, entry-wise Crout by dependency level ----------------
// Entries (rows) with more than CDN_HEAVY terms sit at the head of their level
// and are reduced by the whole team (hub rows/columns).
// The pivot order was validated on the host at the reference point only
// (the reference re-pivots on every iteration, nodalSP.py:281-303).  A pivot
// that has degraded since shows as a zero / non-finite diagonal or a huge
// multiplier: *flag is set (benign race) and the host re-plans at the current
// iterate.
#define CDN_MULT_LIMIT 1e8
__device__ __forceinline__ void pivot_check(double v, int pv, int* flag) {
    if (pv >= 0) { if (!(fabs(v) <= CDN_MULT_LIMIT)) *flag = 1; }
    else if (pv == -2) { if (!(fabs(v) > 0.0) || !isfinite(v)) *flag = 1; }
}

template <class Team>
__device__ __forceinline__ void sparse_factor(Team& tm, const cdn_plan_dev& P, double* lu,
                                              int* flag) {
    for (int l = 0; l < P.nlev_f; ++l) {
        const int e0 = CDN_LD(P.f_lvl_ptr + l), e1 = CDN_LD(P.f_lvl_ptr + l + 1);
        const int nh = CDN_LD(P.f_lvl_nheavy + l);
        for (int h = 0; h < nh; ++h) {
            const int e = e0 + h;
            double part = 0.0;
            const int t1 = CDN_LD(P.e_term_ptr + e + 1);
            for (int t = CDN_LD(P.e_term_ptr + e) + tm.rank(); t < t1; t += tm.size())
                part += lu[CDN_LD(P.e_term_l + t)] * lu[CDN_LD(P.e_term_u + t)];
            part = tm.sum(part);
            if (tm.rank() == 0) {
                double v = lu[e] - part;
                const int pv = CDN_LD(P.e_piv + e);
                if (pv >= 0) v /= lu[pv];
                lu[e] = v;
                pivot_check(v, pv, flag);
            }
        }
        for (int e = e0 + nh + tm.rank(); e < e1; e += tm.size()) {
            double v = lu[e];
            const int t1 = CDN_LD(P.e_term_ptr + e + 1);
            for (int t = CDN_LD(P.e_term_ptr + e); t < t1; ++t)
                v -= lu[CDN_LD(P.e_term_l + t)] * lu[CDN_LD(P.e_term_u + t)];
            const int pv = CDN_LD(P.e_piv + e);
            if (pv >= 0) v /= lu[pv];
            lu[e] = v;
            pivot_check(v, pv, flag);
        }
        tm.sync();
    }
}

// dx <- A^{-1} y   (y in original row order; tmp = permuted work vector)
template <class Team>
__device__ __forceinline__ void sparse_solve(Team& tm, const cdn_plan_dev& P, const double* lu,
                                             const double* y, double* tmp, double* dx) {
    for (int l = 0; l < P.nlev_l; ++l) {
        const int r0 = CDN_LD(P.l_lvl_ptr + l), r1 = CDN_LD(P.l_lvl_ptr + l + 1);
        const int nh = CDN_LD(P.l_lvl_nheavy + l);
        for (int h = 0; h < nh; ++h) {
            const int i = CDN_LD(P.l_rows + r0 + h);
            double part = 0.0;
            const int k1 = CDN_LD(P.l_ptr + i + 1);
            for (int k = CDN_LD(P.l_ptr + i) + tm.rank(); k < k1; k += tm.size())
                part += lu[CDN_LD(P.l_slot + k)] * tmp[CDN_LD(P.l_col + k)];
            part = tm.sum(part);
            if (tm.rank() == 0) tmp[i] = y[CDN_LD(P.rowmap + i)] - part;
        }
        for (int q = r0 + nh + tm.rank(); q < r1; q += tm.size()) {
            const int i = CDN_LD(P.l_rows + q);
            double v = y[CDN_LD(P.rowmap + i)];
            const int k1 = CDN_LD(P.l_ptr + i + 1);
            for (int k = CDN_LD(P.l_ptr + i); k < k1; ++k)
                v -= lu[CDN_LD(P.l_slot + k)] * tmp[CDN_LD(P.l_col + k)];
            tmp[i] = v;
        }
        tm.sync();
    }
    for (int l = 0; l < P.nlev_u; ++l) {
        const int r0 = CDN_LD(P.u_lvl_ptr + l), r1 = CDN_LD(P.u_lvl_ptr + l + 1);
        const int nh = CDN_LD(P.u_lvl_nheavy + l);
        for (int h = 0; h < nh; ++h) {
            const int i = CDN_LD(P.u_rows + r0 + h);
            double part = 0.0;
            const int k1 = CDN_LD(P.u_ptr + i + 1);
            for (int k = CDN_LD(P.u_ptr + i) + tm.rank(); k < k1; k += tm.size())
                part += lu[CDN_LD(P.u_slot + k)] * tmp[CDN_LD(P.u_col + k)];
            part = tm.sum(part);
            if (tm.rank() == 0) {
                const double v = (tmp[i] - part) / lu[CDN_LD(P.u_diag + i)];
                tmp[i] = v;
                dx[CDN_LD(P.colmap + i)] = v;
            }
        }
        for (int q = r0 + nh + tm.rank(); q < r1; q += tm.size()) {
            const int i = CDN_LD(P.u_rows + q);
            double v = tmp[i];
            const int k1 = CDN_LD(P.u_ptr + i + 1);
            for (int k = CDN_LD(P.u_ptr + i); k < k1; ++k)
                v -= lu[CDN_LD(P.u_slot + k)] * tmp[CDN_LD(P.u_col + k)];
            v /= lu[CDN_LD(P.u_diag + i)];
            tmp[i] = v;
            dx[CDN_LD(P.colmap + i)] = v;
        }
        tm.sync();
    }
}

// Same factorisation for n <= team size, one lane per matrix column: no
// integer division, no shuffles, two barriers per pivot step.  Every lane
// scans the pivot column itself (shared-memory broadcast reads).
template <class Team>
__device__ __forceinline__ void dense_factor_cols(Team& tm, double* a, int* piv, int n) {
    const int j = tm.rank();
    for (int k = 0; k < n; ++k) {
        double best = fabs(a[k * n + k]);
        int bi = k;
        if (!(best >= 0.0)) best = -1.0;               // NaN never wins
        for (int i = k + 1; i < n; ++i) {
            const double v = fabs(a[i * n + k]);
            if (v > best) { best = v; bi = i; }
        }
        tm.sync();                                     // all lanes have read column k
        if (j < n && bi != k) {
            const double t = a[k * n + j];
            a[k * n + j] = a[bi * n + j];
            a[bi * n + j] = t;
        }
        if (j == 0) piv[k] = bi;
        tm.sync();
        const double pk = a[k * n + k];
        const double pinv = (pk != 0.0) ? 1.0 / pk : 1.0;      // dgetf2: zero pivot
        if (j > k && j < n) {
            const double akj = a[k * n + j];
            for (int i = k + 1; i < n; ++i) a[i * n + j] -= (a[i * n + k] * pinv) * akj;
        }
        tm.sync();                                     // column k read by everybody
        if (j == k)
            for (int i = k + 1; i < n; ++i) a[i * n + k] *= pinv;
    }
    tm.sync();
}

// b <- A^{-1} b by one lane (tiny systems: the barriers of the cooperative
// version cost more than the 2 n^2 flops)
template <class Team>
__device__ __forceinline__ void dense_solve_serial(Team& tm, const double* a, const int* piv,
                                                   double* b, int n) {
    if (tm.rank() == 0) {
        for (int k = 0; k < n; ++k) {
            const int p = piv[k];
            if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; }
        }
        for (int i = 1; i < n; ++i) {
            double v = b[i];
            for (int k = 0; k < i; ++k) v -= a[i * n + k] * b[k];
            b[i] = v;
        }
        for (int i = n - 1; i >= 0; --i) {
            double v = b[i];
            for (int k = i + 1; k < n; ++k) v -= a[i * n + k] * b[k];
            b[i] = v / a[i * n + i];
        }
    }
    tm.sync();
}

// n <= 8 on a warp tile: the whole matrix lives in registers, lane j holding
// column j; the pivot column is broadcast with shuffles, so the factorisation
// has no barrier and no shared-memory round trip inside (same arithmetic and
// pivot choice as above).  The factors are written back for the solves.
template <int T>
__device__ __forceinline__ void dense_factor_reg8(TileTeam<T>& tm, double* a, int* piv, int n) {
    const int j = tm.rank();
    double col[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) col[i] = (i < n && j < n) ? a[i * n + j] : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < n) {
            double v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (i >= k) ? tm.tile.shfl(col[i], k) : 0.0;
            double best = fabs(v[k]);
            int bi = k;
            if (!(best >= 0.0)) best = -1.0;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i > k && i < n) {
                    const double t = fabs(v[i]);
                    if (t > best) { best = t; bi = i; }
                }
            if (j == 0) piv[k] = bi;
            // swap rows k and bi (own column and the broadcast pivot column)
            const double ck = col[k], vk = v[k];
            double cb = ck, vb = vk;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i > k && i == bi) { cb = col[i]; vb = v[i]; col[i] = ck; v[i] = vk; }
            col[k] = cb;
            const double pinv = (vb != 0.0) ? 1.0 / vb : 1.0;   // dgetf2: zero pivot
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i > k && i < n) {
                    const double l = v[i] * pinv;
                    if (j > k) col[i] -= l * cb;
                    else if (j == k) col[i] = l;
                }
        }
    }
    if (j < n) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < n) a[i * n + j] = col[i];
    }
    tm.sync();
}

// b <- A^{-1} b with the factors of dense_factor_reg8: every lane carries the
// whole right-hand side in registers, the L and U entries come from the lane
// that owns their column.
template <int T>
__device__ __forceinline__ void dense_solve_reg8(TileTeam<T>& tm, const double* a, const int* piv,
                                                 double* b, int n) {
    const int j = tm.rank();
    double col[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        col[i] = (i < n && j < n) ? a[i * n + j] : 0.0;
        x[i] = (i < n) ? b[i] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < n) {
            const int p = piv[k];
            if (p != k) {
                const double t = x[k];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i > k && i == p) { x[k] = x[i]; x[i] = t; }
            }
        }
#pragma unroll
    for (int k = 0; k < 8; ++k)            // L y = P b (unit diagonal)
        if (k + 1 < n) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i > k && i < n) x[i] -= tm.tile.shfl(col[i], k) * x[k];
        }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {       // U x = y
        const int k = 7 - kk;
        if (k < n) {
            x[k] = x[k] / tm.tile.shfl(col[k], k);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < k) x[i] -= tm.tile.shfl(col[i], k) * x[k];
        }
    }
    tm.sync();                              // everybody has read b
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i == j && i < n) b[i] = x[i];
    tm.sync();
}

template <class Team>
__device__ __forceinline__ void factor_small(Team& tm, double* a, int* piv, int n) {
    dense_factor_cols(tm, a, piv, n);
}
template <int T>
__device__ __forceinline__ void factor_small(TileTeam<T>& tm, double* a, int* piv, int n) {
    if (T >= 8 && n <= 8) dense_factor_reg8(tm, a, piv, n);
    else dense_factor_cols(tm, a, piv, n);
}
template <class Team>
__device__ __forceinline__ void solve_small(Team& tm, const double* a, const int* piv, double* b,
                                            int n) {
    dense_solve_serial(tm, a, piv, b, n);
}
template <int T>
__device__ __forceinline__ void solve_small(TileTeam<T>& tm, const double* a, const int* piv,
                                            double* b, int n) {
    if (T >= 8 && n <= 8) dense_solve_reg8(tm, a, piv, b, n);
    else dense_solve_serial(tm, a, piv, b, n);
}

template <class Team>
__device__ __forceinline__ void factor(Team& tm, const cdn_plan_dev& P, const Ws& w) {
    if (P.dense) {
        if (Team::kFlat && P.n <= tm.size()) factor_small(tm, w.lu, w.piv, P.n);
        else dense_factor(tm, w.lu, w.piv, P.n);
    } else {
        sparse_factor(tm, P, w.lu, w.piv + P.n);
    }
}

// Sweeps from shared memory (block team, compact plans): the LU values and
// the work vector are resident; the program of each sweep (level pointers,
// packed row / entry records, 16-bit index streams; see plan.h) is one blob that
// is staged next to them with 16-byte copies, so that between two barriers a
// thread touches shared memory only.
__host__ __device__ inline long compact_stage_bytes(const cdn_plan_dev& P) {
    int w = P.cl_words > P.cu_words ? P.cl_words : P.cu_words;
    if (P.cf_words > w) w = P.cf_words;
    return 4L * w + 32;
}

__device__ __forceinline__ void stage_blob(unsigned* dst, const unsigned* src, int nwords) {
    const uint4* s4 = (const uint4*)src;
    uint4* d4 = (uint4*)dst;
    const int n4 = nwords >> 2;                    // blobs are padded to 16 bytes
    // (batching several loads per thread before the stores was measured slower:
    // the 1024-thread CTA has 64 registers per thread)
    for (int k = threadIdx.x; k < n4; k += blockDim.x) d4[k] = s4[k];
}

// (The shared-memory pointers are derived from the kernel's dynamic shared array
// here, not taken from the generic pointers in Ws: the compiler then emits
// LDS/STS instead of generic loads and stores.)
__device__ __forceinline__ unsigned* stage_area(double* dsm, const cdn_plan_dev& P) {
    // same layout as engine_block_kernel: lu[nlu] | tmp[n] | 16-byte aligned blob
    return (unsigned*)(dsm + ((P.nlu + P.n + 1) & ~1));
}

// One level of a staged sweep: 2^lg lanes share a row (plan.h; the host picks lg
// per level).  The row loop is uniform over each warp -- lanes past the end
// of the level run along with an empty row -- so that the shuffles of
// group_sum can name the whole warp.
//   CDN_GROUP_ROWS(r0, r1, lg) { CDN_GROUP_ROW(r1, lg); ... }   defines g, sub, q, ok
#define CDN_GROUP_ROWS(r0, r1, lg)                                                           \
    for (int qw_ = (r0) + (int)((threadIdx.x & ~31u) >> (lg)); qw_ < (r1);                   \
         qw_ += (int)(blockDim.x >> (lg)))
#define CDN_GROUP_ROW(r1, lg)                                                                \
    const int g = 1 << (lg), sub = (int)(threadIdx.x & (unsigned)(g - 1));                   \
    const int q = qw_ + (int)((threadIdx.x & 31u) >> (lg));                                  \
    const bool ok = q < (r1)
__device__ __forceinline__ double group_sum(double s, int g) {
    for (int off = g >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    return s;
}

__device__ __forceinline__ void staged_solve(BlockTeam& tm, const cdn_plan_dev& P, const Ws& w) {
    extern __shared__ double dsm[];
    const int n = P.n;
    double* lu = dsm;
    double* tmp = dsm + P.nlu;
    unsigned* blob = stage_area(dsm, P);
    long long* dp = w.dprof;
    long long tt = dp ? clock64() : 0;
#define CDN_DP(k) do { if (dp) { const long long t_ = clock64(); if (threadIdx.x == 0) dp[k] += t_ - tt; tt = t_; } } while (0)
    // ---- L sweep -------------------------------------------------------------------
    {
        stage_blob(blob, P.cl_prog, P.cl_words);
        __syncthreads();
        CDN_DP(8);
        const int* lvl = (const int*)(blob + P.cl_off[0]);
        const uint2* rec = (const uint2*)(blob + P.cl_off[1]);
        const unsigned short* col = (const unsigned short*)(blob + P.cl_off[2]);
        const unsigned short* slot = (const unsigned short*)(blob + P.cl_off[3]);
        const unsigned short* rmap = (const unsigned short*)(blob + P.cl_off[4]);
        for (int i = threadIdx.x; i < n; i += blockDim.x) tmp[i] = w.y[rmap[i]];
        __syncthreads();
        CDN_DP(9);
        const int nlev = P.nlev_l;
        const int* glg = lvl + nlev + 1;
        for (int l = 0; l < nlev; ++l) {
            const int r0 = lvl[l], r1 = lvl[l + 1], lg = glg[l];
            CDN_GROUP_ROWS(r0, r1, lg) {
                CDN_GROUP_ROW(r1, lg);
                const uint2 r = ok ? rec[q] : make_uint2(0u, 0u);
                const int i = r.x & 0xffffu, cnt = r.x >> 16, k0 = r.y;
                double s = 0.0;
#pragma unroll 2
                for (int k = sub; k < cnt; k += g) s = fma(lu[slot[k0 + k]], tmp[col[k0 + k]], s);
                s = group_sum(s, g);
                if (ok && sub == 0) tmp[i] -= s;
            }
            __syncthreads();
            if (l < 32) CDN_DP(16 + l);
        }
    }
    CDN_DP(10);
    // ---- U sweep -------------------------------------------------------------------
    {
        stage_blob(blob, P.cu_prog, P.cu_words);
        __syncthreads();
        CDN_DP(11);
        const int* lvl = (const int*)(blob + P.cu_off[0]);
        const uint2* rec = (const uint2*)(blob + P.cu_off[1]);
        const unsigned short* col = (const unsigned short*)(blob + P.cu_off[2]);
        const unsigned short* slot = (const unsigned short*)(blob + P.cu_off[3]);
        const unsigned short* cmap = (const unsigned short*)(blob + P.cu_off[4]);
        const int nlev = P.nlev_u;
        const int* glg = lvl + nlev + 1;
        for (int l = 0; l < nlev; ++l) {
            const int r0 = lvl[l], r1 = lvl[l + 1], lg = glg[l];
            CDN_GROUP_ROWS(r0, r1, lg) {
                CDN_GROUP_ROW(r1, lg);
                const uint2 r = ok ? rec[q] : make_uint2(0u, 0u);
                const int i = r.x & 0xffffu, cnt = r.x >> 16;
                const int k0 = r.y & 0xffffu, dg = r.y >> 16;
                double s = 0.0;
#pragma unroll 2
                for (int k = sub; k < cnt; k += g) s = fma(lu[slot[k0 + k]], tmp[col[k0 + k]], s);
                s = group_sum(s, g);
                if (ok && sub == 0) tmp[i] = (tmp[i] - s) / lu[dg];
            }
            __syncthreads();
            if (l < 32) CDN_DP(48 + l);
        }
        CDN_DP(12);
        for (int i = threadIdx.x; i < n; i += blockDim.x) w.dx[cmap[i]] = tmp[i];
        __syncthreads();
        CDN_DP(13);
    }
#undef CDN_DP
}

template <class Team>
__device__ __forceinline__ bool try_staged_solve(Team&, const cdn_plan_dev&, const Ws&) {
    return false;
}
__device__ __forceinline__ bool try_staged_solve(BlockTeam& tm, const cdn_plan_dev& P, const Ws& w) {
    if (!w.stage || !P.compact) return false;
    staged_solve(tm, P, w);
    return true;
}

// Partial refactorisation (block team, compact plans).  Between two Newton
// iterations only the device contributions change: an entry of L+U changes only
// if it receives one, or if an operand of its Crout update or its pivot does.
// That closure is computed once on the host; all other entries keep the values
// of the last full factorisation (same a0 and gmin).  The affected entries are
// re-assembled from the linear parts and re-eliminated level by level from a
// staged program.
__device__ __forceinline__ void staged_refactor(BlockTeam& tm, const cdn_plan_dev& P, const Ws& w,
                                                const cdn_opts& o) {
    extern __shared__ double dsm[];
    double* lu = dsm;
    unsigned* blob = stage_area(dsm, P);
    stage_blob(blob, P.cf_prog, P.cf_words);
    __syncthreads();
    const int* lvl = (const int*)(blob + P.cf_off[0]);
    const uint2* rec = (const uint2*)(blob + P.cf_off[1]);
    const unsigned* term = blob + P.cf_off[2];
    const double a0 = o.use_q ? o.a0 : 0.0;
    // assembly of the affected slots
    for (int q = threadIdx.x; q < P.naff; q += blockDim.x) {
        const int e = rec[q].x & 0xffffu;
        double v = P.e_g[e];
        if (o.use_q) v += a0 * P.e_c[e];
        if (o.gmin != 0.0) v += o.gmin * P.e_gones[e];
        lu[e] = v + slot_sum<BlockTeam>(P, w, o, a0, e, 0, 1);
    }
    __syncthreads();
    const int nlev = P.nlev_f;
    const int* glg = lvl + nlev + 1;
    for (int l = 0; l < nlev; ++l) {
        const int q0 = lvl[l], q1 = lvl[l + 1], lg = glg[l];
        if (q0 == q1) continue;                    // nothing changes on this level
        CDN_GROUP_ROWS(q0, q1, lg) {
            CDN_GROUP_ROW(q1, lg);
            const uint2 r = ok ? rec[q] : make_uint2(0u, 0xffff0000u);
            const int e = r.x & 0xffffu, cnt = r.x >> 16;
            const int t0 = r.y & 0xffffu, pv = r.y >> 16;
            double s = 0.0;
#pragma unroll 2
            for (int t = sub; t < cnt; t += g) {
                const unsigned tt = term[t0 + t];
                s = fma(lu[tt & 0xffffu], lu[tt >> 16], s);
            }
            s = group_sum(s, g);
            if (ok && sub == 0) {
                double v = lu[e] - s;
                if (pv != 0xffff) {
                    v /= lu[pv];
                    if (!(fabs(v) <= CDN_MULT_LIMIT)) w.piv[P.n] = 1;
                }
                lu[e] = v;
            }
        }
        __syncthreads();
    }
}

// state of the factors held in w.lu during one launch
struct FactState {
    bool valid;
    double a0, gmin;
};

template <class Team>
__device__ __forceinline__ bool try_staged_refactor(Team&, const cdn_plan_dev&, const Ws&,
                                                    const cdn_opts&, const FactState&) {
    return false;
}
__device__ __forceinline__ bool try_staged_refactor(BlockTeam& tm, const cdn_plan_dev& P,
                                                    const Ws& w, const cdn_opts& o,
                                                    const FactState& fs) {
    const double a0 = o.use_q ? o.a0 : 0.0;
    if (!w.stage || !P.compact || !fs.valid || fs.a0 != a0 || fs.gmin != o.gmin) return false;
    staged_refactor(tm, P, w, o);
    return true;
}

// dx <- J^{-1} y  (y is consumed).  `scrub`: condition_array of the dense
// factor_and_solve / solve_linear_system (nodal.py:385-402, 626, 644): nan ->
// 10 abstol, +-inf -> +-0.1 maxdelta; get_chord_deltax (:648) does not scrub.
template <class Team>
__device__ __forceinline__ void lusolve(Team& tm, const cdn_plan_dev& P, const Ws& w,
                                        const cdn_opts& o, bool scrub) {
    if (P.dense) {
        if (Team::kFlat && P.n <= 16) solve_small(tm, w.lu, w.piv, w.y, P.n);
        else dense_solve(tm, w.lu, w.piv, w.y, P.n);
        for (int r = tm.rank(); r < P.n; r += tm.size()) {
            double d = w.y[r];
            if (scrub && !isfinite(d))
                d = isnan(d) ? 10.0 * o.abstol : (d > 0.0 ? 0.1 : -0.1) * o.maxdelta;
            w.dx[r] = d;
        }
        tm.sync();
    } else if (!try_staged_solve(tm, P, w)) {
        sparse_solve(tm, P, w.lu, w.y, w.tmp, w.dx);
    }
}


// ---- damped Newton (fsolve.py:59-128) -------------------------------------------------
// Returns the iteration count; *ok tells whether the tolerances were met.
// On return w.ivec holds the residual evaluated at the LAST-BUT-ONE iterate and
// w.lu the factors of the Jacobian there -- exactly what the reference's chord
// predictor re-uses (tran.py:181) -- unless errfunc is set.
// phase timers (instance 0, thread 0 of the team): where a Newton iteration goes
#define CDN_PROF_T0 long long pt_ = prof ? clock64() : 0
#define CDN_PROF(k)                                                     \
    do {                                                                \
        if (prof) {                                                     \
            const long long t_ = clock64();                             \
            if (tm.rank() == 0) prof[k] += t_ - pt_;                    \
            pt_ = t_;                                                   \
        }                                                               \
    } while (0)

template <int MS, class Team>
__device__ __forceinline__ int newton(Team& tm, const cdn_plan_dev& P, int b, const Ws& w,
                                      const cdn_opts& o, bool active, double* res_out, bool* ok,
                                      long long* prof = nullptr, FactState* fs = nullptr) {
    double res = 0.0;
    bool success = false;
    int it = 0;
    // `active` teams iterate; the others only keep the CTA-level count of
    // cta_any() calls in step (tile teams; see TileTeam::cta_any)
    bool run = active;
    while (tm.cta_any(run && it < o.maxiter)) {
        if (!(run && it < o.maxiter)) continue;
        CDN_PROF_T0;
        eval_devices<MS, true>(tm, P, b, w.x, w, true, o.h, o.use_q ? o.a0 : 1.0);
        CDN_PROF(0);
        bool fin = residual(tm, P, w, o);
        if (fs && try_staged_refactor(tm, P, w, o, *fs)) {
            CDN_PROF(2);
        } else {
            fin = assemble(tm, P, w, o) && fin;
            if (P.dense && !tm.all(fin)) {
                // NaN/inf in the dense Jacobian or right-hand side: this
                // Newton solve has failed (oracle/nodal_ref.py
                // factor_and_solve explains the choice); next helper or
                // bisection step
                ++it;
                res = INFINITY;
                run = false;
                continue;
            }
            tm.sync();
            CDN_PROF(1);
            factor(tm, P, w);
            if (fs) { fs->valid = true; fs->a0 = o.use_q ? o.a0 : 0.0; fs->gmin = o.gmin; }
            CDN_PROF(2);
        }
        lusolve(tm, P, w, o, true);
        CDN_PROF(3);
        double m = 0.0;
        for (int r = tm.rank(); r < P.n; r += tm.size()) m = fmax(m, fabs(w.dx[r]));
        m = tm.max(m);
        const bool damp = m > o.maxdelta;
        const double scale = damp ? o.maxdelta / m : 1.0;
        bool conv = true;
        for (int r = tm.rank(); r < P.n; r += tm.size()) {
            double d = w.dx[r];
            if (damp) d *= scale;
            const double xo = w.x[r], xn = xo + d;
            conv = conv && (fabs(d) < o.reltol * fmax(fabs(xo), fabs(xn)) + o.abstol);
            w.x[r] = xn;
        }
        res = damp ? m * scale : m;
        const bool n1 = tm.all(conv);      // (also orders the x update)
        tm.sync();
        bool n2 = true;
        if (n1 && o.errfunc) {
            eval_devices<MS, false>(tm, P, b, w.x, w, true, o.h);
            residual(tm, P, w, o);
            tm.sync();
            double ef = 0.0;
            for (int r = tm.rank(); r < P.n; r += tm.size()) ef = fmax(ef, fabs(w.y[r]));
            ef = tm.max(ef);
            n2 = ef < o.abstol;
            res = fmax(ef, res);
        }
        ++it;
        if (n1 && n2) { success = true; run = false; }
        CDN_PROF(4);
    }
    *res_out = res;
    *ok = success;
    return it;
}


// Continuation on lambda with bisection and backtracking (nodal.py:545-604):
// mode 0 = gmin stepping on the nonlinear ports (gmin = 1e-3/lambda - 1e-3,
// nodal.py:491-515, 535-543), mode 1 = source stepping (:517-533), mode 2 =
// gmin stepping from ground to the external nodes (:469-489).  w.xb is
// the caller's x0 and, as in the reference, is overwritten by every
// successful intermediate solution.
template <int MS, class Team>
__device__ __noinline__ int homotopy(Team& tm, const cdn_plan_dev& P, int b, const Ws& w,
                                     cdn_opts o, int mode, bool active, double* res_out,
                                     bool* ok) {
    double stack[48];
    int sp = 0;
    stack[sp++] = 0.0;
    stack[sp++] = 1.0;
    double lam = 0.5, step = 0.5, res = 0.0;
    const double small = 1e-4;
    bool success = true;
    int tot = 0;
    if (active) {
        for (int r = tm.rank(); r < P.n; r += tm.size()) w.x[r] = w.xb[r];
        tm.sync();
    }
    bool run = active;
    while (tm.cta_any(run && sp > 0)) {
        const bool go = run && sp > 0;
        if (mode == 0) o.gmin = 1e-3 / lam - 1e-3;
        else if (mode == 1) o.lambda = lam;
        else o.gmin_ext = 1e-3 / lam - 1e-3;
        bool ok1 = false;
        tot += newton<MS>(tm, P, b, w, o, go, &res, &ok1);
        if (!go) continue;
        if (ok1) {
            for (int r = tm.rank(); r < P.n; r += tm.size()) w.xb[r] = w.x[r];
            step = stack[sp - 1] - lam;
            lam = stack[--sp];
        } else {
            for (int r = tm.rank(); r < P.n; r += tm.size()) w.x[r] = w.xb[r];
            if (sp < 48) stack[sp++] = lam;
            step *= 0.5;
            lam -= step;
            if (lam < small || step < small) { success = false; run = false; }
        }
        tm.sync();
    }
    *res_out = res;
    *ok = success && active;
    return tot;
}

// fsolve.solve (fsolve.py:23-56) over the helper lists of the reference: plain
// Newton, gmin stepping on the nonlinear ports, source stepping (sparse engine,
// nodalSP.py:232-235) and, for dense plans, gmin stepping on the external nodes
// (nodal.py:443-447).  x0 is taken from w.x; returns the iteration count of
// the helper that succeeded and its index in *used.
template <int MS, class Team>
__device__ __forceinline__ int solve_helpers(Team& tm, const cdn_plan_dev& P, int b,
                                             const Ws& w, const cdn_opts& o, int helpers,
                                             bool active, double* res_out, bool* ok,
                                             int* used, long long* prof = nullptr,
                                             FactState* fs = nullptr) {
    // `helpers`: bit 0 enables the chain; bits 1..3 (tests) skip plain Newton,
    // gmin stepping on the ports, source stepping
    *used = 0;
    if (helpers && active) {
        for (int r = tm.rank(); r < P.n; r += tm.size()) w.xb[r] = w.x[r];
        tm.sync();
    }
    int it = 0;
    *ok = false;
    *res_out = 0.0;
    if (!(helpers & 2)) it = newton<MS>(tm, P, b, w, o, active, res_out, ok, prof, fs);
    if (!helpers) return it;
    bool need = active && !*ok;
    if (!tm.cta_any(need)) return it;
    if (fs) fs->valid = false;              // the helpers refactor in full
    double res2 = 0.0;
    bool ok2 = false;
    int it2;
    if (!(helpers & 4)) {
        it2 = homotopy<MS>(tm, P, b, w, o, 0, need, &res2, &ok2);
        if (need) { it = it2; *res_out = res2; *ok = ok2; *used = 1; }
        need = need && !ok2;
        if (!tm.cta_any(need)) return it;
    }
    if (!(helpers & 8)) {
        it2 = homotopy<MS>(tm, P, b, w, o, 1, need, &res2, &ok2);
        if (need) { it = it2; *res_out = res2; *ok = ok2; *used = 2; }
        need = need && !ok2;
    }
    const bool has4 = P.dense && o.n_ext > 0;        // team-uniform
    if (!has4 || !tm.cta_any(need)) return it;
    it2 = homotopy<MS>(tm, P, b, w, o, 2, need, &res2, &ok2);
    if (need) { it = it2; *res_out = res2; *ok = ok2; *used = 3; }
    return it;
}

// q(x) with a values-only device sweep (nodalSP.py:719-733)
template <int MS, class Team>
__device__ __forceinline__ void update_q(Team& tm, const cdn_plan_dev& P, int b, const Ws& w) {
    eval_devices<MS, false>(tm, P, b, w.x, w);
    charge(tm, P, w);
    tm.sync();
}

template <class Team>
__device__ __forceinline__ void integ_init(Team& tm, const cdn_plan_dev& P, const Ws& w) {
    for (int r = tm.rank(); r < P.n; r += tm.size()) { w.qprev[r] = w.q[r]; w.dq[r] = 0.0; }
    tm.sync();
}

// integration.py:36-38 (BE), :83-91 (trapezoidal)
template <class Team>
__device__ __forceinline__ void integ_accept(Team& tm, const cdn_plan_dev& P, const Ws& w,
                                             const cdn_opts& o) {
    for (int r = tm.rank(); r < P.n; r += tm.size()) {
        const double q = w.q[r];
        if (o.im == 1) w.dq[r] = -w.dq[r] + o.a0 * (q - w.qprev[r]);
        w.qprev[r] = q;
    }
    tm.sync();
}

// sv = f_n1() + s   (integration.py:47-49, :93-97; nodalSP.py:792-802)
template <class Team>
__device__ __forceinline__ void make_rhs(Team& tm, const cdn_plan_dev& P, const Ws& w,
                                         const cdn_opts& o, const double* s_dense,
                                         int nsrc, const int* src_rows, const double* src_val) {
    for (int r = tm.rank(); r < P.n; r += tm.size()) {
        double f = o.a0 * w.qprev[r];
        if (o.im == 1) f += w.dq[r];
        w.sv[r] = f + (s_dense ? __ldg(s_dense + r) : 0.0);
    }
    tm.sync();
    for (int k = tm.rank(); k < nsrc; k += tm.size()) w.sv[__ldg(src_rows + k)] += __ldg(src_val + k);
    if (nsrc) tm.sync();
}

// dx = LU \ (sv - iVec) with the factors and the residual left by the last
// Newton iteration (nodalSP.py:345-354)
template <class Team>
__device__ __forceinline__ void chord(Team& tm, const cdn_plan_dev& P, const Ws& w,
                                      const cdn_opts& o) {
    for (int r = tm.rank(); r < P.n; r += tm.size()) w.y[r] = w.sv[r] - w.ivec[r];
    tm.sync();
    lusolve(tm, P, w, o, false);
}


// ---- one circuit instance, one operation ------------------------------------------------
// `active` is false for the padding tiles of a ragged batch: they only take
// part in the CTA-level lockstep (TileTeam::cta_any).
template <int MS, class Team>
__device__ __forceinline__ void run_instance(Team& tm, const cdn_engine_args& A,
                                             const cdn_plan_dev& P, int b, const Ws& w,
                                             bool active) {
    const cdn_opts& o = A.o;
    const int n = P.n;
    switch (A.op) {
    case CDN_OP_NEWTON: {
        if (active) {
            for (int r = tm.rank(); r < n; r += tm.size()) w.sv[r] = __ldg(A.sv_in + r);
            tm.sync();
        }
        double res; bool ok; int used;
        if (active && !P.dense && tm.rank() == 0) w.piv[n] = 0;
        const int it = solve_helpers<MS>(tm, P, b, w, o, A.helpers, active, &res, &ok, &used,
                                         b == 0 ? A.prof : nullptr);
        if (active && tm.rank() == 0) {
            A.iters[b] = it; A.res[b] = res; A.status[b] = ok ? 0 : -1;
            if (A.info) { A.info[2 * b] = used; A.info[2 * b + 1] = P.dense ? 0 : w.piv[n]; }
        }
        break;
    }
    case CDN_OP_TRAN: {
        if (active && A.init_ic) {               // set_IC (nodalSP.py:666-696)
            update_q<MS>(tm, P, b, w);
            integ_init(tm, P, w);
            hist_store(tm, P, w, true);
        }
        int status = 0, used_max = 0, piv_at = 0;
        bool run = active;
        if (active && !P.dense && tm.rank() == 0) w.piv[n] = 0;
        FactState fs;
        fs.valid = false; fs.a0 = 0.0; fs.gmin = 0.0;
        for (int i = A.first; i < A.nsamples; ++i) {
            if (!tm.cta_any(run)) break;
            long long* prof = (b == 0) ? A.prof : nullptr;
            if (run) {
                CDN_PROF_T0;
                update_q<MS>(tm, P, b, w);       // accept(xOld)
                integ_accept(tm, P, w, o);
                hist_store(tm, P, w, false);
                make_rhs(tm, P, w, o, A.src_dc, A.nsrc, A.src_rows,
                         A.src_val + (long)i * A.nsrc);
                CDN_PROF(5);
                if (i > 1) {                     // chord predictor (tran.py:181)
                    chord(tm, P, w, o);
                    for (int r = tm.rank(); r < n; r += tm.size()) w.x[r] += w.dx[r];
                    tm.sync();
                }
                CDN_PROF(6);
            }
            double res; bool ok; int used;
            const int it = solve_helpers<MS>(tm, P, b, w, o, A.helpers, run, &res, &ok, &used,
                                             prof, &fs);
            if (!run) continue;
            if (used > used_max) used_max = used;
            if (!P.dense && !piv_at && w.piv[n]) piv_at = i + 1;
            const long at = (long)b * A.nsamples + i;
            if (tm.rank() == 0) { A.iters[at] = it; A.res[at] = res; }
            if (!ok) { status = i; run = false; continue; }
            for (int k = tm.rank(); k < A.nsave; k += tm.size())
                A.wave[at * A.nsave + k] = w.x[__ldg(A.save_rows + k)];
        }
        if (active && tm.rank() == 0) {
            A.status[b] = status;
            if (A.info) { A.info[2 * b] = used_max; A.info[2 * b + 1] = piv_at; }
        }
        break;
    }
    case CDN_OP_UPDATE_Q:
        if (!active) break;
        update_q<MS>(tm, P, b, w);
        break;
    case CDN_OP_SET_IC:
        if (!active) break;
        update_q<MS>(tm, P, b, w);
        integ_init(tm, P, w);
        hist_store(tm, P, w, true);
        break;
    case CDN_OP_ACCEPT:
        if (!active) break;
        update_q<MS>(tm, P, b, w);
        integ_accept(tm, P, w, o);
        hist_store(tm, P, w, false);
        break;
    case CDN_OP_RHS:
        if (!active) break;
        make_rhs(tm, P, w, o, A.sv_in, 0, nullptr, nullptr);
        break;
    case CDN_OP_CHORD:
        if (!active) break;
        chord(tm, P, w, o);
        break;
    case CDN_OP_GET_I:
        if (!active) break;
        eval_devices<MS, false>(tm, P, b, w.x, w, true, o.h);
        residual(tm, P, w, o);
        tm.sync();
        break;
    case CDN_OP_GET_I_JAC:
        if (!active) break;
        eval_devices<MS, true>(tm, P, b, w.x, w, true, o.h, o.use_q ? o.a0 : 1.0);
        assemble(tm, P, w, o);
        residual(tm, P, w, o);
        tm.sync();
        break;
    case CDN_OP_SOLVE:
        if (!active) break;
        factor(tm, P, w);
        lusolve(tm, P, w, o, true);
        break;
    default: break;
    }
}

// ---- shared-memory copy of a small (dense) plan ------------------------------------------
__host__ __device__ inline long plan_smem_bytes(const cdn_plan_dev& P) {
    long b = 0;
    b += 3L * P.nlu * 8;                         // e_g, e_c, e_gones
    b += (2L * P.nlu + 1L) * 4 + (long)P.ncon * 4;    // e_con_ptr, e_bal, e_con_pk
    b += (P.n + 1L) * 4 * 2;                     // r_lin_ptr, r_con_ptr
    b += (long)P.nrlin * (4 + 3 * 8);            // r_lin_col, r_lin_g/c/gones
    b += (long)P.nrcon * 4;                      // r_con_pk
    b += (long)P.ndev * 8;                       // dev_g, dev_i
    if (P.sd) b += 9L * 4 + 4L * P.sd_nterms + 4L * P.sd_nfix;   // sd_ptr, sd_prog, sd_fix
    for (int g = 0; g < P.ngroups; ++g) b += 2L * P.groups[g].nxin * P.groups[g].n * 4;
    return b + 16 * 16;                          // alignment slack
}

template <class T>
__device__ __forceinline__ const T* smem_copy(const T* src, long count, char*& cur) {
    cur = (char*)(((unsigned long long)cur + 15ull) & ~15ull);
    T* dst = (T*)cur;
    for (long k = threadIdx.x; k < count; k += blockDim.x) dst[k] = src[k];
    cur += count * (long)sizeof(T);
    return dst;
}

// all threads of the CTA; ends with a CTA barrier
__device__ __forceinline__ void plan_to_smem(const cdn_plan_dev& G, cdn_plan_dev& S, char* buf) {
    {   // the struct itself
        const int* s = (const int*)&G;
        int* d = (int*)&S;
        for (int k = threadIdx.x; k < (int)(sizeof(cdn_plan_dev) / 4); k += blockDim.x) d[k] = s[k];
    }
    __syncthreads();
    char* cur = buf;
    const double* e_g = smem_copy(G.e_g, G.nlu, cur);
    const double* e_c = smem_copy(G.e_c, G.nlu, cur);
    const double* e_gones = smem_copy(G.e_gones, G.nlu, cur);
    const double* r_lin_g = smem_copy(G.r_lin_g, G.nrlin, cur);
    const double* r_lin_c = smem_copy(G.r_lin_c, G.nrlin, cur);
    const double* r_lin_gones = smem_copy(G.r_lin_gones, G.nrlin, cur);
    const int* e_con_ptr = smem_copy(G.e_con_ptr, G.nlu + 1L, cur);
    const unsigned* e_con_pk = smem_copy(G.e_con_pk, G.ncon, cur);
    const int* r_lin_ptr = smem_copy(G.r_lin_ptr, G.n + 1L, cur);
    const int* r_lin_col = smem_copy(G.r_lin_col, G.nrlin, cur);
    const int* r_con_ptr = smem_copy(G.r_con_ptr, G.n + 1L, cur);
    const unsigned* r_con_pk = smem_copy(G.r_con_pk, G.nrcon, cur);
    const int* e_bal = smem_copy(G.e_bal, G.nlu, cur);
    const int* dev_g = smem_copy(G.dev_g, G.ndev, cur);
    const int* dev_i = smem_copy(G.dev_i, G.ndev, cur);
    const int* sd_ptr = G.sd ? smem_copy(G.sd_ptr, 9L, cur) : nullptr;
    const unsigned* sd_prog = G.sd ? smem_copy(G.sd_prog, (long)G.sd_nterms, cur) : nullptr;
    const int* sd_fix = G.sd ? smem_copy(G.sd_fix, (long)G.sd_nfix, cur) : nullptr;
    const int* vp[CDN_MAX_GROUPS];
    const int* vn[CDN_MAX_GROUPS];
    for (int g = 0; g < G.ngroups; ++g) {
        const long cnt = (long)G.groups[g].nxin * G.groups[g].n;
        vp[g] = smem_copy(G.groups[g].vpos, cnt, cur);
        vn[g] = smem_copy(G.groups[g].vneg, cnt, cur);
    }
    if (threadIdx.x == 0) {
        S.e_g = e_g; S.e_c = e_c; S.e_gones = e_gones;
        S.r_lin_g = r_lin_g; S.r_lin_c = r_lin_c; S.r_lin_gones = r_lin_gones;
        S.e_con_ptr = e_con_ptr; S.e_con_pk = e_con_pk;
        S.r_lin_ptr = r_lin_ptr; S.r_lin_col = r_lin_col;
        S.r_con_ptr = r_con_ptr; S.r_con_pk = r_con_pk; S.e_bal = e_bal;
        S.dev_g = dev_g; S.dev_i = dev_i;
        if (G.sd) { S.sd_ptr = sd_ptr; S.sd_prog = sd_prog; S.sd_fix = sd_fix; }
        for (int g = 0; g < G.ngroups; ++g) { S.groups[g].vpos = vp[g]; S.groups[g].vneg = vn[g]; }
    }
    __syncthreads();
}

#include "engine_sd8.cuh"

template <int MS, int T>
__device__ __forceinline__ void run_tile(TileTeam<T>& tm, const cdn_engine_args& A,
                                         const cdn_plan_dev& P, int b, const Ws& w, bool active) {
    run_instance<MS>(tm, A, P, b, w, active);
}
template <int MS>
__device__ __forceinline__ void run_tile(TileTeam<8>& tm, const cdn_engine_args& A,
                                         const cdn_plan_dev& P, int b, const Ws& w, bool active) {
    if (w.sd_plan) sd8_run<MS>(tm, A, P, b, w, active);
    else run_instance<MS>(tm, A, P, b, w, active);
}

// ---- kernels --------------------------------------------------------------------------------
// Batches of small circuits: T lanes per instance; the state lives in shared
// memory unless the caller wants it kept (A.ws != NULL).
#ifndef CDN_TILE_MAXT
#define CDN_TILE_MAXT 256
#endif
#ifndef CDN_TILE_MINB
#define CDN_TILE_MINB 1
#endif
template <int MS, int T>
__global__ void __launch_bounds__(CDN_TILE_MAXT, CDN_TILE_MINB)
engine_tile_kernel(const __grid_constant__ cdn_engine_args A) {
    extern __shared__ double smem[];
    TileTeam<T> tm;
    const int tiles = blockDim.x / T;
    const int tile_id = threadIdx.x / T;
    // The gather lists of a dense plan are a few KB and are read by every
    // instance in every Newton iteration: keep a copy in shared memory (the
    // per-instance device records streaming through L1 evict the global copy).
    __shared__ cdn_plan_dev Ps;
    __shared__ Sd8Plan sdp;
    const bool shadow = A.plan_smem > 0;
    if (shadow) {
        const long per0 = ws_doubles(A.plan);
        plan_to_smem(A.plan, Ps, (char*)(smem + (long)tiles * per0));
    }
    const cdn_plan_dev& P = shadow ? Ps : A.plan;
    // small dense plans on 8-lane tiles: register-resident Newton linear algebra
    // (sd8_*), which addresses the shared-memory plan copy directly
    const bool use_sd8 = (T == 8) && shadow && P.sd;
    if (use_sd8) {
        if (threadIdx.x == 0) sd8_plan_init(sdp, P);
        __syncthreads();
    }
    const long per = ws_doubles(P);
    // the loop bounds are CTA-uniform: padding tiles run along inactive
    for (int b0 = blockIdx.x * tiles; b0 < A.B; b0 += gridDim.x * tiles) {
        const int b = b0 + tile_id;
        const bool active = b < A.B;
        // always shared memory (callers that keep the state between launches are
        // routed to the block kernel): with the address space known at compile
        // time the state is accessed with LDS/STS, not generic loads and stores
        double* base = smem + (long)tile_id * per;
        Ws w = ws_carve(base, P);
        if (use_sd8) { w.sd_plan = saddr(&sdp); w.sd_base = saddr(base); }
        if (active) {
            for (int r = tm.rank(); r < P.n; r += T) w.x[r] = A.x[(long)b * P.n + r];
            // (small dense plans accumulate device contributions per slot in
            // w.lu: slots without any stay zero; so does the padding source)
            if (P.sd) {
                for (int e = tm.rank(); e < P.nlu; e += T) w.lu[e] = 0.0;
                if (tm.rank() == 0) w.fact[CDN_SD8_PART + CDN_SD8_NFIX] = 0.0;
            }
            tm.sync();
        }
        run_tile<MS>(tm, A, P, b, w, active);
        if (active) {
            tm.sync();
            for (int r = tm.rank(); r < P.n; r += T) A.x[(long)b * P.n + r] = w.x[r];
            tm.sync();
        }
    }
}

// One CTA per circuit instance, state in global memory (L2).  When the caller
// provides enough dynamic shared memory (A.smem_lu), the LU values and the two
// work vectors of the triangular solves live there for the whole launch: the
// level-scheduled factorisation and solves are chains of short dependent
// loads, and their latency, not bandwidth, is what a time step costs.
template <int MS, int TPB>
__global__ void __launch_bounds__(TPB)
engine_block_kernel(const __grid_constant__ cdn_engine_args A) {
    extern __shared__ double dsm[];
    __shared__ double sred[64];
    __shared__ int ired[64];
    const cdn_plan_dev& P = A.plan;
    BlockTeam tm;
    tm.sred = sred; tm.ired = ired;
    for (int b = blockIdx.x; b < A.B; b += gridDim.x) {
        Ws w = ws_carve(A.ws + (long)b * A.ws_stride, P);
        double* const g_lu = w.lu;
        if (A.smem_lu) {
            w.lu = dsm; w.tmp = dsm + P.nlu;
            w.dprof = A.prof;
            if (A.smem_lu > 1) w.stage = stage_area(dsm, P);   // 16-byte aligned
            for (int e = tm.rank(); e < P.nlu; e += tm.size()) w.lu[e] = g_lu[e];
        }
        if (A.op == CDN_OP_NEWTON || A.op == CDN_OP_TRAN) {
            for (int r = tm.rank(); r < P.n; r += tm.size()) w.x[r] = A.x[(long)b * P.n + r];
        }
        tm.sync();
        run_instance<MS>(tm, A, P, b, w, true);
        tm.sync();
        for (int r = tm.rank(); r < P.n; r += tm.size()) A.x[(long)b * P.n + r] = w.x[r];
        if (A.smem_lu)
            for (int e = tm.rank(); e < P.nlu; e += tm.size()) g_lu[e] = w.lu[e];
        tm.sync();
    }
}

// The whole (cooperative) grid on one circuit.
template <int MS, int TPB>
__global__ void __launch_bounds__(TPB)
engine_grid_kernel(const __grid_constant__ cdn_engine_args A) {
    __shared__ double sred[64];
    const cdn_plan_dev& P = A.plan;
    GridTeam tm;
    tm.sred = sred; tm.gred = A.red;
    const Ws w = ws_carve(A.ws, P);
    if (A.op == CDN_OP_NEWTON || A.op == CDN_OP_TRAN) {
        for (int r = tm.rank(); r < P.n; r += tm.size()) w.x[r] = A.x[r];
        tm.sync();
    }
    run_instance<MS>(tm, A, P, 0, w, true);
    tm.sync();
    for (int r = tm.rank(); r < P.n; r += tm.size()) A.x[r] = w.x[r];
}
