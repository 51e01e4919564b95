// Small dense systems (dense plan, n <= 8) on 8-lane warp tiles: the Newton /
// time loop of batches of small circuits (Monte Carlo, corner sweeps) with the
// linear algebra of an instance in the registers of its tile, one matrix ROW
// per lane (lane r: row r of [J | y]).  Included by engine.cuh.
//
//   * the device-Jacobian contributions are summed from flat, equally long
//     per-lane term lists (plan.h sd_prog) into w.lu;
//   * lane r adds the linear part G' = G + a0 C (+ gmin patterns) of its row,
//     which at the same time gives the linear part of the residual, G' x;
//   * Gaussian elimination with partial pivoting (LAPACK getrf arithmetic:
//     largest |a_ik| of the remaining rows, multipliers by the reciprocal of
//     the pivot, a zero pivot leaves its column unscaled) runs on the registers
//     with the right-hand side carried along; rows are not moved -- the lane
//     that holds the pivot row of step k is remembered instead (`perm`);
//   * back substitution, damping and the convergence test follow in registers;
//   * multipliers, U rows and reciprocal pivots are stored (w.fact) so that the
//     chord predictor of the next time step (tran.py:181) re-uses them.
// Ties between equal pivot candidates go to the lowest lane (LAPACK: lowest
// current row position); values agree with getrf/getrs to rounding.
//
// Control flow is uniform over each WARP (four tiles): when any of its tiles
// has work, all four run the linear algebra and only the live ones commit
// their stores.  The tile-local shuffles, votes and barriers can then name the
// whole warp (constant mask, segments of 8 lanes); with per-tile masks ptxas
// guards every one of them with a MATCH/REDUX/VOTE sequence and runs the four
// tiles of a warp one after another.  Device evaluation has no collectives and
// is simply skipped by tiles without work.
//
// Reference functions restated: see the header of engine.cuh (fsolve_Newton,
// solve + helper chain, factor_and_solve with condition_array, integration,
// time loop with chord predictor).
#pragma once

struct Sd8 {
    double row[8];
    double y;        // right-hand side entry, then solution of unknown `step`
    double pinv;     // 1 / pivot of the step at which this lane held the pivot row
    int step;        // that step = the unknown this lane solves for (8: none yet)
    unsigned perm;   // lane holding the pivot row of step k in bits 4k .. 4k + 2
};

#define SD8_FULL 0xffffffffu
// experiment: evaluate q(x) of the accept step with the Jacobian variant of the
// device code (one code path in the hot loop: instruction-cache footprint)
#ifndef SD8_ACCEPT_JAC
#define SD8_ACCEPT_JAC true
#endif
__device__ __forceinline__ double sd8_shfl(double v, int src) { return __shfl_sync(SD8_FULL, v, src, 8); }
__device__ __forceinline__ int sd8_shfl(int v, int src) { return __shfl_sync(SD8_FULL, v, src, 8); }
__device__ __forceinline__ double sd8_shfl_xor(double v, int off) {
    return __shfl_xor_sync(SD8_FULL, v, off, 8);
}
__device__ __forceinline__ int sd8_shfl_xor(int v, int off) {
    return __shfl_xor_sync(SD8_FULL, v, off, 8);
}
// p holds on all eight lanes of this tile
__device__ __forceinline__ bool sd8_all(bool p) {
    const unsigned bal = __ballot_sync(SD8_FULL, p);
    return ((bal >> (threadIdx.x & 24u)) & 0xffu) == 0xffu;
}
__device__ __forceinline__ bool sd8_warp_any(bool p) { return __any_sync(SD8_FULL, p) != 0; }
// Loop control: the warps of a CTA iterate in lockstep (every thread of the CTA
// passes the same sequence of sd8_any calls), so that they walk the ~110 KB of
// hot code together and share instruction fetches.  Measured on B200, 4096
// ring_bsim3 instances: 27.9 ms in lockstep, 49.8 ms with independent warps
// (SD8_LOCKSTEP 0), although a warp then only waits for the slowest of its own
// four instances instead of the slowest of the CTA's 28.
#ifndef SD8_LOCKSTEP
#define SD8_LOCKSTEP 1
#endif
__device__ __forceinline__ bool sd8_any(TileTeam<8>& tm, bool p) {
#if SD8_LOCKSTEP
    return tm.cta_any(p);
#else
    return sd8_warp_any(p);
#endif
}

// The state and the plan copy of these kernels live in shared memory.  They
// are addressed with explicit ld/st.shared on 32-bit addresses: the generic
// pointers of `Ws` / `cdn_plan_dev` lose their address space once the structs
// are spilled (255 registers), and every access became a generic LD.E/ST.E
// behind a pointer re-load.
__device__ __forceinline__ unsigned saddr(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ double lds_f64(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned lds_u32(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ double flip_sign(double v, unsigned signbit) {   // 0 / 0x80000000
    return __hiloint2double(__double2hiint(v) ^ (int)signbit, __double2loint(v));
}

// shared-memory addresses of the plan copy and byte offsets of the state
// vectors inside one instance's state (one per CTA, filled by sd8_plan_init)
struct Sd8Plan {
    unsigned e_g, e_c, e_ee, rptr, rpk, sptr, sprog, sfix;
    unsigned lu, out, fact;          // byte offsets from the instance base (x)
    int n, nlu, maxlen, rmax, nfix;
};

__device__ __forceinline__ void sd8_plan_init(Sd8Plan& sp, const cdn_plan_dev& P) {
    sp.e_g = saddr(P.e_g); sp.e_c = saddr(P.e_c); sp.e_ee = saddr(P.e_gones);
    sp.rptr = saddr(P.r_con_ptr); sp.rpk = saddr(P.r_con_pk);
    sp.sptr = saddr(P.sd_ptr); sp.sprog = saddr(P.sd_prog); sp.sfix = saddr(P.sd_fix);
    sp.lu = 8u * 10u * (unsigned)P.n;
    sp.out = sp.lu + 8u * (unsigned)P.nlu;
    sp.fact = sp.out + 8u * (unsigned)(P.dev_out_size + P.dev_jac_size + P.hist_size);
    sp.n = P.n; sp.nlu = P.nlu; sp.maxlen = P.sd_maxlen; sp.rmax = P.sd_rmax; sp.nfix = P.sd_nfix;
}

__device__ __forceinline__ Sd8Plan sd8_plan_load(unsigned a) {
    Sd8Plan sp;
    unsigned* d = (unsigned*)&sp;
#pragma unroll
    for (int k = 0; k < (int)(sizeof(Sd8Plan) / 4); ++k) d[k] = lds_u32(a + 4u * k);
    return sp;
}

// state vectors (engine.cuh ws_carve): x 0, dx n, ivec 2n, sv 3n, q 4n, qprev 5n,
// dq 6n, y 7n, tmp 8n, xb 9n
#define SD8_VEC(base, n, k, r) ((base) + 8u * (unsigned)((k) * (n) + (r)))
#define SD8_X(base, n, r) SD8_VEC(base, n, 0, r)
#define SD8_DX(base, n, r) SD8_VEC(base, n, 1, r)
#define SD8_IVEC(base, n, r) SD8_VEC(base, n, 2, r)
#define SD8_SV(base, n, r) SD8_VEC(base, n, 3, r)
#define SD8_Q(base, n, r) SD8_VEC(base, n, 4, r)
#define SD8_QPREV(base, n, r) SD8_VEC(base, n, 5, r)
#define SD8_DQ(base, n, r) SD8_VEC(base, n, 6, r)
#define SD8_Y(base, n, r) SD8_VEC(base, n, 7, r)
#define SD8_XB(base, n, r) SD8_VEC(base, n, 9, r)

// lu[e] = sum of the device-Jacobian contributions of slot e.  Every lane
// walks `maxlen` terms (padded with terms that add an always-zero location):
//   term = source (doubles from the instance base) | target << 14 |
//          store-and-reset << 30 | minus << 31
__device__ __forceinline__ void sd8_gather(const Sd8Plan& sp, unsigned base) {
    const int l = threadIdx.x & 7;
    const int maxlen = sp.maxlen, nfix = sp.nfix;
    unsigned pa = sp.sprog + 4u * (unsigned)(l * maxlen);
    double acc = 0.0;
    unsigned pk = (maxlen > 0) ? lds_u32(pa) : 0u;
    for (int t = 0; t < maxlen; ++t) {
        pa += 4u;
        const unsigned nx = (t + 1 < maxlen) ? lds_u32(pa) : 0u;    // one trip ahead
        const double v = lds_f64(base + 8u * (pk & 0x3fffu));
        acc += flip_sign(v, pk & 0x80000000u);
        if (pk & (1u << 30)) {
            sts_f64(base + 8u * ((pk >> 14) & 0x3fffu), acc);
            acc = 0.0;
        }
        pk = nx;
    }
    __syncwarp();
    if (nfix) {
        // partial accumulators of slots cut between two lanes
        if (l == 0)
            for (int j = 0; j < nfix; ++j) {
                const unsigned tgt = base + sp.lu + 8u * lds_u32(sp.sfix + 4u * j);
                const unsigned src = base + sp.fact + 8u * (unsigned)(CDN_SD8_PART + j);
                sts_f64(tgt, lds_f64(tgt) + lds_f64(src));
            }
        __syncwarp();
    }
}

// Rows of J and the residual: S.row = G' row + device part, ivec = i(x),
// S.y = lambda sv - i(x) (ivec and y are stored if `commit`).  Returns whether
// this lane's row and rhs are finite.
__device__ __forceinline__ bool sd8_rows(const Sd8Plan& sp, unsigned base, const cdn_opts& o,
                                         Sd8& S, bool want_jac, bool commit) {
    const int n = sp.n, l = threadIdx.x & 7;
    const int r = l < n ? l : n - 1;              // spare lanes shadow the last row
    const double a0 = o.use_q ? o.a0 : 0.0;
    const unsigned ro = 8u * (unsigned)(r * n);
    const unsigned eg = sp.e_g + ro, ec = sp.e_c + ro, ee = sp.e_ee + ro, lu = base + sp.lu + ro;
    const bool use_q = o.use_q != 0, use_g = o.gmin != 0.0;
    const bool use_ext = o.gmin_ext != 0.0 && r < o.n_ext;
    // device-output list of this row
    const unsigned c0 = lds_u32(sp.rptr + 4u * r), c1 = lds_u32(sp.rptr + 4u * r + 4u);
    double lin = 0.0;
    bool fin = true;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double v = 0.0;
        if (c < n) {
            double b = lds_f64(eg + 8u * c);
            if (use_q) b = fma(a0, lds_f64(ec + 8u * c), b);
            if (use_g) b = fma(o.gmin, lds_f64(ee + 8u * c), b);
            if (use_ext && c == r) b += o.gmin_ext;               // nodal.py:478-479
            lin = fma(b, lds_f64(SD8_X(base, n, c)), lin);
            if (want_jac) {
                v = lds_f64(lu + 8u * c) + b;
                fin = fin && isfinite(v);
            }
        }
        S.row[c] = v;
    }
    // device currents and charges of this row: +-i and a0 * (+-q) on separate
    // accumulators; every lane runs rmax trips (past its end a lane re-reads its
    // first entry without adding it)
    const int len = (int)(c1 - c0), rmax = sp.rmax;
    const unsigned rpk = sp.rpk + 4u * c0, out = base + sp.out;
    double ai = 0.0, aq = 0.0;
    for (int t = 0; t < rmax; ++t) {
        const bool live = t < len;
        const unsigned pk = lds_u32(rpk + 4u * (unsigned)(live ? t : 0));
        const double v = lds_f64(out + 8u * (pk & 0x1fffffffu));
        const double sv = flip_sign(v, (pk << 2) & 0x80000000u);   // bit 29: minus
        const bool chg = (pk & (1u << 30)) != 0u;
        if (live && chg) aq += sv;
        if (live && !chg) ai += sv;
    }
    const double acc = (lin + ai) + a0 * aq;
    const double y = o.lambda * lds_f64(SD8_SV(base, n, r)) - acc;
    if (commit && l < n) { sts_f64(SD8_IVEC(base, n, l), acc); sts_f64(SD8_Y(base, n, l), y); }
    S.y = y;
    return (l >= n) || (fin && isfinite(y));
}

// q = C x + Tq q(x) of this lane's row (device outputs must be current)
__device__ __forceinline__ void sd8_charge(const Sd8Plan& sp, unsigned base, bool commit) {
    const int n = sp.n, l = threadIdx.x & 7;
    const int r = l < n ? l : n - 1;
    const unsigned ec = sp.e_c + 8u * (unsigned)(r * n);
    const unsigned c0 = lds_u32(sp.rptr + 4u * r), c1 = lds_u32(sp.rptr + 4u * r + 4u);
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (c < n) acc = fma(lds_f64(ec + 8u * c), lds_f64(SD8_X(base, n, c)), acc);
    const int len = (int)(c1 - c0), rmax = sp.rmax;
    const unsigned rpk = sp.rpk + 4u * c0, out = base + sp.out;
    double aq = 0.0;
    for (int t = 0; t < rmax; ++t) {
        const bool live = t < len;
        const unsigned pk = lds_u32(rpk + 4u * (unsigned)(live ? t : 0));
        const double v = lds_f64(out + 8u * (pk & 0x1fffffffu));
        if (live && (pk & (1u << 30))) aq += flip_sign(v, (pk << 2) & 0x80000000u);
    }
    if (commit && l < n) sts_f64(SD8_Q(base, n, l), acc + aq);
    __syncwarp();
}

// Elimination with the right-hand side carried along (forward substitution fused).
__device__ __forceinline__ void sd8_factor(Sd8& S, int n) {
    const int l = threadIdx.x & 7;
    S.step = 8;
    S.perm = 0u;
    S.pinv = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < n) {
            const bool cand = (l < n) && (S.step == 8);
            // argmax over the tile in one 64-bit key per lane: |a_ik| with its
            // three lowest mantissa bits replaced by 7 - lane, so that the
            // larger key is the larger magnitude and, among magnitudes equal to
            // 2^-49 relative, the lower lane (keys of non-negative doubles order
            // like the doubles themselves); NaN loses against any number,
            // lanes without a candidate row lose against everything
            double v = fabs(S.row[k]);
            if (!(v >= 0.0)) v = -1.0;
            if (!cand) v = -2.0;
            v = __hiloint2double(__double2hiint(v), (__double2loint(v) & ~7) | (7 - l));
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) v = fmax(v, sd8_shfl_xor(v, off));
            const int p = 7 - (__double2loint(v) & 7);
            S.perm |= (unsigned)p << (4 * k);
            const double pv = sd8_shfl(S.row[k], p);
            // reciprocal of the pivot: MUFU seed + two Newton steps (dual.cuh
            // cdn_rcp: within an ulp of 1 / pv, inf for a zero pivot, branch-free)
            const double pinv = cdn_rcp(pv);
            const double m = S.row[k] * ((pv != 0.0) ? pinv : 1.0);   // dgetf2: zero pivot
            const bool elim = cand && (l != p);
            if (l == p) { S.step = k; S.pinv = pinv; }
            if (elim) S.row[k] = m;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j > k && j < n) {
                    const double pr = sd8_shfl(S.row[j], p);
                    if (elim) S.row[j] = fma(-m, pr, S.row[j]);
                }
            const double py = sd8_shfl(S.y, p);
            if (elim) S.y = fma(-m, py, S.y);
        }
    }
}

// forward substitution of a fresh right-hand side with stored factors
__device__ __forceinline__ void sd8_forward(Sd8& S, int n) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < n) {
            const double yk = sd8_shfl(S.y, (int)((S.perm >> (4 * k)) & 7u));
            if (S.step > k && S.step < 8) S.y = fma(-S.row[k], yk, S.y);
        }
}

// U x = y; afterwards the lane that held the pivot row of step k holds x_k
__device__ __forceinline__ void sd8_backward(Sd8& S, int n) {
    const int l = threadIdx.x & 7;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        const int k = 7 - kk;
        if (k < n) {
            const int p = (int)((S.perm >> (4 * k)) & 7u);
            const double xk = sd8_shfl(S.y * S.pinv, p);
            if (S.step < k) S.y = fma(-S.row[k], xk, S.y);
            if (l == p) S.y = xk;
        }
    }
}

// x_l for lane l (unknown l sits in the lane that held the pivot row of step l)
__device__ __forceinline__ double sd8_unknown(const Sd8& S) {
    const int l = threadIdx.x & 7;
    return sd8_shfl(S.y, (int)((S.perm >> (4 * l)) & 7u));
}

__device__ __forceinline__ void sd8_store(const Sd8& S, const Sd8Plan& sp, unsigned base) {
    const int l = threadIdx.x & 7;
    const unsigned f = base + sp.fact;
#pragma unroll
    for (int c = 0; c < 8; ++c) sts_f64(f + 8u * (unsigned)(l * 8 + c), S.row[c]);
    sts_f64(f + 8u * (unsigned)(64 + l), S.pinv);
    sts_u32(f + 8u * 72u + 4u * (unsigned)l, (unsigned)S.step);
    if (l == 0) sts_u32(f + 8u * 72u + 32u, S.perm);
}

__device__ __forceinline__ void sd8_load(Sd8& S, const Sd8Plan& sp, unsigned base) {
    const int l = threadIdx.x & 7;
    const unsigned f = base + sp.fact;
#pragma unroll
    for (int c = 0; c < 8; ++c) S.row[c] = lds_f64(f + 8u * (unsigned)(l * 8 + c));
    S.pinv = lds_f64(f + 8u * (unsigned)(64 + l));
    S.step = (int)lds_u32(f + 8u * 72u + 4u * (unsigned)l);
    S.perm = lds_u32(f + 8u * 72u + 32u);
}

__device__ __forceinline__ double sd8_max(double v) {
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) v = fmax(v, sd8_shfl_xor(v, off));
    return v;
}

// dx = LU \ (sv - iVec) with the stored factors; x += dx (tran.py:181)
__device__ __forceinline__ void sd8_chord(const Sd8Plan& sp, unsigned base, bool commit) {
    const int n = sp.n, l = threadIdx.x & 7;
    const int r = l < n ? l : n - 1;
    Sd8 S;
    sd8_load(S, sp, base);
    S.y = (l < n) ? lds_f64(SD8_SV(base, n, r)) - lds_f64(SD8_IVEC(base, n, r)) : 0.0;
    sd8_forward(S, n);
    sd8_backward(S, n);
    const double d = sd8_unknown(S);
    if (commit && l < n) {
        sts_f64(SD8_DX(base, n, l), d);
        sts_f64(SD8_X(base, n, l), lds_f64(SD8_X(base, n, l)) + d);
    }
    __syncwarp();
}

// device sweep of one instance by its tile (no collectives inside)
template <int MS, bool DERIV>
__device__ __forceinline__ void sd8_eval(const cdn_plan_dev& P, int b, const Ws& w, double h,
                                         double qs) {
    for (int d = threadIdx.x & 7; d < P.ndev; d += 8)
        eval_one<MS, DERIV>(P.groups[CDN_LD(P.dev_g + d)], b, CDN_LD(P.dev_i + d), w.x, w, 0, h, 0,
                            qs);
}

// ---- damped Newton (fsolve.py:59-128; engine.cuh newton() is the general version) -------
// Every thread of the warp (of the CTA with SD8_LOCKSTEP) calls this the same number of
// times.
template <int MS>
__device__ __forceinline__ int sd8_newton(TileTeam<8>& tm, const cdn_plan_dev& P, int b,
                                          const Ws& w, const cdn_opts& o, bool active,
                                          double* res_out, bool* ok, long long* prof) {
    const int n = P.n, l = threadIdx.x & 7;
    const unsigned base = w.sd_base;
    double res = 0.0;
    bool success = false;
    int it = 0;
    bool run = active;
    while (sd8_any(tm, run && it < o.maxiter)) {
        bool live = run && it < o.maxiter;
        if (!sd8_warp_any(live)) continue;              // warp-uniform from here on
        CDN_PROF_T0;
        if (live) sd8_eval<MS, true>(P, b, w, o.h, o.use_q ? o.a0 : 1.0);
        __syncwarp();
        CDN_PROF(0);
        const Sd8Plan sp = sd8_plan_load(w.sd_plan);
        sd8_gather(sp, base);
        Sd8 S;
        const bool fin = sd8_all(sd8_rows(sp, base, o, S, true, live));
        if (live && !fin) {
            // NaN/inf in the Jacobian or right-hand side: this Newton solve has
            // failed (oracle/nodal_ref.py factor_and_solve explains the
            // choice); next helper or bisection step
            ++it;
            res = INFINITY;
            run = false;
            live = false;
        }
        CDN_PROF(1);
        sd8_factor(S, n);
        if (live) sd8_store(S, sp, base);
        CDN_PROF(2);
        sd8_backward(S, n);
        double d = sd8_unknown(S);
        if (!isfinite(d))               // condition_array (nodal.py:385-402)
            d = isnan(d) ? 10.0 * o.abstol : (d > 0.0 ? 0.1 : -0.1) * o.maxdelta;
        if (l >= n) d = 0.0;
        CDN_PROF(3);
        const double m = sd8_max(fabs(d));
        const bool damp = m > o.maxdelta;
        const double scale = damp ? o.maxdelta / m : 1.0;
        if (damp) d *= scale;
        bool conv = true;
        if (l < n) {
            const double xo = lds_f64(SD8_X(base, n, l)), xn = xo + d;
            conv = fabs(d) < o.reltol * fmax(fabs(xo), fabs(xn)) + o.abstol;
            if (live) {
                sts_f64(SD8_X(base, n, l), xn);
                sts_f64(SD8_DX(base, n, l), d);
            }
        }
        if (live) res = damp ? m * scale : m;
        const bool n1 = sd8_all(conv);
        __syncwarp();
        bool n2 = true;
        if (o.errfunc) {
            const bool chk = live && n1;
            if (sd8_warp_any(chk)) {
                if (chk) sd8_eval<MS, false>(P, b, w, o.h, 1.0);
                __syncwarp();
                Sd8 R;
                sd8_rows(sp, base, o, R, false, chk);
                const double ef = sd8_max(l < n ? fabs(R.y) : 0.0);
                __syncwarp();
                if (chk) {
                    n2 = ef < o.abstol;
                    res = fmax(ef, res);
                }
            }
        }
        if (live) {
            ++it;
            if (n1 && n2) { success = true; run = false; }
        }
        CDN_PROF(4);
    }
    *res_out = res;
    *ok = success;
    return it;
}

// Continuation on lambda (nodal.py:545-604; engine.cuh homotopy() is the general version)
template <int MS>
__device__ __noinline__ int sd8_homotopy(TileTeam<8>& tm, const cdn_plan_dev& P, int b,
                                         const Ws& w, cdn_opts o, int mode, bool active,
                                         double* res_out, bool* ok) {
    const int n = P.n, l = threadIdx.x & 7;
    const unsigned base = w.sd_base;
    double stack[48];
    int sp = 0;
    stack[sp++] = 0.0;
    stack[sp++] = 1.0;
    double lam = 0.5, step = 0.5, res = 0.0;
    const double small = 1e-4;
    bool success = true;
    int tot = 0;
    if (active && l < n) sts_f64(SD8_X(base, n, l), lds_f64(SD8_XB(base, n, l)));
    __syncwarp();
    bool run = active;
    while (sd8_any(tm, run && sp > 0)) {
        const bool go = run && sp > 0;
        if (mode == 0) o.gmin = 1e-3 / lam - 1e-3;
        else if (mode == 1) o.lambda = lam;
        else o.gmin_ext = 1e-3 / lam - 1e-3;
        bool ok1 = false;
        tot += sd8_newton<MS>(tm, P, b, w, o, go, &res, &ok1, nullptr);
        if (go) {
            if (ok1) {
                if (l < n) sts_f64(SD8_XB(base, n, l), lds_f64(SD8_X(base, n, l)));
                step = stack[sp - 1] - lam;
                lam = stack[--sp];
            } else {
                if (l < n) sts_f64(SD8_X(base, n, l), lds_f64(SD8_XB(base, n, l)));
                if (sp < 48) stack[sp++] = lam;
                step *= 0.5;
                lam -= step;
                if (lam < small || step < small) { success = false; run = false; }
            }
        }
        __syncwarp();
    }
    *res_out = res;
    *ok = success && active;
    return tot;
}

// fsolve.solve over the dense engine's helper list (engine.cuh solve_helpers)
template <int MS>
__device__ __forceinline__ int sd8_solve_helpers(TileTeam<8>& tm, const cdn_plan_dev& P, int b,
                                                 const Ws& w, const cdn_opts& o, int helpers,
                                                 bool active, double* res_out, bool* ok,
                                                 int* used, long long* prof) {
    const int n = P.n, l = threadIdx.x & 7;
    const unsigned base = w.sd_base;
    *used = 0;
    if (helpers && active && l < n) sts_f64(SD8_XB(base, n, l), lds_f64(SD8_X(base, n, l)));
    __syncwarp();
    int it = 0;
    *ok = false;
    *res_out = 0.0;
    if (!(helpers & 2)) it = sd8_newton<MS>(tm, P, b, w, o, active, res_out, ok, prof);
    if (!helpers) return it;
    bool need = active && !*ok;
    if (!sd8_any(tm, need)) return it;
    double res2 = 0.0;
    bool ok2 = false;
    int it2;
    if (!(helpers & 4)) {
        it2 = sd8_homotopy<MS>(tm, P, b, w, o, 0, need, &res2, &ok2);
        if (need) { it = it2; *res_out = res2; *ok = ok2; *used = 1; }
        need = need && !ok2;
        if (!sd8_any(tm, need)) return it;
    }
    if (!(helpers & 8)) {
        it2 = sd8_homotopy<MS>(tm, P, b, w, o, 1, need, &res2, &ok2);
        if (need) { it = it2; *res_out = res2; *ok = ok2; *used = 2; }
        need = need && !ok2;
    }
    if (o.n_ext <= 0 || !sd8_any(tm, need)) return it;
    it2 = sd8_homotopy<MS>(tm, P, b, w, o, 2, need, &res2, &ok2);
    if (need) { it = it2; *res_out = res2; *ok = ok2; *used = 3; }
    return it;
}

// ---- one circuit instance: operating point or the whole time loop ------------------------
// (engine.cuh run_instance() is the general version; the tile kernels only run
// CDN_OP_NEWTON and CDN_OP_TRAN from fresh state)
template <int MS>
__device__ __forceinline__ void sd8_run(TileTeam<8>& tm, const cdn_engine_args& A,
                                        const cdn_plan_dev& P, int b, const Ws& w, bool active) {
    const cdn_opts& o = A.o;
    const int n = P.n, l = threadIdx.x & 7;
    const unsigned base = w.sd_base;
    if (A.op == CDN_OP_NEWTON) {
        if (active && l < n) sts_f64(SD8_SV(base, n, l), __ldg(A.sv_in + l));
        __syncwarp();
        double res; bool ok; int used;
        const int it = sd8_solve_helpers<MS>(tm, P, b, w, o, A.helpers, active, &res, &ok, &used,
                                             b == 0 ? A.prof : nullptr);
        if (active && l == 0) {
            A.iters[b] = it; A.res[b] = res; A.status[b] = ok ? 0 : -1;
            if (A.info) { A.info[2 * b] = used; A.info[2 * b + 1] = 0; }
        }
        return;
    }
    if (A.op != CDN_OP_TRAN) return;
    const Sd8Plan sp = sd8_plan_load(w.sd_plan);
    if (A.init_ic) {                               // set_IC (nodalSP.py:666-696)
        if (active) sd8_eval<MS, false>(P, b, w, 0.0, 1.0);
        __syncwarp();
        sd8_charge(sp, base, active);
        if (active && l < n) {                     // integration.py init
            sts_f64(SD8_QPREV(base, n, l), lds_f64(SD8_Q(base, n, l)));
            sts_f64(SD8_DQ(base, n, l), 0.0);
        }
        __syncwarp();
    }
    int status = 0, used_max = 0;
    bool run = active;
    long long* prof = (b == 0) ? A.prof : nullptr;
    for (int i = A.first; i < A.nsamples; ++i) {
        if (!sd8_any(tm, run)) break;
        if (sd8_warp_any(run)) {
            CDN_PROF_T0;
            // accept(xOld): q(x) with a values-only sweep (nodalSP.py:719-758),
            // integration.py:36-38 / :83-91, then sv = f_n1() + s (:47-49, :93-97)
            if (run) sd8_eval<MS, SD8_ACCEPT_JAC>(P, b, w, 0.0, 1.0);
            __syncwarp();
            sd8_charge(sp, base, run);
            if (run && l < n) {
                const double q = lds_f64(SD8_Q(base, n, l)), qp = lds_f64(SD8_QPREV(base, n, l));
                double dq = lds_f64(SD8_DQ(base, n, l));
                if (o.im == 1) {
                    dq = -dq + o.a0 * (q - qp);
                    sts_f64(SD8_DQ(base, n, l), dq);
                }
                sts_f64(SD8_QPREV(base, n, l), q);
                double f = o.a0 * q;
                if (o.im == 1) f += dq;
                f += A.src_dc ? __ldg(A.src_dc + l) : 0.0;
                const double* sval = A.src_val + (long)i * A.nsrc;
                for (int k = 0; k < A.nsrc; ++k)
                    if (__ldg(A.src_rows + k) == l) f += __ldg(sval + k);
                sts_f64(SD8_SV(base, n, l), f);
            }
            __syncwarp();
            CDN_PROF(5);
            if (i > 1) sd8_chord(sp, base, run);    // chord predictor (tran.py:181)
            CDN_PROF(6);
        }
        double res; bool ok; int used;
        const int it = sd8_solve_helpers<MS>(tm, P, b, w, o, A.helpers, run, &res, &ok, &used, prof);
        if (!run) continue;
        if (used > used_max) used_max = used;
        const long at = (long)b * A.nsamples + i;
        if (l == 0) { A.iters[at] = it; A.res[at] = res; }
        if (!ok) { status = i; run = false; continue; }
        for (int k = l; k < A.nsave; k += 8)
            A.wave[at * A.nsave + k] = lds_f64(SD8_X(base, n, __ldg(A.save_rows + k)));
    }
    if (active && l == 0) {
        A.status[b] = status;
        if (A.info) { A.info[2 * b] = used_max; A.info[2 * b + 1] = 0; }
    }
}
