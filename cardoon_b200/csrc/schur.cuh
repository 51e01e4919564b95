// Boundary / interior transient kernel: the fixed-step time loop of ONE sparse
// circuit on a thread-block cluster (host analysis and the mathematics:
// cardoon_b200/analyses/schur.py; numpy emulation: tests/schur_emul.py).
//
// The reference refactors the whole Jacobian with SuperLU on every Newton
// iteration (nodalSP.py:281-303, :838-887).  With a fixed step only the entries
// stamped by nonlinear devices change, J = A + D(x), A = G + a0 C.  The unknowns
// are split into a small boundary B (device ports + cut nodes) that every CTA
// keeps a copy of, and interior pieces dealt to the CTAs.  Per CTA, in shared
// memory for the whole launch:
//   Linv, Uinv   explicit inverses of the LU factors of A[Ic, Ic]: an interior
//                solve is two sparse matrix-vector products, no dependency
//                levels and no barriers other than the two between them
//   X = A[Ic,Ic]^-1 A[Ic,B],  Asi = A[B,Ic],  Csi = C[B,Ic],  rows Ic of A and C
// and replicated: A[B,B], C[B,B], the Schur complement T0 (band storage) and the
// gather lists of device outputs / Jacobians into the boundary system.
//
// One Newton iteration (fsolve.py:99-126 semantics, engine.cuh newton()):
//   device sweep (every CTA, redundantly: the boundary state is replicated and
//   must stay bit-identical) -> ivec_I = A_c x, y_I = sv_I - ivec_I ->
//   t_I = Uinv Linv y_I -> partial sums Asi x_I, Asi t_I -> ONE exchange over
//   distributed shared memory -> r_B, T = T0 + D(x), banded LU without pivoting,
//   dx_B -> dx_I = t_I - X dx_B -> max|dx| / convergence: second exchange.
// The chord predictor (tran.py:181) re-uses the last factors of T and the last
// residual, with one exchange.
//
// What the time goes to is latency, so the kernel is organised around it
// (measurements: profiles/README.md, tools/fp64_latency.cu):
//   * the device sweep runs on the last warps of the CTA beside the interior
//     work of the same phase (named barriers), which never needs device outputs;
//   * operators applied one thread per row are stored as sliced ELLPACK, the
//     long rows of Linv are summed by a warp each;
//   * the cluster barrier is split into arrive / wait around the work that
//     needs no remote data; x is updated on the undamped assumption before the
//     convergence exchange;
//   * the boundary LU runs on one thread in a register window, its triangular
//     sweeps in blocks of eight rows with all loads ahead of the dependent chain.
//
// Anything this kernel cannot do (Newton failure -> helper chain, a pivot of T
// that degrades) is reported through status / info and the host re-runs the
// analysis on the general kernels (nodalCUDA.Engine.tran).
#pragma once
#include "engine.cuh"

#define CDN_SCHUR_MAXCTA 16
#define CDN_SCHUR_NOPS 12
#define CDN_SCHUR_LONG 32     // rows of Linv longer than this are summed by a warp
enum { SOP_LINV = 0, SOP_UINV, SOP_X, SOP_ASI, SOP_CSI, SOP_AROW, SOP_CROW, SOP_ABB, SOP_CBB,
       SOP_DCUR, SOP_DCHG, SOP_TJAC };
// blob header (32-bit words): see schur.py pack_blob
enum { SH_M = 0, SH_NB, SH_BAND, SH_OUT, SH_JAC, SH_HASCSI, SH_LLONG, SH_NLONG, SH_OPS = 8,
       SH_TB = SH_OPS + 4 * CDN_SCHUR_NOPS, SH_IC, SH_B, SH_BYTES, SH_ELL = 64, SH_WORDS = 96 };
// sliced-ELLPACK copies (schur.py Ell) of the operators applied one thread per row
enum { EOP_LINV = 0, EOP_UINV, EOP_X, EOP_AROW };

struct cdn_schur_args {
    cdn_group_dev groups[CDN_MAX_GROUPS];   // vpos / vneg re-mapped to boundary positions
    int ngroups, ndev;
    const int* dev_g;
    const int* dev_i;
    cdn_opts o;
    const unsigned char* blob;              // per-CTA operator blobs
    long cta_off[CDN_SCHUR_MAXCTA + 1];
    int ncta, nB;
    double* x;                              // [n] operating point in, last sample out
    int nsamples, nsrc, nsave;
    const int* src_own;                     // [nsrc] owner CTA of the row (-1: boundary)
    const int* src_loc;                     // [nsrc] position in the owner's local vector
    const double* src_val;                  // [nsamples][nsrc]
    const double* src_dc;                   // [n]
    const int* save_ptr;                    // [ncta + 1]
    const int* save_k;                      // column of wave
    const int* save_loc;                    // position in the CTA's local vector
    double* wave;                           // [nsamples][nsave]
    int* iters;
    double* res;
    int* status;
    int* info;                              // [2]: helper used (0), pivot trouble
    long long* prof;                        // optional [16] phase cycles (CTA 0)
};

struct SOp {
    int nrows;
    const int* ptr;
    const unsigned short* col;
    const double* val;
};

__device__ __forceinline__ SOp sop_get(const unsigned char* blob, int k) {
    const int* h = (const int*)blob + SH_OPS + 4 * k;
    SOp o;
    o.nrows = h[0];
    o.ptr = (const int*)(blob + h[1]);
    o.col = (const unsigned short*)(blob + h[2]);
    o.val = (const double*)(blob + h[3]);
    return o;
}

// one row: eight entries in flight (index and value loads, then the gathers,
// then four independent accumulators) -- every one of these is a shared-memory
// or FP64 latency, and a row is short
__device__ __forceinline__ double sop_row(const SOp& op, int r, const double* in) {
    const int a = op.ptr[r], b = op.ptr[r + 1];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int k = a; k < b; k += 8) {
        int c[8];
        double v[8], x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool ok = k + j < b;
            c[j] = ok ? (int)op.col[k + j] : 0;
            v[j] = ok ? op.val[k + j] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = in[c[j]];
        s0 = fma(v[0], x[0], s0); s1 = fma(v[1], x[1], s1);
        s2 = fma(v[2], x[2], s2); s3 = fma(v[3], x[3], s3);
        s0 = fma(v[4], x[4], s0); s1 = fma(v[5], x[5], s1);
        s2 = fma(v[6], x[6], s2); s3 = fma(v[7], x[7], s3);
    }
    return (s0 + s1) + (s2 + s3);
}

// Sliced ELLPACK operator: slices of 32 rows sorted by length, column-major
// inside a slice, rows padded to a multiple of four entries (value 0, column 0):
// a warp reads consecutive shared-memory words for values and columns (the CSR
// walk was bound by bank conflicts: rows of different length start at scattered
// addresses) and needs no predicates.
struct EOp {
    int nslots;
    const int* sbase;
    const unsigned short* col;
    const double* val;
    const short* rowid;
};

__device__ __forceinline__ EOp eop_get(const unsigned char* blob, int k) {
    const int* h = (const int*)blob + SH_ELL + 5 * k;
    EOp o;
    o.nslots = h[0];
    o.sbase = (const int*)(blob + h[1]);
    o.col = (const unsigned short*)(blob + h[2]);
    o.val = (const double*)(blob + h[3]);
    o.rowid = (const short*)(blob + h[4]);
    return o;
}

__device__ __forceinline__ double ell_slot(const EOp& op, int slot, const double* in) {
    const int sl = slot >> 5;
    const int base = op.sbase[sl];
    const int w = (op.sbase[sl + 1] - base) >> 5;
    const double* v = op.val + base + (slot & 31);
    const unsigned short* c = op.col + base + (slot & 31);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
    for (int j = 0; j < w; j += 4) {
        const int c0 = c[32 * j], c1 = c[32 * j + 32], c2 = c[32 * j + 64], c3 = c[32 * j + 96];
        const double v0 = v[32 * j], v1 = v[32 * j + 32], v2 = v[32 * j + 64], v3 = v[32 * j + 96];
        s0 = fma(v0, in[c0], s0);
        s1 = fma(v1, in[c1], s1);
        s2 = fma(v2, in[c2], s2);
        s3 = fma(v3, in[c3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}

// per-CTA state in shared memory
struct SchurSm {
    const unsigned char* blob;
    int m, nB, band, W, nout, njac, has_csi;
    double *xloc, *ivI, *svI, *qI, *qpI, *dqI, *uI, *tI, *srcI;
    double *ivB, *svB, *qB, *qpB, *dqB, *rB, *dB, *yB, *srcB, *snd, *Tf, *Dinv, *out, *jac, *mail, *red;
    int MB;
};

__host__ __device__ inline long schur_vec_doubles(int m, int nB, int band, int nout, int njac,
                                                  int ncta) {
    const int W = 2 * band + 1, MB = 2 * nB + 2;
    return (long)(nB + m) + 8L * m + 8L * nB + MB + (long)nB * W + nB + nout + njac +
           2L * ncta * MB + 64 + 12 * W + 32 + nB;
}


// ---- boundary system: band LU without pivoting, one thread ---------------------------
// (a single warp issues these loops, so what counts is the number of
// instructions per row: the band storage is padded with BAND + 1 identity rows
// and zero right-hand sides, positions outside the matrix hold zeros, and no
// access needs a guard)
// T is stored by rows, Tf[r * W + BAND + (c - r)], W = 2 BAND + 1.  The rows and
// right-hand side entries the elimination is working on sit in registers (a
// sliding (BAND+1) x (BAND+1) window), so that the chain of dependent
// operations per pivot is reciprocal -> multiplier -> update of the next pivot;
// loads of untouched entries and all stores are off that chain.
template <int BAND>
__device__ __forceinline__ void band_backward(const double* Tf, const double* yB,
                                              const double* Dinv, double* dB, int nB8);

template <int BAND>
__device__ __forceinline__ int band_factor_solve(double* Tf, double* rB, double* Dinv, double* dB,
                                                 int nB) {
    constexpr int W = 2 * BAND + 1;
    double win[BAND + 1][BAND + 1], y[BAND + 1];
    int bad = 0;
#pragma unroll
    for (int i = 0; i <= BAND; ++i) {
#pragma unroll
        for (int j = 0; j <= BAND; ++j)
            win[i][j] = Tf[i * W + BAND + (j - i)];
        y[i] = rB[i];
    }
#pragma unroll 4
    for (int k = 0; k < nB; ++k) {
        // entries entering the window after this step (untouched so far)
        double ncol[BAND + 1], nrow[BAND + 1], ny;
        const int q = k + BAND + 1;
#pragma unroll
        for (int i = 0; i <= BAND; ++i) {
            const int r = k + 1 + i;
            ncol[i] = Tf[r * W + BAND + (q - r)];
            nrow[i] = Tf[q * W + i];                       // columns k+1+i, i < BAND
        }
        ny = rB[q];
        const double piv = win[0][0];
        const double pinv = cdn_rcp1(piv);   // 2^-46: a Newton step is not sensitive to it
        if (!(fabs(piv) > 0.0) || !isfinite(pinv)) bad = 1;
        Dinv[k] = pinv;
#pragma unroll
        for (int j = 1; j <= BAND; ++j)
            Tf[k * W + BAND + j] = win[0][j];
        Tf[k * W + BAND] = piv;
        rB[k] = y[0];
#pragma unroll
        for (int i = 1; i <= BAND; ++i) {
            const double mlt = win[i][0] * pinv;
            if (fabs(mlt) > CDN_MULT_LIMIT) bad = 1;
            Tf[(k + i) * W + BAND - i] = mlt;
#pragma unroll
            for (int j = 1; j <= BAND; ++j) win[i][j] = fma(-mlt, win[0][j], win[i][j]);
            y[i] = fma(-mlt, y[0], y[i]);
        }
        // slide
#pragma unroll
        for (int i = 1; i <= BAND; ++i) {
#pragma unroll
            for (int j = 1; j <= BAND; ++j) win[i - 1][j - 1] = win[i][j];
            win[i - 1][BAND] = ncol[i - 1];
            y[i - 1] = y[i];
        }
#pragma unroll
        for (int j = 0; j < BAND; ++j) win[BAND][j] = nrow[j];
        win[BAND][BAND] = ncol[BAND];
        y[BAND] = ny;
    }
    // U dx = y (rB now holds the forward-substituted right-hand side)
    band_backward<BAND>(Tf, rB, Dinv, dB, (nB + 7) & ~7);
    return bad;
}

// dx_B from r_B with stored factors.  The term with the newest solution entry
// is added last, and the backward sweep multiplies by the reciprocal pivot
// before that: one dependent operation per row on either sweep.
// Triangular sweeps of the band factors, eight rows at a time: all loads of a
// block are issued before its chain of dependent fused multiply-adds and all
// stores after it (ptxas keeps shared-memory loads behind earlier shared-memory
// stores, so a row-by-row loop pays a 29-cycle round trip per row; measured 66
// cycles per row against 8.3 for the DFMA itself).  The term with the newest
// solution entry is added last: one dependent operation per row.  Arrays are
// padded to a multiple of eight rows (zero right-hand side, zero reciprocal
// pivot, identity rows).
template <int BAND>
__device__ __forceinline__ void band_forward(const double* Tf, const double* rB, double* yB,
                                             int nB8) {
    constexpr int W = 2 * BAND + 1;
    double yr[BAND + 1];
#pragma unroll
    for (int j = 0; j <= BAND; ++j) yr[j] = 0.0;
    for (int i0 = 0; i0 < nB8; i0 += 8) {
        double r[8], l[8][BAND + 1];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            r[t] = rB[i0 + t];
#pragma unroll
            for (int j = 1; j <= BAND; ++j) l[t][j] = Tf[(i0 + t) * W + BAND - j];
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            double sacc = r[t];
#pragma unroll
            for (int j = BAND; j >= 1; --j) sacc = fma(-l[t][j], yr[j - 1], sacc);
            r[t] = sacc;
#pragma unroll
            for (int j = BAND; j > 0; --j) yr[j] = yr[j - 1];
            yr[0] = sacc;
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) yB[i0 + t] = r[t];
    }
}

template <int BAND>
__device__ __forceinline__ void band_backward(const double* Tf, const double* yB,
                                              const double* Dinv, double* dB, int nB8) {
    constexpr int W = 2 * BAND + 1;
    double xr[BAND + 1];
#pragma unroll
    for (int j = 0; j <= BAND; ++j) xr[j] = 0.0;
    for (int i0 = nB8 - 8; i0 >= 0; i0 -= 8) {
        double y[8], d[8], u[8][BAND + 1];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            y[t] = yB[i0 + t];
            d[t] = Dinv[i0 + t];
#pragma unroll
            for (int j = 1; j <= BAND; ++j) u[t][j] = Tf[(i0 + t) * W + BAND + j];
        }
#pragma unroll
        for (int t = 7; t >= 0; --t) {
            double sacc = y[t];
#pragma unroll
            for (int j = BAND; j >= 2; --j) sacc = fma(-u[t][j], xr[j - 1], sacc);
            double xv = sacc * d[t];
            if (BAND >= 1) xv = fma(-(u[t][1] * d[t]), xr[0], xv);
            y[t] = xv;
#pragma unroll
            for (int j = BAND; j > 0; --j) xr[j] = xr[j - 1];
            xr[0] = xv;
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) dB[i0 + t] = y[t];
    }
}

// dx_B from r_B with stored factors
template <int BAND>
__device__ __forceinline__ void band_resolve(const double* Tf, const double* rB, const double* Dinv,
                                             double* yB, double* dB, int nB) {
    const int nB8 = (nB + 7) & ~7;
    band_forward<BAND>(Tf, rB, yB, nB8);
    band_backward<BAND>(Tf, yB, Dinv, dB, nB8);
}

template <int MS, int TPB>
__global__ void __launch_bounds__(TPB) engine_schur_kernel(const cdn_schur_args A) {
    namespace cgx = cooperative_groups;
    cgx::cluster_group cluster = cgx::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int ncta = A.ncta;
    const int tid = threadIdx.x;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    // ---- stage the operator blob ------------------------------------------------------
    // layout: [mailboxes | blob | vectors]; the mailboxes come first because peers
    // address them by offset (map_shared_rank) and the blobs differ in size
    const unsigned char* gblob = A.blob + A.cta_off[rank];
    const int nbytes = (int)(A.cta_off[rank + 1] - A.cta_off[rank]);
    const int mail_bytes = 2 * ncta * (2 * A.nB + 2) * 8;
    unsigned char* sblob = smem_raw + mail_bytes;
    {
        const uint4* src = (const uint4*)gblob;
        uint4* dst = (uint4*)sblob;
        for (int k = tid; k < nbytes / 16; k += TPB) dst[k] = __ldg(src + k);
    }
    __syncthreads();
    SchurSm s;
    s.blob = sblob;
    s.mail = (double*)smem_raw;
    {
        const int* h = (const int*)sblob;
        s.m = h[SH_M]; s.nB = h[SH_NB]; s.band = h[SH_BAND]; s.nout = h[SH_OUT];
        s.njac = h[SH_JAC]; s.has_csi = h[SH_HASCSI];
        s.W = 2 * s.band + 1;
        s.MB = 2 * s.nB + 2;
    }
    const int m = s.m, nB = s.nB, band = s.band, W = s.W, MB = s.MB;
    {
        double* p = (double*)(sblob + ((nbytes + 15) / 16) * 16);
        s.xloc = p; p += nB + m;
        s.ivI = p; p += m; s.svI = p; p += m; s.qI = p; p += m; s.qpI = p; p += m;
        s.dqI = p; p += m; s.uI = p; p += m; s.tI = p; p += m; s.srcI = p; p += m;
        s.ivB = p; p += nB; s.svB = p; p += nB; s.qB = p; p += nB; s.qpB = p; p += nB;
        s.dqB = p; p += nB; s.rB = p; p += nB + 8; s.dB = p; p += nB + 8; s.yB = p; p += nB + 8; s.srcB = p; p += nB;
        s.snd = p; p += MB;
        s.Tf = p; p += (nB + 12) * W; s.Dinv = p; p += nB + 8;
        s.out = p; p += s.nout; s.jac = p; p += s.njac;
        s.red = p; p += 64;
    }
    const SOp oLinv = sop_get(s.blob, SOP_LINV), oUinv = sop_get(s.blob, SOP_UINV);
    const SOp oX = sop_get(s.blob, SOP_X), oAsi = sop_get(s.blob, SOP_ASI);
    const SOp oCsi = sop_get(s.blob, SOP_CSI), oArow = sop_get(s.blob, SOP_AROW);
    const SOp oCrow = sop_get(s.blob, SOP_CROW), oAbb = sop_get(s.blob, SOP_ABB);
    const SOp oCbb = sop_get(s.blob, SOP_CBB), oDcur = sop_get(s.blob, SOP_DCUR);
    const SOp oDchg = sop_get(s.blob, SOP_DCHG), oTjac = sop_get(s.blob, SOP_TJAC);
    const EOp eLinv = eop_get(s.blob, EOP_LINV), eUinv = eop_get(s.blob, EOP_UINV);
    const EOp eX = eop_get(s.blob, EOP_X), eArow = eop_get(s.blob, EOP_AROW);
    const double* Tb = (const double*)(s.blob + ((const int*)s.blob)[SH_TB]);
    const int* Ic = (const int*)(s.blob + ((const int*)s.blob)[SH_IC]);
    const int* Bl = (const int*)(s.blob + ((const int*)s.blob)[SH_B]);
    const int* llong = (const int*)(s.blob + ((const int*)s.blob)[SH_LLONG]);
    const int nlong = ((const int*)s.blob)[SH_NLONG];
    double* xB = s.xloc;
    double* xI = s.xloc + nB;
    const cdn_opts& o = A.o;
    const double a0 = o.a0;
    const bool trap = o.im == 1;

    Ws w;
    w.out = s.out; w.jac = s.jac; w.hist = nullptr; w.x = xB;
    w.piv = nullptr; w.stage = nullptr; w.dprof = nullptr; w.sd_plan = 0u; w.sd_base = 0u;

    long long* prof = (rank == 0) ? A.prof : nullptr;
    long long pt = 0;
#define SCH_T0 do { if (prof) pt = clock64(); } while (0)
#define SCH_P(k) do { if (prof) { const long long t_ = clock64(); if (tid == 0) prof[k] += t_ - pt; pt = t_; } } while (0)

    for (int j = tid; j < nB; j += TPB) {
        xB[j] = A.x[Bl[j]];
        s.srcB[j] = A.src_dc ? __ldg(A.src_dc + Bl[j]) : 0.0;
        s.ivB[j] = 0.0;
    }
    for (int r = tid; r < m; r += TPB) {
        xI[r] = A.x[Ic[r]];
        s.srcI[r] = A.src_dc ? __ldg(A.src_dc + Ic[r]) : 0.0;
        s.ivI[r] = 0.0;
    }
    for (int e = tid; e < 12 * W; e += TPB) s.Tf[nB * W + e] = (e % W == band) ? 1.0 : 0.0;
    if (tid < 8) { s.rB[nB + tid] = 0.0; s.Dinv[nB + tid] = 0.0; s.yB[nB + tid] = 0.0; }
    __syncthreads();
    int xc = 0;      // exchange counter (parity of the mailbox in use)

    // Warp roles: the device sweep (a chain of dependent FP64 operations per
    // device, few threads) runs on the last warps of the CTA while the others do
    // the interior work of the same phase, which never needs device outputs
    // (interior rows carry no device stamps).  The workers meet at named barrier
    // 1, everybody at barrier 0.
    int evw = (A.ndev + 31) / 32;
    if (evw * 64 > TPB) evw = 0;                         // too many devices: no split
    const int NWK = TPB - 32 * evw;                      // worker threads
    const bool is_worker = tid < NWK;
    const int etid = tid - NWK;                          // rank among the sweep threads
    const int ENT = evw ? 32 * evw : TPB;
    auto wbar = [&]() {
        if (evw) asm volatile("bar.sync 1, %0;" ::"r"(NWK) : "memory");
        else __syncthreads();
    };

    // all CTAs: mail[par][src][0 .. count) <- snd[0 .. count) of CTA src
    auto exchange = [&](int count) {
        __syncthreads();
        const int par = xc & 1;
        ++xc;
        double* mine = s.mail + ((long)par * ncta + rank) * MB;
        for (int t = tid; t < 256 * ncta; t += TPB) {       // (count <= 2 * 96 + 2)
            const int dst = t >> 8, j = t & 255;
            if (j < count) {
                double* remote = cluster.map_shared_rank(mine, dst);
                remote[j] = s.snd[j];
            }
        }
        cluster.sync();
        return s.mail + (long)par * ncta * MB;
    };
    // the same in two halves, so that work that needs no remote data runs while
    // the partial sums travel: send + arrive, ..., wait
    auto exchange_send = [&](int count) {
        __syncthreads();
        const int par = xc & 1;
        double* mine = s.mail + ((long)par * ncta + rank) * MB;
        for (int t = tid; t < 256 * ncta; t += TPB) {
            const int dst = t >> 8, j = t & 255;
            if (j < count) {
                double* remote = cluster.map_shared_rank(mine, dst);
                remote[j] = s.snd[j];
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    };
    auto exchange_wait = [&]() {
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        const int par = xc & 1;
        ++xc;
        return s.mail + (long)par * ncta * MB;
    };
    // device sweep by the sweep warps (by everybody without the split)
    auto sweep = [&](bool deriv) {
        for (int d = (evw ? etid : tid); d < A.ndev; d += ENT) {
            const cdn_group_dev& G = A.groups[__ldg(A.dev_g + d)];
            if (deriv) eval_one<MS, true>(G, 0, __ldg(A.dev_i + d), xB, w, 0, o.h, 0, a0);
            else eval_one<MS, false>(G, 0, __ldg(A.dev_i + d), xB, w, 0, o.h, 0, 1.0);
        }
    };
    // dx_B from r_B with the stored factors (one thread; band is 0..4)
    auto band_solve = [&]() {
        if (tid == 0) {
            switch (band) {
            case 0: band_resolve<0>(s.Tf, s.rB, s.Dinv, s.yB, s.dB, nB); break;
            case 1: band_resolve<1>(s.Tf, s.rB, s.Dinv, s.yB, s.dB, nB); break;
            case 2: band_resolve<2>(s.Tf, s.rB, s.Dinv, s.yB, s.dB, nB); break;
            case 3: band_resolve<3>(s.Tf, s.rB, s.Dinv, s.yB, s.dB, nB); break;
            default: band_resolve<4>(s.Tf, s.rB, s.Dinv, s.yB, s.dB, nB); break;
            }
        }
        __syncthreads();
    };
    // workers: t_I = Uinv Linv y_I (the rows of Linv that belong to the top
    // separators of a piece are long -- they depend on the whole piece -- and are
    // summed by a warp each)
    auto interior_solve = [&](const double* yI) {
        wbar();
        // (row q of the CSR operator SOP_LINV is row llong[q] of the piece, longest
        // first; four entries per lane in flight)
        for (int q = tid >> 5; q < nlong; q += NWK / 32) {
            const int a = oLinv.ptr[q], b = oLinv.ptr[q + 1];
            double a0_ = 0.0, a1_ = 0.0, a2_ = 0.0, a3_ = 0.0;
            for (int k = a + (tid & 31); k < b; k += 128) {
                const bool k1 = k + 32 < b, k2 = k + 64 < b, k3 = k + 96 < b;
                const int c0 = oLinv.col[k], c1 = k1 ? (int)oLinv.col[k + 32] : 0;
                const int c2 = k2 ? (int)oLinv.col[k + 64] : 0, c3 = k3 ? (int)oLinv.col[k + 96] : 0;
                const double v0 = oLinv.val[k], v1 = k1 ? oLinv.val[k + 32] : 0.0;
                const double v2 = k2 ? oLinv.val[k + 64] : 0.0, v3 = k3 ? oLinv.val[k + 96] : 0.0;
                a0_ = fma(v0, yI[c0], a0_);
                a1_ = fma(v1, yI[c1], a1_);
                a2_ = fma(v2, yI[c2], a2_);
                a3_ = fma(v3, yI[c3], a3_);
            }
            double acc = (a0_ + a1_) + (a2_ + a3_);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if ((tid & 31) == 0) s.uI[llong[q]] = acc;
        }
        for (int slot = tid; slot < eLinv.nslots; slot += NWK) {
            const int r = eLinv.rowid[slot];
            const double acc = ell_slot(eLinv, slot, yI);
            if (r >= 0) s.uI[r] = acc;
        }
        wbar();
        for (int slot = tid; slot < eUinv.nslots; slot += NWK) {
            const int r = eUinv.rowid[slot];
            const double acc = ell_slot(eUinv, slot, s.uI);
            if (r >= 0) s.tI[r] = acc;
        }
        wbar();
    };
    // time-varying sources of sample i on the rows this thread just wrote
    auto add_sources_I = [&](int i, int r, double f) {
        for (int k = 0; k < A.nsrc; ++k)
            if (__ldg(A.src_own + k) == rank && __ldg(A.src_loc + k) - nB == r)
                f += __ldg(A.src_val + (long)i * A.nsrc + k);
        return f;
    };
    auto add_sources_B = [&](int i, int j, double f) {
        for (int k = 0; k < A.nsrc; ++k)
            if (__ldg(A.src_own + k) < 0 && __ldg(A.src_loc + k) == j)
                f += __ldg(A.src_val + (long)i * A.nsrc + k);
        return f;
    };

    // ---- set_IC (nodalSP.py:666-696), integration.py init ------------------------------
    // q = C x + Tq q(x) (nodalSP.py:719-733)
    if (!evw || !is_worker) sweep(false);
    if (!evw || is_worker) {
        for (int r = tid; r < m; r += NWK) { s.qpI[r] = sop_row(oCrow, r, s.xloc); s.dqI[r] = 0.0; }
        if (s.has_csi)
            for (int j = tid; j < nB; j += NWK) s.snd[j] = sop_row(oCsi, j, xI);
    }
    {
        const double* mb = nullptr;
        if (s.has_csi) mb = exchange(nB);
        else __syncthreads();
        for (int j = tid; j < nB; j += TPB) {
            double q = sop_row(oCbb, j, xB) + sop_row(oDchg, j, s.out);
            if (mb)
                for (int c = 0; c < ncta; ++c) q += mb[c * MB + j];
            s.qpB[j] = q;
            s.dqB[j] = 0.0;
        }
        __syncthreads();
    }

    int status = 0, piv_at = 0;
    for (int i = 1; i < A.nsamples; ++i) {
        SCH_T0;
        // ---- accept(xOld) + rhs (nodalSP.py:735-802, integration.py:36-49, :83-97) and the
        // interior part of the chord predictor (tran.py:181: last factors, last residual);
        // the values-only device sweep runs beside them
        const bool chord = i > 1;
        if (!evw || !is_worker) sweep(false);
        if (!evw || is_worker) {
            for (int r = tid; r < m; r += NWK) {
                const double q = sop_row(oCrow, r, s.xloc);
                double dq = s.dqI[r];
                if (trap) { dq = -dq + a0 * (q - s.qpI[r]); s.dqI[r] = dq; }
                s.qpI[r] = q;
                double f = a0 * q;
                if (trap) f += dq;
                f = add_sources_I(i, r, f + s.srcI[r]);
                s.svI[r] = f;
                s.qI[r] = f - s.ivI[r];                  // y_I of the chord step
            }
            if (chord) {
                interior_solve(s.qI);
                for (int j = tid; j < nB; j += NWK) s.snd[j] = sop_row(oAsi, j, s.tI);
            }
            if (s.has_csi)
                for (int j = tid; j < nB; j += NWK) s.snd[nB + j] = sop_row(oCsi, j, xI);
        }
        SCH_P(8);
        {
            const double* mb = nullptr;
            const bool ex = chord || s.has_csi;
            if (ex) exchange_send(2 * nB);
            else __syncthreads();
            for (int j = tid; j < nB; j += TPB)          // local part of q_B meanwhile
                s.qB[j] = sop_row(oCbb, j, xB) + sop_row(oDchg, j, s.out);
            if (ex) mb = exchange_wait();
            SCH_P(9);
            for (int j = tid; j < nB; j += TPB) {
                double q = s.qB[j];
                if (s.has_csi)
                    for (int c = 0; c < ncta; ++c) q += mb[c * MB + nB + j];
                double dq = s.dqB[j];
                if (trap) { dq = -dq + a0 * (q - s.qpB[j]); s.dqB[j] = dq; }
                s.qpB[j] = q;
                double f = a0 * q;
                if (trap) f += dq;
                f = add_sources_B(i, j, f + s.srcB[j]);
                s.svB[j] = f;
                if (chord) {
                    double r = f - s.ivB[j];
                    for (int c = 0; c < ncta; ++c) r -= mb[c * MB + j];
                    s.rB[j] = r;
                }
            }
            __syncthreads();
        }
        SCH_P(5);
        if (chord) {
            band_solve();
            SCH_P(10);
            for (int slot = tid; slot < eX.nslots; slot += TPB) {
                const int r = eX.rowid[slot];
                const double acc = ell_slot(eX, slot, s.dB);
                if (r >= 0) xI[r] += s.tI[r] - acc;
            }
            for (int j = tid; j < nB; j += TPB) xB[j] += s.dB[j];
            __syncthreads();
        }
        SCH_P(6);
        // ---- damped Newton (fsolve.py:59-128) ------------------------------------------------
        int it = 0;
        bool ok = false;
        double res = 0.0;
        while (it < o.maxiter) {
            SCH_T0;
            // device sweep with derivatives beside the interior residual and solve
            if (!evw || !is_worker) sweep(true);
            if (!evw || is_worker) {
                for (int slot = tid; slot < eArow.nslots; slot += NWK) {
                    const int r = eArow.rowid[slot];
                    const double v = ell_slot(eArow, slot, s.xloc);
                    if (r >= 0) { s.ivI[r] = v; s.qI[r] = s.svI[r] - v; }
                }
                interior_solve(s.qI);
                for (int j = tid; j < nB; j += NWK) {
                    s.snd[j] = sop_row(oAsi, j, xI);
                    s.snd[nB + j] = sop_row(oAsi, j, s.tI);
                }
            }
            SCH_P(1);
            exchange_send(2 * nB);
            // what needs no remote data, while the partial sums travel: the local
            // part of i_B and the boundary matrix T = T0 + D(x)
            for (int j = tid; j < nB; j += TPB)
                s.ivB[j] = sop_row(oAbb, j, xB) + sop_row(oDcur, j, s.out) +
                           a0 * sop_row(oDchg, j, s.out);
            for (int e = tid; e < nB * W; e += TPB) s.Tf[e] = Tb[e] + sop_row(oTjac, e, s.jac);
            SCH_P(11);
            const double* mb = exchange_wait();
            SCH_P(7);
            for (int j = tid; j < nB; j += TPB) {
                double iv = s.ivB[j];
                double pb = 0.0;
                for (int c = 0; c < ncta; ++c) { iv += mb[c * MB + j]; pb += mb[c * MB + nB + j]; }
                s.ivB[j] = iv;
                s.rB[j] = (s.svB[j] - iv) - pb;
            }
            __syncthreads();
            // banded LU without pivoting, right-hand side carried along
            if (tid == 0) {
                int bad;
                switch (band) {
                case 0: bad = band_factor_solve<0>(s.Tf, s.rB, s.Dinv, s.dB, nB); break;
                case 1: bad = band_factor_solve<1>(s.Tf, s.rB, s.Dinv, s.dB, nB); break;
                case 2: bad = band_factor_solve<2>(s.Tf, s.rB, s.Dinv, s.dB, nB); break;
                case 3: bad = band_factor_solve<3>(s.Tf, s.rB, s.Dinv, s.dB, nB); break;
                default: bad = band_factor_solve<4>(s.Tf, s.rB, s.Dinv, s.dB, nB); break;
                }
                s.red[32] = (double)bad;
            }
            __syncthreads();
            if (s.red[32] != 0.0 && !piv_at) piv_at = i + 1;
            SCH_P(2);
            // dx_I = t_I - X dx_B (kept in tI), local max |dx|, convergence assuming
            // no damping -- and x is updated on that assumption right away (the old
            // values are kept in qI / yB); a damped step, known after the exchange,
            // re-does its update
            double mloc = 0.0;
            bool conv = true;
            for (int slot = tid; slot < eX.nslots; slot += TPB) {
                const int r = eX.rowid[slot];
                const double acc = ell_slot(eX, slot, s.dB);
                if (r >= 0) {
                    const double d = s.tI[r] - acc;
                    s.tI[r] = d;
                    mloc = fmax(mloc, fabs(d));
                    const double xo = xI[r], xn = xo + d;
                    conv = conv && (fabs(d) < o.reltol * fmax(fabs(xo), fabs(xn)) + o.abstol);
                    s.qI[r] = xo;
                    xI[r] = xn;
                }
            }
            for (int j = tid; j < nB; j += TPB) {
                const double d = s.dB[j];
                mloc = fmax(mloc, fabs(d));
                const double xo = xB[j], xn = xo + d;
                conv = conv && (fabs(d) < o.reltol * fmax(fabs(xo), fabs(xn)) + o.abstol);
                s.yB[j] = xo;
                xB[j] = xn;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mloc = fmax(mloc, __shfl_xor_sync(0xffffffffu, mloc, off));
            if ((tid & 31) == 0) s.red[tid >> 5] = mloc;
            const int call = __syncthreads_and(conv ? 1 : 0);
            if (tid == 0) {
                double mm = 0.0;
                for (int k = 0; k < TPB / 32; ++k) mm = fmax(mm, s.red[k]);
                s.snd[0] = mm;
                s.snd[1] = call ? 1.0 : 0.0;
            }
            SCH_P(3);
            const double* mb2 = exchange(2);
            SCH_P(12);
            double mg = 0.0;
            bool cg_ = true;
            for (int c = 0; c < ncta; ++c) {
                mg = fmax(mg, mb2[c * MB]);
                cg_ = cg_ && (mb2[c * MB + 1] != 0.0);
            }
            const bool damp = mg > o.maxdelta;
            const double scale = damp ? o.maxdelta / mg : 1.0;
            if (damp) {
                // convergence is judged on the damped step (fsolve.py:112-121)
                conv = true;
                for (int r = tid; r < m; r += TPB) {
                    const double d = s.tI[r] * scale;
                    const double xo = s.qI[r], xn = xo + d;
                    conv = conv && (fabs(d) < o.reltol * fmax(fabs(xo), fabs(xn)) + o.abstol);
                    xI[r] = xn;
                }
                for (int j = tid; j < nB; j += TPB) {
                    const double d = s.dB[j] * scale;
                    const double xo = s.yB[j], xn = xo + d;
                    conv = conv && (fabs(d) < o.reltol * fmax(fabs(xo), fabs(xn)) + o.abstol);
                    xB[j] = xn;
                }
                const int c2 = __syncthreads_and(conv ? 1 : 0);
                if (tid == 0) s.snd[0] = c2 ? 1.0 : 0.0;
                const double* mb3 = exchange(1);
                cg_ = true;
                for (int c = 0; c < ncta; ++c) cg_ = cg_ && (mb3[c * MB] != 0.0);
            }
            res = damp ? mg * scale : mg;
            ++it;
            SCH_P(4);
            if (cg_) { ok = true; break; }
        }
        if (rank == 0 && tid == 0) { A.iters[i] = it; A.res[i] = res; }
        if (!ok) { status = i; break; }
        {
            const int k0 = __ldg(A.save_ptr + rank), k1 = __ldg(A.save_ptr + rank + 1);
            for (int k = k0 + tid; k < k1; k += TPB)
                A.wave[(long)i * A.nsave + __ldg(A.save_k + k)] = s.xloc[__ldg(A.save_loc + k)];
        }
    }
    // last sample back to x
    for (int r = tid; r < m; r += TPB) A.x[Ic[r]] = xI[r];
    if (rank == 0) {
        for (int j = tid; j < nB; j += TPB) A.x[Bl[j]] = xB[j];
        if (tid == 0) {
            A.status[0] = status;
            if (A.info) { A.info[0] = 0; A.info[1] = piv_at; }
        }
    }
    cluster.sync();      // no CTA may exit while peers can still write its mailbox
#undef SCH_T0
#undef SCH_P
}

template <int MS>
static cudaError_t launch_schur(const cdn_schur_args& A, long smem, cudaStream_t st) {
    // light model sets: one thread per interior row of a typical piece plus the
    // warps of the device sweep
    constexpr int TPB = (MS == CDN_MS(CDN_MODEL_DIODE)) ? 512 : 256;
    auto kern = engine_schur_kernel<MS, TPB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (A.ncta > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(A.ncta);
    cfg.blockDim = dim3(TPB);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = A.ncta;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, A);
}
