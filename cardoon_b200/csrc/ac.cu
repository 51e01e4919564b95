// Small-signal AC sweep: for every frequency f_k solve
//     Y(w_k) x_k = s,   Y(w) = G + j w C + sum_d (g_d + j w c_d) exp(-j w tau_d)
//                               + Y_fd(w_k)
// Replaces the per-frequency loop of the reference's run_AC
// (analyses/nodal.py:286-381: Y[:] = G + jomega*C, set_Jac of every device
// Jacobian with exp(-jomega*tau) on delayed columns, linalg.solve).
//
// Frequencies are independent, so they are the batch dimension: one CTA
// assembles and factors one Y (dense complex LU with partial pivoting on
// |re|+|im|, the LAPACK zgetrf pivot rule) and solves for the source vector.
// G and C already contain the frequency-independent device stamps
// (dI/dv into G, dQ/dv into C); only delayed control ports are applied per
// frequency from the triplet list, and the Y parameters of frequency-defined
// elements (get_Y_matrix, nodal.py:366-370; evaluated on the host for the whole
// sweep) are added from a per-frequency value table.  Matrices of up to 56x56 live in shared
// memory, larger ones in the caller's workspace (L2-resident).
#include "common.cuh"

struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    cplx r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r;
}
__device__ __forceinline__ cplx csub(cplx a, cplx b) { cplx r; r.re = a.re - b.re; r.im = a.im - b.im; return r; }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    // Smith's algorithm (what the C99 / numpy complex division does)
    cplx r;
    if (fabs(b.re) >= fabs(b.im)) {
        const double t = b.im / b.re, d = b.re + b.im * t;
        r.re = (a.re + a.im * t) / d; r.im = (a.im - a.re * t) / d;
    } else {
        const double t = b.re / b.im, d = b.re * t + b.im;
        r.re = (a.re * t + a.im) / d; r.im = (a.im * t - a.re) / d;
    }
    return r;
}

// first row (>= k) with the largest |re| + |im| in column k of the column-major
// matrix; -1 if the column is zero.  Two barriers.
template <int TPB>
__device__ __forceinline__ int ac_pivot(const cplx* Y, int n, int k, double* red_v, int* red_i, int* piv) {
    const int tid = threadIdx.x;
    double bv = -1.0; int bi = n;
    for (int i = k + tid; i < n; i += TPB) {
        const cplx y = Y[(long)k * n + i];
        const double a = fabs(y.re) + fabs(y.im);
        if (a > bv) { bv = a; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { red_v[tid >> 5] = bv; red_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int q = 1; q < TPB / 32; ++q)
            if (red_v[q] > bv || (red_v[q] == bv && red_i[q] < bi)) { bv = red_v[q]; bi = red_i[q]; }
        *piv = (bv > 0.0) ? bi : -1;
    }
    __syncthreads();
    return *piv;
}

#define AC_SMEM_N 56
#define AC_R 4                   // pivots per pass over the trailing block
#define AC_MAX_N 2048          // panel multipliers + right-hand side in shared memory: 160 KB

// AC_TPB threads per CTA: 256 for the shared-memory path, 512 for matrices in
// the workspace (every CTA reuses one workspace slot for all its frequencies,
// so the matrices in flight stay L2-resident)
template <int AC_TPB>
__global__ void __launch_bounds__(AC_TPB, 1024 / AC_TPB)
ac_solve_kernel(int n, int nf, const double* __restrict__ G, const double* __restrict__ C,
                int ndel, const int* __restrict__ drow, const int* __restrict__ dcol,
                const double* __restrict__ dg, const double* __restrict__ dc,
                const double* __restrict__ dtau, int nfd, const int* __restrict__ frow,
                const int* __restrict__ fcol, const double* __restrict__ fval,
                const double* __restrict__ freqs,
                const double* __restrict__ s, cplx* __restrict__ work, cplx* __restrict__ xout,
                int* __restrict__ status) {
    extern __shared__ double ac_sm[];
    __shared__ double red_v[AC_TPB / 32];
    __shared__ int red_i[AC_TPB / 32];
    __shared__ int piv;
    const int tid = threadIdx.x;
    const bool in_smem = n <= AC_SMEM_N;
    for (int f = blockIdx.x; f < nf; f += gridDim.x) {
        // one workspace slot per CTA, reused for all its frequencies: the
        // matrices in flight stay L2-resident
        cplx* Y = in_smem ? (cplx*)ac_sm : work + (long)blockIdx.x * n * n;
        const double w = 6.283185307179586476925287 * freqs[f];      // 2 pi f
        // column-major storage (AT(r, c)): the pivot search, the multipliers,
        // the rank-1 update (lanes along the rows) and the column-oriented
        // back substitution all walk contiguous memory; only the row swaps
        // are strided
#define AT(r, c) Y[(long)(c) * n + (r)]
        for (int e = tid; e < n * n; e += AC_TPB) {
            const int c = e / n, r = e - c * n;
            cplx y; y.re = G[(long)r * n + c]; y.im = w * C[(long)r * n + c];
            Y[e] = y;
        }
        cplx* colk = (cplx*)ac_sm + (in_smem ? (long)n * n : 0);      // [n] multipliers
        cplx* xs = colk + (long)AC_R * n;                              // [n] right-hand side
        for (int r = tid; r < n; r += AC_TPB) { cplx v; v.re = s[2 * r]; v.im = s[2 * r + 1]; xs[r] = v; }
        __syncthreads();
        if (tid == 0) {
            // delayed control ports: (g + j w c) exp(-j w tau); entries may repeat
            for (int d = 0; d < ndel; ++d) {
                double sn, cs;
                sincos(-w * dtau[d], &sn, &cs);
                cplx a; a.re = dg[d]; a.im = w * dc[d];
                cplx ph; ph.re = cs; ph.im = sn;
                const cplx v = cmul(a, ph);
                cplx* y = &AT(drow[d], dcol[d]);
                y->re += v.re; y->im += v.im;
            }
            // frequency-defined elements: fval[f][e] = (re, im)
            for (int e = 0; e < nfd; ++e) {
                cplx* y = &AT(frow[e], fcol[e]);
                y->re += fval[((long)f * nfd + e) * 2];
                y->im += fval[((long)f * nfd + e) * 2 + 1];
            }
        }
        __syncthreads();
        bool singular = false;
        // AC_R pivots per pass over the trailing block (rank-AC_R update): the
        // block is loaded and stored once for AC_R eliminated columns, which
        // divides the L2 traffic per multiply-add of the unblocked algorithm
        // by AC_R.  The multipliers of the panel stay in shared memory:
        // Lq[q * n + i] = L(i, k + q).
        cplx* Lq = colk;                         // [AC_R][n] (colk is its first column)
        for (int k = 0; k < n && !singular;) {
            const int R = (n - k < AC_R) ? n - k : AC_R;
            for (int m = 0; m < R; ++m) {
                const int km = k + m;
                if (m > 0) {
                    // bring column km up to date with the panel's earlier pivots:
                    // u_q = U(k+q, km) by a tiny forward substitution every
                    // thread repeats for itself, then one pass over the rows
                    cplx u[AC_R];
#pragma unroll
                    for (int q = 0; q < AC_R - 1; ++q) {
                        if (q < m) {
                            cplx v = AT(k + q, km);
#pragma unroll
                            for (int q2 = 0; q2 < AC_R - 1; ++q2)
                                if (q2 < q) v = csub(v, cmul(Lq[(long)q2 * n + k + q], u[q2]));
                            u[q] = v;
                        }
                    }
                    __syncthreads();            // everybody has read rows k..km-1 of the column
                    for (int i = km + tid; i < n; i += AC_TPB) {
                        cplx a = AT(i, km);
#pragma unroll
                        for (int q = 0; q < AC_R - 1; ++q)
                            if (q < m) a = csub(a, cmul(Lq[(long)q * n + i], u[q]));
                        AT(i, km) = a;
                    }
                    if (tid < m) AT(k + tid, km) = u[tid];
                    __syncthreads();
                }
                const int p = ac_pivot<AC_TPB>(Y, n, km, red_v, red_i, &piv);
                if (p < 0) { singular = true; break; }
                if (p != km) {
                    // (the pending update of the columns right of the panel
                    // commutes with the row swap because the panel's
                    // multipliers are swapped too)
                    for (int j = tid; j < n; j += AC_TPB) { const cplx t = AT(km, j); AT(km, j) = AT(p, j); AT(p, j) = t; }
                    if (tid == 0) { const cplx t = xs[km]; xs[km] = xs[p]; xs[p] = t; }
                    if (tid < m) {
                        const cplx t = Lq[(long)tid * n + km];
                        Lq[(long)tid * n + km] = Lq[(long)tid * n + p];
                        Lq[(long)tid * n + p] = t;
                    }
                    __syncthreads();
                }
                const cplx ykk = AT(km, km);
                const cplx xk = xs[km];
                for (int i = km + 1 + tid; i < n; i += AC_TPB) {
                    const cplx l = cdiv(AT(i, km), ykk);
                    AT(i, km) = l;
                    Lq[(long)m * n + i] = l;
                    xs[i] = csub(xs[i], cmul(l, xk));
                }
                __syncthreads();
            }
            if (singular) break;
            // ---- rank-R update of the columns right of the panel: a warp per
            // column, lanes along the contiguous rows
            const int kR = k + R;
            for (int j = kR + (tid >> 5); j < n; j += AC_TPB / 32) {
                cplx* yj = Y + (long)j * n;
                cplx u[AC_R];
#pragma unroll
                for (int q = 0; q < AC_R; ++q) {
                    if (q < R) {
                        cplx v = yj[k + q];
#pragma unroll
                        for (int q2 = 0; q2 < AC_R; ++q2)
                            if (q2 < q) v = csub(v, cmul(Lq[(long)q2 * n + k + q], u[q2]));
                        u[q] = v;
                    }
                }
                __syncwarp();
                for (int i = kR + (tid & 31); i < n; i += 32) {
                    cplx a = yj[i];
#pragma unroll
                    for (int q = 0; q < AC_R; ++q)
                        if (q < R) a = csub(a, cmul(Lq[(long)q * n + i], u[q]));
                    yj[i] = a;
                }
                if ((tid & 31) < R) {
                    const int q = tid & 31;
                    cplx v = u[0];
#pragma unroll
                    for (int q2 = 1; q2 < AC_R; ++q2) if (q2 == q) v = u[q2];
                    yj[k + q] = v;
                }
            }
            __syncthreads();
            k = kR;
        }
        cplx* x = xout + (long)f * n;
        if (singular) {
            if (tid == 0) atomicExch(status, f + 1);
            __syncthreads();
            continue;
        }
        // ---- back substitution, column oriented: every thread forms x_k itself
        // (one barrier per step), the column of U is contiguous
        for (int k = n - 1; k >= 0; --k) {
            const cplx xk = cdiv(xs[k], AT(k, k));
            if (tid == 0) x[k] = xk;
            for (int i = tid; i < k; i += AC_TPB) xs[i] = csub(xs[i], cmul(AT(i, k), xk));
            __syncthreads();
        }
#undef AT
    }
}

// tuning (cdn_set_tuning keys 4, 5): CTAs in flight / threads per CTA of the
// workspace path
extern "C" { int cdn_tune_ac_blocks = 0; int cdn_tune_ac_tpb = 0; }

// Reference: run_AC (analyses/nodal.py:286-381).  All pointers are device
// pointers.  G, C: [n*n] row-major; delayed entries (drow, dcol, dg, dc,
// dtau)[ndel]; frequency-defined entries (frow, fcol)[nfd] with
// fval[nf][nfd][2]; freqs[nf]; s[n][2] (re, im); work: work_matrices n*n
// complex matrices when n > 56 (else unused), one per CTA in flight -- the grid
// is clamped to work_matrices; x: [nf][n][2]; status: 0 or 1 + index of a
// frequency whose matrix is singular.
extern "C" long cdn_ac_work_matrices(cdn_ctx* ctx, int nf) {
    if (!ctx || nf <= 0) return 0;
    long blocks = 2L * ctx->sm_count;
    if (cdn_tune_ac_blocks > 0) blocks = cdn_tune_ac_blocks;
    return blocks < nf ? blocks : nf;
}

extern "C" int cdn_ac_solve(cdn_ctx* ctx, int n, int nf, const double* G, const double* C, int ndel,
                            const int* drow, const int* dcol, const double* dg, const double* dc,
                            const double* dtau, int nfd, const int* frow, const int* fcol,
                            const double* fval, const double* freqs, const double* s, double* work,
                            long work_matrices, double* x, int* status, void* stream) {
    if (!ctx) return CDN_ERR_ARG;
    CDN_USE_DEVICE(ctx);
    if (n <= 0 || nf <= 0) return CDN_OK;
    if (!G || !C || !freqs || !s || !x || !status || (ndel > 0 && (!drow || !dcol || !dg || !dc || !dtau))
        || (nfd > 0 && (!frow || !fcol || !fval)))
        return cdn_fail(ctx, CDN_ERR_ARG, "cdn_ac_solve: null argument");
    if (n > AC_SMEM_N && (!work || work_matrices < 1))
        return cdn_fail(ctx, CDN_ERR_ARG, "cdn_ac_solve: workspace required");
    if (n > AC_MAX_N) return cdn_fail(ctx, CDN_ERR_ARG, "cdn_ac_solve: more than 2048 unknowns");
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= AC_SMEM_N) {
        const int smem = (n * n + (AC_R + 1) * n) * (int)sizeof(cplx);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(ac_solve_kernel<256>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return cdn_fail(ctx, CDN_ERR_CUDA, std::string("cdn_ac_solve: ") + cudaGetErrorString(e));
        }
        const int blocks = nf < 8 * ctx->sm_count ? nf : 8 * ctx->sm_count;
        ac_solve_kernel<256><<<blocks, 256, smem, st>>>(n, nf, G, C, ndel, drow, dcol, dg, dc, dtau, nfd,
                                                        frow, fcol, fval, freqs, s, (cplx*)work,
                                                        (cplx*)x, status);
    } else {
        // exactly the resident CTAs, each looping over its share of the
        // frequencies (a second wave would only start when the first has
        // finished all of its matrices);
        // two resident 512-thread CTAs per SM (64 registers per thread): the
        // L2 round trips of one matrix's column step overlap with the other's
        // (measured on B200, N = 189: 1 x 1024 threads 66.1 k, 2 x 512 76.8 k,
        // 4 x 256 71.0 k frequency points/s)
        long blocks = 2L * ctx->sm_count;
        if (cdn_tune_ac_blocks > 0) blocks = cdn_tune_ac_blocks;
        if (blocks > nf) blocks = nf;
        // one n x n matrix of the workspace per CTA: never more CTAs than the
        // caller provided matrices for
        if (blocks > work_matrices) blocks = work_matrices;
        const int smem = (AC_R + 1) * n * (int)sizeof(cplx);
        const bool big = cdn_tune_ac_tpb == 1024;
        if (smem > 48 * 1024) {
            const cudaError_t e = big
                ? cudaFuncSetAttribute(ac_solve_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                : cudaFuncSetAttribute(ac_solve_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return cdn_fail(ctx, CDN_ERR_CUDA, std::string("cdn_ac_solve: ") + cudaGetErrorString(e));
        }
        if (big)
            ac_solve_kernel<1024><<<(int)blocks, 1024, smem, st>>>(
                n, nf, G, C, ndel, drow, dcol, dg, dc, dtau, nfd, frow, fcol, fval, freqs, s,
                (cplx*)work, (cplx*)x, status);
        else
            ac_solve_kernel<512><<<(int)blocks, 512, smem, st>>>(
                n, nf, G, C, ndel, drow, dcol, dg, dc, dtau, nfd, frow, fcol, fval, freqs, s,
                (cplx*)work, (cplx*)x, status);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cdn_fail(ctx, CDN_ERR_CUDA, std::string("cdn_ac_solve launch: ") + cudaGetErrorString(e));
    return CDN_OK;
}
