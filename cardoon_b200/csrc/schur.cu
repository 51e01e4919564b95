// C ABI of the boundary/interior transient kernel (include/cardoon_b200.h,
// cdn_schur_tran): checks the operator blobs of cardoon_b200/analyses/schur.py
// against the shared-memory budget and launches one thread-block cluster.
#include "engine_launch.cuh"
#include "schur.cuh"
#include "plan_host.h"
#include <cstring>

cudaError_t cdn_launch_diode_schur(const cdn_schur_args& A, long smem, cudaStream_t st);
cudaError_t cdn_launch_bsim3_schur(const cdn_schur_args& A, long smem, cudaStream_t st);
cudaError_t cdn_launch_all_schur(const cdn_schur_args& A, long smem, cudaStream_t st);

extern "C" {

long cdn_schur_smem_bytes(long blob_bytes, int m, int nB, int band, int nout, int njac, int ncta) {
    const long mail = 2L * ncta * (2L * nB + 2) * 8;
    return mail + (blob_bytes + 15) / 16 * 16 + 8 * schur_vec_doubles(m, nB, band, nout, njac, 0);
}

int cdn_schur_tran(cdn_plan* P, const cdn_run_host* r, const cdn_schur_host* s, void* stream) {
    if (!P || !r || !s) return CDN_ERR_ARG;
    cdn_ctx* ctx = cdn_plan_ctx(P);
    const cdn_plan_dev* D = cdn_plan_host_view(P);
    if (!D) return cdn_fail(ctx, CDN_ERR_STATE, "schur: plan is not bound");
    CDN_USE_DEVICE(ctx);
    if (s->ncta < 1 || s->ncta > CDN_SCHUR_MAXCTA) return cdn_fail(ctx, CDN_ERR_ARG, "schur: 1..16 CTAs");
    if (r->B != 1 || r->first != 1 || !r->init_ic || !r->use_q)
        return cdn_fail(ctx, CDN_ERR_ARG, "schur: one circuit, whole transient from the operating point");
    if (r->errfunc) return cdn_fail(ctx, CDN_ERR_ARG, "schur: errfunc is not available on this path");
    if (D->nhist > 0) return cdn_fail(ctx, CDN_ERR_ARG, "schur: delayed control ports are not available on this path");
    if (!r->x || !s->blob || !s->cta_off || !r->wave || !r->iters || !r->res || !r->status)
        return cdn_fail(ctx, CDN_ERR_ARG, "schur: NULL argument");

    cdn_schur_args A;
    memset(&A, 0, sizeof(A));
    static_assert(sizeof(cdn_schur_args) <= 4000, "kernel parameter space");
    unsigned ms = 0;
    A.ngroups = D->ngroups;
    for (int g = 0; g < D->ngroups; ++g) {
        if (!D->groups[g].params) return cdn_fail(ctx, CDN_ERR_STATE, "schur: group parameters not set");
        A.groups[g] = D->groups[g];
        A.groups[g].vpos = s->vpos_b[g];
        A.groups[g].vneg = s->vneg_b[g];
        if (D->groups[g].per_inst) return cdn_fail(ctx, CDN_ERR_ARG, "schur: batched parameters");
        ms |= 1u << D->groups[g].model;
    }
    if (ms & ~(unsigned)CDN_MS_ALL) return cdn_fail(ctx, CDN_ERR_ARG, "schur: model not available");
    A.ndev = D->ndev; A.dev_g = D->dev_g; A.dev_i = D->dev_i;
    A.o.abstol = r->abstol; A.o.reltol = r->reltol; A.o.maxdelta = r->maxdelta;
    A.o.a0 = r->a0; A.o.lambda = 1.0; A.o.h = r->h; A.o.maxiter = r->maxiter;
    A.o.use_q = 1; A.o.im = r->im;
    A.blob = (const unsigned char*)s->blob;
    A.ncta = s->ncta; A.nB = s->nB;
    long smem = 0;
    for (int c = 0; c <= s->ncta; ++c) A.cta_off[c] = s->cta_off[c];
    for (int c = 0; c < s->ncta; ++c) {
        if (A.cta_off[c] % 16 || A.cta_off[c + 1] < A.cta_off[c])
            return cdn_fail(ctx, CDN_ERR_ARG, "schur: blob offsets");
        const long need = cdn_schur_smem_bytes(A.cta_off[c + 1] - A.cta_off[c], s->m[c], s->nB,
                                               s->band, (int)D->dev_out_size, (int)D->dev_jac_size,
                                               s->ncta);
        if (need > smem) smem = need;
    }
    if (smem > 227 * 1024) return cdn_fail(ctx, CDN_ERR_ARG, "schur: operators do not fit shared memory");
    A.x = r->x;
    A.nsamples = r->nsamples; A.nsrc = r->nsrc; A.nsave = r->nsave;
    A.src_own = s->src_own; A.src_loc = s->src_loc; A.src_val = r->src_val; A.src_dc = r->src_dc;
    A.save_ptr = s->save_ptr; A.save_k = s->save_k; A.save_loc = s->save_loc;
    A.wave = r->wave; A.iters = r->iters; A.res = r->res; A.status = r->status;
    A.info = r->info; A.prof = r->prof;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if ((ms & ~(unsigned)CDN_MS_DIODE) == 0) e = cdn_launch_diode_schur(A, smem, st);
    else if ((ms & ~(unsigned)CDN_MS_BSIM3) == 0) e = cdn_launch_bsim3_schur(A, smem, st);
    else e = cdn_launch_all_schur(A, smem, st);
    if (e != cudaSuccess)
        return cdn_fail(ctx, CDN_ERR_CUDA, std::string("schur launch: ") + cudaGetErrorString(e));
    return CDN_OK;
}

}  // extern "C"
