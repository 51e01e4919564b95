// C ABI of the device-resident Newton engine (include/cardoon_b200.h,
// "engine" section): sizes the per-instance state, picks the thread team and
// the model-set instantiation, and launches one kernel per operation.
#include "engine_launch.cuh"
#include "plan_host.h"
#include <cstring>

extern "C" {

long cdn_engine_ws_doubles(cdn_plan* P) {
    const cdn_plan_dev* D = cdn_plan_host_view(P);
    if (!D) return -1;
    return ws_doubles(*D);
}

// offsets (in doubles) of x, dx, ivec, sv, q, qprev, dq, y, tmp, xb, lu, out, jac, hist, piv
int cdn_engine_layout(cdn_plan* P, long* off) {
    const cdn_plan_dev* D = cdn_plan_host_view(P);
    if (!D || !off) return CDN_ERR_ARG;
    const long n = D->n;
    for (int k = 0; k < 10; ++k) off[k] = k * n;
    off[10] = 10 * n;
    off[11] = off[10] + D->nlu;
    off[12] = off[11] + D->dev_out_size;
    off[13] = off[12] + D->dev_jac_size;
    off[14] = off[13] + D->hist_size + (D->sd ? CDN_SD8_FACT : 0);
    return CDN_OK;
}

// tuning (cdn_set_tuning key 3): threads per CTA of the tile team, 0 = automatic
int cdn_tune_tile_tpb = 0;

int cdn_engine_run(cdn_plan* P, const cdn_run_host* r, void* stream) {
    if (!P || !r) return CDN_ERR_ARG;
    cdn_ctx* ctx = cdn_plan_ctx(P);
    const cdn_plan_dev* D = cdn_plan_host_view(P);
    if (!D) return cdn_fail(ctx, CDN_ERR_STATE, "engine: plan is not bound");
    CDN_USE_DEVICE(ctx);
    cudaStream_t st = (cudaStream_t)stream;

    // model set
    unsigned ms = 0;
    for (int g = 0; g < D->ngroups; ++g) {
        if (!D->groups[g].params) return cdn_fail(ctx, CDN_ERR_STATE, "engine: group parameters not set");
        ms |= 1u << D->groups[g].model;
    }
    if (ms & ~(unsigned)CDN_MS_ALL) return cdn_fail(ctx, CDN_ERR_ARG, "engine: model not available in the Newton engine");

    cdn_engine_args A;
    memset(&A, 0, sizeof(A));
    A.plan = *D;
    static_assert(sizeof(cdn_engine_args) <= 4000, "kernel parameter space");
    A.o.abstol = r->abstol; A.o.reltol = r->reltol; A.o.maxdelta = r->maxdelta;
    A.o.a0 = r->a0; A.o.gmin = r->gmin; A.o.lambda = r->lambda; A.o.h = r->h;
    A.o.gmin_ext = r->gmin_ext; A.o.n_ext = r->n_ext;
    if (A.o.n_ext > D->n) A.o.n_ext = D->n;
    A.o.maxiter = r->maxiter; A.o.errfunc = r->errfunc; A.o.use_q = r->use_q; A.o.im = r->im;
    A.op = r->op; A.B = r->B;
    A.x = r->x; A.sv_in = r->sv_in; A.ws = r->ws;
    const long per = ws_doubles(*D);
    A.ws_stride = per;
    A.nsamples = r->nsamples; A.first = r->first; A.init_ic = r->init_ic;
    A.nsrc = r->nsrc; A.nsave = r->nsave; A.helpers = r->helpers;
    A.src_rows = r->src_rows; A.src_val = r->src_val; A.src_dc = r->src_dc;
    A.save_rows = r->save_rows; A.wave = r->wave; A.iters = r->iters; A.res = r->res;
    A.status = r->status; A.red = r->red; A.prof = r->prof; A.info = r->info;
    if (A.B <= 0) return CDN_OK;
    if (!A.x) return cdn_fail(ctx, CDN_ERR_ARG, "engine: x is NULL");

    cdn_launch_cfg c;
    memset(&c, 0, sizeof(c));
    c.sm_count = ctx->sm_count;
    c.team = r->team;
    if (c.team < 0) {
        if (D->dense && D->n <= 64) c.team = CDN_TEAM_TILE;
        else if (A.B > 1 || D->dense || (D->nlu < 50000 && D->ndev < 2048)) c.team = CDN_TEAM_BLOCK;
        else c.team = CDN_TEAM_GRID;
    }
    // the tile kernels keep the state in shared memory only
    if (c.team == CDN_TEAM_TILE && A.ws) c.team = CDN_TEAM_BLOCK;
    if (c.team == CDN_TEAM_TILE) {
        // ... for the duration of one launch: nothing may continue from (or
        // leave behind) per-instance state
        const bool fresh = A.op == CDN_OP_NEWTON ||
                           (A.op == CDN_OP_TRAN && A.first <= 1 && A.init_ic);
        if (!fresh)
            return cdn_fail(ctx, CDN_ERR_STATE, "engine: this operation continues from kept state; "
                                                "pass a workspace (ws)");
    }
    if (c.team == CDN_TEAM_TILE) {
        int T = r->T;
        if (T <= 0) {
            const int want = D->ndev > D->n ? D->ndev : D->n;
            T = 8;
            while (T < want && T < 32) T *= 2;
        }
        c.T = (T <= 8) ? 8 : (T <= 16) ? 16 : 32;
        T = c.T;
        // one CTA per SM when the batch fills the GPU (its tiles run in lockstep
        // and share the instruction cache), sized so that every SM gets one:
        // ceil(B / SMs) instances per CTA, whole warps, at most 256 threads
        // (an instance's chain of dependent instructions is what a launch
        // costs -- a warp alone on an SM needs 23 ms for 400 ring_bsim3 samples,
        // seven warps sharing it 28 ms -- so small batches spread over all SMs:
        // 512 instances 25.7 ms with 128-thread CTAs, 23.0 ms with 32)
        c.tpb = 256;
        {
            const long per_sm = (A.B + ctx->sm_count - 1) / ctx->sm_count;
            long tpb = (per_sm * T + 31) / 32 * 32;
            if (tpb < 32) tpb = 32;
            if (tpb < c.tpb) c.tpb = (int)tpb;
        }
        if (cdn_tune_tile_tpb >= 32 && cdn_tune_tile_tpb <= 1024 && cdn_tune_tile_tpb % 32 == 0)
            c.tpb = cdn_tune_tile_tpb;
        if (!A.ws) {
            // state in shared memory: shrink the CTA until it fits
            while (c.tpb > 32 && (long)(c.tpb / T) * per * 8 > 200 * 1024) c.tpb -= 32;
            c.smem = (long)(c.tpb / T) * per * 8;
            // room for the shared-memory copy of the plan's gather lists
            const long pb = plan_smem_bytes(*D);
            if (D->dense && !D->n_e_heavy && !D->n_r_heavy && pb <= 32 * 1024 &&
                c.smem + pb <= 200 * 1024) {
                A.plan_smem = (int)pb;
                c.smem += pb;
            }
            if (c.smem > 227 * 1024) return cdn_fail(ctx, CDN_ERR_ARG, "engine: circuit state does not fit shared memory; pass a workspace");
        }
        const int tiles = c.tpb / T;
        c.blocks = (A.B + tiles - 1) / tiles;
    } else if (c.team == CDN_TEAM_BLOCK) {
        if (!A.ws) return cdn_fail(ctx, CDN_ERR_ARG, "engine: block team needs a workspace");
        c.blocks = A.B;
        c.tpb = ((ms & ~(unsigned)CDN_MS_DIODE) == 0 && D->n >= 1024) ? 1024 : 256;
        // LU values + two work vectors in shared memory when they fit
        const long need = ((long)D->nlu + D->n) * 8;
        if (!D->dense && need <= 220 * 1024) {
            c.smem = need;
            A.smem_lu = 1;
            const long st = compact_stage_bytes(*D);
            if (D->compact && need + st <= 226 * 1024) { c.smem = need + st; A.smem_lu = 2; }
        }
    } else if (c.team == CDN_TEAM_GRID) {
        if (!A.ws || !A.red) return cdn_fail(ctx, CDN_ERR_ARG, "engine: grid team needs workspace and reduction scratch");
        if (A.B != 1) return cdn_fail(ctx, CDN_ERR_ARG, "engine: grid team runs one circuit");
        if (D->dense) return cdn_fail(ctx, CDN_ERR_ARG, "engine: grid team needs a sparse plan");
        c.blocks = r->blocks;
    } else {
        return cdn_fail(ctx, CDN_ERR_ARG, "engine: unknown team");
    }

    cdn_plan_note_team(P, c.team);
    static const cdn_launch_fn table[CDN_NSET][CDN_NKIND] = {
        {cdn_launch_diode_tile8, cdn_launch_diode_tile16, cdn_launch_diode_tile32,
         cdn_launch_diode_block, cdn_launch_diode_grid},
        {cdn_launch_bsim3_tile8, cdn_launch_bsim3_tile16, cdn_launch_bsim3_tile32,
         cdn_launch_bsim3_block, cdn_launch_bsim3_grid},
        {cdn_launch_all_tile8, cdn_launch_all_tile16, cdn_launch_all_tile32, cdn_launch_all_block,
         cdn_launch_all_grid}};
    const int set = ((ms & ~(unsigned)CDN_MS_DIODE) == 0) ? CDN_SET_DIODE
                    : ((ms & ~(unsigned)CDN_MS_BSIM3) == 0) ? CDN_SET_BSIM3 : CDN_SET_ALL;
    int kind = CDN_KIND_GRID;
    if (c.team == CDN_TEAM_BLOCK) kind = CDN_KIND_BLOCK;
    else if (c.team == CDN_TEAM_TILE)
        kind = (c.T <= 8) ? CDN_KIND_TILE8 : (c.T <= 16) ? CDN_KIND_TILE16 : CDN_KIND_TILE32;
    const cudaError_t e = table[set][kind](A, c, st);
    if (e != cudaSuccess)
        return cdn_fail(ctx, CDN_ERR_CUDA, std::string("engine launch: ") + cudaGetErrorString(e));
    return CDN_OK;
}

}  // extern "C"
