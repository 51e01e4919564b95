/*
 * libcardoon_b200.so -- C ABI of the B200-native (sm_100a) nonlinear Newton
 * hot path of the cardoon circuit simulator.
 *
 * The reference (cechrist/cardoon) is pure Python and has no FFI of its own;
 * each entry point below names the reference interface it replaces
 * (file:line under the reference tree).  All `const double*` / `double*` /
 * `int*` arguments documented as "device" are CUDA device pointers owned by
 * the caller (the Python side keeps them in torch tensors); the library never
 * returns memory it allocated.  Functions return 0 on success or a negative
 * CDN_ERR_* code; cdn_last_error() gives the message.  One context per GPU,
 * not re-entrant per context.  `stream` is a cudaStream_t (0 = default).
 */
#ifndef CARDOON_B200_H
#define CARDOON_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cdn_ctx cdn_ctx;

#define CDN_OK 0
#define CDN_ERR_ARG (-1)
#define CDN_ERR_CUDA (-2)
#define CDN_ERR_STATE (-3)
#define CDN_ERR_SINGULAR (-4)

/* ---- context ----------------------------------------------------------- */
int cdn_version(void);
int cdn_init(int device, cdn_ctx** ctx);
int cdn_free(cdn_ctx* ctx);
const char* cdn_last_error(cdn_ctx* ctx);
int cdn_sm_count(cdn_ctx* ctx);
/* number of packed parameters model `model_id` reads (param_layout.py) */
int cdn_model_nparams(int model_id);

/* ---- (a7) batched device evaluation ------------------------------------
 * Replaces Element.eval_and_deriv / Element.eval
 * (cardoon/devices/cppaddev.py:134-170) over eval_cqs of diode.py:275,
 * bjt.py:485, svbjt.py:308, mosBSIM3v3.py:403 (+mosExt.py:281),
 * mosEKV.py:391, mesfetc.py:192, memristor.py:178.
 *   params[k*ldp + pidx[i]] (device; pidx may be NULL = identity)
 *   xin[c*n + i], out[r*n + i], jac[(r*nxin + c)*n + i]   (device)
 * jac == NULL evaluates values only.                                       */
int cdn_dev_eval(cdn_ctx* ctx, int model_id, unsigned flags, long n, int nxin,
                 const double* params, long ldp, const int* pidx,
                 const double* xin, double* out, double* jac, void* stream);

/* Same evaluation for mosbsim3 / bsim3_i devices that share model cards: the
 * rows of the packed record on which all n instances agree are passed once
 * (card[nrows], device), the others per instance (inst[vrow[k]*ld + i],
 * device); vrow[nrows] (host): row of `inst` holding record row k, or -1 for
 * a card row.  Values and Jacobian are those of cdn_dev_eval on the expanded
 * record, bit for bit.                                                       */
int cdn_dev_eval_split(cdn_ctx* ctx, int model_id, unsigned flags, long n,
                       const double* card, int nrows, const short* vrow,
                       const double* inst, long ld, const double* xin,
                       double* out, double* jac, void* stream);

/* ---- (f2) small-signal AC sweep -------------------------------------------
 * Replaces the per-frequency loop of run_AC (cardoon/analyses/nodal.py:286-381):
 * Y = G + j w C, set_Jac of the device Jacobians with exp(-j w tau) on
 * delayed columns, linalg.solve(Y, sVec).  One CTA per frequency.
 *   G, C        [n*n] row-major, with the frequency-independent device stamps
 *   (drow, dcol, dg, dc, dtau)[ndel]   Y[drow,dcol] += (dg + j w dc) exp(-j w dtau)
 *   (frow, fcol)[nfd], fval[nf][nfd][2]   Y_k[frow,fcol] += fval[k] (Y parameters
 *               of frequency-defined elements, get_Y_matrix, nodal.py:366-370)
 *   freqs[nf] (Hz); s[n][2] (re, im); x[nf][n][2]
 *   work        work_matrices * n*n*16 bytes when n > 56 (else may be NULL):
 *               one column-major matrix per CTA in flight; the launch uses
 *               min(nf, 2 * SMs, work_matrices) CTAs, each looping over its
 *               share of the frequencies (cdn_ac_work_matrices gives the
 *               count that keeps every SM busy); n <= 2048
 *   status      int: 0, or 1 + index of a frequency with a singular matrix
 * All pointers are device pointers.                                          */
int cdn_ac_solve(cdn_ctx* ctx, int n, int nf, const double* G, const double* C,
                 int ndel, const int* drow, const int* dcol, const double* dg,
                 const double* dc, const double* dtau, int nfd, const int* frow,
                 const int* fcol, const double* fval, const double* freqs,
                 const double* s, double* work, long work_matrices, double* x,
                 int* status, void* stream);
/* workspace matrices cdn_ac_solve can use for nf frequencies (honours the
 * tuning knobs)                                                              */
long cdn_ac_work_matrices(cdn_ctx* ctx, int nf);

/* ---- (a1-a4, a12) formulation plan: symbolic analysis, done once ---------
 * Replaces what the reference redoes on every Newton iteration: COO -> CSC
 * conversion with duplicate summing, COLAMD ordering, symbolic and numeric
 * SuperLU factorisation (cardoon/analyses/nodalSP.py:281-303) and the
 * per-element index bookkeeping of create_additional_indexes (:111-164).
 * Input = the index lists of make_nodal_circuit / process_nodal_element
 * (cardoon/analyses/nodal.py:113-282): -1 is the reference node (dropped).
 * Port lists are k-major over the n instances of a group: vpos[k*n + i].   */
typedef struct cdn_group_host {
    int model;            /* CDN_MODEL_* (csrc/param_index.h)                */
    unsigned flags;       /* model structure flags, uniform over the group   */
    int n, nxin, ncs, nqs, ndelay;
    const int* vpos;      /* [nxin][n] row of the + control terminal         */
    const int* vneg;      /* [nxin][n]                                       */
    const int* cpos;      /* [ncs][n]  row the current leaves                */
    const int* cneg;      /* [ncs][n]                                        */
    const int* qpos;      /* [nqs][n]                                        */
    const int* qneg;      /* [nqs][n]                                        */
    const double* delay;  /* [ndelay][n] delay of the last ndelay control
                             ports (delayedContPorts), or NULL               */
} cdn_group_host;

typedef struct cdn_circuit_host {
    int n;                /* unknowns                                        */
    int ordering;         /* 0 nested dissection, 1 natural, 2 dense storage
                             (partial pivoting), 3 colorder given            */
    long nG; const int* g_row; const int* g_col; const double* g_val; /* G   */
    long nC; const int* c_row; const int* c_col; const double* c_val; /* C   */
    long nE; const int* e_row; const int* e_col; const double* e_val; /* gmin
                             stepping pattern (nodal.py:491-515)             */
    int ngroups;
    const cdn_group_host* groups;
    const int* rowmatch;  /* [n] unknown whose pivot row each row is, or NULL */
    const int* colorder;  /* [n] unknown eliminated at each position (3)     */
    int heavy;            /* gather/Crout lists longer than this are reduced by
                             the whole thread team (0 = default 256)         */
    int skip_q;           /* 1: DC plan, charge outputs contribute nothing   */
    int nhist;            /* transient with delayed ports: samples of port-
                             voltage history kept per delayed port (>= ceil(
                             max delay / h) + 2); 0 = ports are not delayed  */
} cdn_circuit_host;

typedef struct cdn_plan cdn_plan;

int cdn_init_host(cdn_ctx** ctx);      /* context without a GPU (plans only) */
int cdn_plan_create(cdn_ctx* ctx, const cdn_circuit_host* c, cdn_plan** out);
long cdn_plan_bytes(cdn_plan* plan);   /* device bytes cdn_plan_bind needs   */
/* info[0..11] = n, nlu, factor levels, L levels, U levels, Crout terms,
 * device contributions, nnz(J), longest term list, longest contribution
 * list, doubles of device outputs, doubles of device Jacobians, thread team
 * of the last cdn_engine_run (-1 none, 0 tile, 1 CTA, 2 grid); 16 longs    */
int cdn_plan_info(cdn_plan* plan, long* info);
int cdn_plan_bind(cdn_plan* plan, void* dev_ws, long bytes, void* stream);
/* params[k*ldp + col] (device); per_inst = 1: instance b of a batch reads
 * columns b*n + i, else every instance reads column i                      */
int cdn_plan_set_group_params(cdn_plan* plan, int group, const double* params,
                              long ldp, const int* pidx, int per_inst);
/* (row, unknown) of every Jacobian slot, original numbering (host arrays)  */
int cdn_plan_slot_coords(cdn_plan* plan, int* rows, int* cols);
/* host copy of a packed plan array by field name (csrc/plan.h)             */
int cdn_plan_array(cdn_plan* plan, const char* name, const void** ptr,
                   long* count);
int cdn_plan_free(cdn_plan* plan);

/* ---- (a6-a16) device-resident Newton engine ------------------------------
 * One launch per operation; the state of every circuit instance (x, dx,
 * iVec, sV, q, integrator history, LU factors, device outputs) lives on the
 * device (caller-provided `ws`, or shared memory for batches of small
 * circuits when ws == NULL).  Replaces, per operation:
 *   CDN_OP_NEWTON     fsolve.solve / fsolve_Newton (cardoon/analyses/fsolve.py:
 *                     23-128) with the helper chain of nodal.py:452-604 over
 *                     get_i_Jac + factor_and_solve (nodal.py:455-457,
 *                     787-821, 606-627; nodalSP.py:494-523, 281-303)
 *   CDN_OP_TRAN       the time loop of cardoon/analyses/tran.py:171-206
 *                     (accept, get_rhs, chord predictor, solve, store)
 *   CDN_OP_UPDATE_Q   TransientNodal.update_q   (nodalSP.py:719-733)
 *   CDN_OP_SET_IC     TransientNodal.set_IC     (nodalSP.py:666-696)
 *   CDN_OP_ACCEPT     TransientNodal.accept     (nodalSP.py:735-758)
 *   CDN_OP_RHS        TransientNodal.get_rhs    (nodalSP.py:792-802)
 *   CDN_OP_CHORD      get_chord_deltax          (nodalSP.py:345-354)
 *   CDN_OP_GET_I      get_i                     (nodalSP.py:471-492, 804-836)
 *   CDN_OP_GET_I_JAC  get_i_Jac                 (nodalSP.py:494-523, 838-887)
 *   CDN_OP_SOLVE      factor_and_solve          (nodalSP.py:281-303)
 * Non-convergence is a status, not an error: status[b] = -1 (NEWTON) or the
 * first failed sample (TRAN); the host raises NoConvergenceError.
 * Dense plans follow factor_and_solve of nodal.py:606-627: LAPACK getrf/getrs
 * semantics (a zero pivot does not stop the factorisation) followed by
 * condition_array (nodal.py:385-402) on the step; a Jacobian or right-hand
 * side that already holds NaN/inf fails that Newton solve at once (next
 * helper / bisection), see oracle/nodal_ref.py factor_and_solve.
 * State: with ws == NULL the state of an instance lives in shared memory for
 * one launch only, so operations that continue from earlier state (TRAN with
 * first > 1 or init_ic == 0, CHORD, ACCEPT, RHS, SOLVE, UPDATE_Q, SET_IC,
 * GET_I, GET_I_JAC) return CDN_ERR_STATE.                                    */
#define CDN_OP_NEWTON 0
#define CDN_OP_TRAN 1
#define CDN_OP_UPDATE_Q 2
#define CDN_OP_SET_IC 3
#define CDN_OP_ACCEPT 4
#define CDN_OP_RHS 5
#define CDN_OP_CHORD 6
#define CDN_OP_GET_I 7
#define CDN_OP_GET_I_JAC 8
#define CDN_OP_SOLVE 9

typedef struct cdn_run_host {
    /* .options read by Newton (cardoon/globalVars.py:44-57) */
    double abstol, reltol, maxdelta;
    double a0;           /* integration coefficient: 2/h trap, 1/h BE, 0 DC  */
    double gmin;         /* gmin stepping (multiplies the E pattern)         */
    double gmin_ext;     /* gmin stepping on the diagonal of the first n_ext
                            unknowns (solve_homotopy_gmin, nodal.py:469-489;
                            dense plans only)                                */
    double lambda;       /* source stepping factor (1 = full sources)        */
    double h;            /* time step (delayed-port interpolation)           */
    int maxiter, errfunc;
    int use_q;           /* 0: DC (currents only), 1: transient              */
    int im;              /* 0 backward Euler, 1 trapezoidal                  */
    int op, B;           /* operation, circuit instances in the batch        */
    int team, T;         /* -1 choose / 0 warp tile / 1 CTA / 2 grid; lanes  */
    int helpers;         /* 1: on failure continue in-kernel with the helper
                            chain of fsolve.py:23-56: gmin stepping on the
                            nonlinear ports, source stepping and -- dense
                            plans with n_ext > 0, nodal.py:443-447 -- gmin
                            stepping on the external nodes                   */
    int n_ext;           /* external terminals (ckt.nD_nterms): they are the
                            first unknowns (nodal.py:472-475)                */
    double* x;           /* [B][n] in/out (device)                           */
    const double* sv_in; /* [n] source vector (device)                       */
    double* ws;          /* [B][cdn_engine_ws_doubles] state, or NULL        */
    int nsamples, first, init_ic, nsrc, nsave, blocks;
    const int* src_rows;     /* [nsrc] rows with a time-varying source       */
    const double* src_val;   /* [nsamples][nsrc]                             */
    const double* src_dc;    /* [n] constant part of the source vector       */
    const int* save_rows;    /* [nsave] unknowns to record                   */
    double* wave;            /* [B][nsamples][nsave]                         */
    int* iters;              /* [B][nsamples] (TRAN) or [B]                  */
    double* res;             /* same shape as iters                          */
    int* status;             /* [B]                                          */
    double* red;             /* grid team: reduction scratch                 */
    long long* prof;         /* optional [16] (device): SM cycles instance 0
                                spent per phase: device sweep, assembly, factor,
                                solve, update, accept (+rhs), chord            */
    int* info;               /* optional [B][2] (device): [0] highest helper
                                that had to be used (0 plain Newton, 1 gmin on
                                the nonlinear ports, 2 source stepping, 3 gmin
                                on the external nodes); [1] 1 + first sample
                                (TRAN) or 1 (NEWTON) at which the static-pivot
                                sparse LU saw a zero / non-finite pivot or a
                                multiplier above 1e8 (the host then re-plans at
                                the current iterate), else 0                   */
} cdn_run_host;

long cdn_engine_ws_doubles(cdn_plan* plan);
/* offsets (doubles) of x, dx, ivec, sv, q, qprev, dq, y, tmp, xb, lu, out,
 * jac, hist, piv inside one instance's state (15 longs)                                          */
int cdn_engine_layout(cdn_plan* plan, long* off);
int cdn_engine_run(cdn_plan* plan, const cdn_run_host* run, void* stream);

/* ---- boundary / interior transient engine (one sparse circuit on a thread-
 * block cluster) ------------------------------------------------------------------
 * Replaces, for circuits whose nonlinear devices touch few unknowns, the
 * refactor-everything-per-iteration time loop of the sparse back-end
 * (cardoon/analyses/tran.py:171-206 over nodalSP.py:281-354, :666-887): the
 * linear interior is solved with precomputed inverse factors, only the banded
 * boundary system on the device ports is assembled and factored per Newton
 * iteration.  The operator blobs come from cardoon_b200/analyses/schur.py
 * (pack_blob); `run` is the same structure cdn_engine_run takes (op
 * CDN_OP_TRAN, B = 1, first = 1, init_ic = 1; src_rows is replaced by
 * src_own / src_loc, save_rows by save_ptr / save_k / save_loc).  status[0] != 0
 * (Newton failed at that sample: the helper chain is not available here) or
 * info[1] != 0 (a pivot of the boundary system degraded) tell the caller to
 * run cdn_engine_run instead.                                                  */
typedef struct cdn_schur_host {
    int ncta;                 /* CTAs of the cluster (1..16)                       */
    int nB, band;             /* boundary unknowns, half bandwidth of T            */
    const void* blob;         /* device: per-CTA operator blobs                    */
    const long* cta_off;      /* host [ncta + 1]: byte offsets into blob           */
    const int* m;             /* host [ncta]: interior unknowns per CTA            */
    const int* const* vpos_b; /* host [ngroups] of device [nxin][n]: control-port
                                 rows as boundary positions (-1 ground)            */
    const int* const* vneg_b;
    const int* src_own;       /* device [nsrc]: owner CTA of a source row (-1: B)  */
    const int* src_loc;       /* device [nsrc]: position in the local vector       */
    const int* save_ptr;      /* device [ncta + 1]                                 */
    const int* save_k;        /* device: column of wave                            */
    const int* save_loc;      /* device: position in the CTA's local vector        */
} cdn_schur_host;
long cdn_schur_smem_bytes(long blob_bytes, int m, int nB, int band, int nout,
                          int njac, int ncta);
int cdn_schur_tran(cdn_plan* plan, const cdn_run_host* run,
                   const cdn_schur_host* s, void* stream);

/* tuning knobs for experiments (key 0: resident CTAs per SM the BSIM3
 * evaluation kernel is register-limited for; 0 = default)                    */
int cdn_set_tuning(int key, int value);

/* measured FP64 FMA throughput of this GPU in FLOP/s (roofline denominator
 * for the device-evaluation kernels; none is published in MEASURED_PEAKS). */
int cdn_fp64_peak(cdn_ctx* ctx, double* flops_per_s);

#ifdef __cplusplus
}
#endif
#endif
