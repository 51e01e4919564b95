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

/* measured FP64 FMA throughput of this GPU in FLOP/s (roofline denominator
 * for the device-evaluation kernels; none is published in MEASURED_PEAKS). */
int cdn_fp64_peak(cdn_ctx* ctx, double* flops_per_s);

#ifdef __cplusplus
}
#endif
#endif
