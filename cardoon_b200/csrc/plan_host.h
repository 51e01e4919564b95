// Host-side circuit description handed to cdn_plan_create (plain C structs,
// mirrored by ctypes in cardoon_b200/_lib.py).  Index lists follow the
// reference's nD_* attributes (analyses/nodal.py:205-282): -1 = reference
// node, dropped.  Port lists are k-major: vpos[k*n + i].
#pragma once

typedef struct cdn_group_host {
    int model;
    unsigned flags;
    int n, nxin, ncs, nqs, ndelay;
    const int* vpos;
    const int* vneg;
    const int* cpos;
    const int* cneg;
    const int* qpos;
    const int* qneg;
} cdn_group_host;

typedef struct cdn_circuit_host {
    int n;
    int ordering;                 /* 0 = nested dissection, 1 = natural,
                                     2 = dense storage, 3 = colorder given */
    long nG; const int* g_row; const int* g_col; const double* g_val;
    long nC; const int* c_row; const int* c_col; const double* c_val;
    long nE; const int* e_row; const int* e_col; const double* e_val;
    int ngroups;
    const cdn_group_host* groups;
    const int* rowmatch;          /* [n] column matched to each row, or NULL */
    const int* colorder;          /* [n] unknown eliminated at each position (ordering 3) */
} cdn_circuit_host;

struct cdn_plan;
struct cdn_plan_dev;
/* host copy of the device view (pointers are device pointers once bound) */
const cdn_plan_dev* cdn_plan_host_view(cdn_plan* P);
/* uploads the view struct if it changed; returns its device address */
int cdn_plan_sync_device(cdn_plan* P, cudaStream_t st, const cdn_plan_dev** out);
cdn_ctx* cdn_plan_ctx(cdn_plan* P);
