// Postfix interpreter for netlist-string device laws (memristor M(q)).
// Reference: devices/memristor.py eval_cqs :178-197 (the reference eval()s
// the string inside the AD tape; here it is compiled on the host by
// devices/exprcode.py and interpreted in dual arithmetic).
// Program layout in the parameter column: [.., n, op0, arg0, op1, arg1, ...].
#pragma once
#include "junction.cuh"

#define EXPR_STACK 16

template <int P>
CDN_HD Dual<P> expr_eval(const PV& p, int base, const Dual<P>& var) {
    Dual<P> st[EXPR_STACK];
    int sp = 0;
    const int n = (int)p[base];
    for (int k = 0; k < n; ++k) {
        const int op = (int)p[base + 1 + 2 * k];
        const double arg = p[base + 2 + 2 * k];
        switch (op) {
        case 0: st[sp++] = Dual<P>(arg); break;
        case 1: st[sp++] = var; break;
        case 2: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;
        case 3: sp--; st[sp - 1] = st[sp - 1] - st[sp]; break;
        case 4: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;
        case 5: sp--; st[sp - 1] = st[sp - 1] / st[sp]; break;
        case 6: {   // power: constant integer/real exponent or general
            sp--;
            const Dual<P> e = st[sp];
            bool cst = true;
#pragma unroll
            for (int i = 0; i < P; ++i) cst = cst && (e.d[i] == 0.0);
            if (cst && e.v == 2.0) st[sp - 1] = st[sp - 1] * st[sp - 1];
            else if (cst) st[sp - 1] = dpow(st[sp - 1], e.v);
            else st[sp - 1] = exp(e * log(st[sp - 1]));
            break;
        }
        case 7: st[sp - 1] = -st[sp - 1]; break;
        case 8: st[sp - 1] = dabs(st[sp - 1]); break;
        case 9: st[sp - 1] = exp(st[sp - 1]); break;
        case 10: st[sp - 1] = log(st[sp - 1]); break;
        case 11: st[sp - 1] = sqrt(st[sp - 1]); break;
        case 12: st[sp - 1] = tanh(st[sp - 1]); break;
        case 13: st[sp - 1] = sin(st[sp - 1]); break;
        case 14: st[sp - 1] = cos(st[sp - 1]); break;
        case 15: st[sp - 1] = tan(st[sp - 1]); break;
        case 16: st[sp - 1] = cosh(st[sp - 1]); break;
        default: break;
        }
    }
    return st[0];
}

// memristor: in v = [v_im, v_vc]; out i = [gyr^2 * M(c*v_vc) * v_im]
template <int P, class S>
CDN_HD void memr_eval(const PV& p, unsigned flags, const Dual<P>* v, S& out) {
    const double gyr = p[P_MEMR_gyr];
    const Dual<P> q = p[P_MEMR_c] * v[1];
    const Dual<P> M = expr_eval(p, NP_MEMR, q);
    out.put(0, (gyr * gyr) * M * v[0]);
}
