// Host/device description of one circuit "plan": the precomputed index
// structures that let the Newton kernels evaluate devices, assemble the MNA
// residual and Jacobian by *gather* (no atomics), factor it with a static
// pivot order and solve, all on the device.
//
// Built once per circuit and analysis kind by symbolic.cpp from the index
// lists of the reference's make_nodal_circuit / process_nodal_element /
// create_additional_indexes (analyses/nodal.py:113-282, nodalSP.py:111-164).
#pragma once

#define CDN_MAX_GROUPS 16
// gather / term lists longer than this are summed cooperatively by the team
#define CDN_HEAVY 256
// small dense tile programs (see cdn_plan_dev::sd): the factor storage of one
// instance (rows 0..63, reciprocal pivots 64..71, steps and perm 72..76) is
// followed by CDN_SD8_NFIX partial accumulators and one always-zero double
#define CDN_SD8_NFIX 10
#define CDN_SD8_PART 77
#define CDN_SD8_FACT (CDN_SD8_PART + CDN_SD8_NFIX + 1)

// one homogeneous group of nonlinear devices (same model, flags, port counts)
struct cdn_group_dev {
    int model;
    unsigned flags;
    int n, nxin, ncs, nqs, ndelay;
    int pad_;
    const double* params;     // [npar][ldp] device, owned by caller
    long ldp;
    const int* pidx;          // column of each instance in params, or NULL
    int per_inst;             // 1: every circuit instance b of a batch has its
    int pad2_;                //    own columns (b*n + i); 0: shared by the batch
    const int* vpos;          // [nxin][n] row of + control terminal (-1 gnd)
    const int* vneg;          // [nxin][n]
    long out_off;             // offset of this group's out[nout][n] in dev_out
    long jac_off;             // offset of jac[nout*nxin][n] in dev_jac
    const double* delay;      // [ndelay][n] delays of the last ndelay control ports
    long hist_off;            // offset of hist[ndelay][n][nhist] in the history buffer
};

struct cdn_plan_dev {
    int n;                    // unknowns
    int nlu;                  // stored entries of L+U (static pattern)
    int nlev_f, nlev_l, nlev_u;
    int ngroups;
    int dense;                // 1: slots are a dense n*n row-major matrix
    int ndev;                 // nonlinear devices of one circuit instance
    int max_terms;
    int ncon, nrcon, nrlin;   // lengths of e_con_*, r_con_*, r_lin_*
    int heavy;                // list length above which the team reduces together
    int nhist;                // history samples per delayed port (0: no delays)
    long hist_size;           // doubles of port-voltage history per circuit instance
    int n_e_heavy, n_r_heavy;
    const int* e_heavy;       // [n_e_heavy] slots with long contribution lists
    const int* r_heavy;       // [n_r_heavy] rows with long gather lists
    const int* f_lvl_nheavy;  // [nlev_f] heavy entries at the head of each level
    const int* l_lvl_nheavy;  // [nlev_l]
    const int* u_lvl_nheavy;  // [nlev_u]
    const int* dev_g;         // [ndev] group of device d (flattened list)
    const int* dev_i;         // [ndev] index inside its group
    // ---- factorisation: entries ("slots") numbered in level order --------
    const int* f_lvl_ptr;     // [nlev_f+1]
    const int* e_term_ptr;    // [nlu+1]   Crout terms of entry e
    const int* e_term_l;      // [nterms]  slot of L(i,k)
    const int* e_term_u;      // [nterms]  slot of U(k,j)
    const int* e_piv;         // [nlu]  L entry: slot of U(j,j); U: -1; diag: -2
    const double* e_g;        // [nlu] linear conductance part
    const double* e_c;        // [nlu] linear capacitance part (times a0)
    const double* e_gones;    // [nlu] gmin-stepping pattern (times gmin)
    const int* e_con_ptr;     // [nlu+1] device-Jacobian contributions
    const int* e_con_src;     // [ncon]  index into dev_jac
    const signed char* e_con_kind;  // +-1 current row, +-2 charge row (x a0)
    const unsigned* e_con_pk;  // src | minus << 29 | charge << 30 (what the kernels read)
    const unsigned* r_con_pk;
    // compact triangular-solve programs (n, nlu < 65536, no hub rows): per
    // level-ordered row two words {row | count << 16, first entry | diag slot
    // << 16} and 16-bit column / slot streams; small enough to be staged in
    // shared memory next to the LU values (block team)
    int compact, nnzl, nnzu;
    // one contiguous blob of 32-bit words per sweep (copied with 16-byte loads):
    //   L / U solve: [level ptr, then log2 of the lanes per row of each level
    //                 | row records (2 words) | columns (16 bit) |
    //                 slots (16 bit) | rowmap / colmap (16 bit)]
    //   partial refactorisation: [level ptr, lanes per entry as above | entry
    //                 records (2 words: slot |
    //                 terms << 16, first term | pivot slot << 16) | terms
    //                 (L slot | U slot << 16)] over the entries that depend on a
    //                 device contribution; all other entries keep the values of the
    //                 last full factorisation
    const unsigned* cl_prog;  int cl_words; int cl_off[5];
    const unsigned* cu_prog;  int cu_words; int cu_off[5];
    const unsigned* cf_prog;  int cf_words; int cf_off[3];
    int naff;                 // entries in the partial refactorisation
    const int* e_bal;         // [nlu] dense plans: slots by decreasing gather length, so
                              // that a strided walk gives every lane the same work
    // ---- small dense systems (dense, n <= 8): programs of the 8-lane warp tiles
    // (engine.cuh sd8_*).  The device-Jacobian contributions of all slots, slot
    // after slot, are cut into eight equally long pieces, one per lane (lane l:
    // sd_prog[l * sd_maxlen .. (l + 1) * sd_maxlen), padded), so that every lane
    // walks one flat list of the same length:
    //   term = source | target << 14 | store target and reset << 30 | minus << 31
    // Source and target are positions (in doubles) inside one instance's state
    // (engine.cuh ws_carve): a device-Jacobian entry, and the slot of w.lu the
    // running sum belongs to.  A slot that straddles a cut continues in the next
    // lane on partial accumulator j (in the factor storage), which lane 0 adds
    // to slot sd_fix[j] afterwards (at most CDN_SD8_NFIX of them); padding terms
    // add the always-zero double.
    int sd;                   // 1: programs present
    int sd_maxlen;            // longest lane list
    int sd_rmax;              // longest per-row list of device-output contributions
    int sd_nterms;            // length of sd_prog
    const int* sd_ptr;        // [9] lane l owns sd_prog[sd_ptr[l] .. sd_ptr[l + 1])
    const unsigned* sd_prog;
    int sd_nfix;              // partial accumulators
    int sd_pad_;
    const int* sd_fix;        // [sd_nfix] slot each partial belongs to
    // ---- triangular solves (permuted index space) --------------------------
    const int* l_lvl_ptr;     // [nlev_l+1]
    const int* l_rows;        // [n] rows ordered by level
    const int* l_ptr;         // [n+1]
    const int* l_col;         // [nnzL]
    const int* l_slot;        // [nnzL]
    const int* u_lvl_ptr;     // [nlev_u+1]
    const int* u_rows;        // [n]
    const int* u_ptr;         // [n+1]
    const int* u_col;         // [nnzU] (strictly upper)
    const int* u_slot;        // [nnzU]
    const int* u_diag;        // [n] slot of U(i,i)
    const int* rowmap;        // [n] permuted row i  <- original row rowmap[i]
    const int* colmap;        // [n] permuted col j  -> original unknown colmap[j]
    // ---- residual (original index space) ------------------------------------
    const int* r_lin_ptr;     // [n+1] CSR of the linear matrices
    const int* r_lin_col;
    const double* r_lin_g;
    const double* r_lin_c;
    const double* r_lin_gones;
    const int* r_con_ptr;     // [n+1] device-output contributions per row
    const int* r_con_src;     // index into dev_out
    const signed char* r_con_kind;   // +-1 current, +-2 charge
    long dev_out_size, dev_jac_size;
    cdn_group_dev groups[CDN_MAX_GROUPS];
};
