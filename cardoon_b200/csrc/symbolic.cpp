// Host-side symbolic analysis for the sparse Newton engine: done once per
// circuit (the reference re-runs COLAMD + symbolic + numeric SuperLU on every
// Newton iteration, analyses/nodalSP.py:281-303).
//
//   1. pattern of J from the linear quads and every device's +/- stamp
//      positions (nodalSP.py:45-85, 111-164)
//   2. static pivoting: the caller supplies a row<->column matching that puts
//      large entries on the diagonal (gyrator-built sources and inductors
//      have structurally zero diagonals, devices/vdc.py:112-123)
//   3. nested-dissection ordering of the symmetrised pattern (level-structure
//      bisection; hub nodes last).  Chain-like circuits (soliton line,
//      inverter chains) have ~n elimination levels under AMD/COLAMD but
//      O(log n) under ND, which is what makes level scheduling pay
//   4. symbolic factorisation (elimination tree + row patterns)
//   5. entry-wise Crout schedule: every stored entry of L+U gets the list of
//      (L(i,k), U(k,j)) products it subtracts and a dependency level; entries
//      are renumbered level by level so one barrier separates levels and all
//      arithmetic is a deterministic gather (no atomics)
//   6. gather lists for the device Jacobian/outputs -> matrix entries / rows
//   7. row level schedules for the two triangular solves
#include "common.cuh"
#include "plan.h"
#include "plan_host.h"
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <map>
#include <string>

namespace {

typedef std::vector<int> ivec;

// ---- nested dissection ----------------------------------------------------------
struct Graph {
    int n;
    ivec ptr, adj;
};

struct NDOrder {
    const Graph& g;
    ivec setid;       // current subset label of each node
    ivec level;       // BFS scratch
    ivec order;       // output: elimination order (new -> old)
    int next_label;
    int leaf;

    NDOrder(const Graph& g_, int leaf_) : g(g_), setid(g_.n, 0), level(g_.n, -1),
                                          next_label(1), leaf(leaf_) {}

    // BFS inside subset `lab` from `start`; returns visit order, fills level[]
    void bfs(int start, int lab, ivec& visit) {
        visit.clear();
        visit.push_back(start);
        level[start] = 0;
        for (size_t h = 0; h < visit.size(); ++h) {
            const int u = visit[h];
            for (int e = g.ptr[u]; e < g.ptr[u + 1]; ++e) {
                const int w = g.adj[e];
                if (setid[w] == lab && level[w] < 0) {
                    level[w] = level[u] + 1;
                    visit.push_back(w);
                }
            }
        }
    }

    void emit_by_degree(ivec& nodes) {
        std::stable_sort(nodes.begin(), nodes.end(), [&](int a, int b) {
            return (g.ptr[a + 1] - g.ptr[a]) < (g.ptr[b + 1] - g.ptr[b]);
        });
        for (int v : nodes) order.push_back(v);
    }

    void run(ivec nodes) {
        // explicit work stack; separators are emitted after both halves, which
        // is arranged by pushing a "emit" item below them
        struct Item { ivec nodes; bool emit_only; };
        std::vector<Item> stack;
        stack.push_back(Item{std::move(nodes), false});
        ivec visit, visit2;
        while (!stack.empty()) {
            Item it = std::move(stack.back());
            stack.pop_back();
            if (it.emit_only) { for (int v : it.nodes) order.push_back(v); continue; }
            ivec& S = it.nodes;
            if ((int)S.size() <= leaf) { emit_by_degree(S); continue; }
            const int lab = next_label++;
            for (int v : S) { setid[v] = lab; level[v] = -1; }
            // start from a minimum-degree node
            int start = S[0];
            for (int v : S)
                if (g.ptr[v + 1] - g.ptr[v] < g.ptr[start + 1] - g.ptr[start]) start = v;
            bfs(start, lab, visit);
            if (visit.size() < S.size()) {
                // disconnected: peel this component off and handle both parts
                ivec comp(visit), rest;
                for (int v : comp) setid[v] = -lab;
                for (int v : S) if (setid[v] == lab) rest.push_back(v);
                stack.push_back(Item{std::move(rest), false});
                stack.push_back(Item{std::move(comp), false});
                continue;
            }
            // pseudo-peripheral node: restart from the farthest node
            const int far = visit.back();
            for (int v : S) level[v] = -1;
            bfs(far, lab, visit2);
            const int nlev = level[visit2.back()] + 1;
            if (nlev < 3) { emit_by_degree(S); continue; }
            ivec cnt(nlev, 0);
            for (int v : S) cnt[level[v]]++;
            // separator level: the interior level where the halves balance
            const long half = (long)S.size() / 2;
            long acc = 0;
            int m = 1;
            for (int l = 0; l < nlev; ++l) {
                if (acc + cnt[l] > half) { m = l; break; }
                acc += cnt[l];
            }
            if (m < 1) m = 1;
            if (m > nlev - 2) m = nlev - 2;
            ivec A, B, Sep;
            for (int v : S) {
                if (level[v] < m) A.push_back(v);
                else if (level[v] > m) B.push_back(v);
                else Sep.push_back(v);
            }
            stack.push_back(Item{std::move(Sep), true});
            stack.push_back(Item{std::move(B), false});
            stack.push_back(Item{std::move(A), false});
        }
    }
};

template <class T>
static long put(std::vector<char>& buf, const std::vector<T>& v) {
    long off = (long)((buf.size() + 255) / 256 * 256);
    buf.resize(off + std::max<size_t>(v.size(), 1) * sizeof(T));
    if (!v.empty()) memcpy(buf.data() + off, v.data(), v.size() * sizeof(T));
    return off;
}

}  // namespace

struct cdn_plan {
    cdn_ctx* ctx;
    cdn_plan_dev dev;             // device pointers valid after bind
    std::vector<char> staging;    // packed index arrays
    struct Off {
        long f_lvl_nheavy, l_lvl_nheavy, u_lvl_nheavy, e_heavy, r_heavy;
        long f_lvl_ptr, e_term_ptr, e_term_l, e_term_u, e_piv, e_g, e_c, e_gones, e_con_ptr,
            e_con_src, e_con_kind, l_lvl_ptr, l_rows, l_ptr, l_col, l_slot, u_lvl_ptr, u_rows,
            u_ptr, u_col, u_slot, u_diag, rowmap, colmap, r_lin_ptr, r_lin_col, r_lin_g,
            r_lin_c, r_lin_gones, r_con_ptr, r_con_src, r_con_kind, dev_g, dev_i, e_con_pk,
            r_con_pk, e_bal, cl_prog, cu_prog, cf_prog, sd_ptr, sd_prog, sd_fix;
        long gv[CDN_MAX_GROUPS][2];
        long gd[CDN_MAX_GROUPS];
    } off;
    long nterms, ncon, nnzA;
    bool bound;
    bool dirty;
    int last_team;
    cdn_plan_dev* dev_struct;     // device copy of `dev`
    int max_terms, max_con;
    ivec slot_row, slot_col;      // original (row, unknown) of every slot
    std::map<std::string, std::pair<long, long> > reg;   // name -> (offset, count) in staging
};

static int find_slot(const ivec& rptr, const ivec& rcol, const ivec& rslot, int i, int j) {
    const int* b = rcol.data() + rptr[i];
    const int* e = rcol.data() + rptr[i + 1];
    const int* p = std::lower_bound(b, e, j);
    if (p == e || *p != j) return -1;
    return rslot[p - rcol.data()];
}

extern "C" int cdn_plan_create(cdn_ctx* ctx, const cdn_circuit_host* c, cdn_plan** out) {
    if (!ctx || !c || !out) return CDN_ERR_ARG;
    const int n = c->n;
    if (n <= 0) return cdn_fail(ctx, CDN_ERR_ARG, "plan: empty circuit");
    if (c->ngroups > CDN_MAX_GROUPS) return cdn_fail(ctx, CDN_ERR_ARG, "plan: too many device groups");
    cdn_plan* P = new cdn_plan();
    P->ctx = ctx;
    P->bound = false;
    P->dirty = true;
    P->last_team = -1;
    P->dev_struct = nullptr;
    memset(&P->dev, 0, sizeof(P->dev));

    // ---- permutation from the matching: row r of A becomes row mrow[r] ----------
    ivec mrow(n);
    {
        std::vector<char> seen(n, 0);
        for (int r = 0; r < n; ++r) {
            const int m = c->rowmatch ? c->rowmatch[r] : r;
            if (m < 0 || m >= n || seen[m]) { delete P; return cdn_fail(ctx, CDN_ERR_ARG, "plan: rowmatch is not a permutation"); }
            seen[m] = 1;
            mrow[r] = m;
        }
    }

    // ---- 1. pattern (in "B" coordinates: (mrow[r], c)) ------------------------------
    std::vector<uint64_t> keys;
    auto addkey = [&](int r, int col) {
        if (r >= 0 && col >= 0) keys.push_back(((uint64_t)(uint32_t)mrow[r] << 32) | (uint32_t)col);
    };
    for (long k = 0; k < c->nG; ++k) addkey(c->g_row[k], c->g_col[k]);
    for (long k = 0; k < c->nC; ++k) addkey(c->c_row[k], c->c_col[k]);
    for (long k = 0; k < c->nE; ++k) addkey(c->e_row[k], c->e_col[k]);
    for (int g = 0; g < c->ngroups; ++g) {
        const cdn_group_host& G = c->groups[g];
        for (int i = 0; i < G.n; ++i)
            for (int o = 0; o < G.ncs + G.nqs; ++o) {
                const int rp = (o < G.ncs) ? G.cpos[(long)o * G.n + i] : G.qpos[(long)(o - G.ncs) * G.n + i];
                const int rn = (o < G.ncs) ? G.cneg[(long)o * G.n + i] : G.qneg[(long)(o - G.ncs) * G.n + i];
                for (int k = 0; k < G.nxin; ++k) {
                    const int cp = G.vpos[(long)k * G.n + i], cn = G.vneg[(long)k * G.n + i];
                    addkey(rp, cp); addkey(rn, cn); addkey(rp, cn); addkey(rn, cp);
                }
            }
    }
    for (int i = 0; i < n; ++i) keys.push_back(((uint64_t)(uint32_t)i << 32) | (uint32_t)i);
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    P->nnzA = (long)keys.size();

    const bool dense = (c->ordering == 2);
    const int HEAVY = c->heavy > 0 ? c->heavy : CDN_HEAVY;
    ivec ip(n), rptr(n + 1, 0), rcol, rslot, ucount(n, 0);
    std::vector<ivec> Lrow(n);
    ivec f_lvl_ptr, f_lvl_nheavy, e_term_ptr, e_term_l, e_term_u, e_piv;
    long nlu = 0, nterms = 0;
    int nlev_f = 0;
    P->max_terms = 0;
    auto diag_id = [&](int i) { return rptr[i] + (int)Lrow[i].size(); };
    if (dense) {
        // dense storage for the partial-pivoting small-system kernels:
        // slot(r, c) = r*n + c in the original numbering, no schedule
        if ((long)n * n > 16000000L) { delete P; return cdn_fail(ctx, CDN_ERR_ARG, "plan: too large for dense storage"); }
        for (int i = 0; i < n; ++i) { ip[i] = i; mrow[i] = i; rptr[i + 1] = (i + 1) * n; }
        nlu = (long)n * n;
        rcol.resize(nlu); rslot.resize(nlu);
        for (long q = 0; q < nlu; ++q) { rcol[q] = (int)(q % n); rslot[q] = (int)q; }
        f_lvl_ptr.assign(1, 0);
        e_term_ptr.assign(nlu + 1, 0);
        e_piv.assign(nlu, -1);
    } else {
    // ---- 3. ordering on the symmetrised graph -----------------------------------------
    Graph gr;
    gr.n = n;
    {
        std::vector<uint64_t> ek;
        ek.reserve(keys.size() * 2);
        for (uint64_t k : keys) {
            const uint32_t a = (uint32_t)(k >> 32), b = (uint32_t)k;
            if (a != b) { ek.push_back(((uint64_t)a << 32) | b); ek.push_back(((uint64_t)b << 32) | a); }
        }
        std::sort(ek.begin(), ek.end());
        ek.erase(std::unique(ek.begin(), ek.end()), ek.end());
        gr.ptr.assign(n + 1, 0);
        gr.adj.resize(ek.size());
        for (uint64_t k : ek) gr.ptr[(k >> 32) + 1]++;
        for (int i = 0; i < n; ++i) gr.ptr[i + 1] += gr.ptr[i];
        for (size_t q = 0; q < ek.size(); ++q) gr.adj[q] = (int)(uint32_t)ek[q];
    }
    ivec perm;   // new -> old (B index)
    {
        const int hub_deg = std::max(64, (int)(10.0 * std::sqrt((double)n)));
        ivec normal, hubs;
        for (int v = 0; v < n; ++v)
            ((gr.ptr[v + 1] - gr.ptr[v]) > hub_deg ? hubs : normal).push_back(v);
        // hubs are removed from the graph seen by ND by giving them a label no
        // subset ever uses
        NDOrder nd(gr, 2);
        for (int v : hubs) nd.setid[v] = -1000000;
        if (c->ordering == 3 && c->colorder) {
            // explicit elimination order (pivot sequence validated on the host)
            hubs.clear();
            std::vector<char> seen(n, 0);
            for (int j = 0; j < n; ++j) {
                const int v = c->colorder[j];
                if (v < 0 || v >= n || seen[v]) { delete P; return cdn_fail(ctx, CDN_ERR_ARG, "plan: colorder is not a permutation"); }
                seen[v] = 1;
                nd.order.push_back(v);
            }
        } else if (c->ordering == 1) {
            for (int v : normal) nd.order.push_back(v);       // natural (debug)
        } else {
            nd.run(normal);
        }
        std::stable_sort(hubs.begin(), hubs.end(), [&](int a, int b) {
            return (gr.ptr[a + 1] - gr.ptr[a]) < (gr.ptr[b + 1] - gr.ptr[b]);
        });
        perm = nd.order;
        for (int v : hubs) perm.push_back(v);
        if ((int)perm.size() != n) { delete P; return cdn_fail(ctx, CDN_ERR_STATE, "plan: ordering lost nodes"); }
    }
    for (int i = 0; i < n; ++i) ip[perm[i]] = i;

    // ---- symmetric pattern of M = B[perm, perm], lower part by rows ----------------
    std::vector<ivec> low(n);    // low[i] = { k < i : M(i,k) or M(k,i) != 0 }
    for (uint64_t k : keys) {
        int a = ip[(uint32_t)(k >> 32)], b = ip[(uint32_t)k];
        if (a == b) continue;
        if (a < b) std::swap(a, b);
        low[a].push_back(b);
    }
    for (int i = 0; i < n; ++i) {
        std::sort(low[i].begin(), low[i].end());
        low[i].erase(std::unique(low[i].begin(), low[i].end()), low[i].end());
    }

    // ---- 4. symbolic factorisation: row patterns of L via the elimination tree ------
    {
        ivec parent(n, -1), mark(n, -1);
        for (int i = 0; i < n; ++i) {
            mark[i] = i;
            for (int k : low[i]) {
                int j = k;
                while (mark[j] != i) {
                    if (parent[j] < 0) parent[j] = i;
                    Lrow[i].push_back(j);
                    mark[j] = i;
                    j = parent[j];
                }
            }
            std::sort(Lrow[i].begin(), Lrow[i].end());
        }
    }
    std::vector<ivec>().swap(low);
    // U row patterns = transpose of L patterns
    for (int i = 0; i < n; ++i) for (int k : Lrow[i]) ucount[k]++;
    // natural numbering: row i -> [L cols..., diag, U cols...]
    for (int i = 0; i < n; ++i) rptr[i + 1] = rptr[i] + (int)Lrow[i].size() + 1 + ucount[i];
    nlu = rptr[n];
    if (nlu > 2000000000L) { delete P; return cdn_fail(ctx, CDN_ERR_STATE, "plan: fill too large"); }
    rcol.assign(nlu, 0);
    {
        ivec ufill(n, 0);
        for (int i = 0; i < n; ++i) {
            int q = rptr[i];
            for (int k : Lrow[i]) rcol[q++] = k;
            rcol[q] = i;
        }
        for (int i = 0; i < n; ++i)
            for (int k : Lrow[i]) {     // (i,k) in L  =>  (k,i) in U, rows k filled in increasing i
                const int q = rptr[k] + (int)Lrow[k].size() + 1 + ufill[k]++;
                rcol[q] = i;
            }
    }

    // ---- 5. Crout terms and levels (natural numbering) -----------------------------------
    ivec tcount(nlu, 0);
    {
        ivec pos(n, -1);
        for (int i = 0; i < n; ++i) {
            for (int q = rptr[i]; q < rptr[i + 1]; ++q) pos[rcol[q]] = q;
            for (int k : Lrow[i]) {
                const int ub = diag_id(k) + 1, ue = rptr[k + 1];
                for (int q = ub; q < ue; ++q) tcount[pos[rcol[q]]]++;
            }
            for (int q = rptr[i]; q < rptr[i + 1]; ++q) pos[rcol[q]] = -1;
        }
    }
    std::vector<long> tptr(nlu + 1, 0);
    for (long e = 0; e < nlu; ++e) tptr[e + 1] = tptr[e] + tcount[e];
    nterms = tptr[nlu];
    if (nterms > 2000000000L) { delete P; return cdn_fail(ctx, CDN_ERR_STATE, "plan: too many LU terms"); }
    ivec tl(nterms), tu(nterms), lev(nlu, 0);
    {
        ivec pos(n, -1);
        std::vector<long> fillp(tptr.begin(), tptr.end() - 1);
        for (int i = 0; i < n; ++i) {
            for (int q = rptr[i]; q < rptr[i + 1]; ++q) pos[rcol[q]] = q;
            int li = rptr[i];
            for (int k : Lrow[i]) {
                const int lid = li++;      // natural id of L(i,k)
                const int ub = diag_id(k) + 1, ue = rptr[k + 1];
                for (int q = ub; q < ue; ++q) {
                    const int e = pos[rcol[q]];
                    tl[fillp[e]] = lid;
                    tu[fillp[e]] = q;
                    fillp[e]++;
                }
            }
            // levels of this row's entries, left to right
            for (int q = rptr[i]; q < rptr[i + 1]; ++q) {
                int l = 0;
                for (long t = tptr[q]; t < tptr[q + 1]; ++t) {
                    l = std::max(l, lev[tl[t]] + 1);
                    l = std::max(l, lev[tu[t]] + 1);
                }
                const int j = rcol[q];
                if (j < i) l = std::max(l, lev[diag_id(j)] + 1);
                lev[q] = l;
            }
            for (int q = rptr[i]; q < rptr[i + 1]; ++q) pos[rcol[q]] = -1;
        }
    }
    for (long e = 0; e < nlu; ++e) nlev_f = std::max(nlev_f, lev[e] + 1);
    // renumber by level (stable)
    f_lvl_ptr.assign(nlev_f + 1, 0);
    ivec slot(nlu);
    for (long e = 0; e < nlu; ++e) f_lvl_ptr[lev[e] + 1]++;
    for (int l = 0; l < nlev_f; ++l) f_lvl_ptr[l + 1] += f_lvl_ptr[l];
    f_lvl_nheavy.assign(nlev_f, 0);
    {
        // inside a level the entries with long term lists come first: the
        // kernels reduce those cooperatively (hub rows/columns)
        ivec cur(f_lvl_ptr.begin(), f_lvl_ptr.end() - 1);
        for (long e = 0; e < nlu; ++e)
            if (tcount[e] > HEAVY) { slot[e] = cur[lev[e]]++; f_lvl_nheavy[lev[e]]++; }
        for (long e = 0; e < nlu; ++e)
            if (tcount[e] <= HEAVY) slot[e] = cur[lev[e]]++;
    }
    e_term_ptr.assign(nlu + 1, 0); e_term_l.assign(nterms, 0); e_term_u.assign(nterms, 0); e_piv.assign(nlu, 0);
    for (long e = 0; e < nlu; ++e) { e_term_ptr[slot[e] + 1] = tcount[e]; P->max_terms = std::max(P->max_terms, tcount[e]); }
    for (long s = 0; s < nlu; ++s) e_term_ptr[s + 1] += e_term_ptr[s];
    for (int i = 0; i < n; ++i)
        for (int q = rptr[i]; q < rptr[i + 1]; ++q) {
            const int s = slot[q];
            long w = e_term_ptr[s];
            for (long t = tptr[q]; t < tptr[q + 1]; ++t, ++w) {
                e_term_l[w] = slot[tl[t]];
                e_term_u[w] = slot[tu[t]];
            }
            const int j = rcol[q];
            e_piv[s] = (j < i) ? slot[diag_id(j)] : (j == i ? -2 : -1);
        }
    ivec().swap(tl); ivec().swap(tu); std::vector<long>().swap(tptr);
    rslot.assign(nlu, 0);
    for (long e = 0; e < nlu; ++e) rslot[e] = slot[e];
    }  // !dense

    P->slot_row.assign(nlu, 0); P->slot_col.assign(nlu, 0);
    {
        ivec rowmap0(n), colmap0(n);
        for (int r = 0; r < n; ++r) rowmap0[ip[mrow[r]]] = r;
        for (int col = 0; col < n; ++col) colmap0[ip[col]] = col;
        for (int i = 0; i < n; ++i)
            for (int q = rptr[i]; q < rptr[i + 1]; ++q) {
                P->slot_row[rslot[q]] = rowmap0[i];
                P->slot_col[rslot[q]] = colmap0[rcol[q]];
            }
    }
    // ---- 7. linear parts per slot -----------------------------------------------------------
    std::vector<double> e_g(nlu, 0.0), e_c(nlu, 0.0), e_gones(nlu, 0.0);
    auto slot_of = [&](int r, int col) { return find_slot(rptr, rcol, rslot, ip[mrow[r]], ip[col]); };
    for (long k = 0; k < c->nG; ++k)
        if (c->g_row[k] >= 0 && c->g_col[k] >= 0) e_g[slot_of(c->g_row[k], c->g_col[k])] += c->g_val[k];
    for (long k = 0; k < c->nC; ++k)
        if (c->c_row[k] >= 0 && c->c_col[k] >= 0) e_c[slot_of(c->c_row[k], c->c_col[k])] += c->c_val[k];
    for (long k = 0; k < c->nE; ++k)
        if (c->e_row[k] >= 0 && c->e_col[k] >= 0) e_gones[slot_of(c->e_row[k], c->e_col[k])] += c->e_val[k];

    // ---- 6. device gather lists ----------------------------------------------------------------
    long out_size = 0, jac_size = 0;
    std::vector<long> out_off(c->ngroups), jac_off(c->ngroups);
    for (int g = 0; g < c->ngroups; ++g) {
        const cdn_group_host& G = c->groups[g];
        out_off[g] = out_size;
        jac_off[g] = jac_size;
        out_size += (long)(G.ncs + G.nqs) * G.n;
        jac_size += (long)(G.ncs + G.nqs) * G.nxin * G.n;
    }
    if (jac_size > 500000000L || out_size > 500000000L) { delete P; return cdn_fail(ctx, CDN_ERR_STATE, "plan: device Jacobian buffer too large"); }
    struct Con { int dst; int src; signed char kind; };
    std::vector<Con> econ, rcon;
    for (int g = 0; g < c->ngroups; ++g) {
        const cdn_group_host& G = c->groups[g];
        const int nout = G.ncs + G.nqs;
        for (int i = 0; i < G.n; ++i)
            for (int o = 0; o < nout; ++o) {
                const bool chg = o >= G.ncs;
                if (chg && c->skip_q) continue;
                const int rp = chg ? G.qpos[(long)(o - G.ncs) * G.n + i] : G.cpos[(long)o * G.n + i];
                const int rn = chg ? G.qneg[(long)(o - G.ncs) * G.n + i] : G.cneg[(long)o * G.n + i];
                const signed char kp = chg ? 2 : 1, kn = chg ? -2 : -1;
                const int osrc = (int)(out_off[g] + (long)o * G.n + i);
                if (rp >= 0) rcon.push_back(Con{rp, osrc, kp});
                if (rn >= 0) rcon.push_back(Con{rn, osrc, kn});
                for (int k = 0; k < G.nxin; ++k) {
                    const int cp = G.vpos[(long)k * G.n + i], cn = G.vneg[(long)k * G.n + i];
                    const int jsrc = (int)(jac_off[g] + ((long)o * G.nxin + k) * G.n + i);
                    if (rp >= 0 && cp >= 0) econ.push_back(Con{slot_of(rp, cp), jsrc, kp});
                    if (rn >= 0 && cn >= 0) econ.push_back(Con{slot_of(rn, cn), jsrc, kp});
                    if (rp >= 0 && cn >= 0) econ.push_back(Con{slot_of(rp, cn), jsrc, kn});
                    if (rn >= 0 && cp >= 0) econ.push_back(Con{slot_of(rn, cp), jsrc, kn});
                }
            }
    }
    auto to_csr = [](std::vector<Con>& v, long ndst, ivec& ptr, ivec& src, std::vector<signed char>& kind, int* maxlen) {
        ptr.assign(ndst + 1, 0);
        for (const Con& x : v) ptr[x.dst + 1]++;
        int mx = 0;
        for (long d = 0; d < ndst; ++d) { mx = std::max(mx, ptr[d + 1]); ptr[d + 1] += ptr[d]; }
        if (maxlen) *maxlen = mx;
        src.resize(v.size());
        kind.resize(v.size());
        ivec cur(ptr.begin(), ptr.end() - 1);
        for (const Con& x : v) { src[cur[x.dst]] = x.src; kind[cur[x.dst]] = x.kind; cur[x.dst]++; }
    };
    ivec e_con_ptr, e_con_src, r_con_ptr, r_con_src;
    std::vector<signed char> e_con_kind, r_con_kind;
    to_csr(econ, nlu, e_con_ptr, e_con_src, e_con_kind, &P->max_con);
    to_csr(rcon, n, r_con_ptr, r_con_src, r_con_kind, nullptr);
    // packed gather entries: source index in the low 29 bits, bit 29 = minus
    // sign, bit 30 = charge row (scaled by a0)
    auto pack = [](const ivec& src, const std::vector<signed char>& kind) {
        std::vector<unsigned> pk(src.size());
        for (size_t k = 0; k < src.size(); ++k) {
            const int kd = kind[k];
            pk[k] = (unsigned)src[k] | ((kd < 0) ? (1u << 29) : 0u) | ((kd == 2 || kd == -2) ? (1u << 30) : 0u);
        }
        return pk;
    };
    std::vector<unsigned> e_con_pk = pack(e_con_src, e_con_kind), r_con_pk = pack(r_con_src, r_con_kind);
    P->ncon = (long)econ.size();
    std::vector<Con>().swap(econ); std::vector<Con>().swap(rcon);

    // residual: CSR of the linear matrices in original numbering
    ivec r_lin_ptr(n + 1, 0), r_lin_col;
    std::vector<double> r_lin_g, r_lin_c, r_lin_gones;
    {
        struct LE { uint64_t key; double g, c, e; };
        std::vector<LE> le;
        auto addl = [&](int r, int col, double g, double cc, double e) {
            if (r >= 0 && col >= 0) le.push_back(LE{((uint64_t)(uint32_t)r << 32) | (uint32_t)col, g, cc, e});
        };
        for (long k = 0; k < c->nG; ++k) addl(c->g_row[k], c->g_col[k], c->g_val[k], 0, 0);
        for (long k = 0; k < c->nC; ++k) addl(c->c_row[k], c->c_col[k], 0, c->c_val[k], 0);
        for (long k = 0; k < c->nE; ++k) addl(c->e_row[k], c->e_col[k], 0, 0, c->e_val[k]);
        std::stable_sort(le.begin(), le.end(), [](const LE& a, const LE& b) { return a.key < b.key; });
        for (size_t q = 0; q < le.size();) {
            size_t q2 = q;
            double g = 0, cc = 0, e = 0;
            while (q2 < le.size() && le[q2].key == le[q].key) { g += le[q2].g; cc += le[q2].c; e += le[q2].e; ++q2; }
            const int r = (int)(le[q].key >> 32);
            r_lin_ptr[r + 1]++;
            r_lin_col.push_back((int)(uint32_t)le[q].key);
            r_lin_g.push_back(g); r_lin_c.push_back(cc); r_lin_gones.push_back(e);
            q = q2;
        }
        for (int i = 0; i < n; ++i) r_lin_ptr[i + 1] += r_lin_ptr[i];
    }

    // ---- triangular-solve schedules ---------------------------------------------------------------
    ivec l_ptr(n + 1, 0), u_ptr(n + 1, 0), u_diag(n);
    for (int i = 0; i < n; ++i) {
        l_ptr[i + 1] = l_ptr[i] + (int)Lrow[i].size();
        u_ptr[i + 1] = u_ptr[i] + ucount[i];
    }
    ivec l_col(l_ptr[n]), l_slot(l_ptr[n]), u_col(u_ptr[n]), u_slot(u_ptr[n]);
    ivec levl(n, 0), levu(n, 0);
    int nlev_l = 0, nlev_u = 0;
    for (int i = 0; i < n; ++i) {
        int q = rptr[i], w = l_ptr[i], l = 0;
        for (size_t t = 0; t < Lrow[i].size(); ++t, ++q, ++w) {
            l_col[w] = rcol[q]; l_slot[w] = rslot[q];
            l = std::max(l, levl[rcol[q]] + 1);
        }
        levl[i] = l;
        nlev_l = std::max(nlev_l, l + 1);
        u_diag[i] = rslot[q];
    }
    for (int i = n - 1; i >= 0; --i) {
        int q = diag_id(i) + 1, w = u_ptr[i], l = 0;
        for (; !dense && q < rptr[i + 1]; ++q, ++w) {
            u_col[w] = rcol[q]; u_slot[w] = rslot[q];
            l = std::max(l, levu[rcol[q]] + 1);
        }
        levu[i] = l;
        nlev_u = std::max(nlev_u, l + 1);
    }
    auto by_level = [&](const ivec& lv, int nl, const ivec& rp, ivec& lptr, ivec& rows,
                        ivec& nheavy) {
        lptr.assign(nl + 1, 0);
        nheavy.assign(nl, 0);
        for (int i = 0; i < n; ++i) lptr[lv[i] + 1]++;
        for (int l = 0; l < nl; ++l) lptr[l + 1] += lptr[l];
        rows.resize(n);
        ivec cur(lptr.begin(), lptr.end() - 1);
        for (int i = 0; i < n; ++i)
            if (rp[i + 1] - rp[i] > HEAVY) { rows[cur[lv[i]]++] = i; nheavy[lv[i]]++; }
        for (int i = 0; i < n; ++i)
            if (rp[i + 1] - rp[i] <= HEAVY) rows[cur[lv[i]]++] = i;
    };
    ivec l_lvl_ptr, l_rows, u_lvl_ptr, u_rows, l_lvl_nheavy, u_lvl_nheavy;
    by_level(levl, nlev_l, l_ptr, l_lvl_ptr, l_rows, l_lvl_nheavy);
    by_level(levu, nlev_u, u_ptr, u_lvl_ptr, u_rows, u_lvl_nheavy);
    // slots / rows whose gather lists are long enough for a cooperative sum
    ivec e_heavy, r_heavy;
    for (long e = 0; e < nlu; ++e)
        if (e_con_ptr[e + 1] - e_con_ptr[e] > HEAVY) e_heavy.push_back((int)e);
    for (int r = 0; r < n; ++r)
        if ((r_con_ptr[r + 1] - r_con_ptr[r]) + (r_lin_ptr[r + 1] - r_lin_ptr[r]) > HEAVY)
            r_heavy.push_back(r);
    // permuted row i <- original row rowmap[i]; permuted col j -> unknown colmap[j]
    ivec rowmap(n), colmap(n);
    for (int r = 0; r < n; ++r) rowmap[ip[mrow[r]]] = r;
    for (int col = 0; col < n; ++col) colmap[ip[col]] = col;

    // compact programs (see plan.h)
    // Lanes that share one row (entry) of a level: a level costs what its longest
    // row costs, one term after the other (two dependent shared-memory loads
    // and a multiply-add, ~70 cycles); 2^lg lanes take every 2^lg-th term and
    // add up with lg shuffles.  Picked per level for a 512-thread CTA from a
    // cycle estimate: passes over the level x (fixed + longest share + shuffles).
    auto group_log2 = [](int nitems, int maxcnt) {
        int best = 0;
        long best_cost = -1;
        for (int lg = 0; lg <= 5; ++lg) {
            const int g = 1 << lg;
            if (lg > 0 && g > 2 * maxcnt) break;
            const long passes = ((long)nitems * g + 511) / 512;
            const long cost = std::max(1L, passes) * (150 + 70L * ((maxcnt + g - 1) / g) + 30L * lg);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = lg; }
        }
        return best;
    };
    std::vector<unsigned> cl_prog, cu_prog, cf_prog;
    int cl_off[5] = {0, 0, 0, 0, 0}, cu_off[5] = {0, 0, 0, 0, 0}, cf_off[3] = {0, 0, 0};
    int naff = 0;
    bool compact = !dense && n < 65535 && nlu < 65535;
    for (int l = 0; compact && l < nlev_l; ++l) if (l_lvl_nheavy[l]) compact = false;
    for (int l = 0; compact && l < nlev_u; ++l) if (u_lvl_nheavy[l]) compact = false;
    for (int l = 0; compact && l < nlev_f; ++l) if (f_lvl_nheavy[l]) compact = false;
    if (compact) {
        auto pad4 = [](std::vector<unsigned>& v) { while (v.size() % 4) v.push_back(0u); };
        auto push16 = [&](std::vector<unsigned>& v, const ivec& a) {
            for (size_t k = 0; k < a.size(); k += 2)
                v.push_back((unsigned)a[k] | ((k + 1 < a.size() ? (unsigned)a[k + 1] : 0u) << 16));
            pad4(v);
        };
        // L and U sweeps
        for (int pass = 0; pass < 2; ++pass) {
            std::vector<unsigned>& pr = pass ? cu_prog : cl_prog;
            int* off = pass ? cu_off : cl_off;
            const ivec& lvl = pass ? u_lvl_ptr : l_lvl_ptr;
            const ivec& rows = pass ? u_rows : l_rows;
            const ivec& ptr = pass ? u_ptr : l_ptr;
            off[0] = 0;
            for (int v : lvl) pr.push_back((unsigned)v);
            // lanes per row, level by level (log2)
            for (size_t l = 0; l + 1 < lvl.size(); ++l) {
                int maxcnt = 0;
                for (int q = lvl[l]; q < lvl[l + 1]; ++q) {
                    const int i = rows[q];
                    maxcnt = std::max(maxcnt, ptr[i + 1] - ptr[i]);
                }
                pr.push_back((unsigned)group_log2(lvl[l + 1] - lvl[l], maxcnt));
            }
            pad4(pr);
            off[1] = (int)pr.size();
            for (int q = 0; q < n; ++q) {
                const int i = rows[q];
                pr.push_back((unsigned)i | ((unsigned)(ptr[i + 1] - ptr[i]) << 16));
                pr.push_back((unsigned)ptr[i] | (pass ? ((unsigned)u_diag[i] << 16) : 0u));
            }
            pad4(pr);
            off[2] = (int)pr.size();
            push16(pr, pass ? u_col : l_col);
            off[3] = (int)pr.size();
            push16(pr, pass ? u_slot : l_slot);
            off[4] = (int)pr.size();
            push16(pr, pass ? colmap : rowmap);
        }
        // partial refactorisation: closure of the slots with device contributions
        std::vector<char> aff(nlu, 0);
        for (long e = 0; e < nlu; ++e) {
            bool f = e_con_ptr[e + 1] > e_con_ptr[e];
            for (int t = e_term_ptr[e]; !f && t < e_term_ptr[e + 1]; ++t)
                f = aff[e_term_l[t]] || aff[e_term_u[t]];
            if (!f && e_piv[e] >= 0) f = aff[e_piv[e]];
            aff[e] = f ? 1 : 0;
        }
        ivec flvl(nlev_f + 1, 0), fgrp(nlev_f, 0);
        std::vector<unsigned> frec, fterm;
        for (int l = 0; l < nlev_f; ++l) {
            int maxcnt = 0, cnt_l = 0;
            for (int e = f_lvl_ptr[l]; e < f_lvl_ptr[l + 1]; ++e)
                if (aff[e]) {
                    maxcnt = std::max(maxcnt, e_term_ptr[e + 1] - e_term_ptr[e]);
                    ++cnt_l;
                }
            fgrp[l] = group_log2(cnt_l, maxcnt);
            for (int e = f_lvl_ptr[l]; e < f_lvl_ptr[l + 1]; ++e) {
                if (!aff[e]) continue;
                const int cnt = e_term_ptr[e + 1] - e_term_ptr[e];
                frec.push_back((unsigned)e | ((unsigned)cnt << 16));
                frec.push_back((unsigned)fterm.size() |
                               ((unsigned)(e_piv[e] >= 0 ? e_piv[e] : 0xffff) << 16));
                for (int t = e_term_ptr[e]; t < e_term_ptr[e + 1]; ++t)
                    fterm.push_back((unsigned)e_term_l[t] | ((unsigned)e_term_u[t] << 16));
                ++naff;
            }
            flvl[l + 1] = naff;
        }
        if (fterm.size() >= 65535) {
            compact = false;             // first-term offsets are 16 bit
        } else {
            for (int v : flvl) cf_prog.push_back((unsigned)v);
            for (int v : fgrp) cf_prog.push_back((unsigned)v);
            pad4(cf_prog);
            cf_off[1] = (int)cf_prog.size();
            cf_prog.insert(cf_prog.end(), frec.begin(), frec.end());
            pad4(cf_prog);
            cf_off[2] = (int)cf_prog.size();
            cf_prog.insert(cf_prog.end(), fterm.begin(), fterm.end());
            pad4(cf_prog);
        }
    }
    if (!compact) { cl_prog.clear(); cu_prog.clear(); cf_prog.clear(); naff = 0; }
    // ---- pack ---------------------------------------------------------------------------------------
    std::vector<char>& B = P->staging;
    cdn_plan::Off& O = P->off;
#define PUT(name) do { O.name = put(B, name); P->reg[#name] = std::make_pair(O.name, (long)name.size()); } while (0)
    PUT(f_lvl_ptr);
    PUT(f_lvl_nheavy); PUT(l_lvl_nheavy); PUT(u_lvl_nheavy); PUT(e_heavy); PUT(r_heavy);
    PUT(e_term_ptr);
    PUT(e_term_l);
    PUT(e_term_u);
    PUT(e_piv);
    PUT(e_g);
    PUT(e_c);
    PUT(e_gones);
    PUT(e_con_ptr);
    PUT(e_con_src);
    PUT(e_con_kind);
    PUT(l_lvl_ptr);
    PUT(l_rows);
    PUT(l_ptr);
    PUT(l_col);
    PUT(l_slot);
    PUT(u_lvl_ptr);
    PUT(u_rows);
    PUT(u_ptr);
    PUT(u_col);
    PUT(u_slot);
    PUT(u_diag);
    PUT(rowmap);
    PUT(colmap);
    PUT(r_lin_ptr);
    PUT(r_lin_col);
    PUT(r_lin_g);
    PUT(r_lin_c);
    PUT(r_lin_gones);
    PUT(r_con_ptr);
    PUT(r_con_src);
    PUT(r_con_kind);
    PUT(e_con_pk); PUT(r_con_pk);
    PUT(cl_prog); PUT(cu_prog); PUT(cf_prog);
    ivec e_bal(nlu);
    for (long e = 0; e < nlu; ++e) e_bal[e] = (int)e;
    if (dense)
        std::stable_sort(e_bal.begin(), e_bal.end(), [&](int a, int b) {
            return (e_con_ptr[a + 1] - e_con_ptr[a]) > (e_con_ptr[b + 1] - e_con_ptr[b]);
        });
    PUT(e_bal);
    // small dense systems: flat per-lane programs of the device-Jacobian
    // contributions (plan.h); slot e = row * n + unknown for dense storage
    ivec sd_ptr(9, 0);
    std::vector<unsigned> sd_prog;
    int sd_maxlen = 0, sd_rmax = 0;
    bool sd = dense && n <= 8 && jac_size < (1L << 20) && e_heavy.empty() && r_heavy.empty();
    if (sd)
        for (long e = 0; e < nlu; ++e)
            if (P->slot_row[e] * n + P->slot_col[e] != e) sd = false;
    ivec sd_fix;
    if (c->nhist > 0) sd = false;                    // (delayed ports: general path)
    if (sd) {
        // positions (doubles) inside one instance's state, engine.cuh ws_carve:
        // 10 n-vectors, lu, device outputs, device Jacobians, factor storage
        const long lu_off = 10L * n, jac_off = lu_off + nlu + out_size;
        const long fact_off = jac_off + jac_size;
        const long zero_at = fact_off + CDN_SD8_PART + CDN_SD8_NFIX;
        if (zero_at >= (1L << 14)) sd = false;
        // all terms, slot after slot (longest lists first), cut into 8 pieces
        struct T { int e; int src; bool minus; };
        std::vector<T> all;
        for (long q = 0; q < nlu; ++q) {
            const int e = e_bal[q];
            for (int k = e_con_ptr[e]; k < e_con_ptr[e + 1]; ++k)
                all.push_back(T{e, e_con_src[k], e_con_kind[k] < 0});
        }
        const long total = (long)all.size();
        const long chunk = (total + 7) / 8;
        sd_maxlen = (int)chunk;
        for (int l = 0; l < 8 && sd; ++l) {
            const long t0 = std::min(total, l * chunk), t1 = std::min(total, (l + 1) * chunk);
            for (long t = t0; t < t1; ++t) {
                const int e = all[t].e;
                // the slot's first term lies in an earlier piece: partial accumulator
                const bool cont = t0 > 0 && all[t0 - 1].e == e;
                long tgt = lu_off + e;
                if (cont) {
                    if (t == t0) sd_fix.push_back(e);
                    tgt = fact_off + CDN_SD8_PART + (long)sd_fix.size() - 1;
                }
                unsigned w = (unsigned)(jac_off + all[t].src) | ((unsigned)tgt << 14);
                if (all[t].minus) w |= 1u << 31;
                if (t + 1 == t1 || all[t + 1].e != e) w |= 1u << 30;
                sd_prog.push_back(w);
            }
            for (long t = t1 - t0; t < chunk; ++t) sd_prog.push_back((unsigned)zero_at);
            sd_ptr[l + 1] = (int)sd_prog.size();
        }
        if (!sd || (long)sd_fix.size() > CDN_SD8_NFIX) {
            sd = false; sd_prog.clear(); sd_fix.clear(); sd_maxlen = 0;
            for (int l = 0; l < 9; ++l) sd_ptr[l] = 0;
        }
        for (int r = 0; r < n; ++r) sd_rmax = std::max(sd_rmax, r_con_ptr[r + 1] - r_con_ptr[r]);
    }
    PUT(sd_ptr); PUT(sd_prog); PUT(sd_fix);
    {
        ivec dg, di;
        for (int g = 0; g < c->ngroups; ++g)
            for (int i = 0; i < c->groups[g].n; ++i) { dg.push_back(g); di.push_back(i); }
        ivec& dev_g = dg; ivec& dev_i = di;
        PUT(dev_g); PUT(dev_i);
        P->dev.ndev = (int)dg.size();
    }
    cdn_plan_dev& D = P->dev;
    D.dense = dense ? 1 : 0;
    D.sd = sd ? 1 : 0; D.sd_maxlen = sd_maxlen; D.sd_rmax = sd_rmax; D.sd_nterms = (int)sd_prog.size(); D.sd_nfix = (int)sd_fix.size(); D.sd_pad_ = 0;
    D.compact = compact ? 1 : 0; D.nnzl = (int)l_col.size(); D.nnzu = (int)u_col.size();
    D.cl_words = (int)cl_prog.size(); D.cu_words = (int)cu_prog.size();
    D.cf_words = (int)cf_prog.size(); D.naff = naff;
    for (int k = 0; k < 5; ++k) { D.cl_off[k] = cl_off[k]; D.cu_off[k] = cu_off[k]; }
    for (int k = 0; k < 3; ++k) D.cf_off[k] = cf_off[k];
    D.nhist = c->nhist > 0 ? c->nhist : 0;
    D.hist_size = 0;
    D.heavy = HEAVY;
    D.n_e_heavy = (int)e_heavy.size(); D.n_r_heavy = (int)r_heavy.size();
    D.max_terms = P->max_terms;
    D.ncon = (int)e_con_pk.size(); D.nrcon = (int)r_con_pk.size(); D.nrlin = (int)r_lin_col.size();
    D.n = n; D.nlu = (int)nlu; D.nlev_f = nlev_f; D.nlev_l = nlev_l; D.nlev_u = nlev_u;
    D.ngroups = c->ngroups;
    D.dev_out_size = out_size; D.dev_jac_size = jac_size;
    for (int g = 0; g < c->ngroups; ++g) {
        const cdn_group_host& G = c->groups[g];
        cdn_group_dev& d = D.groups[g];
        d.model = G.model; d.flags = G.flags; d.n = G.n; d.nxin = G.nxin; d.ncs = G.ncs;
        d.nqs = G.nqs; d.ndelay = G.ndelay;
        d.out_off = out_off[g]; d.jac_off = jac_off[g];
        ivec vp(G.vpos, G.vpos + (long)G.nxin * G.n), vn(G.vneg, G.vneg + (long)G.nxin * G.n);
        O.gv[g][0] = put(B, vp); O.gv[g][1] = put(B, vn);
        O.gd[g] = -1;
        d.hist_off = 0;
        if (G.ndelay > 0 && G.delay) {
            std::vector<double> dl(G.delay, G.delay + (long)G.ndelay * G.n);
            O.gd[g] = put(B, dl);
            if (c->nhist > 0) {
                d.hist_off = D.hist_size;
                D.hist_size += (long)G.ndelay * G.n * c->nhist;
            }
        }
    }
    P->nterms = nterms;
    *out = P;
    return CDN_OK;
}

extern "C" long cdn_plan_bytes(cdn_plan* P) {
    return P ? (long)P->staging.size() + 512 + (long)sizeof(cdn_plan_dev) + 256 : 0;
}

// info[0..9] = n, nlu, nlev_f, nlev_l, nlev_u, nterms, ncon, nnzA, max_terms, max_con
extern "C" int cdn_plan_info(cdn_plan* P, long* info) {
    if (!P || !info) return CDN_ERR_ARG;
    info[0] = P->dev.n; info[1] = P->dev.nlu; info[2] = P->dev.nlev_f; info[3] = P->dev.nlev_l;
    info[4] = P->dev.nlev_u; info[5] = P->nterms; info[6] = P->ncon; info[7] = P->nnzA;
    info[8] = P->max_terms; info[9] = P->max_con;
    info[10] = P->dev.dev_out_size; info[11] = P->dev.dev_jac_size;
    info[12] = P->last_team;
    info[13] = P->dev.hist_size;
    return CDN_OK;
}

extern "C" int cdn_plan_bind(cdn_plan* P, void* dev_ws, long bytes, void* stream) {
    if (!P || !dev_ws) return CDN_ERR_ARG;
    cdn_ctx* ctx = P->ctx;
    CDN_USE_DEVICE(ctx);
    char* base = (char*)(((uintptr_t)dev_ws + 255) / 256 * 256);
    const long struct_off = ((long)P->staging.size() + 255) / 256 * 256;
    if ((long)(base - (char*)dev_ws) + struct_off + (long)sizeof(cdn_plan_dev) > bytes)
        return cdn_fail(ctx, CDN_ERR_ARG, "plan_bind: workspace too small");
    CDN_CUDA(ctx, cudaMemcpyAsync(base, P->staging.data(), P->staging.size(),
                                  cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CDN_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    cdn_plan_dev& D = P->dev;
    const cdn_plan::Off& O = P->off;
#define BIND(f, T) D.f = (const T*)(base + O.f)
    BIND(f_lvl_nheavy, int); BIND(l_lvl_nheavy, int); BIND(u_lvl_nheavy, int);
    BIND(e_heavy, int); BIND(r_heavy, int);
    BIND(f_lvl_ptr, int); BIND(e_term_ptr, int); BIND(e_term_l, int); BIND(e_term_u, int);
    BIND(e_piv, int); BIND(e_g, double); BIND(e_c, double); BIND(e_gones, double);
    BIND(e_con_ptr, int); BIND(e_con_src, int); BIND(e_con_kind, signed char);
    BIND(l_lvl_ptr, int); BIND(l_rows, int); BIND(l_ptr, int); BIND(l_col, int); BIND(l_slot, int);
    BIND(u_lvl_ptr, int); BIND(u_rows, int); BIND(u_ptr, int); BIND(u_col, int); BIND(u_slot, int);
    BIND(u_diag, int); BIND(rowmap, int); BIND(colmap, int);
    BIND(r_lin_ptr, int); BIND(r_lin_col, int); BIND(r_lin_g, double); BIND(r_lin_c, double);
    BIND(r_lin_gones, double); BIND(r_con_ptr, int); BIND(r_con_src, int);
    BIND(r_con_kind, signed char);
    BIND(dev_g, int); BIND(dev_i, int);
    BIND(e_con_pk, unsigned); BIND(r_con_pk, unsigned); BIND(e_bal, int);
    BIND(cl_prog, unsigned); BIND(cu_prog, unsigned); BIND(cf_prog, unsigned);
    BIND(sd_ptr, int); BIND(sd_prog, unsigned); BIND(sd_fix, int);
#undef BIND
    for (int g = 0; g < D.ngroups; ++g) {
        D.groups[g].vpos = (const int*)(base + O.gv[g][0]);
        D.groups[g].vneg = (const int*)(base + O.gv[g][1]);
        D.groups[g].delay = O.gd[g] >= 0 ? (const double*)(base + O.gd[g]) : nullptr;
    }
    P->dev_struct = (cdn_plan_dev*)(base + struct_off);
    P->bound = true;
    P->dirty = true;
    return CDN_OK;
}

extern "C" int cdn_plan_set_group_params(cdn_plan* P, int g, const double* params, long ldp,
                                         const int* pidx, int per_inst) {
    if (!P || g < 0 || g >= P->dev.ngroups) return CDN_ERR_ARG;
    P->dev.groups[g].params = params;
    P->dev.groups[g].ldp = ldp;
    P->dev.groups[g].pidx = pidx;
    P->dev.groups[g].per_inst = per_inst;
    P->dirty = true;
    return CDN_OK;
}

extern "C" int cdn_plan_slot_coords(cdn_plan* P, int* rows, int* cols) {
    if (!P || !rows || !cols) return CDN_ERR_ARG;
    memcpy(rows, P->slot_row.data(), P->slot_row.size() * sizeof(int));
    memcpy(cols, P->slot_col.data(), P->slot_col.size() * sizeof(int));
    return CDN_OK;
}

// host copy of one packed index/value array by field name (tests, tooling)
extern "C" int cdn_plan_array(cdn_plan* P, const char* name, const void** ptr, long* count) {
    if (!P || !name || !ptr || !count) return CDN_ERR_ARG;
    auto it = P->reg.find(name);
    if (it == P->reg.end()) return CDN_ERR_ARG;
    *ptr = P->staging.data() + it->second.first;
    *count = it->second.second;
    return CDN_OK;
}

// update the linear values of the plan in place (parameter sweeps): not needed yet
extern "C" int cdn_plan_free(cdn_plan* P) {
    delete P;
    return CDN_OK;
}

const cdn_plan_dev* cdn_plan_host_view(cdn_plan* P) { return P && P->bound ? &P->dev : nullptr; }

int cdn_plan_sync_device(cdn_plan* P, cudaStream_t st, const cdn_plan_dev** out) {
    if (!P || !P->bound) return CDN_ERR_STATE;
    if (P->dirty) {
        // pageable source: the copy is staged before the call returns
        CDN_CUDA(P->ctx, cudaMemcpyAsync(P->dev_struct, &P->dev, sizeof(cdn_plan_dev),
                                         cudaMemcpyHostToDevice, st));
        P->dirty = false;
    }
    *out = P->dev_struct;
    return CDN_OK;
}
cdn_ctx* cdn_plan_ctx(cdn_plan* P) { return P->ctx; }
void cdn_plan_note_team(cdn_plan* P, int team) { P->last_team = team; }
