// pp_profile_plan.cpp -- see pp_profile.h.  Pure host C++ (no CUDA).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "pp_profile.h"

namespace pp {

namespace {

struct Fail {
    ProfilePlan* plan;
    bool operator()(const std::string& why) const {
        plan->ok = false;
        plan->why = why;
        return false;
    }
};

void emission_records(const EmisRec& E, double* v, double* f) {
    // Viterbi: exactly the constants smem_emission_log reads; forward: linear-space twins (pp_lowering.cpp fill_emissions)
    v[0] = E.a0; v[1] = E.a1; v[2] = E.a2; v[3] = E.a3;
    f[0] = E.a0; f[1] = 0.0; f[2] = 0.0; f[3] = 0.0;
    if (E.kind == EMIS_NORMAL) {
        f[1] = E.k2;
        f[2] = E.c2;
    } else if (E.kind == EMIS_POINT) {
        f[3] = 1.0;                                              // state weight 1 (log weight 0 is a precondition)
    } else {
        f[1] = E.a1;
        f[3] = std::exp(E.a3);                                   // 1 / (hi - lo); 1 when lo == hi (a3 = 0)
    }
}

}  // namespace

void build_profile_plan(const pp_model_desc& d, const Lowered& low, int P, int max_warps, ProfilePlan* plan) {
    *plan = ProfilePlan();
    plan->P = P;
    const Fail fail{plan};
    const int m = d.n_states, ne = d.silent_start;
    const int start = d.start_index, end = d.end_index;
    if (!low.lane_ok) { fail("model is not lane-lowered"); return; }
    if (d.finite != 1 || !low.end_final_only) { fail("needs a finite model whose end state has no out-edges"); return; }
    if (low.state_n_edges[start] != 0) { fail("start state has in-edges"); return; }
    for (int k = 0; k < ne; ++k) {
        if (low.emis[k].logw != 0.0) { fail("a state weight differs from 1"); return; }
        if (low.emis[k].kind != EMIS_NORMAL && low.emis[k].kind != EMIS_POINT && low.emis[k].kind != EMIS_UNIFORM) {
            fail("emission family without a device lowering");
            return;
        }
    }
    auto out_begin = [&](int s) { return d.out_ptr[s]; };
    auto out_end = [&](int s) { return d.out_ptr[s + 1]; };
    auto has_out = [&](int s, int t) {
        for (int e = out_begin(s); e < out_end(s); ++e) if (d.out_dst[e] == t) return true;
        return false;
    };

    // ---- classification ---------------------------------------------------------------------------------------
    std::vector<char> is_m(m, 0);
    int n_m = 0;
    for (int s = 0; s < ne; ++s) if (has_out(s, end)) { is_m[s] = 1; ++n_m; }
    if (n_m < 2) { fail("fewer than two match states"); return; }
    int m0 = -1;
    for (int e = out_begin(start); e < out_end(start); ++e) {
        const int t = d.out_dst[e];
        if (t < ne && is_m[t]) { if (m0 >= 0) { fail("start feeds two match states"); return; } m0 = t; }
    }
    if (m0 < 0) { fail("start feeds no match state"); return; }
    std::vector<int> pos(m, -1), cls(m, -1);     // cls: 0 M, 1 BLIP, 2 INS, 3 DROP, 4 SKIP, 5 BACK
    std::vector<int32_t>&M = plan->M, &BLIP = plan->BLIP, &INS = plan->INS, &DROP = plan->DROP, &SKIP = plan->SKIP, &BACK = plan->BACK;
    for (int cur = m0, prev = -1; cur >= 0;) {
        if (pos[cur] >= 0) { fail("match states form a cycle"); return; }
        pos[cur] = static_cast<int>(M.size());
        cls[cur] = 0;
        M.push_back(cur);
        int nxt = -1;
        for (int e = out_begin(cur); e < out_end(cur); ++e) {
            const int t = d.out_dst[e];
            if (t >= ne || !is_m[t] || t == cur || t == prev) continue;
            if (nxt >= 0) { fail("a match state feeds two later match states"); return; }
            nxt = t;
        }
        prev = cur;
        cur = nxt;
    }
    const int K = static_cast<int>(M.size());
    if (K != n_m) { fail("match states are not one chain"); return; }
    BLIP.assign(K, -1); INS.assign(K, -1); DROP.assign(K, -1); SKIP.assign(K, -1); BACK.assign(K, -1);
    auto lowered_src = [&](int s, int j) { return low.low_src[low.state_edge_begin[s] + j]; };
    for (int u = 0; u < ne; ++u) {
        if (is_m[u]) continue;
        if (low.state_n_edges[u] != 2) { fail("an insert / blip state does not have two in-edges"); return; }
        int a = -1;
        for (int j = 0; j < 2; ++j) { const int s = lowered_src(u, j); if (s != u) a = s; }
        if (a < 0 || a >= ne || !is_m[a] || !has_out(u, u)) { fail("an emitting state is neither match, blip nor insert"); return; }
        int b = -1, n_out = 0;
        for (int e = out_begin(u); e < out_end(u); ++e) { ++n_out; if (d.out_dst[e] != u) b = d.out_dst[e]; }
        if (n_out != 2 || b < 0 || b >= ne || !is_m[b]) { fail("an insert / blip state has unexpected out-edges"); return; }
        if (b == a) {
            if (BLIP[pos[a]] >= 0) { fail("two blip states at one position"); return; }
            BLIP[pos[a]] = u; pos[u] = pos[a]; cls[u] = 1;
        } else if (pos[b] == pos[a] + 1) {
            if (INS[pos[b]] >= 0) { fail("two insert states at one position"); return; }
            INS[pos[b]] = u; pos[u] = pos[b]; cls[u] = 2;
        } else {
            fail("an insert state does not sit between neighbouring positions");
            return;
        }
    }
    {   // the skip chain hangs off the start state
        int cur = -1;
        for (int e = out_begin(start); e < out_end(start); ++e) {
            const int t = d.out_dst[e];
            if (t >= ne && t != end) { if (cur >= 0) { fail("start feeds two silent states"); return; } cur = t; }
        }
        for (int k = 0; cur >= 0; ++k) {
            if (k >= K || cls[cur] >= 0) { fail("skip chain longer than the profile"); return; }
            SKIP[k] = cur; pos[cur] = k; cls[cur] = 4;
            int nxt = -1;
            for (int e = out_begin(cur); e < out_end(cur); ++e) {
                const int t = d.out_dst[e];
                if (t < ne || t == end) continue;
                if (nxt >= 0) { fail("a skip state feeds two silent states"); return; }
                nxt = t;
            }
            cur = nxt;
        }
    }
    for (int s = ne; s < m; ++s) {
        if (s == start || s == end || cls[s] >= 0) continue;
        int src_m = -1, n_emit = 0;
        for (int j = 0; j < low.state_n_edges[s]; ++j) {
            const int t = lowered_src(s, j);
            if (t < ne) { ++n_emit; src_m = t; }
        }
        if (n_emit != 1 || !is_m[src_m]) { fail("a silent state is neither drop-off, skip nor back-slip"); return; }
        const int k = pos[src_m];
        bool drop = true;                                          // drop-off: leads back to its own match state and / or to end
        for (int e = out_begin(s); e < out_end(s); ++e) if (d.out_dst[e] != src_m && d.out_dst[e] != end) drop = false;
        std::vector<int32_t>& slot = drop ? DROP : BACK;
        if (slot[k] >= 0) { fail("two drop-off / back-slip states at one position"); return; }
        slot[k] = s; pos[s] = k; cls[s] = drop ? 3 : 5;
    }

    // the kernels evaluate Normal match states, Uniform blips and Normal / point-mass inserts (what HMM_linear_model builds,
    // before and after Baum-Welch); anything else keeps the run-based kernels
    for (int k = 0; k < K; ++k) {
        if (low.emis[M[k]].kind != EMIS_NORMAL) { fail("a match state is not a Normal with sigma > 0"); return; }
        if (BLIP[k] >= 0 && low.emis[BLIP[k]].kind != EMIS_UNIFORM) { fail("a blip state is not Uniform"); return; }
        if (INS[k] >= 0 && low.emis[INS[k]].kind == EMIS_UNIFORM) { fail("an insert state is Uniform"); return; }
    }

    // ---- geometry ---------------------------------------------------------------------------------------------
    const int NW = (K + P - 1) / P;
    if (NW > max_warps) { fail("profile has more positions than one CTA's warps can hold in registers"); return; }
    if (5 * P > 32) { fail("more than 6 positions per warp"); return; }
    const int Kp = NW * P;
    plan->K = K; plan->NW = NW; plan->Kp = Kp;
    const double ninf = -INFINITY;
    const int Kr = Kp + P;                                          // records: the positions and one phantom block behind them
    plan->Kr = Kr;
    plan->rec_v.assign(static_cast<size_t>(Kr) * kProfRecDoubles, 0.0);
    plan->rec_f.assign(static_cast<size_t>(Kr) * kProfRecDoubles, 0.0);
    const int n_w = 16;                                             // weights live in [0, 16) except the kinds word
    for (int k = 0; k < Kr; ++k) {
        double* v = &plan->rec_v[static_cast<size_t>(k) * kProfRecDoubles];
        for (int j = 0; j < n_w; ++j) if (j != PR_KINDS) v[j] = ninf;
        v[PR_BK] = v[PR_BK + 1] = ninf;
        // padding / absent emitting states: a point mass at 0 (any finite record works: all their in-weights are -inf / 0)
        int32_t kinds = EMIS_POINT | (EMIS_POINT << 8) | (EMIS_POINT << 16);
        std::memcpy(&v[PR_KINDS], &kinds, sizeof kinds);
        double* f = &plan->rec_f[static_cast<size_t>(k) * kProfRecDoubles];
        std::memcpy(&f[PR_KINDS], &kinds, sizeof kinds);
        f[PR_EM + 3] = f[PR_EI + 3] = f[PR_EB + 3] = 1.0;
    }
    auto rec = [&](int k, double** v, double** f) {
        *v = &plan->rec_v[static_cast<size_t>(k) * kProfRecDoubles];
        *f = &plan->rec_f[static_cast<size_t>(k) * kProfRecDoubles];
    };
    auto edge_logp = [&](int s, int j) { return d.in_logp[low.low_csr[low.state_edge_begin[s] + j]]; };
    auto set_w = [&](int k, int field, int s, int j) {
        double *v, *f;
        rec(k, &v, &f);
        v[field] = edge_logp(s, j);
        f[field] = std::exp(v[field]);
    };

    // ---- every in-edge list against the template --------------------------------------------------------------
    plan->slot_begin.assign(m, 0);
    plan->slot_src.clear();
    auto begin_slots = [&](int s, int n) {
        plan->slot_begin[s] = static_cast<int32_t>(plan->slot_src.size());
        plan->slot_src.insert(plan->slot_src.end(), n, -1);
        return &plan->slot_src[plan->slot_begin[s]];
    };
    // matches the lowered in-list of `s` against `want[0 .. n)` (expected source per slot, -1 = no such slot here;
    // -2 = any back-slip state), slots strictly ascending; returns false on the first edge the template cannot place
    std::vector<int> placed;
    auto match = [&](int s, const int* want, int n) -> bool {
        placed.assign(low.state_n_edges[s], -1);
        int next_slot = 0;
        for (int j = 0; j < low.state_n_edges[s]; ++j) {
            const int src = lowered_src(s, j);
            int slot = -1;
            for (int t = next_slot; t < n; ++t)
                if (want[t] == src || (want[t] == -2 && src >= ne && cls[src] == 5)) { slot = t; break; }
            if (slot < 0) return false;
            placed[j] = slot;
            next_slot = slot + 1;
        }
        return true;
    };
    int kinds_word[96];
    for (int k = 0; k < Kr; ++k) kinds_word[k] = EMIS_POINT | (EMIS_POINT << 8) | (EMIS_POINT << 16);
    for (int k = 0; k < K; ++k) {
        double *v, *f;
        rec(k, &v, &f);
        {   // M_k
            const int s = M[k];
            const int want[8] = {k > 0 ? M[k - 1] : -1, k > 0 ? SKIP[k - 1] : start, s, DROP[k], BLIP[k], INS[k],
                                 k + 1 < K ? M[k + 1] : -1, -2};
            if (!match(s, want, 8)) { fail("a match state has an in-edge outside the profile template (or out of order)"); return; }
            int32_t* slots = begin_slots(s, 8);
            for (int j = 0; j < low.state_n_edges[s]; ++j) {
                const int slot = placed[j], src = lowered_src(s, j);
                slots[slot == 6 ? 7 : (slot == 7 ? 6 : slot)] = src;   // stored ordinal: 6 = back-slip, 7 = M_{k+1} (kernel)
                set_w(k, PR_WM + slot, s, j);
                if (slot == 6) kinds_word[k] |= 1 << 24;
                if (slot == 7 && pos[src] != k + 1) {               // reads a back-slip state that is not its neighbour's
                    if (k + 1 < K && BACK[k + 1] >= 0) { fail("a match state reads a non-neighbouring back-slip state"); return; }
                    bool known = false;
                    for (int a = 0; a < plan->n_alias; ++a)
                        if (plan->alias_dst[a] == k + 1) { known = true; if (plan->alias_src[a] != pos[src]) { fail("conflicting back-slip aliases"); return; } }
                    if (!known) {
                        if (plan->n_alias == kProfMaxAlias) { fail("too many back-slip aliases"); return; }
                        plan->alias_dst[plan->n_alias] = k + 1;
                        plan->alias_src[plan->n_alias] = pos[src];
                        ++plan->n_alias;
                    }
                }
            }
            emission_records(low.emis[s], v + PR_EM, f + PR_EM);
            kinds_word[k] = (kinds_word[k] & ~0xff) | low.emis[s].kind;
        }
        if (INS[k] >= 0) {
            const int s = INS[k];
            const int want[2] = {k > 0 ? M[k - 1] : -1, s};
            if (!match(s, want, 2)) { fail("an insert state's in-edges do not fit the template"); return; }
            int32_t* slots = begin_slots(s, 2);
            for (int j = 0; j < low.state_n_edges[s]; ++j) { slots[placed[j]] = lowered_src(s, j); set_w(k, PR_WI + placed[j], s, j); }
            emission_records(low.emis[s], v + PR_EI, f + PR_EI);
            kinds_word[k] = (kinds_word[k] & ~0xff00) | (low.emis[s].kind << 8);
        }
        if (BLIP[k] >= 0) {
            const int s = BLIP[k];
            const int want[2] = {M[k], s};
            if (!match(s, want, 2)) { fail("a blip state's in-edges do not fit the template"); return; }
            int32_t* slots = begin_slots(s, 2);
            for (int j = 0; j < low.state_n_edges[s]; ++j) { slots[placed[j]] = lowered_src(s, j); set_w(k, PR_WB + placed[j], s, j); }
            emission_records(low.emis[s], v + PR_EB, f + PR_EB);
            kinds_word[k] = (kinds_word[k] & ~0xff0000) | (low.emis[s].kind << 16);
        }
        if (DROP[k] >= 0) {
            const int s = DROP[k];
            const int want[1] = {M[k]};
            if (!match(s, want, 1) || low.state_n_edges[s] != 1) { fail("a drop-off state's in-edges do not fit the template"); return; }
            begin_slots(s, 1)[0] = M[k];
            set_w(k, PR_WD, s, 0);
        }
        if (SKIP[k] >= 0) {
            const int s = SKIP[k];
            const int want[2] = {k > 0 ? M[k - 1] : -1, k > 0 ? SKIP[k - 1] : start};
            if (!match(s, want, 2)) { fail("a skip state's in-edges do not fit the template"); return; }
            int32_t* slots = begin_slots(s, 2);
            for (int j = 0; j < low.state_n_edges[s]; ++j) {
                slots[placed[j]] = lowered_src(s, j);
                // the edge from M_{k-1} is kept with its SOURCE position (that warp publishes M_{k-1} + t for the chain warp)
                if (placed[j] == 0) set_w(k - 1, PR_SK, s, j);
                else set_w(k, PR_SK + 1, s, j);
            }
        }
        if (BACK[k] >= 0) {
            const int s = BACK[k];
            const int want[2] = {M[k], k + 1 < K ? BACK[k + 1] : -1};
            if (!match(s, want, 2)) { fail("a back-slip state's in-edges do not fit the template"); return; }
            int32_t* slots = begin_slots(s, 2);
            for (int j = 0; j < low.state_n_edges[s]; ++j) { slots[placed[j]] = lowered_src(s, j); set_w(k, PR_BK + placed[j], s, j); }
        }
    }
    // A match state that reads a back-slip state two positions up (bake merged the one between them away): the missing
    // state becomes a phantom link of the descending chain with weight 0 (1 in linear space) and no pass-1 candidate --
    // x + 0.0 = x, so it carries its upper neighbour's value unchanged.
    for (int a = 0; a < plan->n_alias; ++a) {
        const int dst = plan->alias_dst[a], src = plan->alias_src[a];
        if (dst >= K || src != dst + 1 || BACK[src] < 0 || BACK[dst] >= 0) { fail("a match state reads a back-slip state that is not adjacent to the missing one"); return; }
        double *v, *f;
        rec(dst, &v, &f);
        v[PR_BK + 1] = 0.0;
        f[PR_BK + 1] = 1.0;
    }
    // forward sweep: products of the chain weights over each warp's block (affine block carries, pp_profile_kernels.cuh)
    for (int k0 = 0; k0 < Kr; k0 += P) {
        double up = 1.0, up_tail = 1.0, down = 1.0;
        for (int j = 0; j < P; ++j) {
            const double* f = &plan->rec_f[static_cast<size_t>(k0 + j) * kProfRecDoubles];
            up *= f[PR_SK + 1];
            if (j > 0) up_tail *= f[PR_SK + 1];
            down *= f[PR_BK + 1];
        }
        double* f0 = &plan->rec_f[static_cast<size_t>(k0) * kProfRecDoubles];
        f0[PR_CUP] = up; f0[PR_CUT] = up_tail; f0[PR_CDN] = down;
    }
    for (int k = 0; k < Kr; ++k) {
        std::memcpy(&plan->rec_v[static_cast<size_t>(k) * kProfRecDoubles + PR_KINDS], &kinds_word[k], sizeof(int32_t));
        std::memcpy(&plan->rec_f[static_cast<size_t>(k) * kProfRecDoubles + PR_KINDS], &kinds_word[k], sizeof(int32_t));
    }
    {   // end: any mixture of match / drop-off / skip / back-slip sources, reference visiting order
        const int n = low.state_n_edges[end];
        int32_t* slots = begin_slots(end, n);
        for (int j = 0; j < n; ++j) {
            const int src = lowered_src(end, j);
            int c;
            switch (cls[src]) {
                case 0: c = PROF_M; break;
                case 3: c = PROF_DROP; break;
                case 4: c = PROF_SKIP; break;
                case 5: c = PROF_BACK; break;
                default: fail("the end state is fed by an insert, blip or start state"); return;
            }
            slots[j] = src;
            const double w = edge_logp(end, j);
            plan->end_v.push_back(ProfEndRec{c, pos[src], w});
            plan->end_f.push_back(ProfEndRec{c, pos[src], std::exp(w)});
        }
    }
    plan->slot_begin[start] = static_cast<int32_t>(plan->slot_src.size());

    // ---- where the ordinals sit in a pointer row ----------------------------------------------------------------
    // position warp w: one unit of UB bytes per lane with 5 bits per position (M: 3, INS: 1, BLIP: 1), written with row
    // r + 1's emitting states; behind those a unit of one byte per lane with the ordinals of the warp's chain states
    // (bit j: S_SKIP_{k0+j}, bit P + j: B_BACK_{k0+j}), written when the warp fills in row r's chain values
    const int UB = 5 * P <= 16 ? 2 : 4;
    const int CB = 2 * P <= 8 ? 1 : 2;                              // bytes per lane of a warp's chain ordinals (2 P bits)
    const int chain_base = NW * UB * kLanes;
    plan->row_bytes = chain_base + NW * CB * kLanes;
    plan->unit_off.assign(m, 0); plan->unit_bytes.assign(m, 1); plan->unit_shift.assign(m, 0); plan->unit_mask.assign(m, 0);
    auto place = [&](int s, int unit_off, int stride, int bit, int mask) {
        if (s < 0) return;
        plan->unit_off[s] = unit_off + bit / 8;
        plan->unit_bytes[s] = stride;                               // bytes per lane
        plan->unit_shift[s] = bit % 8;
        plan->unit_mask[s] = mask;
    };
    for (int k = 0; k < K; ++k) {
        const int w = k / P, p = k % P;
        place(M[k], w * UB * kLanes, UB, 5 * p, 7);
        place(INS[k], w * UB * kLanes, UB, 5 * p + 3, 1);
        place(BLIP[k], w * UB * kLanes, UB, 5 * p + 4, 1);
        place(DROP[k], 0, 0, 0, 0);                                 // one in-edge: ordinal 0, nothing stored
        place(SKIP[k], chain_base + w * CB * kLanes, CB, p, 1);
        place(BACK[k], chain_base + w * CB * kLanes, CB, P + p, 1);
    }
    plan->ok = true;

    // ---- forward-backward: the same weights keyed by their SOURCE, and the statistic every term feeds -----------------
    {
        auto fb_fail = [&](const std::string& why) { plan->fb_ok = false; plan->fb_why = why; };
        const int R = kFbRecDoubles;
        plan->rec_b.assign(static_cast<size_t>(Kr) * R, 0.0);
        auto F = [&](int k, int field) -> double {
            return (k >= 0 && k < Kr) ? plan->rec_f[static_cast<size_t>(k) * kProfRecDoubles + field] : 0.0;
        };
        for (int k = 0; k < Kr; ++k) {
            double* b = &plan->rec_b[static_cast<size_t>(k) * R];
            b[FR_WM + 0] = F(k + 1, PR_WM + 0);
            b[FR_WM + 1] = F(k, PR_WM + 2);
            b[FR_WM + 2] = F(k + 1, PR_WI + 0);
            b[FR_WM + 3] = F(k, PR_WB + 0);
            b[FR_WM + 4] = F(k, PR_WD);
            b[FR_WM + 5] = F(k - 1, PR_WM + 6);
            b[FR_WM + 6] = F(k, PR_SK);
            b[FR_WM + 7] = F(k, PR_BK);
            b[FR_WS + 0] = F(k + 1, PR_WM + 1);
            b[FR_WS + 1] = F(k + 1, PR_SK + 1);
            b[FR_WK + 0] = F(k - 1, PR_WM + 7);
            b[FR_WK + 1] = F(k - 1, PR_BK + 1);
            b[FR_WD] = F(k, PR_WM + 3);
            b[FR_WI + 0] = F(k, PR_WM + 5);
            b[FR_WI + 1] = F(k, PR_WI + 1);
            b[FR_WB + 0] = F(k, PR_WM + 4);
            b[FR_WB + 1] = F(k, PR_WB + 1);
            for (int j = 0; j < 4; ++j) {
                b[FR_EM + j] = F(k, PR_EM + j);
                b[FR_EI + j] = F(k, PR_EI + j);
                b[FR_EB + j] = F(k, PR_EB + j);
            }
            b[FR_EM + 3] = 0.0;                                     // PR_CUT lives there in the forward records
            std::memcpy(&b[FR_KINDS], &kinds_word[k], sizeof(int32_t));
        }
        for (const ProfEndRec& E : plan->end_f) {
            const int j = E.cls == PROF_M ? 0 : E.cls == PROF_DROP ? 1 : E.cls == PROF_SKIP ? 2 : 3;
            plan->rec_b[static_cast<size_t>(E.k) * R + FR_WE + j] = E.w;
        }
        auto B = [&](int k, int field) -> double { return (k >= 0 && k < Kr) ? plan->rec_b[static_cast<size_t>(k) * R + field] : 0.0; };
        for (int k0 = 0; k0 < Kr; k0 += P) {
            double cdn = 1.0, cup = 1.0, tdn = k0 >= P ? B(k0 - 1, FR_WS + 0) : 0.0, tup = B(k0 + P, FR_WK + 0);
            for (int j = 0; j < P; ++j) { cdn *= B(k0 + j, FR_WS + 1); cup *= B(k0 + j, FR_WK + 1); }
            for (int j = 0; j < P - 1; ++j) tdn *= B(k0 - P + j, FR_WS + 1);
            for (int j = 1; j < P; ++j) tup *= B(k0 + P + j, FR_WK + 1);
            double* b = &plan->rec_b[static_cast<size_t>(k0) * R];
            b[FR_CDN] = cdn; b[FR_CUP] = cup; b[FR_TDN] = tdn; b[FR_TUP] = tup;
        }
        plan->start_wm = F(0, PR_WM + 1);
        plan->start_ws = F(0, PR_SK + 1);

        // statistic of every term: the out-edge it counts (or the moment it sums); a phantom back-slip position stands
        // for the state whose value it carries
        auto oe = [&](int s, int t) -> int {
            if (s < 0 || t < 0) return -1;
            for (int e = out_begin(s); e < out_end(s); ++e) if (d.out_dst[e] == t) return e;
            return -1;
        };
        auto at = [&](const std::vector<int32_t>& v, int k) { return (k >= 0 && k < K) ? v[k] : -1; };
        auto back_src = [&](int k) {
            const int s = at(BACK, k);
            if (s >= 0) return s;
            for (int a = 0; a < plan->n_alias; ++a) if (plan->alias_dst[a] == k) return at(BACK, plan->alias_src[a]);
            return -1;
        };
        const int E = d.n_edges;
        plan->fb_slots = Kp * kFbChunks + 1;
        plan->fb_map.assign(static_cast<size_t>(plan->fb_slots) * kLanes, -1);
        plan->fb_log2.assign(plan->fb_slots, 3);
        plan->fb_emit_state.assign(static_cast<size_t>(Kp) * 3, -1);
        std::vector<int> seen(E, 0);
        bool weights_ok = true;
        auto put = [&](int k, int term, int stat, double w) {
            if (stat < 0) return;
            if (stat < E) {
                ++seen[stat];
                if (w != std::exp(d.out_logp[stat])) weights_ok = false;
            }
            const size_t slot = static_cast<size_t>(k) * kFbChunks + term / 8;
            for (int q = 0; q < 4; ++q) plan->fb_map[slot * kLanes + (term % 8) * 4 + q] = stat;
        };
        for (int k = 0; k < K; ++k) {
            const int mk = M[k], dk = at(DROP, k), sk = at(SKIP, k), bk = back_src(k), ik = at(INS, k), lk = at(BLIP, k);
            plan->fb_emit_state[k * 3 + 0] = mk;
            plan->fb_emit_state[k * 3 + 1] = ik;
            plan->fb_emit_state[k * 3 + 2] = lk;
            put(k, 0, oe(mk, at(M, k + 1)), B(k, FR_WM + 0));
            put(k, 1, oe(mk, mk), B(k, FR_WM + 1));
            put(k, 2, oe(mk, at(INS, k + 1)), B(k, FR_WM + 2));
            put(k, 3, oe(mk, lk), B(k, FR_WM + 3));
            put(k, 4, oe(mk, dk), B(k, FR_WM + 4));
            put(k, 5, oe(mk, at(M, k - 1)), B(k, FR_WM + 5));
            put(k, 6, oe(mk, end), B(k, FR_WE + 0));
            put(k, 7, oe(dk, mk), B(k, FR_WD));
            put(k, 8, oe(dk, end), B(k, FR_WE + 1));
            put(k, 9, oe(sk, at(M, k + 1)), B(k, FR_WS + 0));
            put(k, 10, oe(sk, end), B(k, FR_WE + 2));
            put(k, 11, oe(bk, at(M, k - 1)), B(k, FR_WK + 0));
            put(k, 12, at(BACK, k) >= 0 ? oe(bk, end) : -1, B(k, FR_WE + 3));
            put(k, 13, oe(ik, mk), B(k, FR_WI + 0));
            put(k, 14, oe(ik, ik), B(k, FR_WI + 1));
            put(k, 15, oe(lk, mk), B(k, FR_WB + 0));
            put(k, 16, oe(lk, lk), B(k, FR_WB + 1));
            for (int j = 0; j < 3; ++j) {
                if (ik >= 0) put(k, 17 + j, E + 3 * ik + j, 0.0);
                if (lk >= 0) put(k, 20 + j, E + 3 * lk + j, 0.0);
                put(k, 28 + j, E + 3 * mk + j, 0.0);
            }
            put(k, 24, oe(mk, at(SKIP, k + 1)), B(k, FR_WM + 6));
            put(k, 25, oe(mk, at(BACK, k)), B(k, FR_WM + 7));
            put(k, 26, oe(sk, at(SKIP, k + 1)), B(k, FR_WS + 1));
            put(k, 27, at(BACK, k - 1) >= 0 ? oe(bk, at(BACK, k - 1)) : -1, B(k, FR_WK + 1));
        }
        {   // the start state: slot Kp * kFbChunks, terms 0 and 1
            const size_t slot = static_cast<size_t>(Kp) * kFbChunks;
            const int e0 = oe(start, M[0]), e1 = oe(start, at(SKIP, 0));
            for (int q = 0; q < 4; ++q) { plan->fb_map[slot * kLanes + q] = e0; plan->fb_map[slot * kLanes + 4 + q] = e1; }
            if (e0 >= 0) { ++seen[e0]; if (plan->start_wm != std::exp(d.out_logp[e0])) weights_ok = false; }
            if (e1 >= 0) { ++seen[e1]; if (plan->start_ws != std::exp(d.out_logp[e1])) weights_ok = false; }
        }
        plan->fb_ok = weights_ok;
        if (!weights_ok) fb_fail("a backward weight differs from its out-edge");
        for (int e = 0; e < E && plan->fb_ok; ++e)
            if (seen[e] != 1) fb_fail("an out-edge is counted " + std::to_string(seen[e]) + " times by the profile terms");
    }
}

}  // namespace pp
