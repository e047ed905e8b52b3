// pp_lowering.cpp -- see pp_lowering.h.  Pure host C++ (no CUDA).
#include "pp_lowering.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace pp {

namespace {
const double kSqrt2Pi = 2.50662827463;          // the reference's truncated constant (yahmm.pyx:33)
const double kLog2e = 1.4426950408889634074;

double ref_log(double x) { return x > 0 ? std::log(x) : -INFINITY; }   // yahmm.pyx:36-42

std::string fill_emissions(const pp_model_desc& d, Lowered* low) {
    low->emis.assign(d.silent_start, EmisRec());
    for (int k = 0; k < d.silent_start; ++k) {
        EmisRec& e = low->emis[k];
        const double p0 = d.emit_p0[k], p1 = d.emit_p1[k], lw = d.emit_logw[k];
        e = EmisRec();
        e.logw = lw;
        if (d.emit_kind[k] == PP_EMIT_NORMAL) {
            e.a0 = p0;
            if (p1 == 0.0) {                                   // point mass (yahmm.pyx:383-387)
                e.kind = EMIS_POINT;
                e.c2 = lw * kLog2e;
            } else {
                e.kind = EMIS_NORMAL;
                e.a1 = 2 * (p1 * p1);                          // 2 * sigma ** 2            (390)
                e.a2 = 1.0 / e.a1;                             // correctly rounded reciprocal
                e.a3 = ref_log(1.0 / (p1 * kSqrt2Pi));         // _log(1/(sigma*SQRT_2_PI)) (389)
                e.c2 = (e.a3 + lw) * kLog2e;
                e.k2 = kLog2e / e.a1;
            }
        } else if (d.emit_kind[k] == PP_EMIT_UNIFORM) {
            e.kind = EMIS_UNIFORM;
            e.a0 = p0;
            e.a1 = p1;
            e.a3 = ref_log(1.0 / (p1 - p0));                   // (261); a == b is caught by x==a==b first
            e.c2 = (e.a3 + lw) * kLog2e;
        } else {
            return "emission kind " + std::to_string(d.emit_kind[k]) + " of state " + std::to_string(k) +
                   " has no device lowering";
        }
    }
    return "";
}

struct UnionFind {
    std::vector<int> p;
    explicit UnionFind(int n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int find(int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void unite(int a, int b) { a = find(a); b = find(b); if (a != b) p[std::max(a, b)] = std::min(a, b); }
};

}  // namespace

std::string lower_model(const pp_model_desc& d, int warps, Lowered* low) {
    if (!d.in_ptr || !d.in_src || !d.in_logp || !d.out_ptr || !d.out_dst || !d.out_logp) return "null CSR array";
    if (d.silent_start > 0 && (!d.emit_kind || !d.emit_p0 || !d.emit_p1 || !d.emit_logw)) return "null emission array";
    const int m = d.n_states, ne = d.silent_start, ns = m - ne;
    if (m <= 0 || ne < 0 || ne > m) return "bad state counts";
    if (d.start_index < ne || d.start_index >= m || d.end_index < ne || d.end_index >= m)
        return "start/end must be silent states";
    if (m > 65535) return "more than 65535 states";
    if (d.in_ptr[m] != d.n_edges || d.out_ptr[m] != d.n_edges) return "CSR pointer / n_edges mismatch";
    for (int e = 0; e < d.n_edges; ++e)
        if (d.in_src[e] < 0 || d.in_src[e] >= m || d.out_dst[e] < 0 || d.out_dst[e] >= m) return "edge endpoint out of range";

    *low = Lowered();
    low->m = m; low->ne = ne; low->ns = ns;
    low->start = d.start_index; low->end = d.end_index; low->finite = d.finite; low->n_edges = d.n_edges;
    low->warps = warps;
    low->n_slots = ns + 2 * ne;
    low->end_final_only = d.finite == 1 && (d.out_ptr[d.end_index + 1] - d.out_ptr[d.end_index]) == 0;

    std::string err = fill_emissions(d, low);
    if (!err.empty()) return err;

    // ---- lowered in-edge runs, state-major, reference visiting order ------------------------
    low->state_edge_begin.assign(m, 0);
    low->state_n_edges.assign(m, 0);
    auto slot_off = [&](int src, bool dst_emitting, int q) -> uint32_t {
        int slot;
        if (src >= ne) slot = src - ne;                          // silent value of "the" row
        else if (dst_emitting) slot = ns + (q ^ 1) * ne + src;   // previous row
        else slot = ns + q * ne + src;                           // row being computed
        return static_cast<uint32_t>(slot) * kSlotBytes;
    };
    auto push_edge = [&](int csr, int src, bool dst_emitting) {
        for (int q = 0; q < 2; ++q) {
            EdgeRec r{};
            r.val = d.in_logp[csr];
            r.off = slot_off(src, dst_emitting, q);
            r.src = static_cast<uint32_t>(src);
            low->edges_logp[q].push_back(r);
            r.val = std::exp(d.in_logp[csr]);
            low->edges_prob[q].push_back(r);
        }
        low->low_src.push_back(src);
        low->low_csr.push_back(csr);
    };
    for (int l = 0; l < m; ++l) {
        low->state_edge_begin[l] = static_cast<int32_t>(low->low_src.size());
        if (l < ne) {
            for (int k = d.in_ptr[l]; k < d.in_ptr[l + 1]; ++k) push_edge(k, d.in_src[k], true);
        } else {
            for (int k = d.in_ptr[l]; k < d.in_ptr[l + 1]; ++k)
                if (d.in_src[k] < ne) push_edge(k, d.in_src[k], false);
            for (int k = d.in_ptr[l]; k < d.in_ptr[l + 1]; ++k)
                if (d.in_src[k] >= ne && d.in_src[k] < l) push_edge(k, d.in_src[k], false);
        }
        low->state_n_edges[l] = static_cast<int32_t>(low->low_src.size()) - low->state_edge_begin[l];
        const bool is_final_end = low->end_final_only && l == d.end_index;
        if (!is_final_end && low->state_n_edges[l] > 255)
            return "state " + std::to_string(l) + " has more than 255 usable in-edges (8-bit traceback pointers)";
        if (!is_final_end) low->max_item_edges = std::max(low->max_item_edges, low->state_n_edges[l]);
    }

    // ---- silent topological levels (used by the state-parallel table kernels) ---------------
    low->sil_level.assign(m, 0);
    low->n_levels = 0;
    for (int l = ne; l < m; ++l) {
        int lev = 1;
        for (int k = d.in_ptr[l]; k < d.in_ptr[l + 1]; ++k) {
            int s = d.in_src[k];
            if (s >= ne && s < l) lev = std::max(lev, low->sil_level[s] + 1);
        }
        low->sil_level[l] = lev;
        low->n_levels = std::max(low->n_levels, lev);
    }

    // ---- deal items to warps ---------------------------------------------------------------
    auto make_item = [&](int l, int q) {
        ItemRec it{};
        it.edge_begin = low->state_edge_begin[l];
        it.n_edges = low->state_n_edges[l];
        it.state = l;
        it.flags = (l == d.start_index) ? ITEM_START : 0;
        if (l < ne) {
            it.emis = l;
            it.dst_off = static_cast<uint32_t>(ns + q * ne + l) * kSlotBytes;
        } else {
            it.emis = -1;
            it.dst_off = static_cast<uint32_t>(l - ne) * kSlotBytes;
        }
        return it;
    };
    std::vector<std::vector<int>> emit_of(warps), sil_of(warps);
    {   // emitting: greedy LPT on (in-degree + fixed per-item cost)
        std::vector<int> order(ne);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return low->state_n_edges[a] > low->state_n_edges[b];
        });
        std::vector<long> load(warps, 0);
        for (int l : order) {
            int w = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
            emit_of[w].push_back(l);
            load[w] += low->state_n_edges[l] + 3;
        }
        for (auto& v : emit_of) std::sort(v.begin(), v.end());
    }
    {   // silent: connected components of the silent sub-DAG, whole components per warp
        UnionFind uf(m);
        for (int l = ne; l < m; ++l) {
            if (low->end_final_only && l == d.end_index) continue;
            for (int k = d.in_ptr[l]; k < d.in_ptr[l + 1]; ++k) {
                int s = d.in_src[k];
                if (s >= ne && s < l && !(low->end_final_only && s == d.end_index)) uf.unite(s, l);
            }
        }
        std::vector<std::vector<int>> comps;
        std::vector<int> comp_of(m, -1);
        for (int l = ne; l < m; ++l) {
            if (low->end_final_only && l == d.end_index) continue;
            int r = uf.find(l);
            if (comp_of[r] < 0) { comp_of[r] = static_cast<int>(comps.size()); comps.emplace_back(); }
            comps[comp_of[r]].push_back(l);     // ascending l == topological order
        }
        std::vector<int> order(comps.size());
        std::iota(order.begin(), order.end(), 0);
        auto cost = [&](int c) {
            long s = 0;
            for (int l : comps[c]) s += low->state_n_edges[l] + 2;
            // a chain is latency-bound: weigh long components more than their edge count
            return s + 4L * static_cast<long>(comps[c].size() > 1 ? comps[c].size() : 0);
        };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
        std::vector<long> load(warps, 0);
        for (int c : order) {
            int w = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
            for (int l : comps[c]) sil_of[w].push_back(l);
            load[w] += cost(c);
        }
    }
    // traceback pointer packing: ordinals of consecutive items of one (warp, phase) share 32-bit words
    low->ptr_bits = low->max_item_edges <= 16 ? 4 : 8;
    const int per_word = 32 / low->ptr_bits;
    low->state_pword.assign(m, -1);
    low->state_pshift.assign(m, 0);
    int next_word = 0;
    auto assign_slots = [&](const std::vector<int>& states) {
        for (size_t i = 0; i < states.size(); ++i) {
            low->state_pword[states[i]] = next_word + static_cast<int>(i) / per_word;
            low->state_pshift[states[i]] = (static_cast<int>(i) % per_word) * low->ptr_bits;
        }
        next_word += (static_cast<int>(states.size()) + per_word - 1) / per_word;
    };
    for (int w = 0; w < warps; ++w) { assign_slots(emit_of[w]); assign_slots(sil_of[w]); }
    low->n_ptr_words = next_word;
    auto finish = [&](std::vector<ItemRec>& v, size_t first) {     // flag the last ordinal of every word
        for (size_t i = first; i < v.size(); ++i) {
            v[i].pword = low->state_pword[v[i].state];
            v[i].pshift = low->state_pshift[v[i].state];
            const bool last = (i + 1 == v.size()) || low->state_pword[v[i + 1].state] != v[i].pword;
            if (last) v[i].flags |= ITEM_PFLUSH;
        }
    };
    low->warp_ranges.assign(static_cast<size_t>(warps) * 4, 0);
    for (int q = 0; q < 2; ++q) {
        low->items[q].clear();
        for (int w = 0; w < warps; ++w) {
            size_t first = low->items[q].size();
            low->warp_ranges[w * 4 + 0] = static_cast<int32_t>(first);
            for (int l : emit_of[w]) low->items[q].push_back(make_item(l, q));
            finish(low->items[q], first);
            first = low->items[q].size();
            low->warp_ranges[w * 4 + 1] = static_cast<int32_t>(first);
            low->warp_ranges[w * 4 + 2] = static_cast<int32_t>(first);
            for (int l : sil_of[w]) low->items[q].push_back(make_item(l, q));
            finish(low->items[q], first);
            low->warp_ranges[w * 4 + 3] = static_cast<int32_t>(low->items[q].size());
        }
        low->end_item = static_cast<int32_t>(low->items[q].size());
        low->items[q].push_back(make_item(d.end_index, q));
    }
    low->emit_of = emit_of;
    low->sil_of = sil_of;
    return "";
}

std::string refresh_params(const pp_model_desc& d, Lowered* low) {
    if (d.n_states != low->m || d.silent_start != low->ne || d.n_edges != low->n_edges) return "topology changed";
    std::string err = fill_emissions(d, low);
    if (!err.empty()) return err;
    for (int q = 0; q < 2; ++q)
        for (size_t i = 0; i < low->low_csr.size(); ++i) {
            low->edges_logp[q][i].val = d.in_logp[low->low_csr[i]];
            low->edges_prob[q][i].val = std::exp(d.in_logp[low->low_csr[i]]);
        }
    return "";
}

}  // namespace pp
