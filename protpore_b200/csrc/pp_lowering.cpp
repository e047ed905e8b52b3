// pp_lowering.cpp -- see pp_lowering.h.  Pure host C++ (no CUDA), unit-testable through the ABI.
#include "pp_lowering.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace pp {

namespace {
const double kSqrt2Pi = 2.50662827463;          // the reference's truncated constant (yahmm.pyx:33)
const double kLog2e = 1.4426950408889634074;

double ref_log(double x) { return x > 0 ? std::log(x) : -INFINITY; }   // yahmm.pyx:36-42

float round_up_f32(double a) {
    float f = static_cast<float>(a);
    if (static_cast<double>(f) < a) f = std::nextafterf(f, INFINITY);
    return f;
}
float round_down_f32(double a) {
    float f = static_cast<float>(a);
    if (static_cast<double>(f) > a) f = std::nextafterf(f, -INFINITY);
    return f;
}

std::string fill_emissions(const pp_model_desc& d, Lowered* low) {
    low->emis64.assign(d.silent_start, Emis64());
    low->emisf.assign(d.silent_start, EmisF());
    for (int k = 0; k < d.silent_start; ++k) {
        Emis64& e = low->emis64[k];
        EmisF& f = low->emisf[k];
        const double p0 = d.emit_p0[k], p1 = d.emit_p1[k], lw = d.emit_logw[k];
        e.logw = lw;
        if (d.emit_kind[k] == PP_EMIT_NORMAL) {
            e.a = p0;
            f.mu_hi = static_cast<float>(p0);
            f.mu_lo = static_cast<float>(p0 - static_cast<double>(f.mu_hi));
            if (p1 == 0.0) {                                   // point mass (yahmm.pyx:383-387)
                e.kind = EMIS_POINT;
                f.kind = EMIS_POINT;
                f.c2 = static_cast<float>(lw * kLog2e);
            } else {
                e.kind = EMIS_NORMAL;
                e.b = 2 * (p1 * p1);                           // 2 * sigma ** 2      (390)
                e.c = ref_log(1.0 / (p1 * kSqrt2Pi));          // _log(1/(sigma*SQRT_2_PI)) (389)
                f.kind = EMIS_NORMAL;
                f.k2 = static_cast<float>(kLog2e / e.b);
                f.c2 = static_cast<float>((e.c + lw) * kLog2e);
            }
        } else if (d.emit_kind[k] == PP_EMIT_UNIFORM) {
            e.kind = EMIS_UNIFORM;
            e.a = p0;
            e.b = p1;
            e.c = (p1 > p0) ? ref_log(1.0 / (p1 - p0)) : 0.0;  // (261); a == b handled by the x==a==b test
            f.kind = EMIS_UNIFORM;
            f.lo = round_up_f32(p0);
            f.hi = round_down_f32(p1);
            f.c2 = static_cast<float>((e.c + lw) * kLog2e);
        } else {
            return "emission kind " + std::to_string(d.emit_kind[k]) + " of state " + std::to_string(k) +
                   " has no device lowering";
        }
        if (lw != 0.0) e.kind |= EMIS_HAS_LOGW;
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
    if (ns + 2 * ne > 65535) return "more than 65535 column slots";
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

    // ---- lowered in-edge runs, state-major ------------------------------------------------
    // slot numbering: silent state l -> l - ne ; emitting k, parity p -> ns + p*ne + k
    low->state_edge_begin.assign(m, 0);
    low->state_n_edges.assign(m, 0);
    std::vector<EdgeRecV>& ev = low->edges_v;
    auto push_edge = [&](int csr, int src, bool dst_emitting) {
        EdgeRecV r{};
        r.logp = d.in_logp[csr];
        if (src >= ne) {
            r.slot[0] = r.slot[1] = static_cast<uint16_t>(src - ne);
        } else if (dst_emitting) {              // reads the previous row: E[par]
            r.slot[0] = static_cast<uint16_t>(ns + src);
            r.slot[1] = static_cast<uint16_t>(ns + ne + src);
        } else {                                // silent destination reads the new row: E[par ^ 1]
            r.slot[0] = static_cast<uint16_t>(ns + ne + src);
            r.slot[1] = static_cast<uint16_t>(ns + src);
        }
        ev.push_back(r);
        low->low_src.push_back(src);
        low->low_csr.push_back(csr);
    };
    for (int l = 0; l < m; ++l) {
        low->state_edge_begin[l] = static_cast<int32_t>(ev.size());
        if (l < ne) {
            for (int k = d.in_ptr[l]; k < d.in_ptr[l + 1]; ++k) push_edge(k, d.in_src[k], true);
        } else {
            for (int k = d.in_ptr[l]; k < d.in_ptr[l + 1]; ++k)
                if (d.in_src[k] < ne) push_edge(k, d.in_src[k], false);
            for (int k = d.in_ptr[l]; k < d.in_ptr[l + 1]; ++k)
                if (d.in_src[k] >= ne && d.in_src[k] < l) push_edge(k, d.in_src[k], false);
        }
        low->state_n_edges[l] = static_cast<int32_t>(ev.size()) - low->state_edge_begin[l];
        const bool is_final_end = low->end_final_only && l == d.end_index;
        if (!is_final_end && low->state_n_edges[l] > 255)
            return "state " + std::to_string(l) + " has more than 255 usable in-edges (8-bit traceback pointers)";
        if (!is_final_end) low->max_item_edges = std::max(low->max_item_edges, low->state_n_edges[l]);
    }
    low->end_edge_begin = low->state_edge_begin[d.end_index];
    low->end_n_edges = low->state_n_edges[d.end_index];
    low->edges_f.resize(ev.size());
    for (size_t i = 0; i < ev.size(); ++i) {
        low->edges_f[i].w = static_cast<float>(std::exp(ev[i].logp));
        low->edges_f[i].slot[0] = ev[i].slot[0];
        low->edges_f[i].slot[1] = ev[i].slot[1];
    }

    // ---- deal items to warps ---------------------------------------------------------------
    auto make_item = [&](int l) {
        ItemRec it{};
        it.edge_begin = low->state_edge_begin[l];
        it.n_edges = static_cast<int16_t>(low->state_n_edges[l]);
        it.state = l;
        it.flags = (l == d.start_index) ? 1 : 0;
        if (l < ne) {
            it.emis = static_cast<int16_t>(l);
            it.dst_slot[0] = static_cast<uint16_t>(ns + ne + l);   // par = 0 writes E[1]
            it.dst_slot[1] = static_cast<uint16_t>(ns + l);
        } else {
            it.emis = -1;
            it.dst_slot[0] = it.dst_slot[1] = static_cast<uint16_t>(l - ne);
        }
        return it;
    };
    std::vector<std::vector<int>> emit_of(warps), sil_of(warps);
    {   // emitting: greedy LPT on (in-degree + fixed emission cost)
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
        for (auto& v : emit_of)                 // equal in-degrees run back to back inside a warp
            std::stable_sort(v.begin(), v.end(), [&](int a, int b) {
                return low->state_n_edges[a] > low->state_n_edges[b];
            });
    }
    {   // silent: connected components of the silent sub-DAG, whole components per warp
        UnionFind uf(m);
        for (int l = ne; l < m; ++l) {
            if (low->end_final_only && l == d.end_index) continue;
            for (int k = d.in_ptr[l]; k < d.in_ptr[l + 1]; ++k) {
                int s = d.in_src[k];
                if (s >= ne && s < l) uf.unite(s, l);
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
        // start from the warps with the lightest emitting load so chains do not queue behind it
        std::vector<long> load(warps, 0);
        for (int w = 0; w < warps; ++w)
            for (int l : emit_of[w]) load[w] += 0 * l;   // silent phase starts after a barrier: loads start equal
        for (int c : order) {
            int w = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
            for (int l : comps[c]) sil_of[w].push_back(l);
            load[w] += cost(c);
        }
    }
    low->warp_emit_begin.assign(warps, 0); low->warp_emit_end.assign(warps, 0);
    low->warp_sil_begin.assign(warps, 0);  low->warp_sil_end.assign(warps, 0);
    for (int w = 0; w < warps; ++w) {
        low->warp_emit_begin[w] = static_cast<int32_t>(low->items.size());
        for (int l : emit_of[w]) low->items.push_back(make_item(l));
        low->warp_emit_end[w] = static_cast<int32_t>(low->items.size());
        low->warp_sil_begin[w] = static_cast<int32_t>(low->items.size());
        for (int l : sil_of[w]) low->items.push_back(make_item(l));
        low->warp_sil_end[w] = static_cast<int32_t>(low->items.size());
    }
    return "";
}

std::string refresh_params(const pp_model_desc& d, Lowered* low) {
    if (d.n_states != low->m || d.silent_start != low->ne || d.n_edges != low->n_edges) return "topology changed";
    std::string err = fill_emissions(d, low);
    if (!err.empty()) return err;
    for (size_t i = 0; i < low->edges_v.size(); ++i) {
        low->edges_v[i].logp = d.in_logp[low->low_csr[i]];
        low->edges_f[i].w = static_cast<float>(std::exp(low->edges_v[i].logp));
    }
    return "";
}

}  // namespace pp
