// pp_lowering_bw.cpp -- run schedule of the BACKWARD lane-per-event sweep (linear space), the mirror of
// pp_lowering.cpp over the out-edge CSR.
//
// Backward recurrence of Model._backward (yahmm.pyx:2961-3156) in linear space, for row r (descending):
//   H[l]  = p_l(x_{r+1}) * b_{r+1}[l]                                      emitting l     (list 0)
//   b_r[k] = sum_{l emitting} w_kl H[l] + sum_{l silent, l > k} w_kl b_r[l]  silent k, descending index (list 1)
//   b_r[k] = sum_{l emitting} w_kl H[l] + sum_{l silent} w_kl b_r[l]        emitting k     (list 2)
// Column slots (32 doubles each): silent [ns] | emitting b [ne] | H [ne].  Silent chains (skip, back-slip) are
// walked by one warp in descending index with the chain edge forwarded in a register.  The end state of a
// finite model has no out-edges; its value is injected by the kernel (1 / P at the event's last row, else 0).
#include <algorithm>
#include <cmath>
#include <numeric>

#include "pp_lowering.h"

namespace pp {

namespace {

struct BDesc {
    int kind, ne, dst, state;
    int src[kMaxRunEdges];
    int wbase;
};

struct UnionFind {
    std::vector<int> p;
    explicit UnionFind(int n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int find(int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void unite(int a, int b) { a = find(a); b = find(b); if (a != b) p[std::max(a, b)] = std::min(a, b); }
};

// maximal affine runs over a sequence of item descriptors
void build_runs(const std::vector<BDesc>& seq, bool silent_list, std::vector<RunRec>* out) {
    size_t i = 0;
    while (i < seq.size()) {
        size_t j = i + 1;
        while (j < seq.size()) {
            const BDesc &a = seq[i], &p = seq[j - 1], &n = seq[j];
            if (n.kind != a.kind || n.ne != a.ne) break;
            if (j - i >= 2) {
                const BDesc& q = seq[j - 2];
                bool ok = n.dst - p.dst == p.dst - q.dst && n.wbase - p.wbase == p.wbase - q.wbase &&
                          n.state - p.state == p.state - q.state;
                for (int k = 0; k < n.ne && ok; ++k) ok = n.src[k] - p.src[k] == p.src[k] - q.src[k];
                if (!ok) break;
            }
            ++j;
        }
        const BDesc &a = seq[i], &b = (j - i > 1) ? seq[i + 1] : seq[i];
        RunRec r{};
        r.kind = a.kind; r.n_edges = a.ne; r.count = static_cast<int32_t>(j - i);
        for (int q = 0; q < 2; ++q) {
            r.dst_off[q] = a.dst * static_cast<int32_t>(kSlotBytes);
            for (int k = 0; k < a.ne; ++k) r.src_off[q][k] = a.src[k] * static_cast<int32_t>(kSlotBytes);
        }
        r.dst_stride = (b.dst - a.dst) * static_cast<int32_t>(kSlotBytes);
        for (int k = 0; k < a.ne; ++k) r.src_stride[k] = (b.src[k] - a.src[k]) * static_cast<int32_t>(kSlotBytes);
        r.w_base = a.wbase; r.w_stride = b.wbase - a.wbase;
        r.e_base = a.state; r.e_stride = b.state - a.state;
        r.fwd = 0;
        if (silent_list && a.ne > 0 && j - i > 1) {           // chain: last edge reads the previous item's value
            bool all = true;
            for (size_t t = i + 1; t < j; ++t) if (seq[t].src[a.ne - 1] != seq[t - 1].dst) { all = false; break; }
            r.fwd = all ? 1 : 0;
        }
        out->push_back(r);
        i = j;
    }
}

}  // namespace

std::string lower_backward(const pp_model_desc& d, Lowered* low) {
    low->bw_ok = false;
    low->bw_runs.clear();
    low->bw_wtab.clear();
    low->bw_edge.clear();
    low->acc_base.clear();
    low->acc_map.clear();
    low->acc_log2.clear();
    low->n_acc_slots = 0;
    const int m = d.n_states, ne = d.silent_start, ns = m - ne, W = low->warps;
    if (!low->lane_ok || !(d.finite == 1 && low->end_final_only)) return "";        // exact kernels serve these
    const int SIL = 0, EB = ns, HB = ns + ne;                      // slot bases
    low->bw_ranges.assign(static_cast<size_t>(W) * 6, 0);

    auto slot_of_value = [&](int l) { return l < ne ? HB + l : SIL + (l - ne); };  // what a predecessor reads for target l
    auto out_items = [&](int k, int chain_prev, BDesc* o) -> bool {
        o->ne = 0;
        o->state = k;
        o->kind = RUN_SILENT;
        o->dst = k < ne ? EB + k : SIL + (k - ne);
        int chain_edge = -1;
        std::vector<std::pair<int, double>> edges;               // (slot, weight)
        std::vector<int> eids;                                   // out-edge CSR index, same order
        for (int pass = 0; pass < 2; ++pass)                      // emitting targets first, then silent ones
            for (int e = d.out_ptr[k]; e < d.out_ptr[k + 1]; ++e) {
                const int l = d.out_dst[e];
                if ((pass == 0) != (l < ne)) continue;
                if (k >= ne && l >= ne && l < k + 1) continue;    // silent -> silent only to larger indices (3026, 3108)
                if (l == chain_prev) { chain_edge = e; continue; }
                edges.emplace_back(slot_of_value(l), std::exp(d.out_logp[e]));
                eids.push_back(e);
            }
        if (chain_edge >= 0) {
            edges.emplace_back(slot_of_value(chain_prev), std::exp(d.out_logp[chain_edge]));
            eids.push_back(chain_edge);
        }
        if (edges.size() > static_cast<size_t>(kMaxRunEdges)) return false;
        o->wbase = static_cast<int>(low->bw_wtab.size());
        for (size_t i = 0; i < edges.size(); ++i) {
            o->src[o->ne++] = edges[i].first;
            low->bw_wtab.push_back(edges[i].second);
            low->bw_edge.push_back(eids[i]);
        }
        return true;
    };

    // ---- list 0: H items (one per emitting state), list 2: emitting b items --------------------------------
    const int one_idx = static_cast<int>(low->bw_wtab.size());
    low->bw_wtab.push_back(1.0);
    low->bw_edge.push_back(-1);
    std::vector<BDesc> h_items, e_items;
    {
        std::vector<int> order(ne);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return low->emis[a].kind < low->emis[b].kind; });
        for (int l : order) {
            BDesc h{};
            h.kind = low->emis[l].kind; h.ne = 1; h.dst = HB + l; h.state = l; h.src[0] = EB + l; h.wbase = one_idx;
            h_items.push_back(h);
        }
        std::vector<BDesc> tmp;
        for (int k = 0; k < ne; ++k) {
            BDesc e{};
            if (!out_items(k, -1, &e)) return "";
            tmp.push_back(e);
        }
        std::vector<int> eo(ne);
        std::iota(eo.begin(), eo.end(), 0);
        std::stable_sort(eo.begin(), eo.end(), [&](int a, int b) { return tmp[a].ne > tmp[b].ne; });
        for (int k : eo) e_items.push_back(tmp[k]);
    }
    // ---- list 1: silent items, components of the silent sub-DAG, descending index ------------------------------
    std::vector<std::vector<int>> comps;
    {
        UnionFind uf(m);
        for (int k = ne; k < m; ++k) {
            if (k == d.end_index) continue;
            for (int e = d.out_ptr[k]; e < d.out_ptr[k + 1]; ++e) {
                const int l = d.out_dst[e];
                if (l >= ne && l > k && l != d.end_index) uf.unite(k, l);
            }
        }
        std::vector<int> comp_of(m, -1);
        for (int k = m - 1; k >= ne; --k) {                       // descending = processing order
            if (k == d.end_index) continue;
            const int r = uf.find(k);
            if (comp_of[r] < 0) { comp_of[r] = static_cast<int>(comps.size()); comps.emplace_back(); }
            comps[comp_of[r]].push_back(k);
        }
    }
    auto cost_items = [&](const std::vector<BDesc>& v) { long s = 0; for (auto& x : v) s += 4L * x.ne + 8; return s; };
    std::vector<std::vector<BDesc>> sil_of(W);
    {
        std::vector<int> order(comps.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return comps[a].size() > comps[b].size(); });
        std::vector<long> load(W, 0);
        std::vector<std::vector<BDesc>> chains(W), singles(W);
        for (int c : order) {
            const int w = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
            int prev = -1;
            for (int k : comps[c]) {
                BDesc s{};
                if (!out_items(k, prev, &s)) return "";
                (comps[c].size() > 1 ? chains[w] : singles[w]).push_back(s);
                prev = k;
                load[w] += 4L * s.ne + 8 + (comps[c].size() > 1 ? 40 : 0);
            }
        }
        for (int w = 0; w < W; ++w) {
            std::stable_sort(singles[w].begin(), singles[w].end(), [](const BDesc& a, const BDesc& b) {
                if (a.ne != b.ne) return a.ne > b.ne;
                return a.state > b.state;
            });
            sil_of[w] = chains[w];
            sil_of[w].insert(sil_of[w].end(), singles[w].begin(), singles[w].end());
        }
    }
    // ---- deal lists 0 and 2 to the warps in contiguous, cost-balanced slices; emit runs -------------------------
    auto slice = [&](const std::vector<BDesc>& items, int w) {
        const long total = cost_items(items);
        std::vector<BDesc> out;
        long acc = 0;
        int cur = 0;
        for (const BDesc& x : items) {
            if (cur + 1 < W && acc >= (total * (cur + 1)) / W) ++cur;
            if (cur == w) out.push_back(x);
            acc += 4L * x.ne + 8;
        }
        return out;
    };
    // Accumulation slots: replay the kernel's group decomposition (fbb_items in pp_lane_fb.cuh) over the runs of the
    // silent and emitting lists; the terms of a group are {edge terms of item 0, [w, wx, wxx], item 1 ...}.
    auto add_chunks = [&](const std::vector<int32_t>& terms) {
        size_t i = 0;
        for (int s = kAccLog2; s >= 0; --s) {
            const size_t c = size_t(1) << s;
            while (terms.size() - i >= c) {
                low->acc_log2.push_back(s);
                for (int lane = 0; lane < kLanes; ++lane) low->acc_map.push_back(terms[i + (lane >> (5 - s))]);
                i += c;
            }
        }
    };
    auto map_run = [&](const RunRec& r, bool emit) {
        const int NE = r.n_edges, G = run_group_size(NE);
        int item = 0;
        auto group = [&](int g) {
            std::vector<int32_t> terms;
            for (int q = 0; q < g; ++q, ++item) {
                for (int j = 0; j < NE; ++j) terms.push_back(low->bw_edge[r.w_base + item * r.w_stride + j]);
                if (emit) {
                    const int k = r.e_base + item * r.e_stride;
                    for (int t = 0; t < 3; ++t) terms.push_back(d.n_edges + 3 * k + t);
                }
            }
            if (!terms.empty()) add_chunks(terms);
        };
        int left = r.count;
        if (!emit && NE > 0 && r.fwd) { group(1); --left; }       // chain head
        for (; left >= G; left -= G) group(G);
        if (G > 2 && left >= 2) { group(2); left -= 2; }
        if (left >= 1) group(1);
    };
    low->acc_base.assign(W, 0);
    for (int w = 0; w < W; ++w) {
        const std::vector<BDesc> lists[3] = {slice(h_items, w), sil_of[w], slice(e_items, w)};
        low->acc_base[w] = static_cast<int32_t>(low->acc_log2.size());
        for (int list = 0; list < 3; ++list) {
            const size_t r0 = low->bw_runs.size();
            low->bw_ranges[w * 6 + list * 2] = static_cast<int32_t>(r0);
            build_runs(lists[list], list == 1, &low->bw_runs);
            low->bw_ranges[w * 6 + list * 2 + 1] = static_cast<int32_t>(low->bw_runs.size());
            if (list > 0)
                for (size_t r = r0; r < low->bw_runs.size(); ++r) map_run(low->bw_runs[r], list == 2);
        }
    }
    low->n_acc_slots = static_cast<int32_t>(low->acc_log2.size());
    low->bw_ok = true;
    return "";
}

}  // namespace pp
