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
        bool chain_mode = false;
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
            if (silent_list) {
                // as in the forward lowering: the items of a run must not read each other's values (all loads of a
                // group are issued before its stores), except the chain pattern -- the last edge reads the item before
                bool dep_other = false, dep_chain = false;
                for (int k = 0; k < n.ne; ++k)
                    for (size_t t = i; t < j; ++t)
                        if (n.src[k] == seq[t].dst) {
                            if (k == n.ne - 1 && t + 1 == j) dep_chain = true;
                            else dep_other = true;
                        }
                if (dep_other) break;
                if (j - i == 1) chain_mode = dep_chain;
                else if (dep_chain != chain_mode) break;
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

    // An item before its weights are placed: out-edges as (slot the edge reads, out-edge CSR index).
    struct Item {
        BDesc d;
        int eid[kMaxRunEdges];
    };
    auto slot_of_value = [&](int l) { return l < ne ? HB + l : SIL + (l - ne); };  // what a predecessor reads for target l
    auto make_item = [&](int k, int chain_prev, int kind, Item* o) -> bool {
        o->d = BDesc{};
        o->d.state = k;
        o->d.kind = kind;
        o->d.dst = k < ne ? EB + k : SIL + (k - ne);
        int chain_edge = -1, n = 0;
        for (int pass = 0; pass < 2; ++pass)                      // emitting targets first, then silent ones
            for (int e = d.out_ptr[k]; e < d.out_ptr[k + 1]; ++e) {
                const int l = d.out_dst[e];
                if ((pass == 0) != (l < ne)) continue;
                if (k >= ne && l >= ne && l < k + 1) continue;    // silent -> silent only to larger indices (3026, 3108)
                if (l == chain_prev) { chain_edge = e; continue; }
                if (n >= kMaxRunEdges) return false;
                o->d.src[n] = slot_of_value(l);
                o->eid[n++] = e;
            }
        if (chain_edge >= 0) {
            if (n >= kMaxRunEdges) return false;
            o->d.src[n] = slot_of_value(chain_prev);
            o->eid[n++] = chain_edge;
        }
        o->d.ne = n;
        return true;
    };
    // family order: items with the same in-degree and the same slots relative to their own go next to each other,
    // so that positions k, k+1, ... of a profile become one affine run
    auto family_less = [](const Item& a, const Item& b) {
        if (a.d.ne != b.d.ne) return a.d.ne > b.d.ne;
        for (int j = 0; j < a.d.ne; ++j) {
            const int ra = a.d.src[j] - a.d.dst, rb = b.d.src[j] - b.d.dst;
            if (ra != rb) return ra < rb;
        }
        return false;
    };
    auto item_cost = [](const Item& x, bool acc, bool emit) {
        long c = 12 + 5L * x.d.ne;
        if (acc) c += 7L * (x.d.ne + (emit ? 3 : 0)) + 8;
        return c;
    };

    // ---- list 0: H items (one per emitting state) ------------------------------------------------------------
    std::vector<Item> h_items;
    {
        std::vector<int> order(ne);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return low->emis[a].kind < low->emis[b].kind; });
        for (int l : order) {
            Item h{};
            h.d.kind = low->emis[l].kind; h.d.ne = 1; h.d.dst = HB + l; h.d.state = l; h.d.src[0] = EB + l; h.eid[0] = -1;
            h_items.push_back(h);
        }
    }
    // ---- list 1: silent VALUES, no accumulation: the components of the silent sub-DAG in descending index ------
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
    std::vector<std::vector<Item>> sil_of(W);
    {
        std::vector<int> order(comps.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return comps[a].size() > comps[b].size(); });
        std::vector<long> load(W, 0);
        std::vector<std::vector<Item>> chains(W), singles(W);
        for (int c : order) {
            const int w = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
            int prev = -1;
            for (int k : comps[c]) {
                Item s{};
                if (!make_item(k, prev, RUN_SILENT, &s)) return "";
                (comps[c].size() > 1 ? chains[w] : singles[w]).push_back(s);
                prev = k;
                load[w] += item_cost(s, false, false) + (comps[c].size() > 1 ? 30 : 0);
            }
        }
        for (int w = 0; w < W; ++w) {
            std::stable_sort(singles[w].begin(), singles[w].end(), family_less);
            sil_of[w] = chains[w];
            sil_of[w].insert(sil_of[w].end(), singles[w].begin(), singles[w].end());
        }
    }
    // ---- list 2: everything that accumulates: emitting items (value + statistics) and, for every silent state, an
    //      accumulate-only item that re-reads the finished column (keeps the serial chains of list 1 short) --------
    std::vector<Item> acc_items;
    {
        std::vector<Item> em, sa;
        for (int k = 0; k < ne; ++k) {
            Item e{};
            if (!make_item(k, -1, RUN_SILENT, &e)) return "";
            em.push_back(e);
        }
        std::vector<Item> singles;                              // chains keep their chain order (affine), singles by family
        std::vector<int> order(comps.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return comps[a].size() > comps[b].size(); });
        for (int c : order)
            for (int k : comps[c]) {
                Item a{};
                if (!make_item(k, -1, RUN_ACC, &a)) return "";
                if (a.d.ne == 0) continue;
                (comps[c].size() > 1 ? sa : singles).push_back(a);
            }
        std::stable_sort(singles.begin(), singles.end(), family_less);
        sa.insert(sa.end(), singles.begin(), singles.end());
        std::stable_sort(em.begin(), em.end(), family_less);
        acc_items = em;
        acc_items.insert(acc_items.end(), sa.begin(), sa.end());
    }
    // ---- place the weights in final list order (constant weight stride inside a run), deal lists 0 and 2 to the
    //      warps in contiguous cost-balanced slices, emit runs and accumulation slots -------------------------------
    const int one_idx = static_cast<int>(low->bw_wtab.size());
    low->bw_wtab.push_back(1.0);
    low->bw_edge.push_back(-1);
    auto place = [&](std::vector<Item>& items, bool is_h) {
        for (Item& x : items) {
            if (is_h) { x.d.wbase = one_idx; continue; }
            x.d.wbase = static_cast<int>(low->bw_wtab.size());
            for (int j = 0; j < x.d.ne; ++j) {
                low->bw_wtab.push_back(std::exp(d.out_logp[x.eid[j]]));
                low->bw_edge.push_back(x.eid[j]);
            }
        }
    };
    auto slice = [&](const std::vector<Item>& items, int w, bool acc) {
        long total = 0;
        for (const Item& x : items) total += item_cost(x, acc, x.d.kind != RUN_ACC);
        std::vector<Item> out;
        long run = 0;
        int cur = 0;
        for (const Item& x : items) {
            if (cur + 1 < W && run >= (total * (cur + 1)) / W) ++cur;
            if (cur == w) out.push_back(x);
            run += item_cost(x, acc, x.d.kind != RUN_ACC);
        }
        return out;
    };
    // Accumulation slots: replay the kernel's group decomposition (fbb_items in pp_lane_fb.cuh) over the runs of
    // list 2; the terms of a group are {edge terms of item 0, [w, wx, wxx], item 1 ...}.
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
    auto map_run = [&](const RunRec& r) {
        const bool emit = r.kind != RUN_ACC;
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
        for (; left >= G; left -= G) group(G);
        if (G > 2 && left >= 2) { group(2); left -= 2; }
        if (left >= 1) group(1);
    };
    low->acc_base.assign(W, 0);
    for (int w = 0; w < W; ++w) {
        std::vector<Item> lists[3] = {slice(h_items, w, false), sil_of[w], slice(acc_items, w, true)};
        low->acc_base[w] = static_cast<int32_t>(low->acc_log2.size());
        for (int list = 0; list < 3; ++list) {
            place(lists[list], list == 0);
            std::vector<BDesc> seq;
            for (const Item& x : lists[list]) seq.push_back(x.d);
            const size_t r0 = low->bw_runs.size();
            low->bw_ranges[w * 6 + list * 2] = static_cast<int32_t>(r0);
            build_runs(seq, list == 1, &low->bw_runs);
            low->bw_ranges[w * 6 + list * 2 + 1] = static_cast<int32_t>(low->bw_runs.size());
            if (list == 2)
                for (size_t r = r0; r < low->bw_runs.size(); ++r) map_run(low->bw_runs[r]);
        }
    }
    low->n_acc_slots = static_cast<int32_t>(low->acc_log2.size());
    mark_shared_c2(*low, &low->bw_runs);                       // only the H runs carry an emission kind
    low->bw_ok = true;
    return "";
}

}  // namespace pp
