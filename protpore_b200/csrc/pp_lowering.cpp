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
            // (261); a == b: the reference returns 0 for x == a == b (257-259), and nothing else is in range
            e.a3 = (p1 == p0) ? 0.0 : ref_log(1.0 / (p1 - p0));
            e.c2 = (e.a3 + lw) * kLog2e;
        } else if (d.emit_kind[k] == PP_EMIT_TABLE) {
            e.kind = EMIS_TABLE;                                // density comes from the caller's table; only logw is kept
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
        // (no limit here: the state-parallel kernels store 16-bit source states; the lane kernels' 4-bit ordinals are
        //  checked below, where such a model only loses the lane path)
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

    // ---- deal states to warps, as RUNS -------------------------------------------------------
    // A run is a sequence of items of one shape (emission kind, in-degree) whose destination slot,
    // source slots, weight index and emission index all advance by constant strides -- position
    // k, k+1, k+2 ... of a profile HMM.  The kernels execute a run with one shape-templated body and
    // a handful of induction registers, so the code a CTA touches per row is a few KB however many
    // states the model has.  A model without any regularity degrades to runs of length 1.
    struct Desc { int dst; int src[kMaxRunEdges]; int wbase; int ne; int kind; };
    auto slot_of_dst = [&](int l) { return l < ne ? ns + l : l - ne; };          // parity-0 slot; parity 1 adds ne for emitting
    auto describe = [&](int l, Desc* o) -> bool {
        o->ne = low->state_n_edges[l];
        if (o->ne > kMaxRunEdges) return false;
        o->dst = slot_of_dst(l);
        o->wbase = low->state_edge_begin[l];
        o->kind = l < ne ? low->emis[l].kind : (l == d.start_index ? RUN_START : RUN_SILENT);
        for (int k = 0; k < o->ne; ++k) {
            const int s = low->low_src[o->wbase + k];
            // encode region in the slot number of parity 0: silent -> s-ne ; emitting, previous row (emitting dst) ->
            // ns + ne + s ; emitting, this row (silent dst) -> ns + s.  (parity 1 swaps the two emitting regions)
            o->src[k] = s >= ne ? s - ne : (l < ne ? ns + ne + s : ns + s);
        }
        return true;
    };
    auto same_region = [&](int a, int b) {        // both slots in the same column region?
        auto reg = [&](int s) { return s < ns ? 0 : (s < ns + ne ? 1 : 2); };
        return reg(a) == reg(b);
    };
    // can `nxt` extend a run whose last two items are prev2 (may be null) and prev?
    auto extends = [&](const Desc* prev2, const Desc& prev, const Desc& nxt) {
        if (nxt.kind != prev.kind || nxt.ne != prev.ne) return false;
        if (!same_region(nxt.dst, prev.dst)) return false;
        for (int k = 0; k < nxt.ne; ++k) if (!same_region(nxt.src[k], prev.src[k])) return false;
        if (!prev2) return true;
        if (nxt.dst - prev.dst != prev.dst - prev2->dst) return false;
        if (nxt.wbase - prev.wbase != prev.wbase - prev2->wbase) return false;
        for (int k = 0; k < nxt.ne; ++k) if (nxt.src[k] - prev.src[k] != prev.src[k] - prev2->src[k]) return false;
        return true;
    };
    auto slot_bytes = [&](int slot0, int q) -> int32_t {       // parity-0 slot number -> byte offset for row parity q
        int s = slot0;
        if (q == 1 && s >= ns) s = (s < ns + ne) ? s + ne : s - ne;
        return static_cast<int32_t>(s) * static_cast<int32_t>(kSlotBytes);
    };
    low->lane_ok = true;
    for (int l = 0; l < m; ++l) {
        if (low->end_final_only && l == d.end_index) continue;
        if (low->state_n_edges[l] > kMaxRunEdges) { low->lane_ok = false; low->lane_why = "state " + std::to_string(l) + " has more than 8 usable in-edges"; }
    }
    for (int k = 0; k < ne; ++k)
        if (low->emis[k].kind == EMIS_TABLE) { low->lane_ok = false; low->lane_why = "state " + std::to_string(k) + " takes its emissions from a table (state-parallel kernels only)"; }
    if (!low->lane_ok) return "";                 // only the state-parallel kernels can take this model
    auto make_runs = [&](const std::vector<int>& states, bool silent, std::vector<RunRec>* out) {
        size_t i = 0;
        while (i < states.size()) {
            Desc a;
            if (!describe(states[i], &a)) { ++i; continue; }
            std::vector<Desc> seq{a};
            size_t j = i + 1;
            bool chain_mode = false;
            while (j < states.size()) {
                Desc nx;
                if (!describe(states[j], &nx)) break;
                // emission indices must advance with a constant stride as well
                if (!silent && seq.size() >= 2 &&
                    states[j] - states[j - 1] != states[j - 1] - states[j - 2]) break;
                if (!extends(seq.size() >= 2 ? &seq[seq.size() - 2] : nullptr, seq.back(), nx)) break;
                if (silent) {
                    // The kernels load ALL sources of a group of items before they store any of them, so the items of
                    // a run must not read each other -- except for the one pattern they forward in a register: the
                    // LAST edge of every item reads the item right before it (a chain).  Anything else starts a new run.
                    bool dep_other = false, dep_chain = false;
                    for (int k = 0; k < nx.ne; ++k) {
                        const int src = low->low_src[nx.wbase + k];
                        for (size_t t = i; t < j; ++t)
                            if (src == states[t]) {
                                if (k == nx.ne - 1 && t + 1 == j) dep_chain = true;
                                else dep_other = true;
                            }
                    }
                    if (dep_other) break;
                    if (seq.size() == 1) chain_mode = dep_chain;
                    else if (dep_chain != chain_mode) break;
                }
                seq.push_back(nx);
                ++j;
            }
            RunRec r{};
            r.kind = a.kind; r.n_edges = a.ne; r.count = static_cast<int32_t>(seq.size());
            const Desc& b = seq.size() > 1 ? seq[1] : seq[0];
            for (int q = 0; q < 2; ++q) {
                r.dst_off[q] = slot_bytes(a.dst, q);
                for (int k = 0; k < a.ne; ++k) r.src_off[q][k] = slot_bytes(a.src[k], q);
            }
            r.dst_stride = (b.dst - a.dst) * static_cast<int32_t>(kSlotBytes);
            for (int k = 0; k < a.ne; ++k) r.src_stride[k] = (b.src[k] - a.src[k]) * static_cast<int32_t>(kSlotBytes);
            r.w_base = a.wbase; r.w_stride = b.wbase - a.wbase;
            r.e_base = silent ? 0 : states[i];
            r.e_stride = (silent || seq.size() < 2) ? 0 : states[i + 1] - states[i];
            r.has_logw = 0;
            if (!silent)
                for (size_t t = i; t < j; ++t)
                    if (low->emis[states[t]].logw != 0.0) r.has_logw = 1;
            // register forward: the LAST edge of every item but the first reads the previous item
            r.fwd = 0;
            if (silent && a.ne > 0 && seq.size() > 1) {
                bool all = true;
                for (size_t t = 1; t < seq.size(); ++t)
                    if (low->low_src[seq[t].wbase + a.ne - 1] != states[i + t - 1]) { all = false; break; }
                r.fwd = all ? 1 : 0;
                // 2: every side source is an emitting state -> the kernels may load the next group of links
                //    before they store the current one (no link reads a silent slot written inside the run)
                if (all) {
                    bool side_emitting = true;
                    for (size_t t = 0; t < seq.size(); ++t)
                        for (int k = 0; k + 1 < a.ne; ++k)
                            if (low->low_src[seq[t].wbase + k] >= ne) side_emitting = false;
                    if (side_emitting) r.fwd = 2;
                }
            }
            out->push_back(r);
            i = j;
        }
    };

    // Phases of one row r (steady state), separated by the two barriers of the row:
    //   A  "late" emitting states of row r  -- those with a silent source, so they need silent(r-1);
    //   B  silent states of row r, and next to them the "early" emitting states of row r+1 -- those
    //      whose sources are all emitting states of row r, already complete after phase A's barrier.
    // Early items fill the warps that have no silent chain to walk: in a profile HMM the chain owners
    // are busy for the whole of phase B while everybody else would idle at the barrier.
    std::vector<std::vector<int>> late_of(warps), early_of(warps), sil_of(warps);
    std::vector<int> late, early;
    for (int l = 0; l < ne; ++l) {
        bool has_silent_src = false;
        for (int k = 0; k < low->state_n_edges[l]; ++k)
            if (low->low_src[low->state_edge_begin[l] + k] >= ne) has_silent_src = true;
        (has_silent_src ? late : early).push_back(l);
    }
    auto shape_sort = [&](std::vector<int>* v) {
        std::stable_sort(v->begin(), v->end(), [&](int a, int b) {
            if (low->state_n_edges[a] != low->state_n_edges[b]) return low->state_n_edges[a] > low->state_n_edges[b];
            return low->emis[a].kind < low->emis[b].kind;
        });
    };
    auto emit_cost = [&](int l) { return 9L * low->state_n_edges[l] + (low->emis[l].kind == EMIS_NORMAL ? 22 : 14); };
    shape_sort(&late);
    shape_sort(&early);
    {   // late: cut the shape-sorted sequence into W contiguous slices of equal cost
        long total = 0, acc = 0;
        for (int l : late) total += emit_cost(l);
        int w = 0;
        for (int l : late) {
            if (w + 1 < warps && acc >= (total * (w + 1)) / warps) ++w;
            late_of[w].push_back(l);
            acc += emit_cost(l);
        }
    }
    std::vector<long> load_b(warps, 0);               // phase-B load per warp (silent work first)
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
            for (int l : comps[c]) s += 9L * low->state_n_edges[l] + 8;
            // a chain is latency-bound: weigh long components far above their instruction count
            return s + (comps[c].size() > 1 ? 60L * static_cast<long>(comps[c].size()) : 0L);
        };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
        std::vector<std::vector<int>> chains(warps), singles(warps);
        for (int c : order) {
            int w = static_cast<int>(std::min_element(load_b.begin(), load_b.end()) - load_b.begin());
            auto& dstv = comps[c].size() > 1 ? chains[w] : singles[w];
            for (int l : comps[c]) dstv.push_back(l);
            load_b[w] += cost(c);
        }
        for (int w = 0; w < warps; ++w) {
            // chains first (the critical path starts at once), then the independent single states sorted by
            // (shape, index) so that they form runs
            std::stable_sort(singles[w].begin(), singles[w].end(), [&](int a, int b) {
                if (low->state_n_edges[a] != low->state_n_edges[b]) return low->state_n_edges[a] > low->state_n_edges[b];
                return a < b;
            });
            sil_of[w] = chains[w];
            sil_of[w].insert(sil_of[w].end(), singles[w].begin(), singles[w].end());
        }
    }
    {   // early: fill the warps up to a common phase-B level (water-filling over the shape-sorted sequence)
        long need = 0;
        for (int l : early) need += emit_cost(l);
        // smallest level such that sum_w max(0, level - load_b[w]) covers the early work
        long lo = *std::min_element(load_b.begin(), load_b.end());
        long hi = *std::max_element(load_b.begin(), load_b.end()) + need + 1;
        while (lo < hi) {
            const long mid = lo + (hi - lo) / 2;
            long cap = 0;
            for (long v : load_b) cap += std::max(0L, mid - v);
            if (cap >= need) hi = mid; else lo = mid + 1;
        }
        const long level = lo;
        int w = 0;
        long acc = 0;
        for (int l : early) {
            while (w + 1 < warps && acc + load_b[w] >= level) { ++w; acc = 0; }
            early_of[w].push_back(l);
            acc += emit_cost(l);
        }
    }

    // Traceback ordinals.  The kernels handle a run in GROUPS of 4 items (in-degree <= 2) or 2 items (more),
    // then a remainder of one pair and/or one single; a chain run handles its head alone first.  Each group
    // stores ITS ordinals with one coalesced store: a UNIT of 1, 2 or 4 bytes per lane, laid out as
    // [unit][lane] -- so there is no packing state carried between items.  Ordinals take 4 bits when every
    // item has <= 16 in-edges, else 8.
    low->ptr_bits = low->max_item_edges <= 16 ? 4 : 8;
    low->state_pword.assign(m, -1);
    low->state_pmul.assign(m, 1);
    low->state_pshift.assign(m, 0);
    int row_bytes = 0;
    low->runs.clear();
    low->warp_run_ranges.assign(static_cast<size_t>(warps) * 6, 0);
    auto place = [&](const std::vector<int>& states, size_t* si, int n_items) {      // one unit holding n_items ordinals
        const int unit = std::max(1, n_items * low->ptr_bits / 8);
        for (int k = 0; k < n_items; ++k, ++*si) {
            const int bit = k * low->ptr_bits;
            low->state_pword[states[*si]] = row_bytes + bit / 8;
            low->state_pmul[states[*si]] = unit;
            low->state_pshift[states[*si]] = bit % 8;
        }
        row_bytes += unit * kLanes;
    };
    for (int w = 0; w < warps; ++w) {
        for (int list = 0; list < 3; ++list) {                  // 0: late emitting (A), 1: silent (B), 2: early emitting (B)
            const std::vector<int>& states = list == 0 ? late_of[w] : (list == 1 ? sil_of[w] : early_of[w]);
            low->warp_run_ranges[w * 6 + list * 2 + 0] = static_cast<int32_t>(low->runs.size());
            make_runs(states, list == 1, &low->runs);
            low->warp_run_ranges[w * 6 + list * 2 + 1] = static_cast<int32_t>(low->runs.size());
            size_t si = 0;
            for (int r = low->warp_run_ranges[w * 6 + list * 2]; r < low->warp_run_ranges[w * 6 + list * 2 + 1]; ++r) {
                RunRec& R = low->runs[r];
                R.pword = row_bytes;                             // byte offset of the run's first unit in a pointer row
                R.pshift = 0;
                int left = R.count;
                if (R.fwd) { place(states, &si, 1); --left; }    // chain head
                const int G = run_group_size(R.n_edges);
                for (; left >= G; left -= G) place(states, &si, G);
                if (left >= 2) { place(states, &si, 2); left -= 2; }
                if (left >= 1) { place(states, &si, 1); left -= 1; }
            }
        }
    }
    low->n_ptr_words = (row_bytes + 127) / 128;                  // pointer-row stride in units of 32 lanes x 4 bytes
    low->late_of = late_of;
    low->early_of = early_of;
    low->sil_of = sil_of;
    mark_shared_c2(*low, &low->runs);
    return lower_backward(d, low);
}

void mark_shared_c2(const Lowered& low, std::vector<RunRec>* runs) {
    for (RunRec& r : *runs) {
        r.shared_c2 = 0;
        if (r.kind != EMIS_POINT && r.kind != EMIS_UNIFORM) continue;
        bool same = true;
        const double c0 = low.emis[r.e_base].c2;
        for (int t = 1; t < r.count && same; ++t) same = low.emis[r.e_base + t * r.e_stride].c2 == c0;
        r.shared_c2 = same ? 1 : 0;
    }
}

std::string refresh_params(const pp_model_desc& d, Lowered* low) {
    if (d.n_states != low->m || d.silent_start != low->ne || d.n_edges != low->n_edges) return "topology changed";
    // A changed emission KIND (a Normal whose sigma was 0 gets the min_std clamp in the M-step, yahmm.pyx:513, and
    // leaves the point-mass branch) changes the shape of the runs: lower again from scratch.
    for (int k = 0; k < d.silent_start; ++k) {
        const int kind = d.emit_kind[k] == PP_EMIT_UNIFORM ? EMIS_UNIFORM : (d.emit_p1[k] == 0.0 ? EMIS_POINT : EMIS_NORMAL);
        if (d.emit_kind[k] == PP_EMIT_NORMAL || d.emit_kind[k] == PP_EMIT_UNIFORM) {
            if (kind != low->emis[k].kind) {
                Lowered fresh;
                std::string e = lower_model(d, low->warps, &fresh);
                if (e.empty()) *low = std::move(fresh);
                return e;
            }
        }
    }
    // a state that switched to / from table emissions changes which kernel family serves the model: lower again
    // (compared BEFORE the records are refreshed: fill_emissions overwrites the kinds)
    for (int k = 0; k < d.silent_start; ++k)
        if ((d.emit_kind[k] == PP_EMIT_TABLE) != (low->emis[k].kind == EMIS_TABLE)) {
            Lowered fresh;
            std::string e = lower_model(d, low->warps, &fresh);
            if (e.empty()) *low = std::move(fresh);
            return e;
        }
    std::string err = fill_emissions(d, low);
    if (!err.empty()) return err;
    for (int q = 0; q < 2; ++q)
        for (size_t i = 0; i < low->low_csr.size(); ++i) {
            low->edges_logp[q][i].val = d.in_logp[low->low_csr[i]];
            low->edges_prob[q][i].val = std::exp(d.in_logp[low->low_csr[i]]);
        }
    if (!low->lane_ok) return "";
    mark_shared_c2(*low, &low->runs);
    return lower_backward(d, low);
}

}  // namespace pp
