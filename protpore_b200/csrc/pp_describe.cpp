// pp_describe.cpp -- text dump of the run schedule (pp_schedule_describe; tests and DESIGN.md tables).
#include <sstream>
#include <vector>

#include "pp_lowering.h"

namespace pp {

// Human-readable dump of the run schedule (tests).
std::string describe_schedule(const Lowered& L) {
    std::ostringstream o;
    o << "states " << L.m << " emitting " << L.ne << " silent " << L.ns << " slots " << L.n_slots << " runs " << L.runs.size()
      << " ptr_bits " << L.ptr_bits << " ptr_row_bytes " << L.n_ptr_words * 128 << " lane_ok " << (L.lane_ok ? 1 : 0) << "\n";
    if (!L.lane_ok) {                                   // no run schedule: the state-parallel kernels serve this model
        o << "state-parallel kernels only: " << L.lane_why << "\n";
        return o.str();
    }
    for (int w = 0; w < L.warps; ++w) {
        for (int list = 0; list < 3; ++list) {
            const int b = L.warp_run_ranges[w * 6 + list * 2], e = L.warp_run_ranges[w * 6 + list * 2 + 1];
            o << "warp " << w << (list == 0 ? " A late-emit" : list == 1 ? " B silent" : " B early-emit") << ":";
            for (int r = b; r < e; ++r) {
                const RunRec& R = L.runs[r];
                o << " [kind " << R.kind << " ne " << R.n_edges << " x" << R.count << (R.fwd ? " fwd" : "") << "]";
            }
            o << "\n";
        }
    }
    // backward schedule of the forward-backward kernel and the coverage of its accumulation slots
    o << "backward bw_ok " << (L.bw_ok ? 1 : 0) << " runs " << L.bw_runs.size() << " weights " << L.bw_wtab.size()
      << " acc_slots " << L.n_acc_slots;
    if (L.bw_ok) {
        std::vector<int> seen(static_cast<size_t>(L.n_edges) + 3 * static_cast<size_t>(L.ne), 0);
        int dup = 0;
        for (int s = 0; s < L.n_acc_slots; ++s) {
            const int width = 32 >> L.acc_log2[s];
            for (int lane = 0; lane < 32; lane += width) {
                const int st = L.acc_map[static_cast<size_t>(s) * 32 + lane];
                if (st >= 0 && seen[st]++) ++dup;
            }
        }
        int edges = 0, moments = 0;
        for (int i = 0; i < L.n_edges; ++i) edges += seen[i] ? 1 : 0;
        for (int i = 0; i < 3 * L.ne; ++i) moments += seen[L.n_edges + i] ? 1 : 0;
        o << " edges_covered " << edges << " of " << L.n_edges << " moments_covered " << moments << " of " << 3 * L.ne
          << " duplicates " << dup;
    }
    o << "\n";
    for (int w = 0; L.bw_ok && w < L.warps; ++w) {
        for (int list = 0; list < 3; ++list) {
            const int b = L.bw_ranges[w * 6 + list * 2], e = L.bw_ranges[w * 6 + list * 2 + 1];
            o << "warp " << w << (list == 0 ? " bw H" : list == 1 ? " bw silent" : " bw emit") << ":";
            for (int r = b; r < e; ++r) {
                const RunRec& R = L.bw_runs[r];
                o << " [ne " << R.n_edges << " x" << R.count << (R.fwd ? " fwd" : "") << "]";
            }
            o << "\n";
        }
    }
    return o.str();
}

}  // namespace pp
