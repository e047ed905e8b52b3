// pp_describe.cpp -- text dump of the run schedule (pp_schedule_describe; tests and DESIGN.md tables).
#include <sstream>

#include "pp_lowering.h"

namespace pp {

// Human-readable dump of the run schedule (tests).
std::string describe_schedule(const Lowered& L) {
    std::ostringstream o;
    o << "states " << L.m << " emitting " << L.ne << " silent " << L.ns << " slots " << L.n_slots << " runs " << L.runs.size()
      << " ptr_bits " << L.ptr_bits << " ptr_row_bytes " << L.n_ptr_words * 128 << " lane_ok " << (L.lane_ok ? 1 : 0) << "\n";
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
    return o.str();
}

}  // namespace pp
