// pp_lowering.h -- host-side lowering of a baked yahmm model into the column-kernel schedule.
//
// The baked CSR (Model.bake, yahmm.pyx:2280-2655) is turned into a static per-warp program for the
// "lane-per-event" column kernels: a CTA of W warps sweeps 32 events, lane j of EVERY
// warp works on event j, and the warps split the STATES of a column between them.
//   * every state becomes one ITEM: a destination slot plus a run of in-edge records;
//   * the DP column of the 32 events lives in shared memory as slots of 32 doubles (one per lane):
//     a SILENT region [ns] and two EMITTING regions [ne] (row parity 0 / 1).  Row r is computed into
//     E[r & 1] from E[(r - 1) & 1]; silent states are single-buffered because emitting states read
//     silent values of the previous row before the silent phase of the new row overwrites them;
//   * emitting items are dealt to warps by greedy longest-processing-time on their in-degree;
//   * silent states are split into the connected components of the silent->silent sub-DAG (the
//     skip chain, the back-slip chain, the isolated drop-off states ...).  A component goes to ONE
//     warp in topological (= baked index) order, so the serial chain inside a column is plain
//     program order inside one thread: no scan, no barrier, and the reference's floating-point
//     association is kept exactly;
//   * per item the in-edges are ordered exactly as the reference visits them: for an emitting
//     state all in-edges in CSR order (3545-3562); for a silent state first the emitting sources
//     (pass 1, 3564-3584) then the silent sources with index < l (pass 2, 3586-3605; edges from
//     silent sources >= l are ignored by the reference, 2861/2918/3530/3593).  The traceback
//     pointer stored per cell is the ordinal in that list, so ties break identically;
//   * a finite model's end state without out-edges is evaluated once, at each event's last row.
// Records are emitted once per row parity so that a kernel only swaps two base pointers per row.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/protpore_b200.h"

namespace pp {

constexpr int kLanes = 32;                       // events per CTA (one per lane)
#ifndef PP_LANE_WARPS
#define PP_LANE_WARPS 8
#endif
constexpr int kLaneWarps = PP_LANE_WARPS;        // warps per CTA of the lane kernels (the states of a column are split over them)
constexpr int kMaxLaneWarps = 16;
#ifndef PP_G_BIG
#define PP_G_BIG 2
#endif
constexpr uint32_t kSlotBytes = kLanes * 8;      // one column slot: 32 doubles

// ---- records shared by host and device ---------------------------------------------------
struct __attribute__((aligned(16))) EdgeRec {
    double val;                                  // log p (Viterbi) or p (forward / backward)
    uint32_t off;                                // byte offset of the source slot in the column
    uint32_t src;                                // baked index of the source state
};
// A RUN: `count` items of one shape whose slots / weights / emission records advance by constant strides
// (position k, k+1, ... of a profile).  Executed by one shape-templated loop body (pp_lane_kernels.cuh).
constexpr int kMaxRunEdges = 8;
// items a kernel handles at once in a run of this in-degree (kernels and lowering must agree)
constexpr int run_group_size(int n_edges) { return n_edges <= 2 ? 4 : (n_edges <= 4 ? 2 : PP_G_BIG); }
// forward-backward accumulation: per-lane terms are lane-transposed in chunks of kAccChunk (then 4, 2, 1)
#ifndef PP_ACC_LOG2
#define PP_ACC_LOG2 3
#endif
constexpr int kAccLog2 = PP_ACC_LOG2;
constexpr int kAccChunk = 1 << kAccLog2;
enum : int32_t { EMIS_NORMAL = 0, EMIS_POINT = 1, EMIS_UNIFORM = 2, RUN_SILENT = 3, RUN_START = 4,
                 RUN_ACC = 5 /* backward: accumulate-only item of a silent state */,
                 EMIS_TABLE = 6 /* log-density read from the caller's emission table (pp_model_set_emission_table) */ };
struct __attribute__((aligned(16))) RunRec {
    int32_t kind;                                // EMIS_* for emitting items, RUN_SILENT, RUN_START (the start state)
    int32_t n_edges;                             // in-edges per item (<= kMaxRunEdges), reference visiting order
    int32_t count;                               // items in the run
    int32_t fwd;                                 // >0: the last edge reads the previous item of the run (a chain link);
                                                 // 2: and every other edge has an emitting source (prefetch-safe)
    int32_t dst_off[2];                          // byte offset of the first destination slot, per row parity
    int32_t dst_stride;                          // bytes per item
    int32_t w_base, w_stride;                    // weight-table index of the first item's first edge; per item
    int32_t e_base, e_stride;                    // emission-record index of the first item; per item
    int32_t pword, pshift;                       // pword: byte offset, inside a traceback pointer row, of the run's first unit
    int32_t has_logw;                            // some item of the run has log(state.weight) != 0
    int32_t shared_c2;                           // emitting run of point-mass / uniform states whose log2 densities are all equal:
                                                 // the linear-space sweeps exponentiate once per run and row
    int32_t pad;
    int32_t src_off[2][kMaxRunEdges];            // byte offset of each edge's first source slot, per row parity
    int32_t src_stride[kMaxRunEdges];            // bytes per item
};
struct __attribute__((aligned(64))) EmisRec {
    // Normal : a0 = mu, a1 = 2*sigma^2, a2 = RN(1/a1), a3 = log(1/(sigma*SQRT_2_PI))
    // Point  : a0 = mu                                  (sigma == 0, yahmm.pyx:383-387)
    // Uniform: a0 = lo, a1 = hi, a3 = log(1/(hi-lo))    (yahmm.pyx:257-262)
    double a0, a1, a2, a3;
    double logw;                                 // log(state.weight) (yahmm.pyx:2515-2517)
    double c2, k2;                               // log2-domain constants of the linear-space sweeps
    int32_t kind;
    int32_t pad;
};

struct Lowered {
    int32_t m = 0, ne = 0, ns = 0, start = 0, end = 0, finite = 0, n_edges = 0;
    bool end_final_only = false;                 // end evaluated only at each event's last row
    int32_t warps = 8;                           // warps per CTA the schedule was built for
    int32_t n_slots = 0;                         // ns + 2*ne column slots
    int32_t max_item_edges = 0;                  // over per-row items (end excluded if final-only)
    int32_t ptr_bits = 8;                        // traceback ordinal width: 4 if every item has <= 16 edges, else 8
    int32_t n_ptr_words = 0;                     // pointer-row stride in units of 128 bytes (32 lanes x 4 bytes)
    // per state (traceback lookup): lane l finds the ordinal at byte state_pword + l * state_pmul of the row, shifted
    // right by state_pshift; state_pword = -1 for a final-only end state
    std::vector<int32_t> state_pword, state_pmul, state_pshift;
    std::vector<RunRec> runs;                    // warp-major: late-emitting, silent, early-emitting runs of each warp
    std::vector<int32_t> warp_run_ranges;        // [warps][6]: {run_begin, run_end} x {late emitting (A), silent (B), early emitting (B)}
    bool lane_ok = true;                         // false: some state has too many in-edges for the run kernels
    std::string lane_why;
    std::vector<std::vector<int>> late_of, sil_of, early_of;   // states per warp and list, in execution order
    std::vector<EdgeRec> edges_logp[2];          // per row parity, val = log p
    std::vector<EdgeRec> edges_prob[2];          // per row parity, val = p
    std::vector<int32_t> low_csr;                // index of each lowered edge in the IN-edge CSR
    std::vector<int32_t> state_edge_begin;       // per state: first lowered edge (traceback lookup)
    std::vector<int32_t> state_n_edges;
    std::vector<int32_t> low_src;                // source state of each lowered edge
    std::vector<int32_t> sil_level;              // per state: 0 for emitting, else 1 + longest silent-pred chain
    int32_t n_levels = 0;
    std::vector<EmisRec> emis;
    // backward sweep (pp_lowering_bw.cpp): runs over the out-edge CSR, lists {H items, silent, emitting} per warp
    bool bw_ok = false;
    std::vector<RunRec> bw_runs;
    std::vector<int32_t> bw_ranges;              // [warps][6]
    std::vector<double> bw_wtab;                 // transition probabilities in backward-lowered edge order (+ a 1.0)
    std::vector<int32_t> bw_edge;                // out-edge CSR index of each entry of bw_wtab (-1: the constant 1.0)
    // accumulation slots of the forward-backward kernel (pp_lane_fb.cuh): every chunk of <= kAccChunk per-lane terms is
    // lane-transposed and added into one 32-double slot; lane L of a slot of 2^s terms holds term L >> (5 - s)
    std::vector<int32_t> acc_base;               // [warps] first slot of the warp (program order: silent list, emitting list)
    std::vector<int32_t> acc_map;                // [n_acc_slots * 32] statistic index per (slot, lane), -1 = unused
    std::vector<int32_t> acc_log2;               // [n_acc_slots] s
    int32_t n_acc_slots = 0;                     // statistic index: out-edge e | n_edges + 3 k + {0, 1, 2} (sum w, w x, w x^2)
};

// Returns "" on success, otherwise why the model cannot be lowered.
std::string lower_model(const pp_model_desc& d, int warps, Lowered* out);
// Refresh weights / emission parameters of an existing lowering (same topology).
std::string refresh_params(const pp_model_desc& d, Lowered* low);
// Sets RunRec::shared_c2 of the emitting runs in `runs` from the current emission records.
void mark_shared_c2(const Lowered& low, std::vector<RunRec>* runs);
// Run schedule of the backward sweep; sets low->bw_ok (false: the exact state-parallel kernels serve the model).
std::string lower_backward(const pp_model_desc& d, Lowered* low);
// One line per warp and phase: the runs it executes (pp_describe.cpp).
std::string describe_schedule(const Lowered& L);

}  // namespace pp
