// pp_lowering.h -- host-side lowering of a baked yahmm model into the column-kernel schedule.
//
// The baked CSR (Model.bake, yahmm.pyx:2280-2655) is turned into a static per-warp program:
//   * every state becomes one ITEM (destination slot + a run of in-edge records);
//   * emitting items are dealt to the W warps of a CTA by greedy longest-processing-time on their
//     in-degree; all lanes of a warp work on the SAME item for 32 (or 64) different events, so the
//     edge records are warp-uniform and column accesses are stride-1 (DESIGN.md "layout");
//   * silent states are split into the connected components of the silent->silent sub-DAG (the
//     delete/skip chain, the back-slip chain, the isolated drop-off states ...).  A component goes
//     to ONE warp, in topological (= baked index) order, so the serial chain inside a column needs
//     no barrier and the value just produced is forwarded in registers;
//   * per item the in-edges are re-ordered exactly as the reference visits them: for an emitting
//     state all in-edges in CSR order (3545-3562); for a silent state first the emitting sources
//     (pass 1, 3564-3584) then the silent sources with index < l (pass 2, 3586-3605; edges from
//     silent sources >= l are ignored by the reference, 3593).  The traceback pointer stored per
//     cell is the ordinal in that list, so ties break identically;
//   * a finite model's end state with no out-edges is evaluated once, at each event's last row.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/protpore_b200.h"

namespace pp {

// ---- records shared by host and device ---------------------------------------------------
struct __attribute__((aligned(16))) EdgeRecV {   // Viterbi (fp64 max-plus)
    double logp;
    uint16_t slot[2];                            // source column slot for parity 0 / 1
    uint32_t pad;
};
struct __attribute__((aligned(8))) EdgeRecF {    // forward / backward (fp32 linear, scaled)
    float w;                                     // exp(logp)
    uint16_t slot[2];
};
struct __attribute__((aligned(16))) ItemRec {
    int32_t edge_begin;                          // first lowered edge record
    int16_t n_edges;
    int16_t emis;                                // emitting state index, -1 for silent items
    uint16_t dst_slot[2];                        // destination column slot for parity 0 / 1
    int32_t state;                               // baked state index (traceback pointer row slot)
    int32_t flags;                               // bit0: this is the start state
};
enum : int32_t { EMIS_NORMAL = 0, EMIS_POINT = 1, EMIS_UNIFORM = 2, EMIS_HAS_LOGW = 8 };
struct __attribute__((aligned(16))) Emis64 {     // exact float64 emission (Viterbi, tables)
    double a, b, c, logw;                        // Normal: mu, 2*sigma^2, log(1/(sigma*SQRT_2_PI));
    int32_t kind;                                // Point: mu; Uniform: lo, hi, log(1/(hi-lo))
    int32_t pad[3];
};
struct __attribute__((aligned(16))) EmisF {      // fp32 log2-domain emission (forward / backward)
    float mu_hi, mu_lo, k2, c2;                  // e2(x) = c2 - ((x-mu_hi)-mu_lo)^2 * k2
    float lo, hi;                                // Uniform support rounded inwards to float32
    int32_t kind;
    int32_t pad;
};

struct Lowered {
    int32_t m = 0, ne = 0, ns = 0, start = 0, end = 0, finite = 0, n_edges = 0;
    bool end_final_only = false;                 // end evaluated only at each event's last row
    int32_t warps = 8;                           // warps per CTA the schedule was built for
    int32_t n_slots = 0;                         // ns + 2*ne column slots
    int32_t max_item_edges = 0;
    std::vector<ItemRec> items;                  // warp-major: [emit items w0 | silent items w0 | emit w1 ...]
    std::vector<int32_t> warp_emit_begin, warp_emit_end, warp_sil_begin, warp_sil_end;
    std::vector<EdgeRecV> edges_v;
    std::vector<EdgeRecF> edges_f;
    std::vector<int32_t> low_src;                // source state of each lowered edge
    std::vector<int32_t> low_csr;                // index of each lowered edge in the IN-edge CSR
    std::vector<int32_t> state_edge_begin;       // per state: first lowered edge (traceback lookup)
    std::vector<int32_t> state_n_edges;
    int32_t end_edge_begin = 0, end_n_edges = 0; // lowered edge run of the end state
    std::vector<Emis64> emis64;
    std::vector<EmisF> emisf;
};

// Returns "" on success, otherwise why the model cannot be lowered.
std::string lower_model(const pp_model_desc& d, int warps, Lowered* out);
// Refresh weights / emission parameters of an existing lowering (same topology).
std::string refresh_params(const pp_model_desc& d, Lowered* low);

}  // namespace pp
