// pp_profile.h -- the "linear profile" fast path: recognition of the topology HMM_linear_model builds
// (peptide_hmm_v0.0.1.py:63-243) in a baked CSR, and the per-position tables its kernels (pp_profile_kernels.cuh) read.
//
// The lane-per-event kernels (pp_lane_kernels.cuh) keep the whole float64 column of 32 events in shared memory and
// interpret a run schedule; every candidate costs two shared-memory loads and its address arithmetic.  A profile HMM
// is a chain of identical POSITIONS: per position k a match state M_k, a blip state I_BLIP_k, an insert state I_INS_k
// (emitting) and a drop-off S_DROPOFF_k, a skip S_SKIP_k and a back-slip B_BACK_SHORT_k state (silent), wired to
// positions k - 1, k, k + 1 only.  The profile kernels give every warp P consecutive positions and keep THEIR emitting
// values in registers from row to row; only what crosses a warp boundary (M_k) and the two silent chains go through
// shared memory.  The arithmetic per state is the reference's, candidate by candidate, in the reference's visiting
// order, so scores and paths stay bit-identical (the same tests cover both kernel families).
//
// Nothing here trusts names: `build_profile_plan` classifies the states from the graph alone, then checks EVERY in-edge
// list of the lowered model (reference visiting order) against the template; one unexplained edge, or an order the
// template cannot express, and the model simply keeps the run-based kernels.
//
// Template, per position k (slot order = order of the candidates in the kernels' first-maximum trees):
//   M_k    : [0] M_{k-1}  [1] S_SKIP_{k-1} (k = 0: the start state)  [2] M_k  [3] S_DROPOFF_k  [4] I_BLIP_k
//            [5] I_INS_k  [6] M_{k+1} (only where bake merged a back-slip state away)  [7] B_BACK_{k+1} (or an alias)
//   I_INS_k: [0] M_{k-1}  [1] I_INS_k           I_BLIP_k: [0] M_k  [1] I_BLIP_k          S_DROPOFF_k: [0] M_k
//   S_SKIP_k: [0] M_{k-1}  [1] S_SKIP_{k-1} (k = 0: start)        B_BACK_k: [0] M_k  [1] B_BACK_{k+1}
//   end    : any list of M / S_DROPOFF / S_SKIP / B_BACK states, evaluated once per event at its last row.
// A slot without an edge carries weight -inf (Viterbi) / 0 (forward): such a candidate can never win against a finite
// one, and a cell whose candidates are all -inf is never on a traced path.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "pp_lowering.h"

namespace pp {

constexpr int kProfRecDoubles = 32;              // one position record: 256 bytes
// offsets (in doubles) inside a position record
enum : int {
    PR_WM = 0,      // [8] weights of M_k's slots
    PR_WI = 8,      // [2] I_INS_k
    PR_WB = 10,     // [2] I_BLIP_k
    PR_WD = 12,     // M_k -> S_DROPOFF_k
    PR_KINDS = 13,  // int32 x 2: {kind M | kind INS << 8 | kind BLIP << 16 | wide << 24, unused}
    PR_SK = 14,     // [2] M_k -> S_SKIP_{k+1} (kept with its source), S_SKIP_{k-1} -> S_SKIP_k
    PR_EM = 16,     // [4] emission record of M_k
    PR_EI = 20,     // [4] I_INS_k
    PR_EB = 24,     // [4] I_BLIP_k
    PR_BK = 28,     // [2] B_BACK_k: from M_k, from B_BACK_{k+1}
    // forward records only, first position of a warp's block: products of the chain weights of the block
    PR_CUP = 30,    //   prod_j t(S_SKIP_{k0+j-1} -> S_SKIP_{k0+j}), j = 0 .. P-1
    PR_CDN = 31,    //   prod_j t(B_BACK_{k0+j+1} -> B_BACK_{k0+j})
    PR_CUT = 19,    //   PR_CUP without j = 0 (the fourth constant of M's emission record is unused in linear space)
};
// Emission records.  Viterbi (log domain, exact):   Normal {mu, 2 sigma^2, RN(1 / (2 sigma^2)), log(1/(sigma sqrt(2 pi)))}
//                                                    Point  {mu, -, -, -}      Uniform {lo, hi, -, log(1/(hi - lo))}
//                    forward (linear, scaled):       Normal {mu, log2(e)/(2 sigma^2), log2 of the peak density, -}
//                                                    Point  {mu, -, -, 1}      Uniform {lo, hi, -, 1/(hi - lo)}
enum : int { PROF_M = 0, PROF_DROP = 1, PROF_SKIP = 2, PROF_BACK = 3 };
struct __attribute__((aligned(16))) ProfEndRec {
    int32_t cls, k;                              // source state of this in-edge of `end`: class and position
    double w;                                    // log p (Viterbi) or p (forward)
};

// ---- forward-backward (Baum-Welch E-step) on the profile, pp_profile_fb.cuh -------------------------------------------
// The backward sweep reads the SAME weights by their source: one record per position with everything that leaves it.
constexpr int kFbRecDoubles = 40;
enum : int {
    FR_WM = 0,      // [8] out of M_k: -> M_{k+1}, M_k, I_INS_{k+1}, I_BLIP_k, S_DROPOFF_k, M_{k-1}, S_SKIP_{k+1}, B_BACK_k
    FR_WE = 8,      // [4] into the end state from M_k, S_DROPOFF_k, S_SKIP_k, B_BACK_k
    FR_WS = 12,     // [2] out of S_SKIP_k: -> M_{k+1}, S_SKIP_{k+1}
    FR_WK = 14,     // [2] out of B_BACK_k: -> M_{k-1}, B_BACK_{k-1}
    FR_WD = 16,     // [2] out of S_DROPOFF_k: -> M_k, (unused)
    FR_WI = 18,     // [2] out of I_INS_k: -> M_k, I_INS_k
    FR_WB = 20,     // [2] out of I_BLIP_k: -> M_k, I_BLIP_k
    FR_EM = 22,     // [4] linear-space emission records (as in the forward records)
    FR_EI = 26,
    FR_EB = 30,
    FR_KINDS = 34,  // int32: kind M | kind INS << 8 | kind BLIP << 16
    // first position of a warp's block:
    FR_CDN = 36,    //   prod_j t(S_SKIP_{k0+j} -> S_SKIP_{k0+j+1}): the descending skip chain across the block
    FR_CUP = 37,    //   prod_j t(B_BACK_{k0+j} -> B_BACK_{k0+j-1}): the ascending back-slip chain
    FR_TDN = 38,    //   what H[M_{k0}] contributes to the LEFT block's way out: t(S_SKIP_{k0-1} -> M_{k0}) prod_{j<P-1} t(S_SKIP_{k0-P+j} -> ..)
    FR_TUP = 39,    //   what H[M_{k0+P-1}] contributes to the RIGHT block's: t(B_BACK_{k0+P} -> M_{k0+P-1}) prod_{j>=1} t(B_BACK_{k0+P+j} -> ..)
};
// Statistic terms per position: 4 chunks of 8 (pp_profile_fb.cuh lists them); after the in-warp reduction lane L of a
// chunk holds a partial sum of term L >> 2.
constexpr int kFbChunks = 4;
constexpr int kFbTermsPerPosition = 8 * kFbChunks;

constexpr int kProfMaxAlias = 4;
struct ProfilePlan {
    bool ok = false;
    std::string why;                             // why not, when !ok
    int K = 0;                                   // positions
    int P = 3;                                   // positions per warp the tables are padded for
    int NW = 0;                                  // position warps = ceil(K / P);  the CTA has NW + 1 warps
    int Kp = 0;                                  // NW * P (records of padding positions hold -inf / 0 weights)
    int Kr = 0;                                  // records in rec_v / rec_f: Kp + P
    std::vector<double> rec_v, rec_f;            // [Kp][32]
    std::vector<ProfEndRec> end_v, end_f;        // in-edges of `end` in the reference's visiting order
    int n_alias = 0;                             // B_BACK slots that mirror another one (M_0 reads B_BACK_2 in pro_001)
    int alias_dst[kProfMaxAlias] = {0, 0, 0, 0}, alias_src[kProfMaxAlias] = {0, 0, 0, 0};
    // traceback tables: where each state's ordinal sits in a pointer row, and which state each slot leads to
    int32_t row_bytes = 0;                       // bytes of one pointer row (all 32 lanes)
    std::vector<int32_t> unit_off, unit_bytes, unit_shift, unit_mask;   // [m]
    std::vector<int32_t> slot_begin, slot_src;   // [m], [sum of slots]
    // forward-backward on the profile (fb_ok): backward records, and which statistic every accumulation lane feeds
    bool fb_ok = false;
    std::string fb_why;
    std::vector<double> rec_b;                   // [Kr][kFbRecDoubles]
    std::vector<int32_t> fb_map;                 // [fb_slots * 32]: out-edge index, n_edges + 3 * state + {0, 1, 2}, or -1
    std::vector<int32_t> fb_log2;                // [fb_slots]: log2 of the terms per slot (3; the start slot too)
    int fb_slots = 0;                            // Kp * kFbChunks + 1 (the last one: the two out-edges of the start state)
    std::vector<int32_t> fb_emit_state;          // [Kp][3]: state index of M_k, I_INS_k, I_BLIP_k or -1
    double start_wm = 0.0, start_ws = 0.0;       // t(start -> M_0), t(start -> S_SKIP_0)
    // classification (tests, describe)
    std::vector<int32_t> M, BLIP, INS, DROP, SKIP, BACK;               // [K] state index or -1
};

// Recognises the linear-profile topology in the lowered model; plan->ok tells whether the fast path applies.
void build_profile_plan(const pp_model_desc& d, const Lowered& low, int positions_per_warp, int max_position_warps,
                        ProfilePlan* plan);

}  // namespace pp
