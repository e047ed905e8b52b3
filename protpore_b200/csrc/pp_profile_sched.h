// pp_profile_sched.h -- kernel-parameter struct of the linear-profile kernels (host and device).
#pragma once
#include <cstdint>

#include "pp_profile.h"

namespace pp {

#ifndef PP_PROF_P
#define PP_PROF_P 4
#endif
#ifndef PP_PROF_NW
#define PP_PROF_NW 11
#endif
constexpr int kProfPositionsPerWarp = PP_PROF_P; // P: positions whose emitting values one warp keeps in registers
constexpr int kProfMaxPositionWarps = PP_PROF_NW; // + 1 chain warp: 12 warps = 384 threads, 2 CTAs per SM at 80 registers

struct ProfSched {
    const double* rec;                           // [Kp][32] position records of this sweep (log p | p)
    const ProfEndRec* end_list;                  // in-edges of the end state, reference visiting order
    int32_t end_n;
    int32_t K, Kp, Kr, NW;                       // positions, padded to whole warps, records (Kp + P), position warps
    int32_t row_bytes;                           // bytes per traceback pointer row
    int32_t end_state;
};

// forward-backward on the profile (pp_profile_fb.cuh): one CTA per SM (the statistics live in registers) with its own plan,
// so its geometry can differ from the sweeps'.  Measured on B200, 200k events: 4 positions per warp x 11 warps (168
// registers) 103 ms; 2 x 22 warps (80 registers, 27 % more instructions, the shared-memory pipe saturates) 122 ms.
#ifndef PP_PFB_P
#define PP_PFB_P 4
#endif
#ifndef PP_PFB_NW
#define PP_PFB_NW 11
#endif
constexpr int kPfbPositionsPerWarp = PP_PFB_P;
constexpr int kPfbMaxPositionWarps = PP_PFB_NW;
constexpr int kPfbThreads = (PP_PFB_NW + 1) * 32;
struct ProfFbSched {
    const double* rec_f;                 // [Kr][32] forward records (probabilities)
    const double* rec_b;                 // [Kr][40] backward records
    const ProfEndRec* end_list;          // in-edges of the end state (probabilities)
    const int32_t* emit_state;           // [Kp][3] state index of M_k, I_INS_k, I_BLIP_k (-1: none)
    int32_t end_n;
    int32_t K, Kp, Kr, NW;
    int32_t n_groups;
    int32_t row_slots;                   // NW * 5 P + 1 (the last slot holds S_r)
    int32_t n_acc_slots;                 // Kp * 4 + 1
    unsigned char* scratch;              // per CTA: (max_len + 1) rows of row_slots * 256 bytes
    size_t scratch_stride;
    double* acc;                         // per CTA: n_acc_slots * 32 doubles
    int32_t* counter;                    // work-stealing group counter (zeroed by the host)
    int32_t* fail;                       // [0] failed events, [1] reason of the first, [2] its event, [3] row; doubles at +7
    double* logp;                        // [n_events]
    double* minmax;                      // [2 * ne]
    double start_wm, start_ws;           // t(start -> M_0), t(start -> S_SKIP_0)
    double* post;                        // optional: log posteriors, event e row i state k at [(offsets[e] + i) * ne + k] (3304-3321)
    int32_t ne;
};

}  // namespace pp
