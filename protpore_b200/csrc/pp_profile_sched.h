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

}  // namespace pp
