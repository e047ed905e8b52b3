// pp_model.h -- the opaque model handle behind the C ABI (host side).
#pragma once
#include <cstdint>
#include <vector>

#include "pp_common.h"
#include "pp_lowering.h"
#include "pp_profile.h"

// Kernel-time slots reported by pp_profile_get (milliseconds of the last batched call, summed over
// its chunks, measured with CUDA events on the model's stream).
enum pp_prof_slot {
    PP_PROF_VITERBI = 0,
    PP_PROF_TRACEBACK = 1,
    PP_PROF_FORWARD = 2,
    PP_PROF_BACKWARD = 3,
    PP_PROF_SAMPLE = 4,
    PP_PROF_H2D = 5,
    PP_PROF_D2H = 6,
    PP_PROF_SCHEDULE_HOST_MS = 7,   // host wall time spent bucketing events
    PP_PROF_SLOTS = 8
};

constexpr int kPipeSlots = 3;   // chunks in flight: copy-in of chunk c overlaps the sweeps of c - 1 and the traceback / copy-out of c - 2

struct pp_model {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    size_t smem_optin = 0;
    int warps = pp::kLaneWarps;

    // host copy of the baked description
    int32_t m = 0, ne = 0, n_edges = 0, start = 0, end = 0, finite = 0;
    std::vector<int32_t> in_ptr, in_src, out_ptr, out_dst, emit_kind;
    std::vector<double> in_logp, out_logp, emit_p0, emit_p1, emit_logw;

    pp::Lowered low;
    // linear-profile fast path (pp_profile.h): plan.ok says whether the register-resident kernels serve this model;
    // kernel_path (pp_model_set_kernel_path): 0 = automatic, 1 = run-based lane kernels only, 2 = like 1 and the exact
    // state-parallel kernel for forward-backward, 3 = like 0 and posteriors from k_profile_fb too
    pp::ProfilePlan plan;
    int kernel_path = 0;
    pp::DevBuf d_prof_rec_v, d_prof_rec_f, d_prof_end_v, d_prof_end_f, d_prof_unit_off, d_prof_unit_bytes, d_prof_unit_shift,
        d_prof_unit_mask, d_prof_slot_begin, d_prof_slot_src, d_state_pmask;
    // profile forward-backward (pp_profile_fb.cuh): its own plan -- narrower warp blocks (pp_profile_sched.h)
    pp::ProfilePlan plan_fb;
    pp::DevBuf d_pfb_rec_f, d_pfb_end_f, d_prof_rec_b, d_prof_fb_map, d_prof_fb_log2, d_prof_emit_state;
    // traceback lookup tables (pp_device_common.cuh TraceRec) of the two kernel families; n entries of the source tables
    pp::DevBuf d_trace_rec[2], d_trace_src[2];
    int32_t n_trace_src[2] = {0, 0};

    // device copy of the schedule
    pp::DevBuf d_edges_logp[2], d_edges_prob[2], d_emis, d_warp_run_ranges;
    std::vector<double> arena;            // host image of the constant arena: [runs | log p | p | emission records]
    int32_t arena_emis = 0;
    pp::DevBuf d_wtab_logp, d_wtab_prob, d_etab_vit, d_etab_fwd;
    pp::DevBuf d_state_edge_begin, d_low_src, d_state_pword, d_state_pshift, d_state_pmul;
    // device copy of the baked CSR (state-parallel kernels, sampler)
    pp::DevBuf d_in_ptr, d_in_src, d_in_logp, d_out_ptr, d_out_dst, d_out_logp, d_out_prob, d_alive;
    pp::DevBuf d_lev_ptr, d_lev_states, d_rlev_ptr, d_rlev_states;
    int32_t n_levels = 0, n_rlevels = 0;

    // per-call scratch
    pp::DevBuf d_ws;          // traceback pointer table / DP tables
    pp::DevBuf d_obs, d_offsets, d_order, d_group_len, d_group_row;
    pp::DevBuf d_score, d_end_state, d_end_ord, d_path_len, d_path_off, d_path, d_scan_tmp, d_logp;
    pp::DevBuf d_tab, d_xbuf, d_tab_off, d_list, d_counter, d_stats;
    // lane-per-event forward-backward (pp_lane_fb.cuh)
    int64_t fb_info[3] = {0, 0, 0};       // pp_forward_backward_info
    std::vector<double> arena_fb;         // constant-arena image: forward runs, then backward runs
    pp::DevBuf d_bw_wtab, d_acc_map, d_acc_log2, d_fb_acc, d_fb_scratch, d_fb_ctl, d_fb_minmax;
    size_t ws_limit = 0;
    // emission table of PP_EMIT_TABLE states (pp_model_set_emission_table): device pointer, rows, owned copy of a host table
    const double* etab = nullptr;
    int64_t etab_rows = 0;
    pp::DevBuf d_etab_own;
    pp::DevBuf d_gcol;        // per-CTA DP columns of the global-column lane kernels (models too large for shared memory)
    bool has_table_states = false;

    // chunk pipeline of the batched host/device calls (pp_api.cu): copy-in and copy-out streams beside the compute
    // stream, two slots of per-chunk buffers, pinned scalars the host polls, a pool of timing events
    cudaStream_t s_in = nullptr, s_out = nullptr;
    struct PipeSlot {
        pp::DevBuf obs, ws, path, order, group_len, group_row, list, counter, sched, path_tmp;
        cudaEvent_t in_done = nullptr, a_done = nullptr, tb_done = nullptr, out_done = nullptr, sch_done = nullptr;
    } ps[3];                              // kPipeSlots
    int64_t* h_pin = nullptr;             // [64] pinned: per slot s -- path totals [2s, 2s+1], exact-forward counter [8+s], schedule read-back [16+4s ..], scratch overflow [32+s]
    struct Span { cudaEvent_t a, b; int slot; };
    std::vector<cudaEvent_t> evpool;
    size_t evnext = 0;
    std::vector<Span> spans;

    cudaEvent_t ev[2] = {nullptr, nullptr};
    double prof[PP_PROF_SLOTS] = {0};
    int64_t launches = 0;
};

namespace pp {

// Length-bucketed schedule of events [e0, e1): 32 events per group, longest first.
struct Schedule {
    std::vector<int32_t> order;       // [n_groups * 32], -1 padded
    std::vector<int32_t> group_len;   // [n_groups]
    std::vector<int64_t> group_row;   // [n_groups] first pointer-table row (group_len + 1 rows each)
    int64_t total_rows = 0;
    int32_t n_groups = 0;
};
void build_schedule(const int64_t* offsets, int64_t e0, int64_t e1, Schedule* out);

int upload_model(pp_model* M);
int upload_state_model(pp_model* M);          // pp_aux.cu: level structure for the state-parallel kernels
void release_state_model(pp_model* M);
// Re-runs, with the exact log-space kernel, every event whose entry in d_logp is -inf or NaN.
// State-parallel batched Viterbi (pp_aux.cu) for models too large for the lane-per-event kernels.
int viterbi_batch_state(pp_model* M, const void* obs, int32_t obs_dtype, const int64_t* offsets, int64_t n_events,
                        int32_t mem, double* logp_out, int64_t* path_len_out, uint16_t* path_out, int64_t path_capacity,
                        int64_t* path_total);
int rerun_exact_forward(pp_model* M, const void* d_obs_base, int32_t obs_dtype, const int64_t* d_off,
                        const int64_t* h_off, int64_t n_events, double* d_logp);
// The same in two halves for the chunk pipeline: enqueue the collection of the flagged events of a chunk (list, counter
// and the pinned copy of the counter belong to pipeline slot `slot`), later -- after the caller has waited for the
// slot's a_done event -- run the exact kernel on them.
int exact_forward_collect(pp_model* M, int slot, const double* d_logp, int64_t n_events);
int exact_forward_run(pp_model* M, int slot, const void* d_obs_base, int32_t obs_dtype, const int64_t* d_off,
                      int64_t maxlen, double* d_logp);

}  // namespace pp

// Models with PP_EMIT_TABLE states need a table covering every observation row of the batch.
inline int pp_check_emission_table(const pp_model* M, int64_t rows_needed) {
    if (!M->has_table_states) return PP_OK;
    if (!M->etab || M->etab_rows < rows_needed) {
        pp::set_error("model has table-emission states: call pp_model_set_emission_table with >= %lld rows first (have %lld)",
                      (long long)rows_needed, (long long)(M->etab ? M->etab_rows : 0));
        return PP_ERR_ARG;
    }
    return PP_OK;
}
