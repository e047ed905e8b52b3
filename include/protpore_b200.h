/*
 * protpore_b200.h -- C ABI of the B200-native profile-HMM scoring/decoding engine.
 *
 * This is the drop-in boundary for ONE hot path of mitenjain/protpore: scoring and decoding many
 * segmented nanopore events against a baked yahmm `Model`.  The reference has no stable native
 * interface for this path (its sweeps are Cython `cdef` methods on private memoryviews,
 * yahmm/yahmm.pyx:2029-2040), so every entry point below names the reference routine it replaces;
 * INTEGRATION.md shows the ctypes binding a maintainer adds to yahmm.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.  Nothing throws across the ABI.
 *   - Every call returns PP_OK (0) or a negative pp_status; pp_last_error() gives the message of
 *     the last failure on the calling thread.
 *   - The caller owns every input/output buffer.  `mem` says where ALL data buffers of that call
 *     live: PP_MEM_HOST (library stages H2D/D2H itself; pinned memory recommended) or
 *     PP_MEM_DEVICE (device pointers on the model's GPU; work is enqueued on the model's stream and
 *     the call returns after that stream is synchronised).  The model's stream is a non-blocking stream of its own
 *     (pp_model_stream): whatever produced a device input on ANOTHER stream must have completed -- or that stream
 *     must have been made to precede pp_model_stream with an event -- before the call; the Python layer synchronises
 *     torch's current stream for every device tensor it passes.
 *   - The library owns only the opaque model handle and its device workspace.  A handle is not
 *     thread-safe; distinct handles are.  One handle per GPU (one process per GPU under torchrun).
 *   - Events are passed CSR-style: `obs` holds the segment means of all events back to back
 *     (float32 or float64, see pp_dtype), `offsets[n_events + 1]` (int64) delimits them.  Every
 *     event must have at least one observation (the reference raises on empty sequences,
 *     yahmm.pyx:2836).
 *   - Impossible events (log-probability -inf) are reported in-band: logp = -inf, path_len = -1;
 *     the Python layer maps them to the reference's (-inf, None) / (None, None).
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     PP_ERR_CUDA.
 */
#ifndef PROTPORE_B200_H
#define PROTPORE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pp_status {
    PP_OK = 0,
    PP_ERR_ARG = -1,         /* bad argument (null pointer, empty event, inconsistent CSR ...) */
    PP_ERR_CUDA = -2,        /* CUDA runtime failure or no device */
    PP_ERR_NOMEM = -3,       /* batch needs more device workspace than the handle's limit */
    PP_ERR_UNSUPPORTED = -4, /* model shape outside what the kernels lower */
    PP_ERR_CAPACITY = -5     /* caller's path buffer too small; *path_total holds the needed size */
} pp_status;

typedef enum pp_mem { PP_MEM_HOST = 0, PP_MEM_DEVICE = 1 } pp_mem;
typedef enum pp_dtype { PP_F32 = 0, PP_F64 = 1 } pp_dtype;
typedef enum pp_emission_kind { PP_EMIT_NORMAL = 0, PP_EMIT_UNIFORM = 1, PP_EMIT_TABLE = 2 } pp_emission_kind;

/*
 * Baked model, exactly the arrays Model.bake produces (yahmm.pyx:2280-2655):
 * states [0, silent_start) emit, [silent_start, n_states) are silent in topological order.
 *   in_ptr/in_src/in_logp    = in_edge_count / in_transitions / in_transition_log_probabilities
 *   out_ptr/out_dst/out_logp = out_edge_count / out_transitions / out_transition_log_probabilities
 *   emit_kind/p0/p1          = distribution family and parameters of each emitting state
 *                              (Normal: mean, std -- std == 0 selects the point-mass branch,
 *                              yahmm.pyx:383-387; Uniform: start, end, yahmm.pyx:257-262)
 *                              PP_EMIT_TABLE: any other Distribution (yahmm.pyx:525-1917) -- its log-densities are
 *                              supplied per batch through pp_model_set_emission_table, p0/p1 unused
 *   emit_logw                = state_weights = log(state.weight) (yahmm.pyx:2515-2517)
 */
typedef struct pp_model_desc {
    int32_t n_states, silent_start, start_index, end_index, finite, n_edges;
    const int32_t *in_ptr, *in_src;
    const double *in_logp;
    const int32_t *out_ptr, *out_dst;
    const double *out_logp;
    const int32_t *emit_kind;
    const double *emit_p0, *emit_p1, *emit_logw;
} pp_model_desc;

typedef struct pp_model pp_model;

/* Build a device-resident, kernel-lowered copy of a baked model on CUDA device `device`.
 * Replaces the implicit "self." state every reference sweep reads (yahmm.pyx:2029-2040). */
int pp_model_create(const pp_model_desc *desc, int device, pp_model **out);
void pp_model_destroy(pp_model *model);

/* New parameters, same topology (after a Baum-Welch M-step, yahmm.pyx:4277-4327). */
int pp_model_update_params(pp_model *model, const pp_model_desc *desc);

/*
 * Emission table for states of kind PP_EMIT_TABLE: the reference fills e[i, k] = dist_k.log_probability(x_i) for
 * every observation before each sweep (yahmm.pyx:2836-2840); the caller does that for the families the kernels do not
 * evaluate themselves and hands the matrix over: table[r * silent_start + k] is the log-density of emitting state k
 * for observation r of the `obs` array of the batched calls that follow (log(state.weight) is added by the kernels;
 * columns of non-table states are ignored).  A host table is copied to the device; a device table must stay valid
 * until it is replaced.  NULL clears it.  Models with table states run on the exact state-parallel kernels.
 */
int pp_model_set_emission_table(pp_model *model, const double *table, int64_t n_rows, int32_t mem);

/* Text dump of the run schedule the kernels execute for this model (one line per warp and phase); needs no
 * GPU.  Returns the length of the full text (buf may be NULL to query it), < 0 on error. */
int64_t pp_schedule_describe(const pp_model_desc *desc, char *buf, int64_t cap);

/*
 * Kernel family of the batched forward / Viterbi sweeps.  A model whose baked graph is a linear profile -- what
 * HMM_linear_model builds (peptide_hmm_v0.0.1.py:63-243): match / blip / insert / drop-off / skip / back-slip states per
 * position, recognised from the CSR alone -- runs on register-resident kernels (pp_profile_kernels.cuh); every other
 * lane-lowered model on the run-based kernels (pp_lane_kernels.cuh).  Both follow the reference's arithmetic candidate by
 * candidate and give identical results; `path` 1 pins a handle to the run-based kernels (tests, comparisons), 2 does the
 * same and also keeps pp_forward_backward_batch on the exact state-parallel kernel, 3 is 0 plus posteriors from the batched
 * profile kernel (see pp_forward_backward_batch), 0 restores the automatic choice.  The environment variable PP_KERNEL_PATH sets the initial value.
 * pp_model_kernel_info: out4 = {family: 2 profile, 1 run-based lane kernels, 0 state-parallel; positions of the profile
 * (0 if none); warps per CTA; bytes of one traceback pointer row (32 events)}.
 */
int pp_model_set_kernel_path(pp_model *model, int32_t path);
int pp_model_kernel_info(const pp_model *model, int32_t *out4);

/* Upper limit for the handle's device workspace in bytes (default: 70% of free memory at creation). */
int pp_model_set_workspace_limit(pp_model *model, int64_t bytes);

/* Device workspace pp_viterbi_batch needs for these events (offsets on the HOST). Lets a host
 * caller split a job into chunks that fit. */
int64_t pp_viterbi_workspace_bytes(const pp_model *model, const int64_t *offsets_host, int64_t n_events);

/*
 * Batched Model.log_probability (yahmm.pyx:3366-3396 over Model._forward, 2810-2930):
 * logp_out[e] = log P(event e | model), float64, -inf if impossible.
 * Computed by the scaled linear-space float64 column kernel with exact power-of-two rescaling
 * (DESIGN.md "forward"; emission factors through MUFU.EX2, rel. error ~2^-22 each); measured worst error
 * < 1e-7 relative against the float64 reference, inside the 1e-3 abs / 1e-6 rel bar.
 */
int pp_forward_batch(pp_model *model, const void *obs, int32_t obs_dtype, const int64_t *offsets,
                     int64_t n_events, int32_t mem, double *logp_out);

/*
 * Batched Model.viterbi (yahmm.pyx:3444-3652): float64 max-plus sweep with the reference's exact
 * operation order and tie-breaking, packed traceback pointers in HBM, traceback on the GPU.
 *   logp_out[e]     : Viterbi score (bit-identical to the reference), -inf if impossible
 *   path_len_out[e] : number of states on the path (start ... end, silent states included),
 *                     -1 if impossible (reference returns (-inf, None), 3621-3623)
 *   path_out        : state indices of all paths back to back in event order, uint16
 *                     (n_states <= 65535); event e occupies [sum_{j<e} max(len_j,0), +len_e)
 *   path_capacity   : number of uint16 elements path_out can hold
 *   path_total      : receives the total number of elements (written or needed)
 * path_out may be NULL (scores and lengths only; path_total still reported).
 */
int pp_viterbi_batch(pp_model *model, const void *obs, int32_t obs_dtype, const int64_t *offsets,
                     int64_t n_events, int32_t mem, double *logp_out, int64_t *path_len_out,
                     uint16_t *path_out, int64_t path_capacity, int64_t *path_total);

/*
 * The pair the peptide script runs per event -- Model.viterbi to decode, Model.log_probability to score
 * (peptide_hmm_v0.0.1.py:251-262) -- for a whole batch in ONE call: the observations are uploaded once, the events are
 * bucketed by length once (on the GPU), then both sweeps and the traceback run back to back on the same buffers.  Outputs as
 * for pp_viterbi_batch, plus forward_logp_out[e] = log P(event e | model) as for pp_forward_batch.  Results are identical
 * to the two single calls.
 */
int pp_score_batch(pp_model *model, const void *obs, int32_t obs_dtype, const int64_t *offsets, int64_t n_events,
                   int32_t mem, double *viterbi_logp_out, int64_t *path_len_out, uint16_t *path_out,
                   int64_t path_capacity, int64_t *path_total, double *forward_logp_out);

/*
 * Full DP tables of ONE event in float64 log space, the reference's return values:
 *   pp_forward_table  -> Model.forward  (yahmm.pyx:2780-2930), table_out[(n+1) x n_states]
 *   pp_backward_table -> Model.backward (yahmm.pyx:2932-3156), table_out[(n+1) x n_states]
 * obs is float64[n]; buffers follow `mem`.
 */
int pp_forward_table(pp_model *model, const double *obs, int64_t n, int32_t mem, double *table_out);
int pp_backward_table(pp_model *model, const double *obs, int64_t n, int32_t mem, double *table_out);

/*
 * Model.forward_backward for a batch (yahmm.pyx:3158-3364, tie == False) and the E-step of
 * Model._train_once_baum_welch (4107-4236) in sufficient-statistic form:
 *   logp_out[e]            : sequence log-probability (-inf => event skipped, contributes nothing)
 *   edge_counts[n_edges]   : += expected transition counts, indexed like the OUT-edge CSR
 *   emit_moments[3 * silent_start] : += (sum w, sum w x, sum w x^2) per emitting state
 *   emit_minmax[2 * silent_start]  : min / max over whole events that gave the state non-zero weight
 *   posteriors_out         : optional (may be NULL): log emission weights, float64, event e row i
 *                            state k at [(offsets[e] + i) * silent_start + k]  (3304-3321)
 * Accumulators are read-modify-write so a caller can sum shards before an allreduce.
 * Without posteriors_out linear profiles run on k_profile_fb (pp_profile_fb.cuh), other lane-lowered models on k_fb_lane
 * (both: scaled linear-space forward and backward sweeps, float64, float64-accurate emissions; statistics within ~1e-11
 * relative of the exact log-space kernel).  Events they cannot resolve, other models, and calls that ask for posteriors
 * take the exact state-parallel log-space kernel, which keeps a finite logarithm for posteriors far below the double
 * range like the reference does.  A handle on kernel path 3 gets the posteriors of a linear profile from k_profile_fb
 * instead (batched): within 1e-10 of the exact kernel wherever the posterior exceeds e^-250 (a scaled linear-space sweep
 * spans ~2^1500 below its row maximum); smaller ones lose precision or come out as -inf.  The statistics are not affected.
 */
int pp_forward_backward_batch(pp_model *model, const void *obs, int32_t obs_dtype, const int64_t *offsets,
                              int64_t n_events, int32_t mem, double *logp_out, double *edge_counts,
                              double *emit_moments, double *emit_minmax, double *posteriors_out);

/* How the LAST pp_forward_backward_batch call on this handle was served: out[0] = 1 lane-per-event kernel
 * (pp_lane_fb.cuh), 0 exact state-parallel kernel only, 2 lane kernel failed its numeric checks and the batch was
 * redone by the exact kernel, 3 linear-profile kernel (pp_profile_fb.cuh), 4 that kernel failed its checks; out[1] = events the lane kernel handed to the exact kernel (scaled sweep could not
 * resolve them); out[2] = events whose backward sweep failed the checks, or -(number of accumulated statistics that
 * were not finite) when that is what sent the batch to the exact kernel. */
int pp_forward_backward_info(const pp_model *model, int64_t *out3);

/*
 * Length-conditioned ancestral sampler used to make synthetic events:
 * event e gets exactly lengths[e] observations; edges into `end` are masked (and the remaining
 * out-probabilities renormalised) until the last emission.  Deterministic in (seed, e).
 * obs_out is float32 on the device or host per `mem`; offsets = exclusive scan of lengths.
 */
int pp_sample_events(pp_model *model, const int64_t *offsets, int64_t n_events, uint64_t seed, int32_t mem,
                     float *obs_out);

/*
 * ---- Segmentation: raw current -> segments -> segment means (the step before the HMM path) ----------------
 * Batched FastStatSplit.parse (PyPore/cparsers.pyx:103-203; the class SpeedyStatSplit wraps it,
 * PyPore/parsers.py:505-534; called per event at peptide_hmm_v0.0.1.py:443-446).  Model-independent: these
 * entry points take a CUDA device ordinal instead of a model handle.
 *
 * pp_statsplit_min_gain: the gain threshold FastStatSplit.__init__ derives (cparsers.pyx:56-101); pass 0 for an
 * argument the Python caller left at None.  Pure host arithmetic.
 */
typedef struct pp_statsplit_params {
    int64_t min_width, max_width, window_width;   /* samples; the reference's defaults are 100, 1000000, 10000 */
    double min_gain;                              /* threshold on the log-variance gain (already doubled, 101) */
} pp_statsplit_params;

double pp_statsplit_min_gain(int64_t window_width, double min_gain_per_sample, double false_positive_rate,
                             double prior_segments_per_second, double sampling_freq, double cutoff_freq);

/*
 * current[offsets[e] .. offsets[e+1]) is the float64 current of event e.  Outputs, CSR-style like the HMM calls:
 *   seg_offsets[n_events + 1] : segments of event e are [seg_offsets[e], seg_offsets[e+1])  -- this array and
 *                               seg_mean / seg_mean_f32 are exactly the `offsets` / `obs` of pp_viterbi_batch
 *   seg_start[j]              : first sample of segment j, relative to its event (0 for an event's first segment;
 *                               the others are the reference's breakpoints, ascending)
 *   seg_mean[j], seg_std[j]   : Segment.mean / Segment.std (PyPore/core.py:209-214), float64; each may be NULL
 *   seg_mean_f32[j]           : the mean rounded to float32 (the batched HMM calls' input type); may be NULL
 *   capacity                  : entries the seg_* arrays can hold;  *seg_total receives the number needed
 * seg_start == NULL runs the counting pass only (seg_offsets and *seg_total are filled).  PP_ERR_CAPACITY if the
 * arrays are too small.  Breakpoints are bit-identical to the reference's except where two candidate gains of one
 * window agree to ~1e-13 relative (CUDA's float64 log is accurate to 1 ulp, glibc's is not the same function).
 */
int pp_statsplit_batch(const pp_statsplit_params *prm, int device, const double *current, const int64_t *offsets,
                       int64_t n_events, int32_t mem, int64_t *seg_offsets, int32_t *seg_start, double *seg_mean,
                       double *seg_std, float *seg_mean_f32, int64_t capacity, int64_t *seg_total);

/* Device time of the last pp_statsplit_batch on `device`: out4 = {k_statsplit ms, k_segments ms, kernels launched
 * since the library was loaded, k_cumsum ms}. */
int pp_statsplit_profile(int device, double *out4);

/*
 * ---- Segment alignment ----------------------------------------------------------------------------
 * Batched cSegmentAligner.align (PyPore/calignment.pyx:20-101; the class SegmentAligner wraps it, PyPore/alignment.py:
 * 26-47): the Karplus semi-local alignment of a sequence of (mean, std, duration) segments to a model, with a penalty
 * per unit of model duration skipped or slipped back over.  Model-independent of the HMM handle: takes a device ordinal.
 *   model            : n >= 2 segments (means, stds, durs) and the two penalties
 *   seq_*[offsets[e] .. offsets[e+1]) : the segments of sequence e (float64), CSR-style like the HMM calls
 *   score_out[e]     : score[s-1, n-1] / numpy.sum(seq_durs)   (float64, bit-identical to the reference)
 *   order_out        : for every segment the model position it is aligned to (the reference's `backtrace` array,
 *                      which it returns as float64), at the segment's own offset
 *   status_out[e]    : 0, or a mask of the two cases in which the reference itself is undefined and this library
 *                      defines the result: bit 0 -- no entry of the last score row exceeds -1, so double_argmax
 *                      (calignment.pyx:11-18) returns an uninitialised int; defined as the first maximum of the row.
 *                      bit 1 -- the backtrace stands at model position 0 in a row >= 1, where the reference reads
 *                      score[i-1, j-1] with an unsigned j = 0 and raises IndexError; defined by skipping that test.
 */
typedef struct pp_segment_model {
    int32_t n;
    const double *means, *stds, *durs;            /* host pointers, n each */
    double skip_penalty, backslip_penalty;
} pp_segment_model;

int pp_segment_align_batch(const pp_segment_model *model, int device, const double *seq_means, const double *seq_stds,
                           const double *seq_durs, const int64_t *offsets, int64_t n_seqs, int32_t mem,
                           double *score_out, int32_t *order_out, int32_t *status_out);

/* The LAST pp_segment_align_batch call of this process: out4 = {device time of its kernel launches in ms (CUDA events),
 * launches, bytes of table workspace, groups of 32 sequences}. */
int pp_segment_align_profile(double *out4);
/* pp_segment_align_batch keeps its device buffers (the three DP tables above all) between calls, per device; this frees
 * them.  PP_SEGALIGN_CACHE=0 in the environment makes every call free them itself. */
int pp_segment_align_trim(int device);

/* Measured float64 issue peaks of `device` (a few milliseconds of independent DADD / DFMA chains): out4 = {float64 adds
 * per second, float64 FMAs per second, SM count, SM clock in Hz}.  bench.py divides the sweeps' operation rates by these. */
int pp_device_peaks(int device, double *out4);

/* Number of usable CUDA devices (0 without a driver/GPU; never fails). */
int pp_device_count(void);

/* Pinned host memory for PP_MEM_HOST buffers (cudaHostAlloc / cudaFreeHost); NULL on failure. */
void *pp_host_alloc(int64_t bytes);
void pp_host_free(void *p);

/* Device time of the LAST batched call on this handle, in milliseconds, measured with CUDA events on
 * the handle's stream and summed over the call's chunks.  Slots: 0 Viterbi sweep, 1 traceback (+scan),
 * 2 forward sweep, 3 forward-backward, 4 sampler, 5 host->device copies, 6 device->host copies,
 * 7 host wall time spent bucketing events by length.  Writes min(n, 8) values. */
int pp_profile_get(const pp_model *model, double *out, int32_t n);

/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
int64_t pp_launch_count(const pp_model *model);

/* Guard mode (environment PP_GUARD=1 when the library is loaded): every device buffer of the library carries 4 KiB of a
 * known pattern in front of and behind its payload.  pp_debug_check_guards synchronises the device and returns the number
 * of buffers whose guards were overwritten (0 = clean; -1 guard mode off; -2 CUDA error), pp_last_error names the first;
 * pp_debug_guard_count = buffers currently guarded.  Stands in for compute-sanitizer's memcheck in the GPU test-suite. */
int pp_debug_check_guards(void);
int pp_debug_guard_count(void);
/* Writes one byte behind and eight bytes in front of a scratch buffer and checks that both are reported: PP_OK if so. */
int pp_debug_guard_selftest(void);

/* CUDA stream the handle enqueues on (a cudaStream_t as an integer), for timing with CUDA events. */
uint64_t pp_model_stream(const pp_model *model);

/* Message of the last failure on this thread ("" if none). */
const char *pp_last_error(void);

/* ABI version, bumped on any incompatible change. */
int pp_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PROTPORE_B200_H */
