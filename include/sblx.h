/* sblx.h - C ABI of libsblx.so: the B200-native batched rectangular Newton-Raphson AC
 * power-flow engine (scenario batches of ONE grid: N-1/N-k branch outages and injection sweeps).
 *
 * This is the drop-in boundary for Sparlectra's data-parallel hot path.  Every entry point
 * names the reference interface it replaces (paths relative to the Sparlectra.jl tree):
 *
 *   sblx_create            <- buildPfModel(net) -> PFModel        src/solver_interface.jl:193-295
 *                             (Ybus, busType, slack_idx, Vset, Sspec, V0, qmin_pu/qmax_pu) plus the
 *                             branch two-port table createYBUS stamps (src/equicircuit.jl:583-645,
 *                             src/branch.jl:297-310), build_qload_pu (src/solver_core.jl:302-313) and
 *                             the runpf! keyword defaults that matter
 *                             (src/powerflow_rectangular/rectangular_network_solver.jl:1783-1883)
 *   sblx_solve_outages     <- the per-case loop of runContingencies! / _run_one_contingency
 *                             src/contingency.jl:183-223,257-319  (removeBranch! + markIsolatedBuses! +
 *                             detect_ac_islands + runpf! + calcNetLosses! + _contingency_metrics)
 *   sblx_solve_injections  <- the serial load-sweep loop of
 *                             examples/powerflow/mc_probabilistic_powerflow.jl:73-99 (runpf! per draw)
 *   sblx_solve_one         <- solvePf(::AbstractExternalSolver, ::PFModel) -> PFSolution
 *                             src/solver_interface.jl:143-145 (a batch of one, no outage)
 *   sblx_solve_injection_sweep / _scale
 *                          <- the Monte-Carlo draw of the same example (mc_probabilistic_powerflow.jl:28-41:
 *                             truncated-normal factor per load) and buildComplexSVec's scaling
 *                             (src/network.jl:2359-2402), generated on the device
 *   sblx_comm_* / sblx_create_multi / sblx_multi_solve_outages
 *                          <- the chunked fan-out of runContingencies! over workers with ordered results
 *                             (src/contingency.jl:299-318): contiguous scenario shards over GPUs and ONE
 *                             ncclAllGather of the per-scenario records inside the call
 *   sblx_results fields    <- ContingencyResult (src/contingency.jl:75-86) + PFSolution
 *                             (src/solver_interface.jl:76-82) + the rejection reasons of
 *                             rectangular_network_solver.jl:702,783,902,951,1034 and
 *                             rectangular_final_status.jl:74
 *
 * Conventions: plain pointers and sizes only; complex numbers are interleaved (re, im) doubles
 * (binary compatible with Julia ComplexF64); indices follow `opts.index_base` (1 for Julia).  The
 * library copies inputs during the call and keeps no host pointer; outputs go to caller-allocated
 * arrays (NULL = not wanted).  Every call returns 0 on success or a negative error code with a
 * message available from sblx_last_error(); a scenario that fails numerically is DATA
 * (status[], converged[]), never an error return.  One handle may be used by one thread at a
 * time.  There is no CPU fallback: without a CUDA device every solve entry point fails.
 */
#ifndef SBLX_H
#define SBLX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBLX_VERSION 200 /* 0.2.0 */

/* bus types (busType::Vector{Symbol} of PFModel) */
enum { SBLX_BUS_ISOLATED = 0, SBLX_BUS_PQ = 1, SBLX_BUS_PV = 2, SBLX_BUS_SLACK = 3 };

/* per-scenario status = rejection reason of the reference */
enum {
  SBLX_ST_NONE = 0,                    /* converged and accepted */
  SBLX_ST_NOT_CONVERGED = 1,           /* :nr_mismatch_not_converged (maxiter exhausted) */
  SBLX_ST_NONFINITE = 2,               /* :nr_nonfinite */
  SBLX_ST_SINGULAR = 3,                /* :singular_newton_step */
  SBLX_ST_REMAINING_PV_Q = 4,          /* :remaining_pv_q_limit_violations */
  SBLX_ST_ACTIVE_PV_VRES = 5,          /* :active_pv_voltage_residual */
  SBLX_ST_MAX_SWITCHING = 6,           /* :max_switching_exceeded */
  SBLX_ST_ISLANDED_NO_REF = 7,         /* "islanded without reference" (contingency.jl:218-220) */
  SBLX_ST_BAD_BRANCH = 8,              /* unknown / out-of-range branch in the case */
  SBLX_ST_ISLANDED_UNSUPPORTED = 9     /* reserved, no longer produced: grids that split into several islands with
                                          a reference each are solved island-wise like the reference does */
};

/* error codes */
enum {
  SBLX_OK = 0,
  SBLX_E_ARG = -1,
  SBLX_E_CUDA = -2,
  SBLX_E_NOMEM = -3,
  SBLX_E_INTERNAL = -4,
  SBLX_E_NODEVICE = -5
};

typedef struct sblx_handle sblx_handle; /* opaque */

/* Mirrors the runpf! keywords the contingency path uses (defaults in brackets). */
typedef struct sblx_options {
  int32_t struct_size;          /* sizeof(sblx_options), for ABI evolution */
  int32_t index_base;           /* 0 or 1 [1] */
  int32_t max_iter;             /* maxIte [30] */
  int32_t qlimits_enabled;      /* [1] */
  int32_t qlimit_start_iter;    /* [2] */
  int32_t cooldown_iters;       /* net.cooldown_iters [0] */
  int32_t guard_max_switches;   /* qlimit_guard_max_switches [10] */
  int32_t freeze_after_repeated_switching; /* [1] */
  int32_t device;               /* CUDA device ordinal, -1 = current [-1] */
  int32_t max_slots;            /* scenario slots resident on the GPU, 0 = auto from free memory [0] */
  int32_t ordering;             /* 0 auto (cheaper of the two by fill + a per-level launch charge), 1 minimum degree,
                                   2 nested dissection [0] */
  int32_t relax_small;          /* merge a child front when the merged front has <= this many buses and its pivot
                                   block stays <= 4 buses (the leanest kernel class) [18] */
  int32_t factor_mode;          /* 0 = one fused CTA per front (k_front, default), 1 = experimental split kernels per level
                                   (k_pivot / k_subst / k_schur; measured slower: profiles/README.md) [0] */
  int32_t reserved1;
  double tol;                   /* [1e-8] */
  double damp;                  /* [1.0] */
  double q_hyst_pu;             /* net.q_hyst_pu [0.0] */
  double vm_min_pu;             /* [0.9] */
  double vm_max_pu;             /* [1.1] */
  double base_mva;              /* [100.0] (loading in percent of sn_MVA) */
  double mem_fraction;          /* share of free device memory the auto slot count may use [0.6] */
  double relax_zero_frac;       /* supernode amalgamation threshold [0.08] */
  double panel_smem_kb;         /* shared-memory budget of one front panel in KB (<= 200) [200] */
  double pivot_tol_rel;         /* static-pivoting monitor: a 2x2 pivot block with |det| <= pivot_tol_rel * max|entry|^2
                                   is counted as a tiny pivot (sblx_results.n_tiny_pivots, sblx_stats.tiny_pivots);
                                   det == 0 / non-finite still means singular_newton_step [1e-10] */
} sblx_options;

/* Caller-allocated result arrays (any may be NULL). B = number of scenarios, n = buses. */
typedef struct sblx_results {
  uint8_t* converged;        /* [B]  ContingencyResult.converged */
  int32_t* status;           /* [B]  SBLX_ST_* */
  int32_t* iterations;       /* [B]  ContingencyResult.iterations */
  int32_t* n_switch_events;  /* [B]  length(net.qLimitLog) */
  int32_t* island_count;     /* [B]  ContingencyResult.island_count */
  double* final_mismatch;    /* [B]  last entry of the mismatch history */
  double* vmin_pu;           /* [B]  min_vm_pu (NaN when not converged) */
  double* vmax_pu;           /* [B]  max_vm_pu */
  double* max_loading_pct;   /* [B]  max_branch_loading_pct (NaN when unrated / not converged) */
  double* V;                 /* [B*n*2] final bus voltages (re,im), scenario-major */
  uint8_t* bus_type_final;   /* [B*n] final active set (SBLX_BUS_*) */
  /* --- since 0.2.0: what _contingency_metrics hands to ContingencyResult, computed on the device ------------- */
  int32_t* n_voltage_violations; /* [B] length(voltage_violations): non-isolated buses outside [vm_min_pu, vm_max_pu] */
  int32_t* n_overloaded;         /* [B] length(overloaded_branches): in-service rated branches with loading > 100 % */
  int32_t* n_tiny_pivots;        /* [B] near-singular 2x2 pivot blocks met by the static-pivoting LU (0 = none) */
  /* compacted lists over all converged scenarios, sorted by (scenario, element); ids follow index_base.  The caller
   * provides the capacity; *_count receives the TOTAL number of entries (may exceed the capacity: then only the
   * per-scenario counts above are complete and the call can be repeated with larger arrays). */
  int64_t viol_capacity;         /* entries available in viol_scenario / viol_bus */
  int32_t* viol_scenario;        /* [viol_capacity] scenario index (0-based position in the batch) */
  int32_t* viol_bus;             /* [viol_capacity] bus id */
  int64_t* viol_count;           /* [1] */
  int64_t over_capacity;
  int32_t* over_scenario;        /* [over_capacity] */
  int32_t* over_branch;          /* [over_capacity] branch id */
  double* over_loading_pct;      /* [over_capacity] its loading */
  int64_t* over_count;           /* [1] */
  int64_t qlog_capacity;         /* Q-limit event log of all scenarios = net.qLimitLog (src/limits.jl:75-84): one entry per */
  int32_t* qlog_scenario;        /* [qlog_capacity] PV->PQ switch, sorted by (scenario, iteration, bus) - the reference's push order */
  int32_t* qlog_iter;            /* [qlog_capacity] NR iteration of the event (1-based) */
  int32_t* qlog_bus;             /* [qlog_capacity] bus id */
  int32_t* qlog_side;            /* [qlog_capacity] +1 = :max, -1 = :min */
  int64_t* qlog_count;           /* [1] total entries (may exceed the capacity, see above) */
  double* flows;                 /* [B*nbranch*4] per branch S_from (re,im), S_to (re,im) in p.u. (calcNetLosses!,
                                    src/losses.jl:53-85); zero for outaged / out-of-service branches and for
                                    scenarios that did not converge.  Optional and large: B*nbranch*32 bytes */
} sblx_results;

typedef struct sblx_info {
  int64_t n, m, nnz_y, nnz_j_blocks, nnz_j_scalar, nnz_lu_scalar;
  int64_t n_panels, n_levels, u_blocks, c_blocks;
  int64_t max_slots, bytes_per_slot;
  double block_ops;             /* 2x2 block multiply-adds per numeric factorisation */
  double algorithmic_bytes_per_iteration; /* 16*nnzJ + 16*nnzLU + 128*n   */
  char ordering[16];
} sblx_info;

/* Timing / counters of the last solve call (CUDA events on the engine's stream). */
typedef struct sblx_stats {
  double total_ms;          /* whole device run (first launch to last kernel) */
  double factor_ms;         /* sum over launches of the front factorisation kernels */
  double solve_ms;          /* back-substitution kernels */
  double mismatch_ms;       /* mismatch + q-limit + norm kernels */
  double jacobian_ms;       /* Jacobian fill */
  double other_ms;
  double h2d_ms, d2h_ms, host_prep_ms;
  int64_t launches;         /* kernels launched */
  int64_t factor_launches;
  int64_t ticks;            /* batch iterations (host round trips) */
  int64_t scenario_iterations;   /* Newton steps summed over scenarios (= factorisations) */
  int64_t mismatch_evaluations;  /* iterations summed over scenarios */
  int64_t scenarios_solved;      /* scenarios that went to the GPU */
  int64_t h2d_bytes, d2h_bytes;
  /* since 0.2.0 */
  int64_t tiny_pivots;           /* sum of n_tiny_pivots over the batch */
  int64_t n_ranks;               /* ranks (GPUs) that shared the last sharded solve, 1 otherwise */
  int64_t shard_lo, shard_hi;    /* this rank's contiguous scenario range of the last sharded solve */
  double allgather_ms;           /* device time of the one ncclAllGather of the last sharded solve */
  int64_t refine_passes;         /* iterative-refinement passes (second factor + solve on the residual) run for steps whose
                                    factorisation met a tiny pivot and whose residual exceeded 1e-6 |F| */
} sblx_stats;

int sblx_version(void);
void sblx_default_options(sblx_options* opts);
const char* sblx_last_error(const sblx_handle* h); /* h may be NULL: last create() error */

/* Ybus: n x n CSC (colptr[n+1], rowval[nnz], nzval[2*nnz]); per-bus arrays of length n; branch
 * two-port table of length nbranch (from/to bus, Y11,Y12,Y21,Y22 interleaved complex, rating
 * sn_MVA or NaN, status 0/1).  Ybus must equal the stamp sum of the in-service branches plus
 * bus shunts on the diagonal (what createYBUS produces). */
int sblx_create(sblx_handle** out, int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                int64_t slack_idx, const uint8_t* bus_type, const double* vset, const double* s_spec,
                const double* v0, const double* qmin_pu, const double* qmax_pu, const double* qload_pu,
                int64_t nbranch, const int32_t* br_from, const int32_t* br_to, const double* br_y11,
                const double* br_y12, const double* br_y21, const double* br_y22, const double* br_rating,
                const uint8_t* br_status, const sblx_options* opts);
void sblx_destroy(sblx_handle* h);

/* lock_pv_to_pq_buses (runpf! keyword): PV buses the active set must not switch. */
int sblx_set_locked_pv(sblx_handle* h, int64_t nlock, const int32_t* buses);

/* Replace the start vector V0 (the warm start every scenario begins from; runContingencies! uses
 * the solved base case: src/contingency.jl:275-287). v0 = [2n] interleaved complex. */
int sblx_set_v0(sblx_handle* h, const double* v0);

int sblx_get_info(const sblx_handle* h, sblx_info* info);
int sblx_get_stats(const sblx_handle* h, sblx_stats* stats);

/* N-k outages: scenario s takes branches outage_branch[outage_ptr[s] .. outage_ptr[s+1]) out of
 * service (an empty range is the intact grid).  outage_ptr is 0-based regardless of index_base;
 * branch ids follow index_base.  Duplicate ids within a scenario count once (a branch can only be removed once).  At
 * most 8 distinct branches per scenario; a scenario with more, or with an id that is not a branch of the model, is
 * reported with status SBLX_ST_BAD_BRANCH (never an error return); outage_ptr must start at 0 and be non-decreasing
 * (SBLX_E_ARG otherwise).  Outages that
 * isolate buses or split the grid are handled like the reference does (identity rows; island-wise solves with
 * reference promotion; SBLX_ST_ISLANDED_NO_REF when an island with branches has no slack or PV bus). */
int sblx_solve_outages(sblx_handle* h, int64_t B, const int32_t* outage_ptr, const int32_t* outage_branch,
                       sblx_results* res);

/* Injection sweep: scenario s uses s_spec[s*n .. (s+1)*n) (interleaved complex) and, when given,
 * qload_pu[s*n ..] instead of the model's vectors.  The matrices are streamed through the device in chunks of
 * scenarios (pinned staging, about 4 GB per chunk): B x n is never resident at once. */
int sblx_solve_injections(sblx_handle* h, int64_t B, const double* s_spec, const double* qload_pu,
                          sblx_results* res);

/* Device-side injection sweeps: nothing but the parameters crosses PCIe.
 * _sweep: scenario s scales P, Q and the Q-limit load term of every load bus (Re(s_spec) < 0) by an independent factor
 *   f = clip(1 + sigma * z, lo, hi), z ~ N(0,1) by Box-Muller from Philox4x32-10(key = seed, counter = (s + first, bus)):
 *   the reference's Monte-Carlo recipe (examples/powerflow/mc_probabilistic_powerflow.jl:28-41) with a counter-based
 *   generator, so a scenario's draw does not depend on batching, slot or rank.  `first` = global index of scenario 0
 *   (lets a caller split one study into several calls / ranks).
 * _scale: one factor per scenario for all load buses (uniform load scaling, src/network.jl:2359-2402 semantics). */
int sblx_solve_injection_sweep(sblx_handle* h, int64_t B, int64_t first, uint64_t seed, double sigma, double lo, double hi,
                               sblx_results* res);
int sblx_solve_injection_scale(sblx_handle* h, int64_t B, const double* scale /* [B] */, sblx_results* res);
/* The factors sblx_solve_injection_sweep applies, evaluated by the same device code: out[B*n], 1.0 on non-load buses.
 * Lets a caller / the parity tests rebuild the exact per-scenario injections. */
int sblx_sweep_factors(sblx_handle* h, int64_t B, int64_t first, uint64_t seed, double sigma, double lo, double hi,
                       double* out);

/* One power flow on the intact grid from an explicit start vector (solvePf analogue). */
int sblx_solve_one(sblx_handle* h, const double* v_start /* [2n] or NULL = model V0 */, double* v_out /* [2n] */,
                   uint8_t* converged, int32_t* iters, double* residual_inf, int32_t* status);

/* Split form used for device-resident timing: stage() does host analysis + H2D, run() is device
 * only (inputs resident in HBM), fetch() copies results back.  solve_* = stage + run + fetch. */
int sblx_stage_outages(sblx_handle* h, int64_t B, const int32_t* outage_ptr, const int32_t* outage_branch,
                       int want_v, int want_types);
int sblx_run_staged(sblx_handle* h);
int sblx_fetch_results(sblx_handle* h, sblx_results* res);
/* device pointer + byte size of the packed per-scenario summary records of the last run, one per CALLER scenario in
 * input order (island-wise composites and host verdicts included): record = {int32 converged, status, iterations,
 * n_switch; double final_mismatch, vmin, vmax, max_loading; int32 n_voltage_violations, n_overloaded, island_count,
 * n_tiny_pivots} = SBLX_RECORD_BYTES.  The buffer stays valid until the next stage / solve call on the handle. */
#define SBLX_RECORD_BYTES 64
int sblx_device_summary(sblx_handle* h, void** dev_ptr, int64_t* nbytes);

/* ---- multi-GPU: scenario shards + ONE all-gather inside the call (src/contingency.jl:299-318) ----------------
 * NCCL is loaded at run time (dlopen of libnccl.so.2; SBLX_NCCL_LIB overrides the path) - a build or a process that
 * never calls these entries does not need it.
 *
 * (a) one process per GPU (torchrun / MPI style): every rank creates its own handle on its own device with the SAME
 *     model, rank 0 makes an id (sblx_comm_unique_id) and ships its 128 bytes to the others by any means, every rank
 *     calls sblx_comm_init_rank.  sblx_solve_outages_sharded is then a COLLECTIVE call: every rank passes the SAME full
 *     batch, solves its contiguous cost-balanced shard, and after one ncclAllGather of the 64-byte records every rank
 *     returns the summary arrays of ALL B scenarios in input order.  V / bus_type_final / flows / the lists are filled
 *     for the rank's own shard only (rows outside [shard_lo, shard_hi) of sblx_stats are left untouched).
 * (b) one process, several GPUs (what a Julia caller uses): sblx_create_multi builds one engine per device and an
 *     ncclCommInitAll communicator; sblx_multi_solve_outages runs one host thread per GPU, the same shard + all-gather,
 *     and returns everything for all B scenarios (V / types / flows / lists are copied from every device). */
int sblx_comm_unique_id(uint8_t id128[128]);
int sblx_comm_init_rank(sblx_handle* h, int32_t nranks, int32_t rank, const uint8_t id128[128]);
int sblx_comm_destroy(sblx_handle* h);
int sblx_solve_outages_sharded(sblx_handle* h, int64_t B, const int32_t* outage_ptr, const int32_t* outage_branch,
                               sblx_results* res);
/* the shard boundaries the sharded calls use: nranks + 1 offsets into the batch - contiguous ranges in input order
 * like the reference's chunking (contingency.jl:304), balanced so that shard sizes differ by at most one; identical
 * on every rank.  Host only. */
int sblx_shard_bounds(const sblx_handle* h, int64_t B, const int32_t* outage_ptr, const int32_t* outage_branch,
                      int32_t nranks, int64_t* bounds /* [nranks + 1] */);

typedef struct sblx_multi sblx_multi; /* opaque: ndev engines + one communicator */
int sblx_create_multi(sblx_multi** out, int32_t ndev, const int32_t* devices /* [ndev] or NULL = 0..ndev-1 */, int64_t n,
                      const int64_t* colptr, const int64_t* rowval, const double* nzval, int64_t slack_idx,
                      const uint8_t* bus_type, const double* vset, const double* s_spec, const double* v0,
                      const double* qmin_pu, const double* qmax_pu, const double* qload_pu, int64_t nbranch,
                      const int32_t* br_from, const int32_t* br_to, const double* br_y11, const double* br_y12,
                      const double* br_y21, const double* br_y22, const double* br_rating, const uint8_t* br_status,
                      const sblx_options* opts);
void sblx_destroy_multi(sblx_multi* m);
const char* sblx_multi_last_error(const sblx_multi* m);
int sblx_multi_set_v0(sblx_multi* m, const double* v0);
int sblx_multi_set_locked_pv(sblx_multi* m, int64_t nlock, const int32_t* buses);
int sblx_multi_solve_outages(sblx_multi* m, int64_t B, const int32_t* outage_ptr, const int32_t* outage_branch,
                             sblx_results* res);
int sblx_multi_get_stats(const sblx_multi* m, int32_t dev_index, sblx_stats* stats);
sblx_handle* sblx_multi_handle(sblx_multi* m, int32_t dev_index); /* borrowed: e.g. sblx_solve_one on device 0 */

/* ---- MATPOWER `.m` case reader (host only; text rules of src/MatpowerIO.jl:268-514) -----------------------------------
 * baseMVA and the bus / gen / branch matrices padded or truncated to the MATPOWER v2 widths 13 / 21 / 13 (row-major),
 * `%` comments stripped, rows split at `;` / CR / LF, bus rows sorted by BUS_I when they are not, raw TAP kept (0 = line).
 * Errors: SBLX_E_ARG with the text in sblx_mpc_last_error(). */
typedef struct sblx_mpc sblx_mpc; /* opaque */
int sblx_mpc_read(const char* path, sblx_mpc** out);
int sblx_mpc_parse(const char* text, sblx_mpc** out);
int sblx_mpc_dims(const sblx_mpc* m, double* base_mva, int64_t* nbus, int64_t* ngen, int64_t* nbranch);
int sblx_mpc_tables(const sblx_mpc* m, double* bus /* [nbus*13] */, double* gen /* [ngen*21] */, double* branch /* [nbranch*13] */);
void sblx_mpc_free(sblx_mpc* m);
const char* sblx_mpc_last_error(void);

/* ---- test / parity hooks (one scenario, every intermediate exposed) ---------------------- */
/* Block pattern of the Jacobian: CSR over the m = n-1 non-slack buses in ascending bus order;
 * col entries are bus indices (0-based). */
int sblx_debug_pattern(const sblx_handle* h, int64_t* rowptr /* [m+1] */, int32_t* col /* [nnzB] */);
/* One Newton step at voltage v_in for the given outage set and bus types (NULL = model types):
 * F [2m] (rows interleaved per non-slack bus: dP, dQ|dV), jblocks [4*nnzB] (row-major 2x2 blocks
 * (dP,dQ|dV) x (dVr,dVi) in pattern order), dx [2m] (dVr,dVi per non-slack bus) with J dx = -F. */
int sblx_debug_newton_step(sblx_handle* h, int64_t nout, const int32_t* outage_branch, const uint8_t* bus_type,
                           const double* v_in, double* F, double* jblocks, double* dx, double* max_mismatch);
/* Host-only self-test of the symbolic maps: factor + solve a random diagonally-dominant block
 * matrix on the pattern with a plain C++ loop nest that walks the same panels / index maps as
 * the CUDA kernels; returns the relative residual.  Never used by a solve entry point. */
int sblx_debug_host_selftest(const sblx_handle* h, uint64_t seed, double* rel_residual);

/* Symbolic-only constructor (no CUDA needed): lets CPU tests inspect the analysis. */
int sblx_analyze_only(int64_t n, const int64_t* colptr, const int64_t* rowval, int64_t slack_idx,
                      const sblx_options* opts, sblx_info* info, double* selftest_residual);

/* Host-only (no CUDA): the per-scenario topology analysis of sblx_solve_outages - island count over the in-service
 * branches, number of isolated buses, and the verdict: -1 solved as one system on the device, -2 solved island-wise,
 * SBLX_ST_ISLANDED_NO_REF, SBLX_ST_BAD_BRANCH.  bus_type as in sblx_create (SBLX_BUS_*). */
int sblx_debug_topology(int64_t n, int64_t slack_idx, const uint8_t* bus_type, int64_t nbranch, const int32_t* br_from,
                        const int32_t* br_to, const uint8_t* br_status, int32_t index_base, int64_t B,
                        const int32_t* outage_ptr, const int32_t* outage_branch, int32_t* island_count, int32_t* status,
                        int32_t* n_isolated);

/* Tuning probe (tools only): sets the k_front prefetch / phase-ablation mode bits (SBLX_PF_MODE) on a live handle, so
 * that a probe can solve the base case normally and switch the experiment on afterwards. */
int sblx_debug_set_pf_mode(sblx_handle* h, int32_t mode);

/* Instrumentation: per size class (5) x phase (10) cycle sums of k_front, thread 0's view; only filled by a build with
 * -DSBLX_PHASE_TIMING (tools/phase_timing.py), SBLX_E_ARG otherwise.  `out50` receives 50 doubles. */
int sblx_debug_phase_cycles(double* out50);

#ifdef __cplusplus
}
#endif
#endif /* SBLX_H */
