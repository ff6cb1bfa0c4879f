/* gpulin.h -- C ABI of the B200-native linear bound propagation library (libgpulin.so).
 *
 * This is the drop-in boundary for the reference path
 *      propagationRound (solve.c:437) -> consPropLinear (cons_linear.c:16126) -> propagateCons (:7621)
 *      -> tightenBounds (:6980) -> tightenVarBoundsEasy / tightenVarBounds (:5380 / :6700)
 * The reference has no FFI for this path (it is in-process C); the entry points below are what the SCIP
 * propagator plugin scip_b200/plugin/prop_gpulinear.c (SCIP_DECL_PROPEXEC, type_prop.h:217) binds.
 * No SCIP, torch or C++ types cross this boundary: plain pointers and sizes only.
 *
 * All functions return GPULIN_OK (0) or a negative error code; gpulin_last_error() gives the text.
 * There is NO CPU fallback: if CUDA is unavailable every call fails with GPULIN_ERR_CUDA.
 */
#ifndef GPULIN_H
#define GPULIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPULIN_OK            0
#define GPULIN_ERR_CUDA     -1   /* CUDA runtime / driver failure (SCIP side: SCIP_ERROR) */
#define GPULIN_ERR_ARG      -2   /* invalid argument */
#define GPULIN_ERR_NOMEM    -3   /* host or device allocation failed */
#define GPULIN_ERR_STATE    -4   /* call sequence violated (e.g. propagate before set_bounds) */

/* propagation verdicts (gpulin_result.status) */
#define GPULIN_FIXPOINT      0   /* no bound changed in the last round */
#define GPULIN_CUTOFF        1   /* infeasibility proven (SCIP_CUTOFF) */
#define GPULIN_ROUNDLIMIT    2   /* maxrounds reached while bounds were still moving */

/* variable types (SCIPvarIsIntegral, pub_var.h:1209) */
#define GPULIN_VAR_CONTINUOUS 0
#define GPULIN_VAR_INTEGRAL   1

typedef struct gpulin gpulin_t;

/** numerical tolerances; replaces SCIPinfinity/SCIPepsilon/SCIPsumepsilon/SCIPfeastol/SCIPgetHugeValue
 *  (scip_numerics.c:140-546), numerics/boundstreps (set.c:2404) and constraints/linear/maxeasyactivitydelta
 *  (cons_linear.c:135) */
typedef struct gpulin_numerics
{
   double infinity;             /* 1e20 */
   double epsilon;              /* 1e-9 */
   double sumepsilon;           /* 1e-6 */
   double feastol;              /* 1e-6 */
   double boundstreps;          /* 0.05 */
   double hugeval;              /* 1e15 */
   double maxeasyactivitydelta; /* 1e6  */
} gpulin_numerics;

typedef struct gpulin_result
{
   int32_t status;        /* GPULIN_FIXPOINT / _CUTOFF / _ROUNDLIMIT */
   int32_t nrounds;       /* propagation rounds executed (the last one found no change unless cutoff/limit) */
   int64_t nchanges;      /* accepted bound changes, counted per variable bound and round */
   int64_t nnz_processed; /* nonzeros swept over all rounds (dirty rows only after the first round) */
   double  device_ms;     /* device time of the whole call (CUDA events on the library's stream) */
} gpulin_result;

/** one accepted bound change, in round order (for SCIPinferVarLbProp/UbProp, scip_var.c:7589/7705) */
typedef struct gpulin_change
{
   int32_t var;        /* column index */
   int32_t round;      /* round in which the change was accepted (0-based) */
   double  newbound;
   int32_t is_upper;   /* 0: lower bound raised, 1: upper bound lowered */
   int32_t reserved;
} gpulin_change;

/** fills the reference's default tolerances (def.h:172-184) */
void gpulin_default_numerics(gpulin_numerics* num);

/** text of the last error on the calling thread */
const char* gpulin_last_error(void);

/** number of CUDA devices visible, or a negative error code */
int gpulin_device_count(void);

/** builds the device copy (tiled CSR stream + column->row map) of the linear rows
 *      lhs[i] <= sum_k vals[k] * x[colidx[k]] <= rhs[i],   k in [rowptr[i], rowptr[i+1])
 *  replaces the per-constraint SCIP_CONSDATA arrays (cons_linear.c:187-266) read through
 *  SCIPgetVarsLinear/SCIPgetValsLinear/SCIPgetLhsLinear/SCIPgetRhsLinear (cons_linear.c:18326-18452).
 *  Rows may be a slice of a larger problem (multi-GPU row partition): ncols is always the global count. */
int gpulin_create(
   int                    device,     /* CUDA device ordinal */
   int64_t                nrows,
   int64_t                ncols,
   int64_t                nnz,
   const int64_t*         rowptr,     /* host, nrows+1 */
   const int32_t*         colidx,     /* host, nnz */
   const double*          vals,       /* host, nnz; no zeros (cons_linear.c:5422) */
   const double*          lhs,        /* host, nrows; <= -infinity: none */
   const double*          rhs,        /* host, nrows; >= +infinity: none */
   const uint8_t*         vartype,    /* host, ncols; GPULIN_VAR_* */
   const gpulin_numerics* num,        /* NULL: defaults */
   gpulin_t**             out
   );

void gpulin_destroy(gpulin_t* h);

/** (re)loads all variable bounds from host memory and marks every row for propagation
 *  (SCIPvarGetLbLocal/UbLocal, pub_var.h:877,889) */
int gpulin_set_bounds(gpulin_t* h, const double* lb, const double* ub);

/** same, bounds already on the device of this handle */
int gpulin_set_bounds_device(gpulin_t* h, const double* d_lb, const double* d_ub);

/** bounds as 2 bits per column against resident reference bounds.  At a node of a MIP almost every column sits at its
 *  global bounds or is fixed to one of them: 2 bits of information instead of 16 bytes over PCIe.
 *  gpulin_set_reference_bounds uploads the reference (e.g. SCIPvarGetLbGlobal/UbGlobal, pub_var.h:853,865) once;
 *  gpulin_set_bounds_packed then takes codes[(ncols + 15) / 16] words, column j in bits 2 (j % 16) .. of word j / 16:
 *  0 = reference bounds, 1 = fixed to the reference lower bound, 2 = fixed to the reference upper bound, 3 = explicit: the
 *  bounds follow in (idx, lb, ub)[nexplicit].  Marks every row, like gpulin_set_bounds */
int gpulin_set_reference_bounds(gpulin_t* h, const double* lb, const double* ub);
int gpulin_set_bounds_packed(gpulin_t* h, const uint32_t* codes, int64_t nexplicit, const int32_t* idx, const double* lb,
   const double* ub);

/** ranged-row propagation on (enable != 0) or off (the default): the gcd rule for equations and ranged rows with at least
 *  three nonzeros -- rangedRowPropagation, cons_linear.c:5715-6696 (constraints/linear/rangedrowpropagation), without its
 *  branches that add constraints (constraints/linear/rangedrowartcons = FALSE) -- runs on every marked ranged row beside the
 *  activity rules.  The rule depends on the order in which it walks a row, and the reference has sorted the row by then
 *  (consdataCompVarProp, cons_linear.c:3191-3257): glb / gub = the global bounds the reference sorts by
 *  (SCIPvarGetLbGlobal / UbGlobal), tie = SCIPvarGetProbindex of every column (NULL: the column index).  Call it before
 *  the handle is connected to peers or cloned */
int gpulin_set_rangedrow(gpulin_t* h, int enable, const double* glb, const double* gub, const int32_t* tie);

/** overwrites n bounds (tightened or relaxed: branching, backtracking) and marks only the rows of those
 *  columns -- the counterpart of eventExecLinear's SCIPmarkConsPropagate (cons_linear.c:17229) */
int gpulin_update_bounds(gpulin_t* h, int64_t n, const int32_t* idx, const double* lb, const double* ub);

/** runs propagation rounds on the device until fixpoint, cutoff or maxrounds (<= 0: unlimited); the round
 *  loop runs inside one persistent kernel, no host round trip per round (propagateDomains, solve.c:766) */
int gpulin_propagate(gpulin_t* h, int maxrounds, gpulin_result* res);

/** the two halves of gpulin_propagate: enqueue the whole fixpoint loop on the handle's stream and return at once;
 *  wait for it and fetch the verdict.  Lets one host thread keep several handles (clones) busy concurrently */
int gpulin_propagate_async(gpulin_t* h, int maxrounds);
int gpulin_propagate_wait(gpulin_t* h, gpulin_result* res);

/** another set of bound vectors on the SAME matrix (probing, SCIPstartProbing scip_probing.c:120): the clone shares the
 *  read-only device arrays with `h` and owns bounds, keys, marks, work lists, stream and graph.  Destroy in any order */
int gpulin_clone(gpulin_t* h, gpulin_t** out);

/** probing start: `h` (a clone of `base` or vice versa) takes over the bounds of `base` with no row marked -- the state
 *  of a node whose propagation is complete; follow with gpulin_update_bounds (SCIPchgVarLb/UbProbing :302/:346) and
 *  gpulin_propagate (SCIPpropagateProbing :581); calling it again is SCIPbacktrackProbing (:226) */
int gpulin_reset_from(gpulin_t* h, gpulin_t* base);

/** a batch of probes on the node held by `base` (BASELINE config 5; SCIPapplyProbingVar, prop_probing.c:1254-1279):
 *  probe i sets column var[i] to [lb[i], ub[i]] and propagates to its fixpoint; `nworkers` clones of the base handle
 *  (kept inside it) hold that many probes in flight.  Outputs (any may be NULL): verdict, rounds and accepted bound
 *  changes per probe.  The base handle's bounds are not modified */
int gpulin_probe_batch(gpulin_t* base, int nworkers, int64_t nprobes, const int32_t* var, const double* lb,
   const double* ub, int maxrounds, int32_t* status, int32_t* nrounds, int64_t* nchanges);

/** the same, and what every probe implied: probe i's accepted bound changes in round order are chg[chgbeg[i] .. chgbeg[i+1])
 *  (chgbeg has nprobes + 1 entries) -- the proplbs / propubs that SCIPapplyProbingVar returns per candidate
 *  (prop_probing.c:1203-1303), in sparse form; the bounds at the probe's fixpoint are the node's bounds overwritten by
 *  these entries in order.  *nchg = entries produced; if they exceed maxchg the call fails with GPULIN_ERR_ARG and *nchg
 *  says how many are needed */
int gpulin_probe_batch_changes(gpulin_t* base, int nworkers, int64_t nprobes, const int32_t* var, const double* lb,
   const double* ub, int maxrounds, int32_t* status, int32_t* nrounds, int64_t* nchanges, int64_t* chgbeg, gpulin_change* chg,
   int64_t maxchg, int64_t* nchg);

/** copies the current bounds to host memory */
int gpulin_get_bounds(gpulin_t* h, double* lb, double* ub);

/** copies the current bounds to device memory of this handle's device */
int gpulin_get_bounds_device(gpulin_t* h, double* d_lb, double* d_ub);

/** enables (capacity > 0) or disables the round-ordered change log kept for the SCIP plugin */
int gpulin_set_change_log(gpulin_t* h, int64_t capacity);

/** copies up to maxn log entries of the last gpulin_propagate call; *n gets the number of entries that
 *  were produced (if *n > capacity the log overflowed and only the first `capacity` are valid) */
int gpulin_get_changes(gpulin_t* h, gpulin_change* out, int64_t maxn, int64_t* n);

/** the same log as 12-byte records { uint32 column | is_upper << 31, new bound as 2 x uint32 (low, high word) }; the round
 *  of an entry follows from its position: the log is round ordered, gpulin_get_round_stats gives the counts per round */
int gpulin_get_changes_packed(gpulin_t* h, void* out, int64_t maxn, int64_t* n);

/** the same log as ONE 32-bit word per entry, in log (= round) order: column (bits 0-28) | code << 29 | is_upper << 31, where
 *  code 0 / 1 says the new bound is 0.0 / 1.0 -- what the bound of a binary moves to -- and code 2 that the bound is in the
 *  side list xout: { position of the entry in the log, low word, high word of the bound } per entry, 12 bytes each, in no
 *  particular order.  *n = entries produced (as gpulin_get_changes), *nx = side-list entries produced (only min(*nx, maxx)
 *  are written).  A propagation of a binary program comes back in 4 bytes per change instead of 24; needs ncols < 2^29.
 *  (The bounds are what SCIPinferVarLbProp/UbProp get, scip_var.c:7589/7705.) */
int gpulin_get_changes_compact(gpulin_t* h, uint32_t* out, int64_t maxn, int64_t* n, uint32_t* xout, int64_t maxx, int64_t* nx);

/** per-round statistics of the last gpulin_propagate call (up to maxn rounds; any output array may be NULL):
 *  device time [ms] (%globaltimer stamps taken by the kernels), nonzeros swept, bound changes accepted */
int gpulin_get_round_stats(gpulin_t* h, double* ms, int64_t* nnz, int64_t* nchg, int32_t maxn, int32_t* n);

/** measurement aid: the kernel starts of the last gpulin_propagate call (the first 256), time ordered: ids[i] names the
 *  kernel (1 begin, 2 thread-per-row sweep, 3 tile sweep, 4 block-per-row sweep, 5 exact rules, 6 / 7 exchange push start /
 *  end, 8 / 9 exchange merge start / all peers arrived, 10 apply, 11 sparse rounds), us[i] = microseconds since the call
 *  began on the device (%globaltimer).  Works inside the CUDA-graph loop, where a profiler sees no launches */
int gpulin_get_trace(gpulin_t* h, int32_t* ids, double* us, int32_t maxn, int32_t* n);

/** several GPUs: per round of the last call (up to maxn) the device time [ms] from the start of the round to the start of
 *  its exchange (sweeps + exact rules of this rank's share; -1: the round had no exchange -- it ran redundantly on every
 *  rank) and how long the merge then waited for the slowest peer */
int gpulin_get_exchange_stats(gpulin_t* h, double* before_ms, double* wait_ms, int32_t maxn, int32_t* n);

/** redundancy feedback -- replaces the verdict at the end of propagateCons (cons_linear.c:7743-7753: a row whose activity
 *  bounds lie inside its sides, GE(minactivity, lhs) and LE(maxactivity, rhs), is deleted locally with SCIPdelConsLocal):
 *  the rows (caller's numbering, ascending) that are redundant for the bounds on the device, e.g. after gpulin_propagate.
 *  *n = their number (may exceed maxn; only maxn are written). */
int gpulin_get_redundant_rows(gpulin_t* h, int32_t* rows, int64_t maxn, int64_t* n);

/** storage statistics: [0] nnz, [1] stored nonzeros incl. padding, [2] rows swept thread-per-row (SELL-32 slices),
 *  [3] rows in the tiled CSR stream, [4] rows swept block-per-row, [5] bytes on device, [6] tiles of the stream,
 *  [7..8] persistent blocks of the thread-per-row / tile sweep, [9] longest row, [10] thread-per-row rows whose
 *  coefficients are all +1 / -1 (stored first; the filter sweep does not read their values), [11] blocks of the
 *  bit-table variant of the thread-per-row sweep (0: the gather variant is used) */
int gpulin_get_layout(gpulin_t* h, int64_t* stats, int32_t nstats);

/** what the last gpulin_propagate call launched: [0] kernel launches, [1] rounds run by the dense kernels (filter sweep(s)
 *  + exact + apply [+ barriers with peers]), [2] rounds run inside the persistent sparse-rounds kernel, [3] 1 if the call
 *  was started by the one-block kernel (few updated bounds since the last fixpoint), [4] 1 if that kernel handed over to
 *  the general loop, [5] rows whose exact activities came along from the filter sweep and that took the thread-per-row
 *  phase of the exact kernel (integer rows over integral columns, see gpulin_kernels.cuh: FastAcc) */
int gpulin_get_call_stats(gpulin_t* h, int64_t* stats, int32_t nstats);

/** algorithmic bytes of one full round: nnz*12 + nrows*20 + ncols*17 */
int gpulin_algorithmic_bytes(gpulin_t* h, int64_t* bytes);

/** measurement aid: runs ONE full round (every row marked for propagation) on the current bounds and returns the
 *  CUDA-event time [ms] of its three stages: filter sweep kernel(s), exact kernel, apply kernel */
int gpulin_profile_round(gpulin_t* h, double* sweep_ms, double* exact_ms, double* apply_ms);

/** all following work of this handle is enqueued on the given cudaStream_t (default: a private stream) */
int gpulin_set_stream(gpulin_t* h, void* stream);

/** blocks until all work enqueued by this handle is done */
int gpulin_sync(gpulin_t* h);

/* ---- single rounds: multi-GPU (rows partitioned over ranks, bounds replicated) and profiling ----
 * One sharded round = gpulin_round_sweep on every rank, ONE ncclMin all-reduce (int64) over the buffer returned by
 * gpulin_exchange_buffer (done by the host language on the handle's stream), then gpulin_round_apply(dense=1).
 * All calls are asynchronous on the handle's stream except gpulin_round_apply with non-NULL outputs. */

/** device pointer to the candidate keys: 2*ncols order-preserving int64 keys, [2j] = ~key(lb_j), [2j+1] = key(ub_j)
 *  (both tighten by MIN), followed by 2 spare keys that carry the cutoff verdict (negative = cutoff) */
int gpulin_exchange_buffer(gpulin_t* h, int64_t** d_keys, int64_t* nkeys);

/** copies the key vector to / from host memory (nkeys int64 each) -- for hosts that merge the ranks' candidates
 *  without device-side collectives (tests; an MPI based host) */
int gpulin_get_keys(gpulin_t* h, int64_t* keys);
int gpulin_set_keys(gpulin_t* h, const int64_t* keys);

/** marks every row for propagation */
int gpulin_mark_all(gpulin_t* h);

/** resets round counter, verdict and statistics (what gpulin_propagate does at its start) */
int gpulin_round_begin(gpulin_t* h);

/** one sweep over the rows marked for propagation; candidates are merged into the key vector */
int gpulin_round_sweep(gpulin_t* h);

/** accepts the key vector as the new bounds and marks the rows of changed columns.  dense = 0: only columns
 *  flagged by this handle's own sweep are examined; dense = 1: all columns (keys may have been moved by other
 *  ranks).  If nchanges or cutoff is non-NULL the call synchronises and returns the round's change count / verdict */
int gpulin_round_apply(gpulin_t* h, int dense, int64_t* nchanges, int32_t* cutoff);

/* ---- several GPUs of ONE node: dense rounds are shared, the exchange goes through peer memory (NVLink) -----------------
 * Every rank creates its handle on the WHOLE problem (the matrix is 0.25 - 0.9 GB of 180) and holds the same bounds.  In a
 * dense round a rank sweeps its share of the rows only; the columns its candidates touched travel once, packed, into an
 * inbox on every other rank (stores through peer memory from inside the round's kernels -- no NCCL call, no host round
 * trip), every rank merges what it received with atomicMin and applies the same changes; rounds with few marked rows run
 * on every rank redundantly without any exchange.  gpulin_set_bounds / gpulin_update_bounds / gpulin_propagate become
 * COLLECTIVE calls: the same arguments on every rank, results identical on every rank.  The counterpart in the reference
 * is the min/max merge of bounds of syncstore.c:921. */

/** one process per GPU: allocates this rank's inbox for `nranks` ranks and writes its CUDA IPC handle (*nbytes bytes;
 *  call with out = NULL to query the size); the host language all-gathers the handles */
int gpulin_peer_handles(gpulin_t* h, int nranks, void* out, int64_t* nbytes);

/** opens the other ranks' inboxes; allhandles = the nranks handle blobs in rank order */
int gpulin_peer_connect(gpulin_t* h, int rank, int nranks, const void* allhandles);

/** one process, n handles of the same problem on n different devices (the SCIP plugin: SCIP is single threaded, one
 *  process -- prop.c:646): handle i becomes rank i; peer access between the devices is enabled.  Drive the group with
 *  gpulin_set_bounds + gpulin_propagate_async on every handle, then gpulin_propagate_wait on every handle */
int gpulin_group_connect(gpulin_t** handles, int n);

#ifdef __cplusplus
}
#endif

#endif
