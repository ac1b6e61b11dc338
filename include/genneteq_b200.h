/* genneteq_b200 — C ABI of the B200 (sm_100a) pricing / relabel path of GenNetEq's generalized
 * network simplex.  Plain pointers and sizes only; every function returns a gne_status (0 = ok)
 * and never exits or throws across the boundary (the reference printf+exit()s instead,
 * trees.cpp:310-313, gnePivot.cpp:342-345).  One host thread per handle; each handle owns one
 * CUDA stream and all of its device memory until *_destroy.  Caller owns every host buffer.
 *
 * The reference has no FFI: its "operator interface" is the set of protected, non-virtual
 * members GenNetEq::solve calls every iteration (genneteq.cpp:97-100).  Each entry point below
 * names the reference code it replaces.  INTEGRATION.md shows the derived-class stub a
 * maintainer of the reference would add to route those three calls here.
 *
 * Layer 1 (gne_dev_*)   the device operators: nonbasic-arc structure-of-arrays in setNBarc
 *                       position order, node potentials, pricing passes.
 * Layer 2 (gne_solver_*) a host C++ mirror of GenNetEq (initialize / solve) that drives layer 1.
 * There is no CPU fallback: every call fails with GNE_ERR_CUDA when no device is usable.
 */
#ifndef GENNETEQ_B200_H
#define GENNETEQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GNE_OK = 0,
    GNE_ERR_CUDA = 1,          /* CUDA runtime error or no device (gne_last_error() has the text) */
    GNE_ERR_ARG = 2,           /* bad argument */
    GNE_ERR_CAPACITY = 3,      /* caller's output buffer too small; *count still holds the need  */
    GNE_ERR_TREE = 4,          /* basis forest invariant broken (reference: exit(-1), trees.cpp:310-313,327-330,469-472) */
    GNE_ERR_CYCLING = 5,       /* >5 consecutive cycles (reference: exit(-1), gnePivot.cpp:342-345) */
    GNE_ERR_UNSUPPORTED = 6,   /* rule SD; with equal-flow sets: batches, windows, forest access  */
    GNE_ERR_COMM = 7           /* NCCL error                                                      */
} gne_status;

/* pivot-rule ids: main.h:17-34 */
enum { GNE_LV = 0, GNE_LVW = 1, GNE_LVA = 2, GNE_LVE = 3, GNE_FV = 4, GNE_SD = 5, GNE_RV = 6, GNE_RWV = 7,
       GNE_RLV = 8, GNE_RLVW = 9, GNE_CL = 10, GNE_CLW = 11, GNE_SAA = 12, GNE_SAE = 13, GNE_CL2 = 14,
       GNE_CLW2 = 15, GNE_QS = 16, GNE_QSW = 17 };
/* Variable::setLoc, variables.h:23 */
enum { GNE_AT_LOWER = 1, GNE_AT_UPPER = 2 };

#define GNE_ROUNDOFF_TOLERANCE 1.0e-13      /* main.h:36-43 */

const char *gne_last_error(void);
const char *gne_version(void);
int gne_cuda_device_count(void);            /* 0 when no usable device (never falls back to the CPU) */

/* =============================== layer 1: device operators =============================== */
typedef struct gne_dev gne_dev;

/* n nodes, room for nbCapacity nonbasic arcs.  Multi-GPU: this handle prices global positions
 * [shardLo, shardHi) of the nonbasic list (pass 0, nbCapacity for a single GPU).            */
int gne_dev_create(gne_dev **out, int cudaDevice, int n, int64_t nbCapacity, int64_t shardLo, int64_t shardHi);
int gne_dev_destroy(gne_dev *d);
int gne_dev_sync(gne_dev *d);

/* --- nonbasic set mirror: GenNetEq::setNBarc + insertIntoL/U, removeFromL/U (gnePivot.cpp:527-600) --- */
/* whole list, position order; arrays hold the GLOBAL list, the handle keeps its shard        */
int gne_nb_reset(gne_dev *d, int64_t N, const int32_t *tail, const int32_t *head, const double *cost,
                 const double *mult, const int8_t *state);
/* setVariableCosts (genneteq.cpp:304-323): new cost per position after a phase switch         */
int gne_nb_set_costs(gne_dev *d, int64_t N, const double *cost);
/* one record; swap-remove = write(pos, record of the last position) then set_count(N-1);
 * push = write(N, record) then set_count(N+1).  Writes outside the shard are ignored.        */
int gne_nb_write(gne_dev *d, int64_t pos, int32_t tail, int32_t head, double cost, double mult, int8_t state);
int gne_nb_set_count(gne_dev *d, int64_t N);
/* read back (tests): any of the output pointers may be NULL                                   */
int gne_nb_read(gne_dev *d, int64_t lo, int64_t hi, int32_t *tail, int32_t *head, double *cost, double *mult, int8_t *state);

/* --- potentials: computePotentials (gnePotential.cpp:28-61) ---
 * The device keeps, per node, the root-relative pair (a0, a1) of computePotentialsInTermsOfRoot
 * (gnePotential.cpp:64-110) and the slot of its tree; per slot the tree's theta
 * (gnePotential.cpp:123-124).  finalize computes pi[i] = a0[i] + a1[i] * theta[slot[i]]
 * (gnePotential.cpp:127-135; multiply then add, never fused).  Only dirty entries cross PCIe. */
int gne_pot_update_nodes(gne_dev *d, int64_t count, const int32_t *node, const double *a0, const double *a1, const int32_t *slot);
int gne_pot_update_thetas(gne_dev *d, int64_t count, const int32_t *slot, const double *theta);
int gne_pot_finalize(gne_dev *d);
/* the same for every node of the given tree slots only (a tree whose theta changed): one sweep over the
 * slot column, no node list crosses PCIe                                                       */
int gne_pot_finalize_slots(gne_dev *d, int count, const int32_t *slots);
/* all of the above for one pivot in one call (what the solver uses): queued gne_nb_write records,
 * dirty node records, dirty thetas, then the finalize of the nodes in fnode[0..nFinal)
 * (nFinal < 0: every node).  Small updates are one one-CTA kernel reading pinned host memory.  */
int gne_pot_update_fused(gne_dev *d, int64_t nNodes, const int32_t *node, const double *a0, const double *a1,
                         const int32_t *slot, int64_t nTheta, const int32_t *tslot, const double *theta,
                         int64_t nFinal, const int32_t *fnode);
/* the same with the finalized potentials supplied: fpi[i] = the new pi of fnode[i], computed by the caller with the
 * reference's unfused operations (a0 + a1 * theta).  The device then only stores them - the relabel folded into the
 * pricing pass needs no dependent loads before its release flag.  fpi == NULL: as gne_pot_update_fused.          */
int gne_pot_update_fused_pi(gne_dev *d, int64_t nNodes, const int32_t *node, const double *a0, const double *a1,
                            const int32_t *slot, int64_t nTheta, const int32_t *tslot, const double *theta,
                            int64_t nFinal, const int32_t *fnode, const double *fpi);
/* direct write of pi (callers that computed potentials themselves)                            */
int gne_pot_set(gne_dev *d, int64_t count, const int32_t *node, const double *pi);
int gne_pot_set_all(gne_dev *d, const double *pi);
int gne_pot_get_all(gne_dev *d, double *pi);

/* --- pricing: rc = (cost - pi[tail]) + (mult * pi[head]) (computeArcRedCost, genneteq.h:312-318);
 * violated <=> (L && rc < -tol) || (U && rc > tol) (gnePivotRules.cpp:406-407).  Positions are
 * global positions in the nonbasic list. --- */

/* getVarWithLargestViolation's arc pass: computeArcsWithLargestViolation + checkVarForOffending
 * (gnePivotRules.cpp:401-413, 431-452).  Outputs exactly what the sequential tolerance-band
 * rule leaves behind: the final running largestViolation and offendingVars (as positions,
 * in list order).  *count = list length; GNE_ERR_CAPACITY if > cap.  *any = !isOptimal.      */
int gne_price_lv(gne_dev *d, double *largestViolation, int64_t *count, int64_t *listPos, int64_t cap, int *any);
/* the same in two steps, for callers that only need the list's length and one member (the reference
 * picks offendingVars[drand48() * size], gnePivotRules.cpp:152-153): begin runs the pass and keeps
 * the list inside the handle — long exact-tie lists stay on the device — at(i) returns member i.   */
int gne_price_lv_begin(gne_dev *d, double *largestViolation, int64_t *count, int *any);
/* the same in two halves, so that one host thread can drive several instances: launch starts the pass, ready says
 * without blocking whether its result has arrived (1 / 0), complete waits for it and resolves the list              */
int gne_price_lv_launch(gne_dev *d);
int gne_price_lv_ready(gne_dev *d);
int gne_price_lv_complete(gne_dev *d, double *largestViolation, int64_t *count, int *any);
int gne_price_lv_at(gne_dev *d, int64_t index, int64_t *pos);

/* Launch groups (configs[3], SURVEY 8e "batches: replicas, 32 instances per GPU"): the LV passes of several independent
 * handles of ONE GPU go out as one grid, each instance on its own share of the CTAs, instead of one launch per instance.
 * A group belongs to one host thread.  gne_lv_group_add(g, d) stands for gne_price_lv_launch(d): the pass is prepared and
 * parked; gne_lv_group_launch sends everything parked (add launches by itself once `maxPending` passes - at most 8, 0 = 8 -
 * are parked).  gne_price_lv_ready / _complete then work per handle as usual.  A pass the batched kernel does not cover
 * (sharded handle, per-tile kernel, another device) is launched on its own by add.                                    */
typedef struct gne_lv_group gne_lv_group;
int gne_lv_group_create(gne_lv_group **out, int cudaDevice, int maxPending);
int gne_lv_group_add(gne_lv_group *g, gne_dev *d);
int gne_lv_group_launch(gne_lv_group *g);
int gne_lv_group_pending(const gne_lv_group *g);
int64_t gne_lv_group_launch_count(const gne_lv_group *g);
int gne_lv_group_destroy(gne_lv_group *g);

/* quickSelect's arc loop (gnePivotRules.cpp:742-756): scan from `cursor` (wrapping) until `want`
 * violators were seen or the list is exhausted; best = first strictly-largest |rc|.
 * *bestPos = -1 when there is no violator.  *newCursor is the cursor after the scan (the
 * caller applies the loopIter % majorIterationLength restore, :758-761).                     */
int gne_price_qs(gne_dev *d, int64_t cursor, int64_t want, int64_t *bestPos, double *bestViol,
                 int64_t *newCursor, int64_t *scanned, int64_t *violations);

/* updateCandidateList2 / tryToAddArc (gnePivotRules.cpp:613-689): scan from the rolling cursor — at
 * least minScan arcs (numArcsToAlwaysCheck) — until `want` (clMaxSize) violators are listed or the
 * list wrapped; the violators come back in scan order with |rc|.                              */
int gne_price_scan_list(gne_dev *d, int64_t cursor, int64_t want, int64_t minScan, int64_t *count,
                        int64_t *listPos, double *listViol, int64_t cap, int64_t *newCursor, int64_t *scanned);

/* every violator with |rc| >= threshold, in position order (updateCandidateList
 * gnePivotRules.cpp:504-516; computeReducedCosts gnePotential.cpp:226-234;
 * getRandomVarWithLargeViolation :311-328).  listViol may be NULL.                           */
int gne_price_violators(gne_dev *d, double threshold, int64_t *count, int64_t *listPos, double *listViol, int64_t cap);

/* getFirstVarWithViolation (gnePivotRules.cpp:206-228): lowest violating position or -1       */
int gne_price_first(gne_dev *d, int64_t *pos);

/* computeEqfRedCost (genneteq.h:321-330) for p equal-flow variables: rc[k] = cost[k] - sum nodeVals_k[i] * pi[i], the
 * subtractions in node order as the reference's dense loop does them (zero entries change nothing, so the rows are
 * passed sparse: CSR, entries of a row sorted by node).  SURVEY 8 a12; the equal-flow engine calls it every pivot.    */
int gne_price_eqf(gne_dev *d, int p, const int64_t *rowStart, const int32_t *node, const double *val, const double *cost, double *rc);

/* The same without a copy or a stream synchronisation per call (what the equal-flow engine uses every pivot): the rows of all
 * p sets go to the device once (gne_eqf_set_rows); gne_price_eqf_launch names the q <= 128 sets to price (setIds: row numbers),
 * their costs and how many of the potentials the caller last set are not finite - as kernel arguments -, the kernel is queued
 * behind whatever the handle's stream holds and writes into mapped host memory; gne_price_eqf_fetch waits for it (another pass
 * may be launched in between: one wait then covers both).  GNE_ERR_CAPACITY from launch: use gne_price_eqf.            */
int gne_eqf_set_rows(gne_dev *d, int p, const int64_t *rowStart, const int32_t *node, const double *val);
int gne_price_eqf_launch(gne_dev *d, int q, const int32_t *setIds, const double *cost, int64_t nonFinitePotentials);
int gne_price_eqf_fetch(gne_dev *d, int q, double *rc);

/* reduced cost of every position in [lo, hi) (tests, verifyOptimality)                        */
int gne_reduced_costs(gne_dev *d, int64_t lo, int64_t hi, double *rc);

/* device time of the last pricing call's kernels in milliseconds (CUDA events on the handle's stream);
 * 0 unless timing is on (gne_dev_set_timing, or GNE_TIMING=1 in the environment at create time): the
 * events cost ~2 us per pivot, so the solver path leaves them off by default                    */
int gne_dev_last_kernel_ms(gne_dev *d, float *ms);
int gne_dev_set_timing(gne_dev *d, int on);
/* which kernel prices the LV pass: 0 = automatic (the TMA-streamed persistent kernel for shards of at
 * least ~2.4M arcs, the one-CTA-per-tile kernel below that or while streamed lists keep overflowing),
 * 1 = per-tile kernel, 2 = TMA-streamed kernel.  gne_dev_last_lv_kernel: 1 or 2, the one last used.  */
int gne_dev_set_lv_kernel(gne_dev *d, int mode);
/* CTAs of the persistent LV pass: 0 = two per SM of the whole device (default: one instance owns the GPU); a batch
 * of B independent instances on one GPU gives each handle ~(2 x SMs) / B so that all their passes are resident
 * together instead of queueing behind each other (GNE_STREAM_CTAS in the environment at create time does the same) */
int gne_dev_set_stream_ctas(gne_dev *d, int ctas);
/* "fast, non-parity" arithmetic: on != 0 lets the device fuse the multiply-add of the reduced cost
 * (genneteq.h:315-316) and of the potential finalize (gnePotential.cpp:130).  One rounding less per value: the
 * pivot sequence may leave the reference's at a tie; flows / potentials / objective agree to ~1e-9 relative.
 * Off by default (GNE_FAST_MATH=1 in the environment turns it on at create time); per device, not per handle. */
int gne_dev_set_fast_math(gne_dev *d, int on);
int gne_dev_last_lv_kernel(gne_dev *d);
/* number of kernels this handle has launched since creation                                   */
int64_t gne_dev_launch_count(gne_dev *d);
/* the CUDA device the handle lives on                                                          */
int gne_dev_cuda_device(const gne_dev *d);
/* replay timing helper for bench.py: runs `reps` device steps on the resident state — the relabel
 * (withFinalize 0: none, 1: finalize of every node, 2: replay of the handle's last argument-only
 * update, i.e. what a pivot of the current window really launches; falls back to 1), the LV pricing
 * pass, its reduction and, when sharded, the exchange — each step bracketed by CUDA events on the handle's stream (msStep) with a second
 * event pair around the tile kernel alone (msTiles).  flushL2 != 0 writes a >L2 scratch buffer
 * between steps, outside the events.                                                          */
int gne_bench_price_lv(gne_dev *d, int reps, int flushL2, int withFinalize, float *msStep, float *msTiles);

/* the LV pass `reps` times back to back between ONE pair of events: *msTotal / reps = the average duration of a launch in
 * a burst (no per-launch event bracket; single GPU)                                                              */
int gne_bench_price_lv_burst(gne_dev *d, int reps, int withFinalize, float *msTotal);

/* the same for the block rule: `reps` quickSelect passes as gne_price_qs issues them (cursor advancing), each timed
 * with CUDA events around its kernels; arcsLaunched = arcs the kernels covered, arcsScanned = arcs the rule consumed */
int gne_bench_price_qs(gne_dev *d, int reps, int flushL2, int64_t want, float *msStep, int64_t *arcsLaunched, int64_t *arcsScanned);

/* the device leg of a batch: `reps` rounds of LV passes over the B handles (each replaying its last argument-only relabel),
 * `perLaunch` instances per launch of the batched kernel, launches spread over `nStreams` launch groups; *msTotal = device
 * time of all rounds, *launches (may be NULL) = launches of the batched kernel, warm-up round included           */
int gne_bench_lv_groups(gne_dev **devs, int B, int reps, int perLaunch, int nStreams, float *msTotal, int64_t *launches);

/* --- multi-GPU exchange (one process per GPU; NCCL over NVLink) --- */
/* 128-byte opaque id made on rank 0 and broadcast by the caller (torch.distributed)           */
int gne_comm_unique_id(uint8_t id[128]);
int gne_comm_init(gne_dev *d, const uint8_t id[128], int rank, int nranks);
int gne_comm_destroy(gne_dev *d);
/* The same without NCCL: the host program supplies the all-gather (every rank contributes `bytes` bytes at `send`,
 * `recv` receives nranks x bytes in rank order; 0 = success).  Used for the bootstrap (CUDA IPC handles of the
 * mailboxes) and, when the mailboxes cannot be mapped on every rank or GNE_EXCHANGE=host, for the per-pivot
 * records themselves.  Lets several ranks share one device (NCCL refuses that): the sharded path can then be
 * exercised on a single-GPU box.                                                                      */
typedef int (*gne_allgather_fn)(void *ctx, const void *send, void *recv, int64_t bytes);
int gne_comm_init_host(gne_dev *d, int rank, int nranks, gne_allgather_fn allgather, void *ctx);
/* 0 = single GPU, 1 = per-pivot exchange through ncclAllGather, 2 = through peer mailboxes (NVLink stores from
 * the exchange kernel into CUDA-IPC-mapped buffers; default when IPC mapping succeeds on every rank;
 * GNE_EXCHANGE=nccl forces 1), 3 = through the host program's all-gather (gne_comm_init_host without mailboxes) */
int gne_comm_mode(gne_dev *d);

/* =============================== layer 2: solver mirror =============================== */
typedef struct gne_solver gne_solver;

/* GenNetEq::initialize (gneInit.cpp:25-44) from a sparse arc list sorted by (tail, head) —
 * the varIndex order of gneInit.cpp:82-107.  bigM as graph.cpp:43, or < 0 to have it computed. */
int gne_solver_create(gne_solver **out, int cudaDevice, int n, int64_t m, const int32_t *tail, const int32_t *head,
                      const double *ub, const double *cost, const double *mult, const double *supply, double bigM);
/* The same for an instance WITH equal-flow sets: eqf[a] = 1-based set of arc a, 0 for an arc outside
 * every set - the last field of an `a` line (graph.cpp:382-419); numSets = the `p` line's fourth number.  Each set is
 * one variable behind the arc variables (varIndex = arc variables + set, gneInit.cpp:149-162): its flow is the common
 * flow of its member arcs.  Such an instance runs in the equal-flow engine (Type II trees, pivotTII gnePivot.cpp:122-160,
 * the s x s systems of gnePotential.cpp:142-214 / gneFlow.cpp:212-256 with GSL's LU order restated); arcs and sets are
 * priced on the device.  Rule SD, batches and the window / progress / forest accessors return
 * GNE_ERR_UNSUPPORTED for it.  d_r(i) is accumulated in the order the arcs are given (a file: in file order).      */
int gne_solver_create_eqf(gne_solver **out, int cudaDevice, int n, int64_t m, const int32_t *tail, const int32_t *head,
                          const double *ub, const double *cost, const double *mult, const int32_t *eqf, int numSets,
                          const double *supply, double bigM);
/* the reference's text format (graph.cpp:358-447) read sparsely; a file with equal-flow sets gives the equal-flow engine */
int gne_solver_create_from_col(gne_solver **out, int cudaDevice, const char *path);
int gne_solver_destroy(gne_solver *s);
/* globals phaseIpivot / phaseIIpivot (main.cpp:21-22)                                        */
int gne_solver_set_rules(gne_solver *s, int phaseIpivot, int phaseIIpivot);
/* shard the pricing over ranks (call before the first solve); id from gne_comm_unique_id     */
int gne_solver_set_comm(gne_solver *s, const uint8_t id[128], int rank, int nranks);
int gne_solver_set_comm_host(gne_solver *s, int rank, int nranks, gne_allgather_fn allgather, void *ctx);
/* record (eVar, lVar) varIndex pairs of every pivot                                           */
int gne_solver_set_trace(gne_solver *s, int32_t *e, int32_t *l, int64_t cap);
int64_t gne_solver_trace_len(const gne_solver *s);
/* GenNetEq::solve(usePresolve) (genneteq.cpp:60-117).  maxPivots < 0: to termination.
 * *iters = loopIter; *finished = 0 when the pivot budget stopped it.                          */
int gne_solver_solve(gne_solver *s, int presolve, int64_t maxPivots, int64_t *iters, int *finished);
/* gne_solver_solve for B independent solvers at once (configs[3]: a batch of instances on one GPU), driven by `threads`
 * host threads (0 = one per core, at most B): each thread interleaves its share of the solvers - while one solver's
 * pricing pass runs on the device it does the host work (pivot, relabel) of another - so the host never oversubscribes
 * its cores and every instance stream stays busy.  iters / finished: arrays of B.                                  */
int gne_solver_solve_batch(gne_solver **s, int B, int presolve, int64_t maxPivots, int threads, int64_t *iters, int *finished);
/* the same with the LV passes of a thread's solvers sent in launch groups of up to `perLaunch` instances (1..8; 0 = one
 * launch per instance as above); *groupLaunches (may be NULL): launches of the batched kernel                      */
int gne_solver_solve_batch_grouped(gne_solver **s, int B, int presolve, int64_t maxPivots, int threads, int perLaunch,
                                   int64_t *iters, int *finished, int64_t *groupLaunches);

/* the same with a starting line: alignPivot >= 0 holds every solver before its pivot number alignPivot (of this call) until
 * all solvers of the batch have got there or ended, so that what follows runs side by side whatever the instances' first
 * pivots cost (a phase's first relabel sweeps every node)                                                          */
int gne_solver_solve_batch_ex(gne_solver **s, int B, int presolve, int64_t maxPivots, int threads, int perLaunch, int64_t alignPivot,
                              int64_t *iters, int *finished, int64_t *groupLaunches);

int gne_solver_num_nodes(const gne_solver *s);
int64_t gne_solver_num_vars(const gne_solver *s);
int gne_solver_num_sets(const gne_solver *s);      /* equal-flow variables (the last ones of the variable vector) */
double gne_solver_objective(const gne_solver *s);
int gne_solver_feasible(const gne_solver *s);
int gne_solver_get_flows(gne_solver *s, double *out);
int gne_solver_get_potentials(gne_solver *s, double *out);
/* GenNetEq::computePotentials() (gnePotential.cpp:28-61) for the CURRENT basis: relabels the nodes the last pivot
 * touched and brings the device's potentials and nonbasic mirror up to date, i.e. the state the next pricing pass
 * will read.  Idempotent, consumes no drand48 draw, does not change the pivot sequence: it is the hook for the
 * same-state A/B of SURVEY 8(d)(ii) (read the state back with gne_nb_read / gne_pot_get_all on gne_solver_device). */
int gne_solver_compute_potentials(gne_solver *s);
int gne_solver_get_states(gne_solver *s, int32_t *out);
int gne_solver_get_forest(gne_solver *s, int32_t *tree, int32_t *pred, int32_t *predArc, int32_t *depth, int32_t *thread);
/* counters: [0] pricing s (host wall incl. sync)  [1] potentials s  [2] pivot s  [3] total s
 *           [4] arcs priced  [5] pricing kernel ms (device)  [6] H2D bytes  [7] D2H bytes
 *           [8] kernel launches  [9] pivots  [10] tree surgery s  [11] flow/ratio s           */
int gne_solver_get_counters(const gne_solver *s, double *out12);
gne_dev *gne_solver_device(gne_solver *s);
/* wall-clock window over pivots [startPivot, startPivot + numPivots) of the next gne_solver_solve
 * call, taken inside the loop with the device synchronised at both edges (bench.py's e2e):
 * out27 = { seconds (-1 if the window was not completed), the 12 counters at its start, at its end,
 *          CLOCK_MONOTONIC seconds of the two edges } */
int gne_solver_set_window(gne_solver *s, int64_t startPivot, int64_t numPivots);
int gne_solver_get_window(const gne_solver *s, double *out27);
/* progress log: every `every` pivots of a solve call append a row of 14 doubles to buf —
 * { wall-clock seconds, the 12 counters, size of the largest basis tree }                     */
int gne_solver_set_progress(gne_solver *s, int64_t every, double *buf, int64_t capRows);
int64_t gne_solver_progress_rows(const gne_solver *s);

/* Host-only replay of basis exchanges (no device work, usable without a GPU): applies the given
 * (entering, leaving) varIndex pairs to a fresh forest and returns its thread/pred arrays —
 * lets the forest engine be checked against the oracle on CPU.                               */
int gne_forest_replay(int n, int64_t m, const int32_t *tail, const int32_t *head, const double *supply,
                      int64_t pivots, const int32_t *e, const int32_t *l,
                      int32_t *tree, int32_t *pred, int32_t *predArc, int32_t *depth, int32_t *thread);

/* Host-only replay of a recorded solve (no device, no pricing): entering arcs come from the trace,
 * flows / ratio test / leaving arc / basis update / potential sweep run as in gne_solver_solve and
 * the leaving arc is compared with the trace (*mismatch = first differing pivot or -1).
 * drawsPerPricing = drand48 calls the rule makes per call (LV 1, QS/CL/FV 0, SAA/SAE 2, RV/RWV/RLV 1). */
int gne_host_replay(int n, int64_t m, const int32_t *tail, const int32_t *head, const double *ub, const double *cost,
                    const double *mult, const double *supply, double bigM, int drawsPerPricing, int sparseFlows,
                    int64_t phase1Pivots, int64_t totalPivots, const int32_t *e, const int32_t *l,
                    int64_t *mismatch, double *flowsOut, double *potentialsOut, double *objective);

/* the same for an instance with equal-flow sets (arguments as gne_solver_create_eqf): Type II trees, pivotTII, both s x s
 * systems, tree ids - everything but the pricing - against a recorded trace, without a device                     */
int gne_host_replay_eqf(int n, int64_t m, const int32_t *tail, const int32_t *head, const double *ub, const double *cost,
                        const double *mult, const int32_t *eqf, int numSets, const double *supply, double bigM,
                        int drawsPhaseI, int drawsPhaseII, int64_t phase1Pivots, int64_t totalPivots, const int32_t *e, const int32_t *l,
                        int64_t *mismatch, double *flowsOut, double *potentialsOut, double *objective);

/* the same from a text file, read the way gne_solver_create_from_col reads it (d_r(i) in file order, duplicate lines
 * overwriting); numVars = size of flowsOut (arc variables + sets), potentialsOut holds n values                       */
int gne_host_replay_eqf_col(const char *path, int drawsPhaseI, int drawsPhaseII, int64_t phase1Pivots, int64_t totalPivots,
                            const int32_t *e, const int32_t *l, int64_t *mismatch, int64_t numVars, double *flowsOut,
                            double *potentialsOut, double *objective);

/* the s x s solve of the equal-flow engine on its own (host-only; tests): LU with partial pivoting in GSL 1.x's operation
 * order (gsl_linalg_LU_decomp + gsl_linalg_LU_solve, as gnePotential.cpp:196-197 and gneFlow.cpp:236-237 call them).
 * a: row-major s x s, b: right-hand side, x: solution; lu (s x s) and perm (s) may be NULL                          */
int gne_host_lu_solve(int s, const double *a, const double *b, double *x, double *lu, int32_t *perm);

/* the contiguous position range [*lo, *hi) of the nonbasic list (m real arcs) that `rank` of `nranks` prices:
 * boundaries on multiples of 2048, ranges never overlap, *lo == *hi for a rank left without positions
 * (lists shorter than 2048 x nranks).  What gne_solver_set_comm uses; host-only.                     */
int gne_shard_range(int64_t m, int rank, int nranks, int64_t *lo, int64_t *hi);

/* merge of per-shard LV records (what every rank does after the all-gather); host-only.
 * maxima[r], any[r], counts[r], and lists (pos, viol) concatenated rank-major with stride cap. */
int gne_merge_lv_shards(int nranks, const double *maxima, const int *any, const int64_t *counts,
                        const int64_t *listPos, const double *listViol, int64_t stride,
                        double *largestViolation, int64_t *count, int64_t *outPos, int64_t cap, int *outAny, int *needFallback);

/* instance files.  Text: the reference's format (graph.cpp:358-447), read sparsely, written like
 * Graph::exportColFile (graph.cpp:94-121).  gne_read_col: call with tail == NULL for the sizes (*n, *m,
 * *bigM as graph.cpp:43), then with arrays of those sizes; arcs come back sorted by (tail, head), arcs
 * with capacity <= 0 dropped (gneInit.cpp:84).  Binary: "GNESOA1" + n, m + the six columns.          */
int gne_write_col(const char *path, int n, int64_t m, const int32_t *tail, const int32_t *head, const double *ub,
                  const double *cost, const double *mult, const double *supply);
int gne_read_col(const char *path, int *n, int64_t *m, int32_t *tail, int32_t *head, double *ub, double *cost,
                 double *mult, double *supply, double *bigM);
/* the same for a file with equal-flow sets (gne_read_col answers GNE_ERR_UNSUPPORTED for one): eqf[a] = 1-based set of
 * arc a or 0, *numSets = the problem line's count.  (d_r(i) is summed in file order only by gne_solver_create_from_col;
 * the arrays come back in (tail, head) order.)                                                                  */
int gne_read_col_eqf(const char *path, int *n, int64_t *m, int *numSets, int32_t *tail, int32_t *head, double *ub,
                     double *cost, double *mult, int32_t *eqf, double *supply, double *bigM);
int gne_write_bin(const char *path, int n, int64_t m, const int32_t *tail, const int32_t *head, const double *ub,
                  const double *cost, const double *mult, const double *supply);
int gne_read_bin_header(const char *path, int *n, int64_t *m);
int gne_read_bin(const char *path, int n, int64_t m, int32_t *tail, int32_t *head, double *ub, double *cost,
                 double *mult, double *supply);

/* deterministic large-instance generator (bench workloads C2-C5):
 * ring i->i+1 plus random distinct heads, arcs sorted by (tail, head).  Arrays sized m / n.   */
int gne_generate_instance(int n, int64_t m, uint64_t seed, int mode /*0 wide 1 lossy 2 near1 3 pow2*/,
                          int32_t *tail, int32_t *head, double *ub, double *cost, double *mult, double *supply);

#ifdef __cplusplus
}
#endif
#endif
