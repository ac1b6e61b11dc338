/* TEST INFRASTRUCTURE — C interface of the CPU restatement (oracle/gne_oracle.cpp).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this.  Nothing under genneteq_b200/ includes or links it.
 */
#ifndef GNE_ORACLE_H
#define GNE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcSolver OrcSolver;

/* graph.cpp:43  bigM = numArcs * maxArcCost * maxArcCapacity */
double orc_big_m(long m, const double *cost, const double *ub);
/* computeEqfRedCost (genneteq.h:321-330) over dense nodeVals rows (p x n) */
void orc_price_eqf_dense(int p, int n, const double *nodeVals, const double *cost, const double *pi, double *rc);

/* arcs must be sorted by (tail, head): that is the varIndex order of gneInit.cpp:82-107 */
OrcSolver *orc_create(int n, long m, const int *tail, const int *head, const double *ub,
                      const double *cost, const double *mult, const double *supply, double bigM);
void orc_destroy(OrcSolver *s);
void orc_set_rules(OrcSolver *s, int phaseIpivot, int phaseIIpivot);
void orc_set_trace(OrcSolver *s, int *e, int *l, long cap);
long orc_trace_len(const OrcSolver *s);
/* one phase (genneteq.cpp:60-117); maxPivots < 0 = run to termination.  Returns loopIter. */
long orc_solve(OrcSolver *s, int presolve, long maxPivots, int *finished);
/* status: 0 ok, <0 = the reference would have exit()ed (trees.cpp:310-313 etc.) */
int orc_status(const OrcSolver *s);

int orc_num_nodes(const OrcSolver *s);
long orc_num_vars(const OrcSolver *s);
long orc_num_nonbasic(const OrcSolver *s);
double orc_objective(const OrcSolver *s);
int orc_feasible(const OrcSolver *s);
void orc_get_flows(const OrcSolver *s, double *out);
void orc_get_potentials(const OrcSolver *s, double *out);
void orc_get_states(const OrcSolver *s, int *out);            /* var_loc: 0 B, 1 L, 2 U */
void orc_get_nonbasic_order(const OrcSolver *s, int *out);    /* setNBarc[pos] -> varIndex */
void orc_get_costs(const OrcSolver *s, double *out);
void orc_get_timers(const OrcSolver *s, double *out5);        /* pricing, potentials, pivot, total s; arcs priced */
/* the timers as they stood when pivot `start` of the next orc_solve call began (bench windows) */
void orc_set_window(OrcSolver *s, long start);
void orc_get_window_timers(const OrcSolver *s, double *out5);
/* forest snapshot for checking a host tree engine: per node tree id, pred, predArc, depth, thread */
void orc_get_forest(const OrcSolver *s, int *treeId, int *pred, int *predArc, int *depth, int *thread);
/* recompute potentials from the current basis (gnePotential.cpp:28-61) */
void orc_compute_potentials(OrcSolver *s);

/* glibc drand48, restated (X' = 0x5DEECE66D * X + 0xB mod 2^48, result X'/2^48).  The reference
 * never calls srand48; glibc's never-seeded state is X = 0 (first draw 0xB / 2^48), which the
 * tests check against this image's libc.  state = the 48-bit X. */
double orc_drand48(unsigned long long *state);
unsigned long long orc_drand48_initial_state(void);

/* ---- stateless kernels on a nonbasic structure-of-arrays in setNBarc position order -------
 * state[pos]: 1 = at lower bound (L_var), 2 = at upper bound (U_var).  tol = 1e-13 (main.h:36-43).
 * rc = (cost - pi[tail]) + (mult * pi[head]), no FMA (genneteq.h:312-318).                   */

void orc_reduced_costs(long N, const int *tail, const int *head, const double *cost,
                       const double *mult, const double *pi, double *rc);

/* gnePivotRules.cpp:401-413 + 431-452: the sequential tolerance-band rule.  Returns the list
 * length; writes positions (up to cap) and the final running largestViolation.            */
long orc_price_lv(long N, const int *tail, const int *head, const double *cost, const double *mult,
                  const signed char *state, const double *pi, long *listPos, long cap,
                  double *largest, int *anyViolation);

/* gnePivotRules.cpp:710-764 (arc loop): scan from cursor until `want` violators were seen or
 * the list is exhausted; strictly-greater max.  Outputs best position (-1 if none), its
 * violation, the cursor after the scan (before the majorIteration restore) and arcs scanned. */
void orc_price_qs(long N, const int *tail, const int *head, const double *cost, const double *mult,
                  const signed char *state, const double *pi, long cursor, long want,
                  long *bestPos, double *bestViol, long *newCursor, long *scanned, long *violations);

/* all violators in position order (gnePivotRules.cpp:504-516, gnePotential.cpp:226-234);
 * optionally only those with |rc| >= threshold.  Returns the count; writes up to cap.      */
long orc_price_violators(long N, const int *tail, const int *head, const double *cost,
                         const double *mult, const signed char *state, const double *pi,
                         double threshold, long *listPos, double *listViol, long cap);

/* gnePotential.cpp:127-135: pi[node[i]] = a0[i] + a1[i] * theta (mul then add, no FMA)        */
void orc_finalize_potentials(long count, const int *node, const double *a0, const double *a1,
                             double theta, double *pi);

/* gnePivot.cpp:79-84: flow += delta * deltaFlow over a list of arcs; returns the flowCost
 * increment accumulated in list order.                                                     */
double orc_flow_apply(long count, const int *arc, const double *deltaFlow, const double *cost,
                      double delta, double *flow);

#ifdef __cplusplus
}
#endif
#endif
