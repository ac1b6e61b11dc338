// TEST INFRASTRUCTURE — driver around the UNMODIFIED reference solver.
//
// This file is compiled together with the reference's own translation units taken where
// they lie under /root/reference (see oracle/Makefile); nothing of the reference is copied
// into this repository.  It only adds what the reference lacks for use as an oracle:
//
//   * a per-pivot trace "P <loopIter> <eVar->varIndex> <lVar->varIndex>" — the line the
//     reference itself keeps commented out at gnePivot.cpp:49-50;
//   * a sparse initialiser, because Graph stores dense n x n matrices (graph.cpp:214-227)
//     and GenNetEq::initializeVars walks all n^2 pairs (gneInit.cpp:82-107);
//   * access to the end state (flows, potentials, objective) and timing of the pricing call.
//
// All of GenNetEq's members are protected and none is virtual, so a derived class can drive
// the reference's compiled pricing / potential / pivot code without editing any file.  The
// only logic restated here is the body of GenNetEq::solve (genneteq.cpp:60-117) — needed to
// place the trace line after pivot() — and, for the sparse path, gneInit.cpp:25-44, 46-56,
// 82-127, 149-161, 174-185 with the (i,j) pair walk replaced by a walk over the arc list
// sorted by (tail, head), which visits existing arcs in exactly the same order.
//
// Nothing under genneteq_b200/ may include, link or call this file.
#include "main.h"
#include "matrix.h"
#include "graph.h"
#include "variables.h"
#include "trees.h"
#include "genneteq.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <limits>
#include <vector>
#include <time.h>

// The reference defines these in main.cpp:21-22; main.cpp is not part of the driver build.
int phaseIpivot = LVW;
int phaseIIpivot = LVW;

static double wallNow()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1.0e-9 * (double) ts.tv_nsec;
}

class RefDriver : public GenNetEq
{
  public:
    RefDriver() : traceE(NULL), traceL(NULL), traceCap(0), traceLen(0), traceFile(NULL),
                  secsPricing(0.0), secsPotentials(0.0), secsPivot(0.0), secsTotal(0.0),
                  arcsPriced(0), pivotsDone(0) {}

    // ---- sparse restatement of GenNetEq::initialize (gneInit.cpp:25-186, SELFLOOPS path)
    void initializeSparse(int n, long m, const int *tail, const int *head,
                          const double *ub, const double *cost, const double *mult,
                          const double *supply, double bigMIn)
    {
        numNodes = n;                                   // gneInit.cpp:27
        numArcs = 0;
        numEqualFlowSets = 0;
        bigM = bigMIn;                                  // graph.cpp:43 (computed by caller)
        initialType = SELFLOOPS;                        // gneInit.cpp:31 (hard-coded there)

        nodes.resize(numNodes);                         // gneInit.cpp:46-56
        for (int i = 0; i < numNodes; ++i) {
            nodes[i].id = i;
            nodes[i].potential = 0.0;
            nodes[i].supply = supply[i];
        }
        Tree::setStaticVariables(&nodes, &setBtI, &setBtII);   // gneInit.cpp:62

        for (long a = 0; a < m; ++a) {                  // gneInit.cpp:82-107
            if (ub[a] <= 0.0) continue;                 // "arc doesn't exist" (:84)
            ArcVar *av = new ArcVar();
            av->tail = tail[a];
            av->head = head[a];
            av->LB = 0.0;
            av->UB = ub[a];
            av->actualCost = cost[a];
            av->mult = mult[a];
            av->varType = ARC_VAR;
            av->isArtificial = false;
            arcVars.push_back(av);
        }
        for (int i = 0; i < numNodes; ++i) {            // gneInit.cpp:110-127
            ArcVar *av = new ArcVar();
            av->tail = i;
            av->head = i;
            av->LB = 0.0;
            av->UB = std::numeric_limits<double>::max();
            av->actualCost = bigM;
            av->mult = (supply[i] >= 0.0) ? 0.5 : 2.0;
            av->varType = ARC_VAR;
            av->isArtificial = true;
            arcVars.push_back(av);
        }
        numArcs = arcVars.size();                       // gneInit.cpp:146
        for (size_t a = 0; a < arcVars.size(); ++a) {   // gneInit.cpp:149-155
            arcVars[a]->varIndex = (int) a;
            vars.push_back(arcVars[a]);
        }
        for (size_t a = 0; a < arcVars.size(); ++a) {   // gneInit.cpp:174-185
            if (arcVars[a]->isArtificial) insertIntoB(arcVars[a]);
            else insertIntoL(arcVars[a]);
        }
        Tree::updateTrees();
        phaseIiters = 0;                                // gneInit.cpp:40-41
        phaseIIiters = 0;
    }

    // ---- GenNetEq::solve (genneteq.cpp:60-117) with the trace line and a pivot budget.
    // Returns the number of pivots of this phase; *finished = 0 when the budget cut it short.
    long solveTraced(bool usePresolve, long maxPivots, int *finished)
    {
        bool pIwasInfeasible = usePresolve ? false : isInfeasible;
        initializeStats(usePresolve);
        setVariableCosts();
        *finished = 1;
        if (!presolve && pIwasInfeasible) {
            isInfeasible = true;
            printTerminationStatus();
            return loopIter;
        }
        double t0 = wallNow();
        computeFlows();
        long budget = maxPivots;
        long callPivots = 0;
        while ((!isOptimal) && (!isUnbounded)) {
            if (callPivots == windowStart) { winTimers[0] = secsPricing; winTimers[1] = secsPotentials; winTimers[2] = secsPivot; winTimers[3] = 0.0; winTimers[4] = (double) arcsPriced; }
            if (maxPivots >= 0 && budget == 0) { *finished = 0; break; }
            if ((loopIter % 1000) == 0) {
                computeFlows(2);
            }
            double a = wallNow();
            computePotentials();
            double b = wallNow();
            long nbBefore = (long) setNBarc.size();
            identifyEnteringVariable();
            double c = wallNow();
            secsPotentials += b - a;
            secsPricing += c - b;
            arcsPriced += nbBefore;
            if (eVar != NULL) {
                pivot();
                secsPivot += wallNow() - c;
                if (lVar != NULL) {
                    if (traceFile) fprintf(traceFile, "P %d %d %d\n", loopIter, eVar->varIndex, lVar->varIndex);
                    if (traceLen < traceCap) {
                        traceE[traceLen] = eVar->varIndex;
                        traceL[traceLen] = lVar->varIndex;
                    }
                    ++traceLen;
                }
                ++loopIter;
                ++pivotsDone;
                ++callPivots;
                presolve ? ++phaseIiters : ++phaseIIiters;
                if (budget > 0) --budget;
            }
        }
        if (*finished) {
            computeFlows();
            if (presolve) checkFeasibility();
            verifyOptimality();
            printTerminationStatus();
        }
        secsTotal += wallNow() - t0;
        return loopIter;
    }

    // ---- replay of a recorded pivot sequence with the reference's OWN pivot(): the entering arc of each pivot comes
    // from the trace (no pricing pass, no potentials - pivot() reads neither), flows / ratio test / leaving arc / basis
    // and tree updates are the reference's, and the leaving arc it picks is compared with the trace.  `draws` = drand48
    // calls the rule would have made per pricing call (LV 1, QS / CL / FV 0); the cycling override draws once
    // (gnePivotRules.cpp:307-362).  Reaches a state deep in a solve in minutes instead of hours; returns the index of
    // the first pivot whose leaving arc differs from the trace, or -1.
    long replayTraced(bool usePresolve, const int *e, const int *l, long count, int draws)
    {
        initializeStats(usePresolve);
        setVariableCosts();
        computeFlows();
        for (long k = 0; k < count; ++k) {
            if ((loopIter % 1000) == 0) computeFlows(2);
            if (cyclingDetected) (void) drand48();
            else for (int d = 0; d < draws; ++d) (void) drand48();
            eVar = vars[e[k]];
            isOptimal = false;
            pivot();
            if (lVar == NULL || lVar->varIndex != l[k]) return k;
            ++loopIter;
            ++pivotsDone;
            presolve ? ++phaseIiters : ++phaseIIiters;
        }
        return -1;
    }

    // ---- one pricing call at the current state, leaving the solver state untouched
    // (drand48 stream and rule cursors saved/restored), for same-state A/B timing.
    int priceOnce(double *secs)
    {
        unsigned short saved[3];
        unsigned short *prev = seed48(saved);           // returns pointer to the old state
        unsigned short old[3] = { prev[0], prev[1], prev[2] };
        seed48(old);
        int c1 = nbArcPivotInd, c2 = nbEqfPivotInd, c3 = itersSinceUpdate;
        double lv = lastViolation;
        bool opt = isOptimal;
        double a = wallNow();
        Variable *v = selectEnteringVariable();
        *secs = wallNow() - a;
        seed48(old);
        nbArcPivotInd = c1; nbEqfPivotInd = c2; itersSinceUpdate = c3;
        lastViolation = lv; isOptimal = opt;
        return v ? v->varIndex : -1;
    }

    void potentialsOnce(double *secs)
    {
        double a = wallNow();
        computePotentials();
        *secs = wallNow() - a;
    }

    int nNodes() const { return numNodes; }
    long nVars() const { return (long) vars.size(); }
    long nNonbasic() const { return (long) setNBarc.size(); }
    double objective() const { return flowCost; }
    bool feasible() const { return !isInfeasible; }
    bool optimal() const { return isOptimal; }
    void getFlows(double *out) const { for (size_t i = 0; i < vars.size(); ++i) out[i] = vars[i]->flow; }
    void getPotentials(double *out) const { for (int i = 0; i < numNodes; ++i) out[i] = nodes[i].potential; }
    void getStates(int *out) const { for (size_t i = 0; i < vars.size(); ++i) out[i] = (int) vars[i]->setLoc; }
    void getNonbasicOrder(int *out) const { for (size_t i = 0; i < setNBarc.size(); ++i) out[i] = setNBarc[i]->varIndex; }

    long windowStart = -1;
    double winTimers[5] = { 0, 0, 0, 0, 0 };
    int *traceE, *traceL;
    long traceCap, traceLen;
    FILE *traceFile;
    double secsPricing, secsPotentials, secsPivot, secsTotal;
    long arcsPriced, pivotsDone;
};

// ------------------------------------------------------------------------------------------
// C interface (ctypes from tests/ and bench.py's reference arm).  Tree keeps static pointers
// to one solver (trees.h:83-86), so there is exactly one live solver per process.
// ------------------------------------------------------------------------------------------
static RefDriver *g_solver = NULL;

extern "C" {

int ref_init_sparse(int n, long m, const int *tail, const int *head, const double *ub,
                    const double *cost, const double *mult, const double *supply, double bigM)
{
    if (g_solver) return -1;
    g_solver = new RefDriver();
    g_solver->initializeSparse(n, m, tail, head, ub, cost, mult, supply, bigM);
    return 0;
}

int ref_init_col(const char *path)
{
    if (g_solver) return -1;
    Config conf;
    conf.initType = INFEASIBLE;     // main.cpp:74; ignored by gneInit.cpp:31
    conf.debug = 0;
    conf.inFile = path;
    Graph *g = new Graph();
    g->initialize(&conf);
    g_solver = new RefDriver();
    g_solver->initialize(g, conf.initType);
    delete g;
    return 0;
}

void ref_set_rules(int p1, int p2) { phaseIpivot = p1; phaseIIpivot = p2; }

void ref_set_trace(int *e, int *l, long cap)
{
    g_solver->traceE = e; g_solver->traceL = l; g_solver->traceCap = cap; g_solver->traceLen = 0;
}

long ref_trace_len(void) { return g_solver->traceLen; }

long ref_solve(int presolve, long maxPivots, int *finished)
{
    return g_solver->solveTraced(presolve != 0, maxPivots, finished);
}

long ref_replay(int presolve, const int *e, const int *l, long count, int draws)
{
    return g_solver->replayTraced(presolve != 0, e, l, count, draws);
}

int ref_price_once(double *secs) { return g_solver->priceOnce(secs); }
void ref_potentials_once(double *secs) { g_solver->potentialsOnce(secs); }

void ref_get_timers(double *out5)
{
    out5[0] = g_solver->secsPricing; out5[1] = g_solver->secsPotentials;
    out5[2] = g_solver->secsPivot;   out5[3] = g_solver->secsTotal;
    out5[4] = (double) g_solver->arcsPriced;
}

/* timers as they stood when pivot `start` of the next ref_solve call began */
void ref_set_window(long start) { g_solver->windowStart = start; }
void ref_get_window_timers(double *out5) { for (int i = 0; i < 5; ++i) out5[i] = g_solver->winTimers[i]; }

int ref_num_nodes(void) { return g_solver->nNodes(); }
long ref_num_vars(void) { return g_solver->nVars(); }
long ref_num_nonbasic(void) { return g_solver->nNonbasic(); }
double ref_objective(void) { return g_solver->objective(); }
int ref_feasible(void) { return g_solver->feasible() ? 1 : 0; }
void ref_get_flows(double *out) { g_solver->getFlows(out); }
void ref_get_potentials(double *out) { g_solver->getPotentials(out); }
void ref_get_states(int *out) { g_solver->getStates(out); }
void ref_get_nonbasic_order(int *out) { g_solver->getNonbasicOrder(out); }

void ref_destroy(void) { delete g_solver; g_solver = NULL; }

} // extern "C"

// ------------------------------------------------------------------------------------------
// CLI:  gne_ref_driver [-1 rule] [-2 rule] [-K maxPivotsPerPhase] [-T trace.txt] [-D dump.txt] file.col
// ------------------------------------------------------------------------------------------
#ifdef GNE_REF_DRIVER_MAIN
int main(int argc, char **argv)
{
    const char *tracePath = NULL, *dumpPath = NULL, *inPath = NULL;
    long maxPivots = -1;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-1") && i + 1 < argc) phaseIpivot = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-2") && i + 1 < argc) phaseIIpivot = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-K") && i + 1 < argc) maxPivots = atol(argv[++i]);
        else if (!strcmp(argv[i], "-T") && i + 1 < argc) tracePath = argv[++i];
        else if (!strcmp(argv[i], "-D") && i + 1 < argc) dumpPath = argv[++i];
        else inPath = argv[i];
    }
    if (!inPath) {
        fprintf(stderr, "usage: %s [-1 r] [-2 r] [-K pivots] [-T trace] [-D dump] file.col\n", argv[0]);
        return 2;
    }
    ref_init_col(inPath);
    if (tracePath) g_solver->traceFile = fopen(tracePath, "w");
    int fin1 = 0, fin2 = 0;
    long p1 = ref_solve(1, maxPivots, &fin1);
    long p2 = fin1 ? ref_solve(0, maxPivots, &fin2) : 0;
    if (g_solver->traceFile) fclose(g_solver->traceFile);
    printf("REF pivots %ld %ld finished %d %d objective %.17g\n", p1, p2, fin1, fin2, ref_objective());
    if (dumpPath) {
        FILE *f = fopen(dumpPath, "w");
        long nv = ref_num_vars();
        int nn = ref_num_nodes();
        std::vector<double> x(nv), pi(nn);
        ref_get_flows(&x[0]);
        ref_get_potentials(&pi[0]);
        fprintf(f, "objective %.17g\n", ref_objective());
        for (long i = 0; i < nv; ++i) fprintf(f, "x %ld %.17g\n", i, x[i]);
        for (int i = 0; i < nn; ++i) fprintf(f, "pi %d %.17g\n", i, pi[i]);
        fclose(f);
    }
    return 0;
}
#endif
