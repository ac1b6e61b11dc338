// TEST INFRASTRUCTURE — the drop-in claim, executed: the UNMODIFIED reference solver (its own
// Graph loader, initialisation, potentials, flows, ratio test, tree surgery, libc drand48) with
// only the pricing pass routed through genneteq_b200's C ABI (layer 1), exactly as INTEGRATION.md
// section A describes.  Compiled by oracle/Makefile (`ref` target) against the reference's headers
// and objects where they lie under /root/reference; nothing of the reference is copied.
//
//   gne_gpu_driver -1 <rule> -2 <rule> -T trace.txt [-A every] file.col        rules: 0/1 (LV/LVW), 16/17 (QS/QSW)
//
// Instances WITH equal-flow sets run too (the reference's own Type II trees, pivotTII and s x s solves - GSL's LU through
// oracle/gsl_mini): the arcs are priced by gne_price_lv / gne_price_qs, the nonbasic sets by gne_price_eqf, and the
// reference's own checkVarForOffending continues over the sets' values (computeEqfsWithLargestViolation, :415-429).
//
// -A every: same-state A/B.  At every `every`-th iteration the reference's OWN
// selectEnteringVariable() (host, gnePivotRules.cpp:36-127) and the GPU path price the same state - the drand48
// state (seed48), the quickSelect cursor and the optimality flag are saved and restored in between - the two
// entering variables must be the same object, and both calls are timed.
//
// The pivot trace must equal the reference's own (tests/test_integration.py).
#include "main.h"
#include "matrix.h"
#include "graph.h"
#include "variables.h"
#include "trees.h"
#include "genneteq.h"
#include "../../include/genneteq_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

int phaseIpivot = LVW;          // main.cpp:21-22 (main.cpp is not part of this build)
int phaseIIpivot = LVW;

#define GCK(call) do { int rc_ = (call); if (rc_) { fprintf(stderr, "%s failed: %d %s\n", #call, rc_, gne_last_error()); exit(3); } } while (0)

class GpuGenNetEq : public GenNetEq
{
    gne_dev *dev;
    std::vector<int64_t> pos;

    void upload()                                           // after setVariableCosts(), genneteq.cpp:66
    {
        const int64_t N = (int64_t) setNBarc.size();
        if (!dev) GCK(gne_dev_create(&dev, 0, numNodes, N + 16, 0, N + 16));
        std::vector<int32_t> t(N), h(N);
        std::vector<double> c(N), m(N);
        std::vector<int8_t> s(N);
        for (int64_t p = 0; p < N; ++p) {                   // setNBarc POSITION order (gnePivotRules.cpp:401-413)
            ArcVar *a = setNBarc[p];
            t[p] = a->tail; h[p] = a->head; c[p] = a->cost; m[p] = a->mult; s[p] = (int8_t) a->setLoc;
        }
        GCK(gne_nb_reset(dev, N, t.data(), h.data(), c.data(), m.data(), s.data()));
    }
    void mirrorWrite(int64_t p)                             // after the set moves of updateBasis (gnePivot.cpp:254-309)
    {
        if (p < 0 || p >= (int64_t) setNBarc.size()) return;
        ArcVar *a = setNBarc[p];
        GCK(gne_nb_write(dev, p, a->tail, a->head, a->cost, a->mult, (int8_t) a->setLoc));
    }
    // after computePotentials(), genneteq.cpp:97.  The reference recomputes every node's potential each iteration, but
    // only the trees the last pivot touched get new values: a host-side shadow copy finds them (one pass over n doubles,
    // cheap next to the reference's own O(n) sweep) and only those cross PCIe - the dirty-set hook of mode A.
    void pushPotentials()
    {
        if ((int) shadowPi.size() != numNodes) {            // first call: everything
            shadowPi.resize(numNodes);
            for (int i = 0; i < numNodes; ++i) shadowPi[i] = nodes[i].potential;
            GCK(gne_pot_set_all(dev, shadowPi.data()));
            return;
        }
        dirtyIdx.clear(); dirtyVal.clear();
        for (int i = 0; i < numNodes; ++i) {
            const double v = nodes[i].potential;
            if (memcmp(&v, &shadowPi[i], sizeof(double)) != 0) { shadowPi[i] = v; dirtyIdx.push_back(i); dirtyVal.push_back(v); }
        }
        potentialBytes += 12 * (long) dirtyIdx.size();
        if (!dirtyIdx.empty()) GCK(gne_pot_set(dev, (int64_t) dirtyIdx.size(), dirtyIdx.data(), dirtyVal.data()));
    }
    std::vector<double> shadowPi, dirtyVal;
    std::vector<int32_t> dirtyIdx;
    long potentialBytes = 0;
    // computeEqfRedCost (genneteq.h:321-330) of every nonbasic set on the device: rows = the sets' nodeVals, sparse
    std::vector<double> setRc;
    void priceSets()
    {
        const int q = (int) setNBeqf.size();
        setRc.assign(q, 0.0);
        if (q == 0) return;
        std::vector<int64_t> rowStart(q + 1, 0);
        std::vector<int32_t> node;
        std::vector<double> val, cost(q);
        for (int k = 0; k < q; ++k) {
            EqFlowVar *v = setNBeqf[k];
            for (int i = 0; i < numNodes; ++i) if (v->nodeVals[i] != 0.0) { node.push_back(i); val.push_back(v->nodeVals[i]); }
            rowStart[k + 1] = (int64_t) node.size();
            cost[k] = v->cost;
        }
        GCK(gne_price_eqf(dev, q, rowStart.data(), node.data(), val.data(), cost.data(), setRc.data()));
    }
    static bool violates(Variable *v, double rc)
    {
        return ((v->setLoc == L_var) && (rc < (-1.0 * ROUNDOFF_TOLERANCE))) || ((v->setLoc == U_var) && (rc > ROUNDOFF_TOLERANCE));
    }
    Variable *gpuLargestViolation(bool useWeighted)         // getVarWithLargestViolation, gnePivotRules.cpp:133-157
    {
        double L; int64_t cnt; int any;
        pos.resize(setNBarc.size() + 16);
        GCK(gne_price_lv(dev, &L, &cnt, pos.data(), (int64_t) pos.size(), &any));
        isOptimal = !any;
        offendingVars.clear();
        for (int64_t i = 0; i < cnt; ++i) offendingVars.push_back(setNBarc[pos[i]]);
        if (!setNBeqf.empty()) {                            // computeEqfsWithLargestViolation (:415-429) with the device's values
            priceSets();
            for (size_t k = 0; k < setNBeqf.size(); ++k) {
                EqFlowVar *nbeq = setNBeqf[k];
                const double rc = setRc[k];
                if (!violates(nbeq, rc)) continue;
                isOptimal = false;
                const double frc = useWeighted ? fabs(rc) / nbeq->numOrigArcs : fabs(rc);
                // checkVarForOffending (:431-452; defined inline in gnePivotRules.cpp, so not linkable - its two lines, with
                // trackLowestVarIndex off as genneteq.cpp:144 sets it), continuing the arc pass's state
                if (L < frc - ROUNDOFF_TOLERANCE) offendingVars.clear();
                if (L <= frc + ROUNDOFF_TOLERANCE) { L = frc; offendingVars.push_back(nbeq); }
            }
        }
        lastViolation = L;
        if (offendingVars.empty()) return NULL;
        return offendingVars[drand48() * offendingVars.size()];          // the reference's own draw (:153)
    }
    Variable *gpuQuickSelect(bool useWeighted)              // quickSelect, gnePivotRules.cpp:710-764
    {
        isOptimal = true;
        Variable *bestVar = NULL;
        double largestViol = 0.0;
        if (nbArcPivotInd >= (int) setNBarc.size()) nbArcPivotInd = 0;
        if (nbEqfPivotInd >= (int) setNBeqf.size()) nbEqfPivotInd = 0;
        const int start = nbArcPivotInd, startE = nbEqfPivotInd;
        int eqfViolations = 0;
        bool exhaustedEqfs = setNBeqf.empty();
        if (!exhaustedEqfs) priceSets();
        while ((eqfViolations < 1) && !exhaustedEqfs) {     // the set loop (:723-738) over the device's values
            EqFlowVar *eqfv = setNBeqf[nbEqfPivotInd];
            const double rc = setRc[nbEqfPivotInd];
            if (violates(eqfv, rc)) {
                isOptimal = false;
                ++eqfViolations;
                const double frc = useWeighted ? fabs(rc) / eqfv->numOrigArcs : fabs(rc);
                if (frc > largestViol) { largestViol = frc; bestVar = eqfv; }
            }
            nbEqfPivotInd = (nbEqfPivotInd + 1) % setNBeqf.size();
            exhaustedEqfs = (nbEqfPivotInd == startE);
        }
        int64_t best = -1, cur = start, scanned, viol = 0;
        double bv = 0.0;
        if (!setNBarc.empty()) {
            GCK(gne_price_qs(dev, start, numNodes, &best, &bv, &cur, &scanned, &viol));
            if (viol > 0) isOptimal = false;
            if (best >= 0 && bv > largestViol) bestVar = setNBarc[best];   // an arc replaces the set only when strictly larger
            nbArcPivotInd = (int) cur;
        }
        if ((loopIter % majorIterationLength) > 0) { nbArcPivotInd = start; nbEqfPivotInd = startE; }
        return bestVar;
    }
    Variable *gpuSelect()
    {
        if (cyclingDetected && useRandomJumpWhenCycling) return getRandomVarWithLargeViolation();   // host (:43-46), rare
        const int rule = presolve ? phaseIpivot : phaseIIpivot;
        if (rule == QS || rule == QSW) return gpuQuickSelect(rule == QSW);
        return gpuLargestViolation(rule == LVW);
    }

    static double nowS() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

    // the reference's own rule and the GPU path from the same state
    Variable *selectBothWays()
    {
        unsigned short zero[3] = { 0, 0, 0 }, saved[3];
        memcpy(saved, seed48(zero), sizeof(saved));         // read the generator state ...
        seed48(saved);                                      // ... and put it back
        const int cursor = nbArcPivotInd, cursorE = nbEqfPivotInd, sinceUpdate = itersSinceUpdate;
        const double viol = lastViolation;
        const bool opt = isOptimal;
        const double t0 = nowS();
        Variable *hostVar = selectEnteringVariable();       // reference, host (also writes av->redCost: harmless)
        const double t1 = nowS();
        const bool hostOpt = isOptimal;
        seed48(saved); nbArcPivotInd = cursor; nbEqfPivotInd = cursorE; itersSinceUpdate = sinceUpdate; lastViolation = viol; isOptimal = opt;
        offendingVars.clear();
        const double t2 = nowS();
        Variable *gpuVar = gpuSelect();
        const double t3 = nowS();
        ++abSamples; abHostS += t1 - t0; abGpuS += t3 - t2;
        if (hostVar != gpuVar || hostOpt != isOptimal) {
            ++abMismatch;
            fprintf(stderr, "A/B mismatch at %s iteration %d: host %d gpu %d\n", presolve ? "Phase I" : "Phase II", loopIter,
                    hostVar ? hostVar->varIndex : -1, gpuVar ? gpuVar->varIndex : -1);
        }
        return gpuVar;
    }

  public:
    int abEvery, abSamples, abMismatch;
    double abHostS, abGpuS;
    GpuGenNetEq() : dev(NULL), abEvery(0), abSamples(0), abMismatch(0), abHostS(0.0), abGpuS(0.0) {}
    ~GpuGenNetEq() { if (dev) gne_dev_destroy(dev); }

    // body of GenNetEq::solve (genneteq.cpp:60-117); the three marked lines are the only changes
    int solveGpu(bool usePresolve, FILE *trace)
    {
        bool pIwasInfeasible = usePresolve ? false : isInfeasible;
        initializeStats(usePresolve);
        setVariableCosts();
        upload();                                                           // <-- device mirror of setNBarc
        if (!presolve && pIwasInfeasible) { isInfeasible = true; return loopIter; }
        computeFlows();
        while ((!isOptimal) && (!isUnbounded)) {
            if ((loopIter % 1000) == 0) computeFlows(2);
            computePotentials();
            pushPotentials();                                               // <-- potentials to the device
            eVar = (abEvery > 0 && (loopIter % abEvery) == 0) ? selectBothWays() : gpuSelect();   // <-- identifyEnteringVariable() on the GPU
            if (eVar != NULL) {
                // (positions are those of setNBarc: a set's setInd points into setNBeqf and is none of the mirror's business)
                const int64_t pe = (eVar->varType == ARC_VAR) ? eVar->setInd : -1, last = (int64_t) setNBarc.size() - 1;
                pivot();
                if (lVar != NULL) {
                    if (trace) fprintf(trace, "P %d %d %d\n", loopIter, eVar->varIndex, lVar->varIndex);
                    // the positions updateBasis touched: eVar's old slot (the last element moved in),
                    // the last slot(s) (pushes), and wherever the leaving / flipped variable landed
                    mirrorWrite(pe); mirrorWrite(last); mirrorWrite(last - 1); mirrorWrite(last + 1);
                    if (lVar->varType == ARC_VAR && lVar->setLoc != B_var) mirrorWrite(lVar->setInd);
                    if (eVar->varType == ARC_VAR && eVar->setLoc != B_var) mirrorWrite(eVar->setInd);
                    GCK(gne_nb_set_count(dev, (int64_t) setNBarc.size()));
                }
                ++loopIter;
                presolve ? ++phaseIiters : ++phaseIIiters;
            }
        }
        computeFlows();
        if (presolve) checkFeasibility();
        verifyOptimality();
        printTerminationStatus();
        return loopIter;
    }
    double objective() const { return flowCost; }
};

int main(int argc, char **argv)
{
    const char *tracePath = NULL, *inPath = NULL;
    int abEvery = 0;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-1") && i + 1 < argc) phaseIpivot = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-2") && i + 1 < argc) phaseIIpivot = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-T") && i + 1 < argc) tracePath = argv[++i];
        else if (!strcmp(argv[i], "-A") && i + 1 < argc) abEvery = atoi(argv[++i]);
        else inPath = argv[i];
    }
    if (!inPath) { fprintf(stderr, "usage: %s [-1 r] [-2 r] [-T trace] file.col\n", argv[0]); return 2; }
    Config conf;
    conf.initType = INFEASIBLE;
    conf.debug = 0;
    conf.inFile = inPath;
    Graph *g = new Graph();
    g->initialize(&conf);                                   // the reference's own dense loader
    GpuGenNetEq *solver = new GpuGenNetEq();
    solver->initialize(g, conf.initType);
    solver->abEvery = abEvery;
    delete g;
    FILE *trace = tracePath ? fopen(tracePath, "w") : NULL;
    const int p1 = solver->solveGpu(USE_PRESOLVE, trace);
    const int p2 = solver->solveGpu(USE_SOLVE, trace);
    if (trace) fclose(trace);
    if (abEvery > 0)
        printf("A/B samples %d mismatches %d host_ms %.4f gpu_ms %.4f\n", solver->abSamples, solver->abMismatch,
               1e3 * solver->abHostS / (solver->abSamples ? solver->abSamples : 1), 1e3 * solver->abGpuS / (solver->abSamples ? solver->abSamples : 1));
    printf("GPU-REF pivots %d %d objective %.17g\n", p1, p2, solver->objective());
    delete solver;
    return 0;
}
