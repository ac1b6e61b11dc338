// genneteq_b200 — the solver for instances WITH equal-flow sets (p > 0): SURVEY 8(f3) / a12.
//
// With basic equal-flow variables the basis holds Type II trees (no extra arc, root potential free), the potentials of
// their nodes come out of an s x s system (s = |setBeqf|, gnePotential.cpp:142-214), the flow change of a pivot that
// touches them is an (s+1)-vector per node and arc plus a second s x s system (gneFlow.cpp:188-306, 563-592), and both
// the row/column order of those systems and the tree ids the reference hands out (trees.cpp:183-284) decide the last
// bit of every value.  None of the sparsity the p = 0 engine (gne_solver.cpp) lives on survives that - every pivot is
// O(n s) on the host in the reference too - so this engine restates the reference's data flow directly on flat arrays
// (same member names, same operation order), and keeps on the device what stays data-parallel:
//   * pricing of the nonbasic arcs   - the same device-resident structure-of-arrays mirror and kernels as at p = 0
//                                      (gne_price_lv_* / gne_price_qs / gne_price_first / gne_price_violators);
//   * pricing of the equal-flow sets - gne_price_eqf (computeEqfRedCost, genneteq.h:321-330: one row per set);
//   * potentials                     - uploaded whole after each computePotentials (all of them can change through the
//                                      s x s solve).
// The s x s systems use an LU with partial pivoting in GSL's operation order (gsl_linalg_LU_decomp / _solve of GSL 1.x-2.5
// with gslcblas' reference dtrsv loops; the reference links -lgsl -lgslcblas, version not pinned): see lu_decomp / lu_solve.
// Parity is pinned against the unmodified reference compiled with the same restatement of those two GSL calls
// (the gsl_mini header tree of the test infrastructure), fixtures tests/golden/q_*.npz - "reference + restated GSL", not a
// GSL build.
//
// Pricing rules with sets: every rule id of main.h:17-34 except SD (stale in the reference too): GNE_ERR_UNSUPPORTED.
#pragma once
#include "../../include/genneteq_b200.h"

#include <list>
#include <queue>
#include <stdint.h>
#include <utility>
#include <vector>

namespace gneb200 {

// What the loader hands over (gneInit.cpp:60-105 and graph.cpp:382-419 already applied).
struct EqfInstance {
    int n = 0, p = 0;
    double bigM = 0.0;
    // arc variables: arcs outside every set with capacity > 0, sorted by (tail, head)      [gneInit.cpp:82-95]
    std::vector<int32_t> tail, head;
    std::vector<double> cap, cost, mult;
    std::vector<double> supply;
    // per set r: member arcs with capacity > 0, UB = their smallest capacity, actualCost = sum of their costs in
    // (tail, head) order [gneInit.cpp:96-103]; d_r(i) = +1 per member arc leaving i, -mult per member arc entering i,
    // added up in FILE order [graph.cpp:403-410] - dense rows, n values each
    std::vector<int32_t> numOrigArcs;
    std::vector<double> setUB, setCost;
    std::vector<double> nodeVals;                       // [p * n]
};

class EqfSolver {
  public:
    ~EqfSolver();
    int initialize(const EqfInstance &g, int cudaDevice, bool hostOnlyReplay = false);       // gneInit.cpp:25-44 (SELFLOOPS)
    // host-only replay of a recorded solve (no device, no pricing): the entering variable of every pivot comes from the
    // trace; flows, the s x s systems, ratio test, leaving variable, tree surgery and potentials run as in solve() and
    // the leaving variable is compared (*mismatch = first differing pivot or -1).  draws: drand48 calls per pricing, per phase.
    int replayTrace(int draws1, int draws2, int64_t p1, int64_t total, const int32_t *e, const int32_t *l, int64_t *mismatch);
    int solve(bool usePresolve, int64_t maxPivots, int64_t *iters, int *finished);           // genneteq.cpp:60-117
    // shard the pricing of the nonbasic arcs over ranks (before the first solve): every rank runs this same host engine and
    // prices its position range; the device layer merges (gne_price_lv_* / _qs / _first / _violators are shard-aware), the
    // sets are priced on every rank.  hostGather != NULL: the caller's all-gather instead of NCCL (ranks may share a device)
    int setComm(const uint8_t id[128], int rank, int nranks, gne_allgather_fn hostGather, void *hostCtx);
    int phaseIpivot = GNE_LVW, phaseIIpivot = GNE_LVW;

    void setTrace(int32_t *e, int32_t *l, int64_t cap) { traceE = e; traceL = l; traceCap = cap; traceLen = 0; }
    int64_t getTraceLen() const { return traceLen; }
    int getNumNodes() const { return numNodes; }
    int64_t getNumVars() const { return V; }
    int getNumSets() const { return numSets; }
    double getFlowCost() const { return flowCost; }
    bool getIsInfeasible() const { return isInfeasible; }
    void getFlows(double *out) const { for (int64_t v = 0; v < V; ++v) out[v] = flow[(size_t) v]; }
    void getPotentials(double *out) const { for (int i = 0; i < numNodes; ++i) out[i] = potential[(size_t) i]; }
    void getStates(int32_t *out) const { for (int64_t v = 0; v < V; ++v) out[v] = setLoc[(size_t) v]; }
    void getCounters(double *out12) const;
    int64_t typeIIPivots() const { return pivotsTII; }
    // the s x s solve on its own (tests): a = row-major s x s (overwritten by its LU), perm = row permutation, x = solution
    static void luSolve(std::vector<double> &a, int s, const std::vector<double> &b, std::vector<int> &perm, std::vector<double> &x)
    { lu_decomp(a, s, perm); lu_solve(a, s, perm, b, x); }
    gne_dev *device() { return dev; }

  private:
    enum { B_var = 0, L_var = 1, U_var = 2, NIL = 3 };                 // variables.h:23
    enum { T_NONE = 0, T_I = 1, T_II = 2 };                            // variables.h:25

    struct ETree {                                                     // trees.h:24-109
        int id = -1;
        int8_t type = T_NONE;
        int root = -1;
        bool upToDate = false;
        int32_t extraArc = -1;
        std::vector<int32_t> nodesInTree, arcsInTree, threads;
        // not in the reference: a Type I tree whose arcs, rooting, extra arc and costs did not change since its potentials
        // were last computed would get the same bits again (they depend on nothing else) - it is skipped
        bool potDirty = true;
    };
    struct VarWithViol { int32_t var; double violation; };            // variables.h:210-222
    struct violcomp { bool operator()(const VarWithViol &a, const VarWithViol &b) const { return a.violation < b.violation; } };

    // ---- sizes (variables: arcs 0..A-1 with the artificial self-loops last, then the sets A..A+p-1)
    int numNodes = 0, numSets = 0;
    int64_t A = 0, V = 0;
    double bigM = 0.0;
    bool presolve = true;
    double flowCost = 0.0;
    bool isOptimal = false, isInfeasible = false, isUnbounded = false;
    int phaseIiters = 0, phaseIIiters = 0, loopIter = 0;
    int consecutiveCycles = 0, consecutiveDegenPivots = 0, totalCyclesDetected = 0, totalDegenPivots = 0;
    int numLargestViolVarsToTrack = 5, majorIterationLength = 3, itersSinceUpdate = -1;
    int64_t nbArcPivotInd = 0, nbEqfPivotInd = 0;
    // ---- variables
    std::vector<int32_t> tail, head, setInd, numOrigArcs;
    std::vector<double> mult, UB, cost, actualCost, flow, deltaFlow, flowA0, flowA1;
    std::vector<int8_t> setLoc, isArtificial;
    std::vector<double> nodeVals;                                      // [p * n], row r = set r
    std::vector<int64_t> csrStart;                                     // the same rows sparse (device pricing)
    std::vector<int32_t> csrNode;
    std::vector<double> csrVal;
    // ---- nodes
    std::vector<double> supply, potential, potA0, potA1, supA0, supA1;
    std::vector<int32_t> potTheta;
    int S = 1;                                                         // s + 1 of the flow computation under way
    std::vector<double> swt;                                           // supplyWithThetas: [n * S]
    std::vector<double> fwt;                                           // flowWithThetas of predArc(j): [n * S], row j
    std::vector<int32_t> memberOfTreeID, depth, thread, pred, predArc;
    // not in the reference: a node's index in its tree's thread vector and an arc's index in its tree's arcsInTree - what the
    // path-sparse flow computation of a Type I pivot orders its few active nodes and arcs by
    std::vector<int32_t> threadPos, posInTree;
    std::vector<int32_t> nodeStamp;                                    // active-node marks (value = stampNow)
    int32_t stampNow = 0;
    std::vector<int32_t> activeNodes, activeArcs, activeArcsU, activeArcsV;
    bool sparseFlows = true;                                           // GNE_EQF_DENSE=1 turns the path-sparse computation off
    std::vector<int8_t> memberOfTreeType;
    std::vector<std::vector<int32_t> > incidentArcs;
    // ---- basis
    std::vector<int32_t> setBarc, setBeqf, setNBarc, setNBeqf;
    std::vector<ETree *> setBtI, setBtII;
    // ---- pivoting state
    double delta = 0.0;
    int32_t eVar = -1, lVar = -1;
    std::vector<int32_t> blockingVars;
    bool cyclingDetected = false;
    std::list<std::pair<int32_t, int32_t> > cycleList;
    int lastEVarInd = -1;
    double lastViolation = 100000, baselineViolation = 100000;
    std::priority_queue<VarWithViol, std::vector<VarWithViol>, violcomp> candidateList;
    std::vector<int32_t> candidateList2;                               // :550-705
    std::vector<double> candidateViol2;                                // |redCost| a member had when the list was built
    uint64_t rand48State = 0;
    double drand48();
    // ---- device
    gne_dev *dev = nullptr;
    int cudaDevice = 0;
    bool costsOnDevice = false, hostOnly = false, potsOnDevice = false;
    std::vector<int64_t> listPos;
    std::vector<double> listViol, eqfRc;
    std::vector<int32_t> upNodes;                                      // nodes whose potential the last computePotentials rewrote
    std::vector<double> upPi;
    std::vector<int32_t> offendingVars;
    std::vector<double> offendingViol;
    // ---- trace / counters
    int32_t *traceE = nullptr, *traceL = nullptr;
    int64_t traceCap = 0, traceLen = 0, pivots = 0, pivotsTII = 0;
    double secsPricing = 0, secsPotentials = 0, secsPivot = 0, secsTotal = 0, arcsPriced = 0;
    double secsFlowsTI = 0, secsFlowsTII = 0, secsBasis = 0;           // inside secsPivot
    double secsTIIsets = 0, secsTIItrees = 0;                          // inside secsFlowsTII
    int status = GNE_OK;

    // ---- trees.cpp
    ETree *createNewTree(int8_t type, int rNode);
    void setTreeIDAndType(ETree *t, int newID, int8_t newType);
    void removeTree(ETree *t);
    void convertTItoTII(ETree *t);
    void convertTIItoTI(ETree *t, int32_t av);
    int insertIntoTrees(int32_t av);
    void mergeTxAndTII(ETree *tX, ETree *tII, int32_t av);
    void expandTreeUsingArc(ETree *t, int32_t av);
    int removeFromTrees(int32_t av);
    void splitTree(ETree *t, int32_t removedArc);
    void updateTrees();
    void initializeTreeIndices(ETree *t);
    // ---- genneteq.cpp / gneFlow.cpp / gnePotential.cpp / gnePivot.cpp / gnePivotRules.cpp
    void initializeStats(bool usePresolve);
    int setVariableCosts();
    void checkFeasibility();
    void checkVarForBlocking(int32_t v);
    double computeDeltaVar(int32_t v) const;
    void computeFlows(int updateOption);
    void setFlowValues(int updateOption);
    void computeFlowsPivotingTII(int32_t var);
    void computeFlowsWithEqFlowSets();
    void computeNumericFlowsAndSupply(const std::vector<double> &x);
    void computeBoundedFlows(bool withEqFlowSets);
    void setSupplyWithThetas();
    void setSupplyWithThetas(int32_t var);
    void setSupplyWithEqs();
    void setSupplyWithEqs(int32_t var);
    void computeFlowsTypeI(ETree *t);
    bool computeFlowsTypeISparse(ETree *t, const int32_t *sources, int numSources, std::vector<int32_t> &arcsOut);
    int computeFlowsTypeII(ETree *t);
    int computePotentials();
    void computePotentialsInTermsOfRoot(ETree *t);
    void finalizePotentialsTypeI(ETree *t);
    void computePotentialsWithEqFlowSets();
    int verifyOptimality();
    int pivot();
    void pivotTI(int32_t eav);
    void pivotTII();
    void identifyLeavingVar();
    int updateBasis();
    void checkForCycling();
    int insertIntoB(int32_t v);
    int removeFromB(int32_t v);
    int insertIntoNB(int32_t v, int8_t loc);
    int removeFromNB(int32_t v);
    int selectEnteringVariable(int32_t *out);
    int priceSets();                                                   // reduced cost of every nonbasic set -> eqfRc (setNBeqf order)
    int priceSetsLaunch();                                             // the same in two halves: the arc pass can be queued in between
    int priceSetsFetch();
    bool setsResident = false, setsLaunched = false;                   // rows on the device (gne_eqf_set_rows); a launch awaits its fetch
    int64_t nonFinitePots = 0;                                         // potentials that are NaN or infinite right now
    bool violates(int32_t v, double rc) const;
    double hostRedCost(int32_t v) const;
    int fetchViolators();                                              // computeReducedCosts(false): offendingVars / offendingViol
    int getVarWithLargestViolation(int order, bool useWeighted, int32_t *out);
    int getFirstVarWithViolation(int32_t *out);
    int getRandomVarWithViolation(int32_t *out);
    int getRandomVarWeightedByViolation(int32_t *out);
    int getRandomVarWithLargeViolation(bool useWeighted, int32_t *out);
    int getVarBySimAnnealing(bool arcsFirst, int32_t *out);
    int getNextCandidate(bool useWeighted, int32_t *out);
    int updateCandidateList2();
    int getNextCandidate2(bool useWeighted, int32_t *out);
    int quickSelect(bool useWeighted, int32_t *out);
    // ---- the s x s systems
    static void lu_decomp(std::vector<double> &a, int s, std::vector<int> &perm);
    static void lu_solve(const std::vector<double> &lu, int s, const std::vector<int> &perm, const std::vector<double> &b, std::vector<double> &x);
    bool isArc(int32_t v) const { return v < A; }
};

} // namespace gneb200
