// genneteq_b200 — host mirror of the reference's solver shell (genneteq.h / genneteq.cpp) that
// drives the device operators of include/genneteq_b200.h.  Same public surface as the
// reference's GenNetEq (genneteq.h:41-42: initialize, solve) and Graph (graph.h:35-50), same
// member names where a member survives, so the parity tests read like the reference.
//
// What runs where:
//   device  pricing of every nonbasic arc + entering-arc selection (gnePivotRules.cpp), the
//           potential finalize pi = a0 + a1*theta (gnePotential.cpp:127-135), the setNBarc mirror
//   host    the basis forest (trees.cpp), the root-relative potential sweep of the trees a pivot
//           touched (gnePotential.cpp:64-110), flows and the ratio test (gneFlow.cpp,
//           genneteq.h:334-403), drand48 tie picks, cycling detection — all inherently sequential
// Only dirty (node, a0, a1, slot) records, the touched trees' thetas, at most three 25-byte
// nonbasic records and the pricing result cross PCIe per pivot.
#pragma once
#include "gne_forest.h"
#include "gne_eqf.h"
#include "../../include/genneteq_b200.h"

#include <list>
#include <queue>
#include <string>
#include <set>
#include <vector>

namespace gneb200 {

// phaseIpivot / phaseIIpivot are file-scope globals in the reference (main.cpp:21-22); here they
// are per-solver so that several solvers can live in one process (Tree's statics, trees.h:83-86,
// forbid that in the reference).

struct Config {                     // main.h:50-55
    int initType = 2;
    int debug = 0;
    const char *inFile = nullptr;
};

// Sparse replacement of the reference's dense Graph (graph.h:30-94; 4 n x n matrices).
class Graph {
  public:
    int initialize(Config *conf);                       // graph.cpp:31-53 (returns gne_status)
    int initializeFromArrays(int n, int64_t m, const int32_t *tail, const int32_t *head, const double *ub,
                             const double *cost, const double *mult, const double *supply, double bigMIn);
    // the same with equal-flow sets: eqf[a] = 1-based set of arc a (0: none), as the last field of an `a` line
    int initializeFromArraysEqf(int n, int64_t m, const int32_t *tail, const int32_t *head, const double *ub,
                                const double *cost, const double *mult, const int32_t *eqf, int numSets,
                                const double *supply, double bigMIn);
    // what GenNetEq::initializeVars makes of the sets (gneInit.cpp:64-103)
    void toEqfInstance(EqfInstance &out) const;
    int getNumNodes() const { return numNodes; }
    int64_t getNumArcs() const { return (int64_t) tail.size(); }
    int getNumEqualFlowSets() const { return numEqualFlowSets; }
    double getBigM() const { return bigM; }
    double supply(int i) const { return supplies[i]; }

    // arcs that exist (capacity > 0, gneInit.cpp:84), sorted by (tail, head)
    std::vector<int32_t> tail, head;
    std::vector<double> cap, cost, mult, supplies;
    // with equal-flow sets: per kept arc its set (-1: none), and d_r(i) (graph.cpp:403-410), row r = set r, n values each
    std::vector<int32_t> eqfOf;
    std::vector<double> eqfNodeVals;

  private:
    int numNodes = 0;
    int numEqualFlowSets = 0;
    double bigM = 0.0;
};

struct VarWithViol { int32_t var; double violation; };              // variables.h:210-214
struct violcomp {                                                   // variables.h:216-222
    bool operator()(const VarWithViol &a, const VarWithViol &b) const { return a.violation < b.violation; }
};

class GenNetEq {
  public:
    GenNetEq();
    ~GenNetEq();

    int initialize(Graph *g, int initType, int cudaDevice);          // gneInit.cpp:25-44
    int solve(bool usePresolve, int64_t maxPivots, int64_t *iters, int *finished);   // genneteq.cpp:60-117
    // the same loop in pieces (gne_solver_solve_batch: one host thread interleaves several solvers)
    int solveBegin(bool usePresolve, int64_t maxPivots);
    int solveStep(bool block);                          // 1: more to do (2: still waiting for the device), 0: loop ended, < 0: -(gne_status)
    int solveEnd(int64_t *iters, int *finished);
    gne_lv_group *lvGroup = nullptr;                    // set by the batch driver: LV passes go out through this launch group
    double secsComplete = 0.0;                          // step-wise loop: time inside completeLargestViolation (result already there or awaited)
    // step-wise loop: at the loop head (nothing in flight) with this many pivots of the current call behind it
    bool atLoopHead() const { return stStage == 0; }
    int64_t pivotsThisCall() const { return callPivots; }
    double pivotSeconds() const { return secsPivot; }
    double potentialSeconds() const { return secsPotentials; }

    int phaseIpivot = GNE_LVW, phaseIIpivot = GNE_LVW;               // main.cpp:72-73 defaults

    int setComm(const uint8_t id[128], int rank, int nranks, gne_allgather_fn hostGather = nullptr, void *hostCtx = nullptr);
    // global position range [lo, hi) of rank's shard of a list of m arcs (lo == hi: empty shard)
    static void shardRange(int64_t m, int rank, int nranks, int64_t *lo, int64_t *hi);
    void setTrace(int32_t *e, int32_t *l, int64_t cap) { traceE = e; traceL = l; traceCap = cap; traceLen = 0; }
    int64_t getTraceLen() const { return traceLen; }

    int getNumNodes() const { return numNodes; }
    int64_t getNumVars() const { return M; }
    double getFlowCost() const { return flowCost; }
    bool getIsInfeasible() const { return isInfeasible; }
    void getFlows(double *out) const;
    int getPotentials(double *out);
    int refreshPotentials() { return computePotentials(); }     // computePotentials() for the CURRENT basis (idempotent)
    void getStates(int32_t *out) const;
    const Forest &getForest() const { return forest; }
    Forest &getForestMutable() { return forest; }
    void getCounters(double *out12) const;
    int replayTrace(int draws, int64_t p1, int64_t total, const int32_t *e, const int32_t *l, int64_t *mismatch);
    void setSparseFlows(bool on) { sparseFlows = on; }
    void getHostPotentials(double *out) const { for (int i = 0; i < numNodes; ++i) out[i] = hostPotential(i); }
    // wall-clock window over pivots [start, start + len) of the next solve() call, taken inside the loop
    // with the device synchronised at both edges; out = {seconds, 12 counters at start, 12 at end, t0, t1}
    void setWindow(int64_t start, int64_t len) { winStart = start; winLen = len; winT0 = winT1 = 0.0; }
    void getWindow(double *out27) const;
    // every `every` pivots append {wall seconds, 12 counters, largest tree size} to buf (13+1 doubles per row)
    void setProgress(int64_t every, double *buf, int64_t capRows) { progEvery = every; progBuf = buf; progCap = capRows; progRows = 0; }
    int64_t getProgressRows() const { return progRows; }
    gne_dev *device() { return dev; }

  protected:
    // ---- sizes / parameters (genneteq.h:44-75)
    int numNodes = 0;
    int64_t m = 0, M = 0;
    double bigM = 0.0;
    bool presolve = true;
    double flowCost = 0.0;
    bool isOptimal = false, isInfeasible = false, isUnbounded = false;
    int phaseIiters = 0, phaseIIiters = 0, loopIter = 0;
    int consecutiveCycles = 0, consecutiveDegenPivots = 0, totalCyclesDetected = 0, totalDegenPivots = 0;
    int numLargestViolVarsToTrack = 5, majorIterationLength = 3;
    int clMaxSize = 0, numArcsToAlwaysCheck = 5;

    // ---- arc variables, structure-of-arrays by varIndex (variables.h:89-163)
    std::vector<int32_t> tail, head, setInd;
    std::vector<double> mult, UB, cost, actualCost, flow, deltaFlow, flowA0, flowA1;
    std::vector<int8_t> setLoc, isArtificial;
    // ---- nodes (variables.h:37-87)
    std::vector<double> supply, supA0, supA1;
    std::vector<double> potA0, potA1;                   // persisted potentialWithEq (the reference zeroes it)
    std::vector<int32_t> potSlot;                       // slot the device holds for the node
    std::vector<double> thetaOfSlot;
    // ---- basis sets (genneteq.h:86-90)
    std::vector<int32_t> setBarc, setNBarc;
    Forest forest;

    // ---- pivoting state (genneteq.h:102-128)
    double delta = 0.0;
    int32_t eVar = -1, lVar = -1;
    std::vector<int32_t> blockingVars, offendingVars, candidateList2;
    std::vector<double> offendingViol, candidateViol2;
    bool cyclingDetected = false;
    std::list<std::pair<int32_t, int32_t> > cycleList;
    int lastEVarInd = -1;
    double lastViolation = 100000, baselineViolation = 100000;      // genneteq.cpp:22-23
    std::priority_queue<VarWithViol, std::vector<VarWithViol>, violcomp> candidateList;
    int itersSinceUpdate = -1;
    int64_t nbArcPivotInd = 0;
    uint64_t rand48State = 0;                           // glibc's never-seeded drand48 state
    double drand48();

    // ---- device
    // computeFlows (every 1000 iterations) without the O(|setNBarc|) walk: the positions of the nonbasic arcs at
    // their upper bound, in position order (the order of the supply sums, gneFlow.cpp:343-353), and the nonbasic
    // arcs at the lower bound whose stored flow is not exactly +0.0 (the only ones the reset changes)
    std::set<int32_t> uPos;
    std::vector<int32_t> lDirty;
    // ... and without re-sweeping trees whose flows cannot have changed: a tree is stale when a pivot ran over it,
    // its structure changed, or a nonbasic U arc at one of its nodes changed state or position since the last computeFlows
    std::vector<uint8_t> flowStale;
    bool inFlowCheck = false;                           // GNE_FLOW_CHECK=1: verify the skipping computeFlows against the full walk
    void markFlowStale(int32_t slot) { if (slot < 0) return; if (slot >= (int32_t) flowStale.size()) flowStale.resize((size_t) slot + 1, 1); flowStale[slot] = 1; }
    gne_dev *dev = nullptr;
    int cudaDevice = 0;
    int commRank = 0, commRanks = 1;
    bool costsOnDevice = false;
    std::vector<int64_t> listPos;
    std::vector<double> listViol;
    std::vector<int32_t> upNode, upSlot, upThetaSlot, upFinal, upFinalSlots;
    std::vector<double> upA0, upA1, upTheta, upPi;
    int status = GNE_OK;
    bool allPotDirty = true, hostOnly = false, sparseFlows = true;
    std::vector<int32_t> nodeMark;
    std::vector<int32_t> actHead, actNext, actOrder, actStack, actKids;
    std::vector<int32_t> activeNodes, activeArcsU, activeArcsV;
    std::vector<std::pair<int32_t, int32_t> > activeKeys, ordKids;
    std::vector<int32_t> orderedArcs;

    // ---- state of the step-wise solve loop
    int stStage = 0, stFinished = 1;
    bool stSkip = false;
    int64_t stMaxPivots = -1, stBudget = -1;
    double stT0 = 0.0, stB = 0.0;

    // ---- trace / counters
    int32_t *traceE = nullptr, *traceL = nullptr;
    int64_t traceCap = 0, traceLen = 0;
    double secsPricing = 0, secsPotentials = 0, secsPivot = 0, secsTotal = 0, secsTrees = 0, secsFlows = 0;
    double arcsPriced = 0, kernelMs = 0;
    int64_t pivots = 0;
    int64_t winStart = -1, winLen = 0, callPivots = 0;
    int64_t progEvery = 0, progCap = 0, progRows = 0;
    double *progBuf = nullptr;
    double winT0 = 0.0, winT1 = 0.0, winC0[12] = { 0 }, winC1[12] = { 0 };

    // ---- functions, named as in genneteq.h:153-271
    void initializeStats(bool usePresolve);
    int setVariableCosts();
    void checkFeasibility();
    double computeArcRedCost(int32_t av);               // host, single arcs (candidate re-pricing)
    double hostPotential(int32_t node) const;
    void checkVarForBlocking(int32_t v);
    double computeDeltaVar(int32_t v) const;

    void computeFlows(int updateOption = 0);
    void computeFlowsPivotingTI(int32_t eav, int32_t tU, int32_t tV);
    void computeFlowsTypeI(int32_t slot);

    int computePotentials();
    bool relabelNode(int32_t j, int32_t slot);
    void computeFlowsSparse(int32_t slot, const int32_t *starts, int nStarts, std::vector<int32_t> &activeArcs);
    void applyFlowSparse(int32_t slot, const std::vector<int32_t> &activeArcs);
    int computeReducedCosts(bool includeBasicVars);
    int verifyOptimality();

    int pivot();
    int pivotTI(int32_t eav);
    void applyTreeFlow(int32_t slot);
    void identifyLeavingVar();
    int updateBasis();
    void checkForCycling();

    int insertIntoB(int32_t v);
    int removeFromB(int32_t v);
    int insertIntoNB(int32_t v, int8_t loc);            // insertIntoL / insertIntoU
    int removeFromNB(int32_t v);                        // removeFromL / removeFromU
    int writeNbRecord(int64_t pos, int32_t v);

    int identifyEnteringVariable();
    int selectEnteringVariable(int32_t *out);
    int getVarWithLargestViolation(int32_t *out);
    int completeLargestViolation(int32_t *out);
    int getFirstVarWithViolation(int32_t *out);
    int getRandomVarWithViolation(int32_t *out);
    int getRandomVarWeightedByViolation(int32_t *out);
    int getRandomVarWithLargeViolation(int32_t *out);
    int getVarBySimAnnealing(bool arcsFirst, int32_t *out);
    int getNextCandidate(int32_t *out);
    int updateCandidateList();
    int quickSelect(int32_t *out);
    int getNextCandidate2(int32_t *out);
    int updateCandidateList2();
    int fetchViolators(double threshold);
};

} // namespace gneb200
