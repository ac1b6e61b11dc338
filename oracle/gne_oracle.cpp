// TEST INFRASTRUCTURE — CPU restatement of GenNetEq's generalized network simplex hot path.
//
// This is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it; nothing under genneteq_b200/ does.
//
// PINNED: the pivot traces of this restatement are compared with those of the unmodified
// reference compiled from /root/reference (oracle/Makefile `ref`, oracle/ref_driver.cpp) in
// tests/test_oracle_*.py and against the md5-pinned golden traces in tests/golden/.
// The reference ships no tests or golden vectors of its own.
//
// It restates, on flat index arrays instead of heap objects, with zero equal-flow sets:
//   gneInit.cpp:25-186 (SELFLOOPS path), genneteq.cpp:60-167, 227-323, genneteq.h:312-403,
//   gnePivotRules.cpp (all rules except SD), gnePotential.cpp:28-138, 217-287,
//   gneFlow.cpp:30-185, 312-369, 387-395, 441-560, gnePivot.cpp:28-209, 254-356, 487-600,
//   trees.cpp:183-644.
// The arithmetic keeps the reference's operation order; build with -ffp-contract=off so that
// no multiply-add is fused (the reference's -O3 x86-64 build contains no FMA).
// Equal-flow sets (Type II pivots, GSL LU) are out of scope: parity unpinned there.
#include "gne_oracle.h"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <queue>
#include <vector>
#include <time.h>

namespace {

const double TOL = 1.0e-13;                    // main.h:36-43 ROUNDOFF_TOLERANCE
enum { LOC_B = 0, LOC_L = 1, LOC_U = 2, LOC_NIL = 3 };      // variables.h:23 var_loc
enum { T_NULL = 0, T_I = 1, T_II = 2 };                     // variables.h:25 tree_type
enum { R_LV = 0, R_LVW = 1, R_LVA = 2, R_LVE = 3, R_FV = 4, R_SD = 5, R_RV = 6, R_RWV = 7,
       R_RLV = 8, R_RLVW = 9, R_CL = 10, R_CLW = 11, R_SAA = 12, R_SAE = 13, R_CL2 = 14,
       R_CLW2 = 15, R_QS = 16, R_QSW = 17 };                // main.h:17-34

double wallNow()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1.0e-9 * (double) ts.tv_nsec;
}

struct VarViol { int var; double violation; };
struct ViolLess {                                           // variables.h:216-222 violcomp
    bool operator()(const VarViol &a, const VarViol &b) const { return a.violation < b.violation; }
};
typedef std::priority_queue<VarViol, std::vector<VarViol>, ViolLess> ViolHeap;

struct OTree {                                              // trees.h:24-109
    int id, type, root, extraArc;
    bool upToDate;
    std::vector<int> nodesIn, arcsIn, threads;
};

} // namespace

struct OrcSolver {
    // ---- sizes
    int n; long m; long M;
    double bigM;
    // ---- per arc variable (variables.h:89-163), index = varIndex
    std::vector<int> tail, head, setInd;
    std::vector<double> mult, ub, cost, actualCost, flow, deltaFlow, redCost, fa0, fa1;
    std::vector<signed char> setLoc, isArt;
    // ---- per node (variables.h:37-87)
    std::vector<double> potential, supply, pa0, pa1, sa0, sa1;
    std::vector<int> memberTree, memberType, depth, thread, pred, predArc;
    std::vector<std::vector<int> > incident;
    // ---- basis sets (genneteq.h:86-93)
    std::vector<int> setB, setNB;
    std::vector<OTree *> setBtI, setBtII;
    // ---- solver state (genneteq.h:44-137)
    bool presolve, isOptimal, isInfeasible, isUnbounded, cyclingDetected;
    int phaseIiters, phaseIIiters, loopIter;
    int consecutiveCycles, consecutiveDegenPivots, totalCyclesDetected, totalDegenPivots;
    int numLargestViolVarsToTrack, majorIterationLength, clMaxSize, numArcsToAlwaysCheck;
    int itersSinceUpdate, nbArcPivotInd, lastEVarInd;
    double flowCost, delta, lastViolation, baselineViolation;
    int eVar, lVar;
    std::vector<int> blockingVars, offendingVars, candidateList2;
    std::list<std::pair<int, int> > cycleList;
    ViolHeap candidateList;
    int ruleI, ruleII;
    unsigned long long rng;
    int status;
    // ---- trace / timers
    int *traceE, *traceL; long traceCap, traceLen;
    double secsPricing, secsPotentials, secsPivot, secsTotal; long arcsPriced;
    long windowStart; double winTimers[5];

    OrcSolver() : lastViolation(100000), baselineViolation(100000) {}      // genneteq.cpp:20-24

    double rand48() { return orc_drand48(&rng); }
    void fail(int code, const char *msg) { if (status == 0) { status = code; fprintf(stderr, "oracle: %s\n", msg); } }

    // ======================= trees.cpp =======================
    OTree *createNewTree(int type, int rNode = -1)          // trees.cpp:183-205
    {
        OTree *t = new OTree();
        std::vector<OTree *> &set = (type == T_I) ? setBtI : setBtII;
        t->id = (int) set.size(); t->type = type; t->root = -1; t->upToDate = false; t->extraArc = -1;
        set.push_back(t);
        if (rNode >= 0) {
            t->root = rNode;
            t->nodesIn.push_back(rNode);
            memberTree[rNode] = t->id;
            memberType[rNode] = t->type;
        }
        return t;
    }
    void setTreeIDAndType(OTree *t, int newID, int newType) // trees.cpp:220-230
    {
        t->id = newID; t->type = newType;
        for (size_t k = 0; k < t->nodesIn.size(); ++k) {
            memberTree[t->nodesIn[k]] = newID;
            memberType[t->nodesIn[k]] = newType;
        }
    }
    void removeTree(OTree *t)                               // trees.cpp:233-256
    {
        int old = t->id;
        std::vector<OTree *> &set = (t->type == T_I) ? setBtI : setBtII;
        int type = t->type;
        if (old < (int) set.size() - 1) {
            OTree *last = set.back();
            setTreeIDAndType(last, old, type);
            set[old] = last;
        }
        set.pop_back();
    }
    void convertTItoTII(OTree *t)                           // trees.cpp:259-270
    {
        int newID = (int) setBtII.size();
        removeTree(t);
        setTreeIDAndType(t, newID, T_II);
        t->extraArc = -1;
        setBtII.push_back(t);
    }
    void convertTIItoTI(OTree *t, int av)                   // trees.cpp:273-284
    {
        int newID = (int) setBtI.size();
        removeTree(t);
        setTreeIDAndType(t, newID, T_I);
        t->extraArc = av;
        setBtI.push_back(t);
    }
    void mergeTxAndTII(OTree *tX, OTree *tII, int av)       // trees.cpp:365-395
    {
        removeTree(tII);
        setTreeIDAndType(tII, tX->id, tX->type);
        for (size_t k = 0; k < tII->nodesIn.size(); ++k) tX->nodesIn.push_back(tII->nodesIn[k]);
        for (size_t k = 0; k < tII->arcsIn.size(); ++k) tX->arcsIn.push_back(tII->arcsIn[k]);
        tX->arcsIn.push_back(av);
        incident[tail[av]].push_back(av);
        incident[head[av]].push_back(av);
        tX->upToDate = false;
        delete tII;
    }
    void expandTreeUsingArc(OTree *tX, int av)              // trees.cpp:398-437
    {
        int u = tail[av], v = head[av];
        if (memberTree[u] != tX->id) {
            tX->nodesIn.push_back(u); memberTree[u] = tX->id; memberType[u] = tX->type;
        } else if (memberTree[v] != tX->id) {
            tX->nodesIn.push_back(v); memberTree[v] = tX->id; memberType[v] = tX->type;
        }
        if (u != v) {
            tX->arcsIn.push_back(av);
            incident[u].push_back(av);
            incident[v].push_back(av);
        } else {
            tX->extraArc = av;
        }
        tX->upToDate = false;
    }
    void insertIntoTrees(int av)                            // trees.cpp:290-362
    {
        int u = tail[av], v = head[av];
        int tUID = memberTree[u], tUType = memberType[u];
        int tVID = memberTree[v], tVType = memberType[v];
        if (tUID != -1 && tVID != -1) {
            if (tUID == tVID && tUType == tVType) {
                if (tUType == T_I) { fail(-2, "Insertion adds cycle to type I tree"); return; }
                convertTIItoTI(setBtII[tUID], av);
            } else {
                if (tUType == T_II && tVType == T_II) mergeTxAndTII(setBtII[tUID], setBtII[tVID], av);
                else if (tUType == T_I && tVType == T_II) mergeTxAndTII(setBtI[tUID], setBtII[tVID], av);
                else if (tUType == T_II && tVType == T_I) mergeTxAndTII(setBtI[tVID], setBtII[tUID], av);
                else { fail(-3, "Insertion merges two type I trees"); return; }
            }
        } else {
            OTree *t;
            if (tUID == -1 && tVID == -1) t = createNewTree(u == v ? T_I : T_II, u);
            else if (tUID != -1) t = (tUType == T_I) ? setBtI[tUID] : setBtII[tUID];
            else t = (tVType == T_I) ? setBtI[tVID] : setBtII[tVID];
            expandTreeUsingArc(t, av);
        }
    }
    void splitTree(OTree *tX, int removedArc)               // trees.cpp:502-566
    {
        OTree *tTop = createNewTree(T_II);
        OTree *tBottom = createNewTree(T_II);
        OTree *tCur = tTop;
        std::vector<int> curStack, prevStack;
        curStack.push_back(tX->root);
        prevStack.push_back(-1);
        while (!curStack.empty()) {
            int cur = curStack.back(); curStack.pop_back();
            if (cur < 0) {
                tCur = tTop;
                if (curStack.empty()) break;
                cur = curStack.back(); curStack.pop_back();
            }
            int prev = prevStack.back(); prevStack.pop_back();
            memberTree[cur] = tCur->id;
            memberType[cur] = tCur->type;
            if (tCur->nodesIn.empty()) tCur->root = cur;
            tCur->nodesIn.push_back(cur);
            int removedInd = -1;
            std::vector<int> &inc = incident[cur];
            for (int k = 0; k < (int) inc.size(); ++k) {
                int a = inc[k];
                if (a == removedArc) { removedInd = k; continue; }
                int next = (tail[a] == cur) ? head[a] : tail[a];
                if (next != prev) {
                    tCur->arcsIn.push_back(a);
                    curStack.push_back(next);
                    prevStack.push_back(cur);
                }
            }
            if (removedInd >= 0) {
                inc[removedInd] = inc.back();
                inc.pop_back();
                if (prev > -2) {
                    int next = (tail[removedArc] == cur) ? head[removedArc] : tail[removedArc];
                    curStack.push_back(-1);
                    curStack.push_back(next);
                    prevStack.push_back(-2);
                    tCur = tBottom;
                }
            }
        }
    }
    void removeFromTrees(int av)                            // trees.cpp:456-499
    {
        int u = tail[av], v = head[av];
        if (memberTree[u] != memberTree[v] || memberType[u] != memberType[v]) {
            fail(-4, "Attempting to remove arc spanning two trees"); return;
        }
        OTree *t = (memberType[u] == T_I) ? setBtI[memberTree[u]] : setBtII[memberTree[u]];
        if (t->type == T_I && av == t->extraArc) {
            convertTItoTII(t);
        } else {
            removeTree(t);
            splitTree(t, av);
            if (t->type == T_I) insertIntoTrees(t->extraArc);
            delete t;
        }
    }
    void initializeTreeIndices(OTree *t)                    // trees.cpp:597-644
    {
        std::vector<int> curStack, arcStack;
        t->threads.clear();
        curStack.push_back(t->root);
        while (!curStack.empty()) {
            int cur = curStack.back(); curStack.pop_back();
            int predNode = -1;
            if (!t->threads.empty()) {
                int pa = arcStack.back(); arcStack.pop_back();
                predNode = (cur == tail[pa]) ? head[pa] : tail[pa];
                depth[cur] = depth[predNode] + 1;
                predArc[cur] = pa;
                thread[t->threads.back()] = cur;
            } else {
                depth[cur] = 0;
                predArc[cur] = -1;
            }
            pred[cur] = predNode;
            t->threads.push_back(cur);
            std::vector<int> &inc = incident[cur];
            for (int k = (int) inc.size() - 1; k >= 0; --k) {
                int a = inc[k];
                int u = tail[a], v = head[a];
                if (u == predNode || v == predNode) continue;
                curStack.push_back(cur == u ? v : u);
                arcStack.push_back(a);
            }
        }
        thread[t->threads.back()] = t->root;
    }
    void updateTrees()                                      // trees.cpp:572-591
    {
        for (size_t k = 0; k < setBtI.size(); ++k) {
            OTree *t = setBtI[k];
            if (!t->upToDate) {
                t->root = tail[t->extraArc];
                initializeTreeIndices(t);
                t->upToDate = true;
            }
        }
        for (size_t k = 0; k < setBtII.size(); ++k) {
            OTree *t = setBtII[k];
            if (!t->upToDate) { initializeTreeIndices(t); t->upToDate = true; }
        }
    }

    // ======================= basis sets: gnePivot.cpp:487-600 =======================
    void insertIntoB(int v)
    {
        setB.push_back(v); setLoc[v] = LOC_B; setInd[v] = (int) setB.size() - 1;
        insertIntoTrees(v);
    }
    void removeFromB(int v)
    {
        int prevInd = setInd[v];
        int mv = setB.back();
        setB[prevInd] = mv; setInd[mv] = prevInd; setB.pop_back();
        removeFromTrees(v);
        setLoc[v] = LOC_NIL; setInd[v] = -1;
    }
    void insertIntoNB(int v, int loc)                       // insertIntoL / insertIntoU
    {
        setNB.push_back(v); setLoc[v] = (signed char) loc; setInd[v] = (int) setNB.size() - 1;
    }
    void removeFromNB(int v)                                // removeFromL / removeFromU
    {
        int prevInd = setInd[v];
        int mv = setNB.back();
        setNB[prevInd] = mv; setInd[mv] = prevInd; setNB.pop_back();
        setLoc[v] = LOC_NIL; setInd[v] = -1;
    }

    // ======================= genneteq.h inline helpers =======================
    double computeArcRedCost(int a)                         // genneteq.h:312-318
    {
        double t = mult[a] * potential[head[a]];
        redCost[a] = (cost[a] - potential[tail[a]]) + t;
        return redCost[a];
    }
    bool violates(int a, double rc) const
    {
        return (setLoc[a] == LOC_L && rc < (-1.0 * TOL)) || (setLoc[a] == LOC_U && rc > TOL);
    }
    double computeDeltaVar(int v) const                     // genneteq.h:364-403 (initialType == SELFLOOPS)
    {
        double x = flow[v], y = deltaFlow[v];
        if (y > TOL) {
            if (fabs(ub[v] - x) < TOL) return 0.0;
            return (ub[v] - x) / y;
        } else if (y < (-1.0 * TOL)) {
            if (fabs(x) < TOL) return 0.0;
            return x / (-1.0 * y);
        }
        return DBL_MAX;
    }
    void checkVarForBlocking(int v)                         // genneteq.h:334-360
    {
        double dv = computeDeltaVar(v);
        if (delta > dv + TOL) blockingVars.clear();
        if (delta >= dv - TOL) { delta = dv; blockingVars.push_back(v); }
    }
    void checkVarForOffending(int v, double viol, double *largest)   // gnePivotRules.cpp:431-452
    {
        if (*largest < viol - TOL) offendingVars.clear();
        if (*largest <= viol + TOL) { *largest = viol; offendingVars.push_back(v); }   // trackLowestVarIndex == false
    }

    // ======================= gneFlow.cpp =======================
    void computeFlowsTypeI(OTree *t)                        // gneFlow.cpp:480-560
    {
        int ex = t->extraArc;
        int alpha = tail[ex], beta = head[ex];
        double muAB = mult[ex];
        fa0[ex] = 0.0; fa1[ex] = 1.0;
        sa1[alpha] -= 1.0;
        sa1[beta] += muAB;
        int h = t->root;
        for (long k = (long) t->threads.size() - 1; k >= 0; --k) {
            int j = t->threads[k];
            if (j == h) break;                              // getPrevThread / while (j != h)
            int a = predArc[j];
            if (tail[a] == j) {
                int i = head[a];
                double mu = mult[a];
                fa0[a] = sa0[j]; fa1[a] = sa1[j];
                sa0[i] += sa0[j] * mu;
                sa1[i] += sa1[j] * mu;
            } else {
                int i = tail[a];
                double mu = mult[a];
                fa0[a] = sa0[j] / (-1.0 * mu);
                fa1[a] = sa1[j] / (-1.0 * mu);
                sa0[i] += sa0[j] / mu;
                sa1[i] += sa1[j] / mu;
            }
            sa0[j] = 0.0; sa1[j] = 0.0;
        }
        double theta = (-1.0 * sa0[h]) / sa1[h];
        sa0[h] = 0.0; sa1[h] = 0.0;
        for (size_t k = 0; k < t->arcsIn.size(); ++k) {
            int a = t->arcsIn[k];
            double p = fa1[a] * theta;
            deltaFlow[a] = fa0[a] + p;
            fa0[a] = 0.0; fa1[a] = 0.0;
            checkVarForBlocking(a);
        }
        double p = fa1[ex] * theta;
        deltaFlow[ex] = fa0[ex] + p;
        fa0[ex] = 0.0; fa1[ex] = 0.0;
        checkVarForBlocking(ex);
    }
    void computeFlows(int updateOption)                     // gneFlow.cpp:30-56
    {
        flowCost = 0.0;
        for (int i = 0; i < n; ++i) { sa0[i] = supply[i]; sa1[i] = 0.0; }       // :387-395
        for (size_t p = 0; p < setNB.size(); ++p) {                              // :343-353
            int a = setNB[p];
            if (setLoc[a] == LOC_L) {
                flow[a] = 0.0;
            } else {
                flow[a] = ub[a];
                flowCost += flow[a] * cost[a];
                sa0[tail[a]] -= flow[a];
                sa0[head[a]] += flow[a] * mult[a];
            }
        }
        for (size_t k = 0; k < setBtI.size(); ++k) computeFlowsTypeI(setBtI[k]);
        for (size_t k = 0; k < setB.size(); ++k) {                               // :59-149
            int a = setB[k];
            if (updateOption == 0 || updateOption > 1) flow[a] = deltaFlow[a];
            flowCost += flow[a] * cost[a];
            deltaFlow[a] = 0.0;
        }
    }
    void computeFlowsPivotingTI(int eav, OTree *tU, OTree *tV)   // gneFlow.cpp:155-185
    {
        if (setLoc[eav] == LOC_L) {
            if (tail[eav] == head[eav]) sa0[tail[eav]] = -1.0 * (1.0 - mult[eav]);
            else { sa0[tail[eav]] = -1.0; sa0[head[eav]] = mult[eav]; }
        } else {
            if (tail[eav] == head[eav]) sa0[tail[eav]] = (1.0 - mult[eav]);
            else { sa0[tail[eav]] = 1.0; sa0[head[eav]] = -1.0 * mult[eav]; }
        }
        computeFlowsTypeI(tU);
        if (tU->id != tV->id) computeFlowsTypeI(tV);
    }

    // ======================= gnePotential.cpp =======================
    void computePotentials()                                // gnePotential.cpp:28-61
    {
        for (size_t k = 0; k < setBtI.size(); ++k) {
            OTree *t = setBtI[k];
            // computePotentialsInTermsOfRoot :64-110
            int h = t->threads[0];
            pa0[h] = 0.0; pa1[h] = 1.0;
            for (size_t q = 1; q < t->threads.size(); ++q) {
                int j = t->threads[q];
                int i = pred[j];
                int a = predArc[j];
                double c = cost[a], mu = mult[a];
                if (tail[a] == i) {
                    pa0[j] = (pa0[i] - c) / mu;
                    pa1[j] = pa1[i] / mu;
                } else {
                    double p = mu * pa0[i];
                    pa0[j] = p + c;
                    pa1[j] = mu * pa1[i];
                }
            }
            // finalizePotentialsTypeI :113-138
            int ex = t->extraArc;
            int alpha = tail[ex], beta = head[ex];
            double cAB = cost[ex], muAB = mult[ex];
            double pb0 = muAB * pa0[beta];
            double pb1 = muAB * pa1[beta];
            double theta = ((cAB - pa0[alpha]) + pb0) / (pa1[alpha] - pb1);
            for (size_t q = 0; q < t->nodesIn.size(); ++q) {
                int i = t->nodesIn[q];
                double p = pa1[i] * theta;
                potential[i] = pa0[i] + p;
                pa0[i] = 0.0; pa1[i] = 0.0;
            }
        }
        if (!setBtII.empty()) fail(-5, "type II tree at computePotentials (equal-flow sets are out of scope)");
    }
    void computeReducedCosts(bool includeBasic)             // gnePotential.cpp:217-258
    {
        isOptimal = true;
        offendingVars.clear();
        for (size_t p = 0; p < setNB.size(); ++p) {
            int a = setNB[p];
            double rc = computeArcRedCost(a);
            if (violates(a, rc)) { isOptimal = false; offendingVars.push_back(a); }
        }
        if (includeBasic) for (size_t k = 0; k < setB.size(); ++k) computeArcRedCost(setB[k]);
    }

    // ======================= gnePivotRules.cpp =======================
    int pickOffending()
    {
        if (offendingVars.empty()) return -1;
        return offendingVars[(size_t) (rand48() * offendingVars.size())];   // selectRandomOffendingVar == true
    }
    int getVarWithLargestViolation()                        // :133-157 (also LVA :160-180, LVE :183-203 at p = 0)
    {
        isOptimal = true;
        offendingVars.clear();
        double largest = -1.0;
        for (size_t p = 0; p < setNB.size(); ++p) {         // :401-413
            int a = setNB[p];
            double rc = computeArcRedCost(a);
            if (violates(a, rc)) { isOptimal = false; checkVarForOffending(a, fabs(rc), &largest); }
        }
        lastViolation = largest;
        return pickOffending();
    }
    int getFirstVarWithViolation()                          // :206-228
    {
        isOptimal = true;
        for (size_t p = 0; p < setNB.size(); ++p) {
            int a = setNB[p];
            double rc = computeArcRedCost(a);
            if (violates(a, rc)) { isOptimal = false; return a; }
        }
        return -1;
    }
    int getRandomVarWithViolation()                         // :273-281
    {
        computeReducedCosts(false);
        if (offendingVars.empty()) return -1;
        return offendingVars[(size_t) (rand48() * offendingVars.size())];
    }
    int getRandomVarWeightedByViolation()                   // :284-304
    {
        computeReducedCosts(false);
        if (!offendingVars.empty()) {
            double sum = 0.0;
            for (size_t i = 0; i < offendingVars.size(); ++i) sum += fabs(redCost[offendingVars[i]]);
            double level = rand48() * sum;
            double run = 0.0;
            for (size_t i = 0; i < offendingVars.size(); ++i) {
                run += fabs(redCost[offendingVars[i]]);
                if (level <= run) return offendingVars[i];
            }
        }
        return -1;
    }
    int getRandomVarWithLargeViolation()                    // :307-362
    {
        isOptimal = true;
        ViolHeap q;
        for (size_t p = 0; p < setNB.size(); ++p) {
            int a = setNB[p];
            double rc = computeArcRedCost(a);
            if (violates(a, rc)) {
                isOptimal = false;
                VarViol vv; vv.var = a; vv.violation = -1.0 * fabs(rc);
                q.push(vv);
                while ((int) q.size() > numLargestViolVarsToTrack) q.pop();
            }
        }
        std::vector<int> fin;
        while (!q.empty()) { fin.push_back(q.top().var); q.pop(); }
        if (fin.empty()) return -1;
        return fin[(size_t) (rand48() * fin.size())];
    }
    int getVarBySimAnnealing(bool arcsFirst)                // :364-396
    {
        double r = rand48();
        int best = getVarWithLargestViolation();            // both branches reduce to it at p = 0
        (void) r;
        if (arcsFirst) {
            if (lastViolation * 100 < baselineViolation && baselineViolation > 5000) baselineViolation = lastViolation;
        }
        return best;
    }
    void updateCandidateList()                              // :504-535
    {
        while (!candidateList.empty()) candidateList.pop();
        for (size_t p = 0; p < setNB.size(); ++p) {
            int a = setNB[p];
            double rc = computeArcRedCost(a);
            if (violates(a, rc)) {
                isOptimal = false;
                VarViol vv; vv.var = a; vv.violation = fabs(rc);
                candidateList.push(vv);
            }
        }
    }
    int getNextCandidate()                                  // :457-501
    {
        isOptimal = true;
        if (itersSinceUpdate < 0 || itersSinceUpdate > majorIterationLength) {
            updateCandidateList();
            itersSinceUpdate = 0;
        } else {
            ++itersSinceUpdate;
        }
        int ev = -1;
        while (!candidateList.empty()) {
            int c = candidateList.top().var;
            candidateList.pop();
            bool atLB = (setLoc[c] == LOC_L);
            double rc = computeArcRedCost(c);
            if ((atLB && rc < (-1.0 * TOL)) || (!atLB && rc > TOL)) { isOptimal = false; ev = c; break; }
        }
        if (ev == -1 && itersSinceUpdate > 0) {
            itersSinceUpdate = -1;
            return getNextCandidate();
        }
        return ev;
    }
    void tryToAddArc()                                      // :679-689
    {
        int a = setNB[nbArcPivotInd];
        double rc = computeArcRedCost(a);
        if (violates(a, rc)) { isOptimal = false; candidateList2.push_back(a); }
    }
    void updateCandidateList2()                             // :613-677 with setNBeqf empty
    {
        candidateList2.clear();
        int numNB = (int) setNB.size();
        if (nbArcPivotInd >= numNB) nbArcPivotInd = 0;
        int start = nbArcPivotInd;
        bool exhausted = (numNB == 0);
        for (int i = 0; i < numArcsToAlwaysCheck && !exhausted; ++i) {
            tryToAddArc();
            nbArcPivotInd = (nbArcPivotInd + 1) % numNB;
            exhausted = (nbArcPivotInd == start);
        }
        // exhaustedEqfs is true from the start, so the drand48 interleave loop (:643-653) never runs
        while ((int) candidateList2.size() < clMaxSize && !exhausted) {
            tryToAddArc();
            nbArcPivotInd = (nbArcPivotInd + 1) % numNB;
            exhausted = (nbArcPivotInd == start);
        }
    }
    int getNextCandidate2()                                 // :550-610
    {
        isOptimal = true;
        if (itersSinceUpdate < 0 || itersSinceUpdate > majorIterationLength) {
            updateCandidateList2();
            itersSinceUpdate = 0;
        } else {
            ++itersSinceUpdate;
        }
        int ev = -1, evInd = -1;
        double largest = -1.0;
        size_t k = 0;
        while (k < candidateList2.size()) {
            int c = candidateList2[k];
            bool atLB = (setLoc[c] == LOC_L);
            double rc = (itersSinceUpdate > 0) ? computeArcRedCost(c) : redCost[c];
            if ((atLB && rc < (-1.0 * TOL)) || (!atLB && rc > TOL)) {
                isOptimal = false;
                if (largest < fabs(rc)) { largest = fabs(rc); ev = c; evInd = (int) k; }
                ++k;
            } else {
                candidateList2[k] = candidateList2.back();
                candidateList2.pop_back();
            }
        }
        if (ev == -1 && itersSinceUpdate > 0) {
            itersSinceUpdate = -1;
            return getNextCandidate2();
        }
        if (evInd >= 0) {
            candidateList2[evInd] = candidateList2.back();
            candidateList2.pop_back();
        }
        return ev;
    }
    int quickSelect()                                       // :710-764 (arc loop only)
    {
        isOptimal = true;
        int best = -1;
        double largest = 0.0;
        if (nbArcPivotInd >= (int) setNB.size()) nbArcPivotInd = 0;
        int start = nbArcPivotInd;
        int viol = 0;
        bool exhausted = setNB.empty();
        while (viol < n && !exhausted) {
            int a = setNB[nbArcPivotInd];
            double rc = computeArcRedCost(a);
            if (violates(a, rc)) {
                isOptimal = false;
                ++viol;
                if (fabs(rc) > largest) { largest = fabs(rc); best = a; }
            }
            nbArcPivotInd = (nbArcPivotInd + 1) % (int) setNB.size();
            exhausted = (nbArcPivotInd == start);
        }
        if ((loopIter % majorIterationLength) > 0) nbArcPivotInd = start;
        return best;
    }
    int selectEnteringVariable()                            // :36-127
    {
        if (cyclingDetected) return getRandomVarWithLargeViolation();   // useRandomJumpWhenCycling == true
        int rule = presolve ? ruleI : ruleII;
        switch (rule) {
          case R_LV: case R_LVW: case R_LVA: case R_LVE: return getVarWithLargestViolation();
          case R_FV: return getFirstVarWithViolation();
          case R_RV: return getRandomVarWithViolation();
          case R_RWV: return getRandomVarWeightedByViolation();
          case R_RLV: case R_RLVW: return getRandomVarWithLargeViolation();
          case R_CL: case R_CLW: return getNextCandidate();
          case R_SAA: return getVarBySimAnnealing(true);
          case R_SAE: return getVarBySimAnnealing(false);
          case R_CL2: case R_CLW2: return getNextCandidate2();
          case R_QS: case R_QSW: return quickSelect();
        }
        fail(-6, "pivot rule not restated (SD uses the stale computeMinDelta, gnePivot.cpp:406)");
        return -1;
    }

    // ======================= gnePivot.cpp =======================
    void identifyLeavingVar()                               // :166-209
    {
        if (ub[eVar] <= delta + TOL) {
            delta = ub[eVar];
            lVar = eVar;
        } else {
            if (blockingVars.empty()) { isUnbounded = true; lVar = -1; return; }   // (:172-177 would index an empty vector)
            lVar = blockingVars[(size_t) (rand48() * blockingVars.size())];        // selectRandomBlockingVar == true
        }
        if (delta <= TOL) { ++totalDegenPivots; ++consecutiveDegenPivots; }
        else consecutiveDegenPivots = 0;
    }
    void applyTreeFlow(OTree *t)                            // :79-91
    {
        for (size_t k = 0; k < t->arcsIn.size(); ++k) {
            int a = t->arcsIn[k];
            double d = delta * deltaFlow[a];
            flow[a] += d;
            flowCost += cost[a] * d;
            deltaFlow[a] = 0.0;
        }
        int ex = t->extraArc;
        double d = delta * deltaFlow[ex];
        flow[ex] += d;
        flowCost += cost[ex] * d;
        deltaFlow[ex] = 0.0;
    }
    void pivotTI(int eav)                                   // :63-119
    {
        OTree *tU = setBtI[memberTree[tail[eav]]];
        OTree *tV = setBtI[memberTree[head[eav]]];
        delta = DBL_MAX;
        computeFlowsPivotingTI(eav, tU, tV);
        identifyLeavingVar();
        if (lVar == -1) return;
        applyTreeFlow(tU);
        if (tU->id != tV->id) applyTreeFlow(tV);
        if (setLoc[eav] == LOC_L) { flow[eav] = delta; flowCost += cost[eav] * delta; }
        else { flow[eav] = ub[eav] - delta; flowCost -= cost[eav] * delta; }
    }
    void updateBasis()                                      // :254-309
    {
        if (eVar == lVar) {
            int was = setLoc[eVar];
            removeFromNB(eVar);
            insertIntoNB(eVar, was == LOC_L ? LOC_U : LOC_L);
        } else {
            removeFromNB(eVar);
            removeFromB(lVar);
            insertIntoB(eVar);
            if (flow[lVar] > (0.5 * ub[lVar])) insertIntoNB(lVar, LOC_U);
            else insertIntoNB(lVar, LOC_L);
            updateTrees();
        }
        if (setLoc[lVar] == LOC_L) {
            if (fabs(flow[lVar] - 0.0) > TOL) flow[lVar] = 0.0;
        } else {
            if (fabs(flow[lVar] - ub[lVar]) > TOL) flow[lVar] = ub[lVar];
        }
    }
    void checkForCycling()                                  // :312-356
    {
        if (cycleList.size() >= 10) cycleList.pop_back();
        cycleList.push_front(std::make_pair(eVar, lVar));
        cyclingDetected = false;
        std::list<std::pair<int, int> >::iterator it = cycleList.begin();
        for (++it; it != cycleList.end(); ++it) {
            if (cycleList.front().first == it->first && cycleList.front().second == it->second) {
                cyclingDetected = true;
                break;
            }
        }
        if (cyclingDetected) {
            ++totalCyclesDetected;
            ++consecutiveCycles;
            if (consecutiveCycles > 5) fail(-7, "Stuck in cycling loop");
        } else {
            consecutiveCycles = 0;
        }
        lastEVarInd = eVar;
    }
    void pivot()                                            // :28-60
    {
        if (memberType[tail[eVar]] == T_I && memberType[head[eVar]] == T_I) pivotTI(eVar);
        else { fail(-8, "pivotTII reached (equal-flow sets are out of scope)"); lVar = -1; }
        if (lVar == -1) return;
        updateBasis();
        checkForCycling();
    }

    // ======================= genneteq.cpp =======================
    void initializeStats(bool usePresolve)                  // :123-167
    {
        presolve = usePresolve;
        isOptimal = false; isInfeasible = false; isUnbounded = false;
        loopIter = 0;
        consecutiveCycles = 0; consecutiveDegenPivots = 0; totalCyclesDetected = 0; totalDegenPivots = 0;
        cyclingDetected = false;
        lastEVarInd = -1;
        itersSinceUpdate = -1;
        numLargestViolVarsToTrack = 5;
        majorIterationLength = 3;
        clMaxSize = n;
        nbArcPivotInd = 0;
        numArcsToAlwaysCheck = 5;
        baselineViolation = lastViolation;
    }
    void setVariableCosts()                                 // :304-323 (initialType == SELFLOOPS)
    {
        if (presolve) for (long i = 0; i < M; ++i) cost[i] = isArt[i] ? 1.0 : 0.0;
        else for (long i = 0; i < M; ++i) cost[i] = actualCost[i];
    }
    void checkFeasibility()                                 // :227-288
    {
        isInfeasible = false;
        for (size_t k = 0; k < setB.size(); ++k) {
            int a = setB[k];
            if ((isArt[a] && flow[a] > TOL) || flow[a] < 0.0 - TOL || flow[a] > ub[a] + TOL) { isInfeasible = true; return; }
        }
        for (size_t p = 0; p < setNB.size(); ++p) {
            int a = setNB[p];
            if (isArt[a] && setLoc[a] == LOC_U) { isInfeasible = true; return; }
        }
    }
    long solve(bool usePresolve, long maxPivots, int *finished)   // :60-117
    {
        bool pIwasInfeasible = usePresolve ? false : isInfeasible;
        initializeStats(usePresolve);
        setVariableCosts();
        *finished = 1;
        if (!presolve && pIwasInfeasible) { isInfeasible = true; return loopIter; }
        double t0 = wallNow();
        computeFlows(0);
        long budget = maxPivots;
        long callPivots = 0;
        while (!isOptimal && !isUnbounded && status == 0) {
            if (callPivots == windowStart) { winTimers[0] = secsPricing; winTimers[1] = secsPotentials; winTimers[2] = secsPivot; winTimers[3] = 0.0; winTimers[4] = (double) arcsPriced; }
            if (maxPivots >= 0 && budget == 0) { *finished = 0; break; }
            if ((loopIter % 1000) == 0) computeFlows(2);
            double a = wallNow();
            computePotentials();
            double b = wallNow();
            arcsPriced += (long) setNB.size();
            eVar = selectEnteringVariable();
            double c = wallNow();
            secsPotentials += b - a; secsPricing += c - b;
            if (eVar != -1) {
                pivot();
                secsPivot += wallNow() - c;
                if (lVar != -1) {
                    if (traceLen < traceCap) { traceE[traceLen] = eVar; traceL[traceLen] = lVar; }
                    ++traceLen;
                }
                ++loopIter; ++callPivots;
                if (presolve) ++phaseIiters; else ++phaseIIiters;
                if (budget > 0) --budget;
            }
        }
        if (*finished && status == 0) {
            computeFlows(0);
            if (presolve) checkFeasibility();
            computePotentials();                            // verifyOptimality :261-287
            computeReducedCosts(true);
        }
        secsTotal += wallNow() - t0;
        return loopIter;
    }
};

// ------------------------------------------------------------------------------------------
extern "C" {

double orc_big_m(long m, const double *cost, const double *ub)
{
    double maxCost = 0.0, maxCap = 0.0;                     // graph.cpp:36-37, 412-417
    for (long a = 0; a < m; ++a) {
        if (maxCost < cost[a]) maxCost = cost[a];
        if (maxCap < ub[a]) maxCap = ub[a];
    }
    int numArcs = (int) m;
    return numArcs * maxCost * maxCap;                      // graph.cpp:43
}

// computeEqfRedCost (genneteq.h:321-330), dense as the reference runs it: rc = cost; for every node i in order
// rc -= nodeVals[i] * potential[i].  nodeVals: p rows of n doubles.
void orc_price_eqf_dense(int p, int n, const double *nodeVals, const double *cost, const double *pi, double *rc)
{
    for (int k = 0; k < p; ++k) {
        double x = cost[k];
        for (int i = 0; i < n; ++i) x -= nodeVals[(size_t) k * n + i] * pi[i];
        rc[k] = x;
    }
}

double orc_drand48(unsigned long long *state)
{
    const unsigned long long a = 0x5DEECE66DULL, c = 0xBULL, mask = (1ULL << 48) - 1ULL;
    *state = (a * (*state) + c) & mask;
    return ldexp((double) *state, -48);
}

unsigned long long orc_drand48_initial_state(void) { return 0ULL; }   // glibc never-seeded state (verified in tests)

OrcSolver *orc_create(int n, long m, const int *tailIn, const int *headIn, const double *ubIn,
                      const double *costIn, const double *multIn, const double *supplyIn, double bigM)
{
    OrcSolver *s = new OrcSolver();
    // gneInit.cpp:25-44
    s->n = n; s->bigM = bigM; s->status = 0;
    s->rng = orc_drand48_initial_state();
    s->ruleI = R_LVW; s->ruleII = R_LVW;                    // main.cpp:72-73
    s->traceE = s->traceL = NULL; s->traceCap = 0; s->traceLen = 0;
    s->windowStart = -1; for (int k = 0; k < 5; ++k) s->winTimers[k] = 0.0;
    s->secsPricing = s->secsPotentials = s->secsPivot = s->secsTotal = 0.0; s->arcsPriced = 0;
    s->isInfeasible = false; s->isOptimal = false; s->isUnbounded = false; s->presolve = true;
    s->cyclingDetected = false; s->eVar = s->lVar = -1; s->delta = 0.0; s->flowCost = 0.0;
    s->loopIter = 0; s->itersSinceUpdate = -1; s->nbArcPivotInd = 0; s->lastEVarInd = -1;
    // nodes: gneInit.cpp:46-56
    s->potential.assign(n, 0.0); s->supply.assign(supplyIn, supplyIn + n);
    s->pa0.assign(n, 0.0); s->pa1.assign(n, 0.0); s->sa0.assign(n, 0.0); s->sa1.assign(n, 0.0);
    s->memberTree.assign(n, -1); s->memberType.assign(n, T_NULL);
    s->depth.assign(n, -1); s->thread.assign(n, -1); s->pred.assign(n, -1); s->predArc.assign(n, -1);
    s->incident.resize(n);
    // arcs: gneInit.cpp:82-107 (walk over the (tail, head)-sorted list == the n^2 pair walk)
    for (long a = 0; a < m; ++a) {
        if (ubIn[a] <= 0.0) continue;
        s->tail.push_back(tailIn[a]); s->head.push_back(headIn[a]);
        s->ub.push_back(ubIn[a]); s->actualCost.push_back(costIn[a]); s->mult.push_back(multIn[a]);
        s->isArt.push_back(0);
    }
    s->m = (long) s->tail.size();
    for (int i = 0; i < n; ++i) {                           // gneInit.cpp:110-127
        s->tail.push_back(i); s->head.push_back(i);
        s->ub.push_back(DBL_MAX); s->actualCost.push_back(bigM);
        s->mult.push_back(supplyIn[i] >= 0.0 ? 0.5 : 2.0);
        s->isArt.push_back(1);
    }
    s->M = (long) s->tail.size();
    s->cost.assign(s->M, 0.0); s->flow.assign(s->M, 0.0); s->deltaFlow.assign(s->M, 0.0);
    s->redCost.assign(s->M, 0.0); s->fa0.assign(s->M, 0.0); s->fa1.assign(s->M, 0.0);
    s->setLoc.assign(s->M, LOC_NIL); s->setInd.assign(s->M, -1);
    for (long a = 0; a < s->M; ++a) {                       // gneInit.cpp:174-185
        if (s->isArt[a]) s->insertIntoB((int) a);
        else s->insertIntoNB((int) a, LOC_L);
    }
    s->updateTrees();
    s->phaseIiters = 0; s->phaseIIiters = 0;
    return s;
}

void orc_destroy(OrcSolver *s)
{
    if (!s) return;
    for (size_t k = 0; k < s->setBtI.size(); ++k) delete s->setBtI[k];
    for (size_t k = 0; k < s->setBtII.size(); ++k) delete s->setBtII[k];
    delete s;
}

void orc_set_rules(OrcSolver *s, int p1, int p2) { s->ruleI = p1; s->ruleII = p2; }
void orc_set_trace(OrcSolver *s, int *e, int *l, long cap) { s->traceE = e; s->traceL = l; s->traceCap = cap; s->traceLen = 0; }
long orc_trace_len(const OrcSolver *s) { return s->traceLen; }
long orc_solve(OrcSolver *s, int presolve, long maxPivots, int *finished) { return s->solve(presolve != 0, maxPivots, finished); }
int orc_status(const OrcSolver *s) { return s->status; }
int orc_num_nodes(const OrcSolver *s) { return s->n; }
long orc_num_vars(const OrcSolver *s) { return s->M; }
long orc_num_nonbasic(const OrcSolver *s) { return (long) s->setNB.size(); }
double orc_objective(const OrcSolver *s) { return s->flowCost; }
int orc_feasible(const OrcSolver *s) { return s->isInfeasible ? 0 : 1; }
void orc_get_flows(const OrcSolver *s, double *out) { memcpy(out, &s->flow[0], sizeof(double) * s->M); }
void orc_get_potentials(const OrcSolver *s, double *out) { memcpy(out, &s->potential[0], sizeof(double) * s->n); }
void orc_get_states(const OrcSolver *s, int *out) { for (long i = 0; i < s->M; ++i) out[i] = s->setLoc[i]; }
void orc_get_nonbasic_order(const OrcSolver *s, int *out) { for (size_t i = 0; i < s->setNB.size(); ++i) out[i] = s->setNB[i]; }
void orc_get_costs(const OrcSolver *s, double *out) { memcpy(out, &s->cost[0], sizeof(double) * s->M); }
void orc_get_timers(const OrcSolver *s, double *o)
{
    o[0] = s->secsPricing; o[1] = s->secsPotentials; o[2] = s->secsPivot; o[3] = s->secsTotal; o[4] = (double) s->arcsPriced;
}
void orc_set_window(OrcSolver *s, long start) { s->windowStart = start; }
void orc_get_window_timers(const OrcSolver *s, double *o) { for (int k = 0; k < 5; ++k) o[k] = s->winTimers[k]; }
void orc_get_forest(const OrcSolver *s, int *treeId, int *pred, int *predArc, int *depth, int *thread)
{
    for (int i = 0; i < s->n; ++i) {
        treeId[i] = s->memberTree[i]; pred[i] = s->pred[i]; predArc[i] = s->predArc[i];
        depth[i] = s->depth[i]; thread[i] = s->thread[i];
    }
}
void orc_compute_potentials(OrcSolver *s) { s->computePotentials(); }

// ---- stateless kernels ------------------------------------------------------------------
static inline double arcRedCost(long p, const int *tail, const int *head, const double *cost,
                                const double *mult, const double *pi)
{
    double t = mult[p] * pi[head[p]];                       // genneteq.h:315-316, no FMA
    return (cost[p] - pi[tail[p]]) + t;
}
static inline bool arcViolates(signed char st, double rc)
{
    return (st == LOC_L && rc < (-1.0 * TOL)) || (st == LOC_U && rc > TOL);
}

void orc_reduced_costs(long N, const int *tail, const int *head, const double *cost,
                       const double *mult, const double *pi, double *rc)
{
    for (long p = 0; p < N; ++p) rc[p] = arcRedCost(p, tail, head, cost, mult, pi);
}

long orc_price_lv(long N, const int *tail, const int *head, const double *cost, const double *mult,
                  const signed char *state, const double *pi, long *listPos, long cap,
                  double *largest, int *anyViolation)
{
    std::vector<long> list;
    double L = -1.0;                                        // gnePivotRules.cpp:137
    *anyViolation = 0;
    for (long p = 0; p < N; ++p) {                          // :401-413
        double rc = arcRedCost(p, tail, head, cost, mult, pi);
        if (!arcViolates(state[p], rc)) continue;
        *anyViolation = 1;
        double v = fabs(rc);
        if (L < v - TOL) list.clear();                      // :431-452
        if (L <= v + TOL) { L = v; list.push_back(p); }
    }
    *largest = L;
    for (size_t i = 0; i < list.size() && (long) i < cap; ++i) listPos[i] = list[i];
    return (long) list.size();
}

void orc_price_qs(long N, const int *tail, const int *head, const double *cost, const double *mult,
                  const signed char *state, const double *pi, long cursor, long want,
                  long *bestPos, double *bestViol, long *newCursor, long *scanned, long *violations)
{
    *bestPos = -1; *bestViol = 0.0; *scanned = 0; *violations = 0;
    if (cursor >= N) cursor = 0;                            // gnePivotRules.cpp:716
    long start = cursor;
    bool exhausted = (N == 0);
    while (*violations < want && !exhausted) {              // :742-756
        double rc = arcRedCost(cursor, tail, head, cost, mult, pi);
        if (arcViolates(state[cursor], rc)) {
            ++*violations;
            if (fabs(rc) > *bestViol) { *bestViol = fabs(rc); *bestPos = cursor; }
        }
        ++*scanned;
        cursor = (cursor + 1) % N;
        exhausted = (cursor == start);
    }
    *newCursor = cursor;
}

long orc_price_violators(long N, const int *tail, const int *head, const double *cost,
                         const double *mult, const signed char *state, const double *pi,
                         double threshold, long *listPos, double *listViol, long cap)
{
    long cnt = 0;
    for (long p = 0; p < N; ++p) {
        double rc = arcRedCost(p, tail, head, cost, mult, pi);
        if (!arcViolates(state[p], rc)) continue;
        double v = fabs(rc);
        if (v < threshold) continue;
        if (cnt < cap) { listPos[cnt] = p; if (listViol) listViol[cnt] = v; }
        ++cnt;
    }
    return cnt;
}

void orc_finalize_potentials(long count, const int *node, const double *a0, const double *a1,
                             double theta, double *pi)
{
    for (long i = 0; i < count; ++i) {
        double p = a1[i] * theta;
        pi[node[i]] = a0[i] + p;
    }
}

double orc_flow_apply(long count, const int *arc, const double *deltaFlow, const double *cost,
                      double delta, double *flow)
{
    double inc = 0.0;
    for (long i = 0; i < count; ++i) {
        double d = delta * deltaFlow[i];
        flow[arc[i]] += d;
        inc += cost[arc[i]] * d;
    }
    return inc;
}

} // extern "C"
