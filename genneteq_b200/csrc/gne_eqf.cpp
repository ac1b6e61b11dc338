// genneteq_b200 — solver for instances with equal-flow sets (see gne_eqf.h).  All `file:line` citations are into the
// reference.  Host arithmetic is compiled -ffp-contract=off: every a*b+c below rounds twice, as the reference's does.
#include "gne_eqf.h"
#include "gne_internal.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stack>
#include <time.h>

namespace gneb200 {

static const double TOL = GNE_ROUNDOFF_TOLERANCE;                  // main.h:36-43
#define GNE_TRY(call) do { int rc_ = (call); if (rc_ != GNE_OK) return rc_; } while (0)

static double wallNow()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1.0e-9 * (double) ts.tv_nsec;
}

double EqfSolver::drand48()                                        // glibc, never seeded by the reference: X0 = 0
{
    rand48State = (0x5DEECE66DULL * rand48State + 0xBULL) & ((1ULL << 48) - 1ULL);
    return ldexp((double) rand48State, -48);
}

EqfSolver::~EqfSolver()
{
    for (size_t i = 0; i < setBtI.size(); ++i) delete setBtI[i];
    for (size_t i = 0; i < setBtII.size(); ++i) delete setBtII[i];
    if (dev) gne_dev_destroy(dev);
}

void EqfSolver::getCounters(double *out12) const
{
    int64_t h2d = 0, d2h = 0;
    if (dev) gne_dev_traffic(dev, &h2d, &d2h);
    out12[0] = secsPricing; out12[1] = secsPotentials; out12[2] = secsPivot; out12[3] = secsTotal;
    out12[4] = arcsPriced; out12[5] = 0.0; out12[6] = (double) h2d; out12[7] = (double) d2h;
    out12[8] = dev ? (double) gne_dev_launch_count(dev) : 0.0; out12[9] = (double) pivots;
    out12[10] = (double) pivotsTII; out12[11] = 0.0;
}

// ------------------------------------------------------------------------------------------
// gneInit.cpp:25-186 (initialType is forced to SELFLOOPS there, :31)
// ------------------------------------------------------------------------------------------
int EqfSolver::initialize(const EqfInstance &g, int cudaDev, bool hostOnlyReplay)
{
    hostOnly = hostOnlyReplay;
    if (g.n <= 0 || g.p <= 0) { gne_set_error("EqfSolver: needs nodes and at least one equal-flow set"); return GNE_ERR_ARG; }
    cudaDevice = cudaDev;
    numNodes = g.n; numSets = g.p; bigM = g.bigM;
    const int64_t m = (int64_t) g.tail.size();
    A = m + numNodes;
    V = A + numSets;
    tail.assign((size_t) A, 0); head.assign((size_t) A, 0); mult.assign((size_t) A, 0.0);
    UB.assign((size_t) V, 0.0); actualCost.assign((size_t) V, 0.0); isArtificial.assign((size_t) V, 0);
    for (int64_t a = 0; a < m; ++a) {                   // :82-95
        tail[a] = g.tail[a]; head[a] = g.head[a]; UB[a] = g.cap[a]; actualCost[a] = g.cost[a]; mult[a] = g.mult[a];
    }
    supply = g.supply;
    for (int i = 0; i < numNodes; ++i) {                // :110-127 artificial self-loops
        const int64_t a = m + i;
        tail[a] = i; head[a] = i; UB[a] = DBL_MAX; actualCost[a] = bigM;
        mult[a] = (supply[i] >= 0.0) ? 0.5 : 2.0;
        isArtificial[a] = 1;
    }
    numOrigArcs = g.numOrigArcs;
    for (int r = 0; r < numSets; ++r) {                 // :64-79, 96-103 (isArtificial never set for a set: `r == numEqualFlowSets` is never true)
        UB[A + r] = g.setUB[r]; actualCost[A + r] = g.setCost[r];
    }
    nodeVals = g.nodeVals;
    csrStart.assign((size_t) numSets + 1, 0); csrNode.clear(); csrVal.clear();
    for (int r = 0; r < numSets; ++r) {
        for (int i = 0; i < numNodes; ++i) {
            const double d = nodeVals[(size_t) r * numNodes + i];
            if (d != 0.0) { csrNode.push_back(i); csrVal.push_back(d); }
        }
        csrStart[(size_t) r + 1] = (int64_t) csrNode.size();
    }
    cost.assign((size_t) V, 0.0); flow.assign((size_t) V, 0.0); deltaFlow.assign((size_t) V, 0.0);
    flowA0.assign((size_t) A, 0.0); flowA1.assign((size_t) A, 0.0);
    setLoc.assign((size_t) V, NIL); setInd.assign((size_t) V, -1);
    potential.assign(numNodes, 0.0); potA0.assign(numNodes, 0.0); potA1.assign(numNodes, 0.0); potTheta.assign(numNodes, -2);
    supA0.assign(numNodes, 0.0); supA1.assign(numNodes, 0.0);
    memberOfTreeID.assign(numNodes, -1); memberOfTreeType.assign(numNodes, T_NONE);
    depth.assign(numNodes, -1); thread.assign(numNodes, -1); pred.assign(numNodes, -1); predArc.assign(numNodes, -1);
    incidentArcs.assign(numNodes, std::vector<int32_t>());
    threadPos.assign(numNodes, -1); posInTree.assign((size_t) A, -1); nodeStamp.assign(numNodes, 0); stampNow = 0;
    sparseFlows = !(getenv("GNE_EQF_DENSE") && atoi(getenv("GNE_EQF_DENSE")) != 0);
    // the nonbasic arc list never holds more than m + p arcs (|setBarc| = n - |setBeqf| >= n - p)
    if (!hostOnly) {
        GNE_TRY(gne_dev_create(&dev, cudaDevice, numNodes, m + numSets + 16, 0, m + numSets + 16));
        GNE_TRY(gne_eqf_set_rows(dev, numSets, csrStart.data(), csrNode.data(), csrVal.data()));
        setsResident = true;
    }
    for (int64_t a = 0; a < A; ++a) {                   // formInitialBFS :174-186
        if (isArtificial[a]) GNE_TRY(insertIntoB((int32_t) a));
        else GNE_TRY(insertIntoNB((int32_t) a, L_var));
    }
    for (int r = 0; r < numSets; ++r) GNE_TRY(insertIntoNB((int32_t) (A + r), L_var));
    updateTrees();
    phaseIiters = 0; phaseIIiters = 0;
    return GNE_OK;
}

int EqfSolver::setComm(const uint8_t id[128], int rank, int nranks, gne_allgather_fn hostGather, void *hostCtx)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) { gne_set_error("setComm: bad rank %d of %d", rank, nranks); return GNE_ERR_ARG; }
    if (hostOnly || costsOnDevice) { gne_set_error("setComm must precede the first solve"); return GNE_ERR_ARG; }
    // the handle made by initialize() covers the whole list: replace it by one for this rank's position range
    const int64_t list = A - numNodes + numSets;        // the nonbasic arc list never grows beyond m + p
    int64_t lo = 0, hi = 0;
    GNE_TRY(gne_shard_range(list, rank, nranks, &lo, &hi));
    if (dev) { gne_dev_destroy(dev); dev = nullptr; }
    GNE_TRY(gne_dev_create(&dev, cudaDevice, numNodes, list + 16, lo, hi));
    GNE_TRY(gne_eqf_set_rows(dev, numSets, csrStart.data(), csrNode.data(), csrVal.data()));
    if (nranks > 1) {
        if (hostGather) GNE_TRY(gne_comm_init_host(dev, rank, nranks, hostGather, hostCtx));
        else GNE_TRY(gne_comm_init(dev, id, rank, nranks));
    }
    return GNE_OK;
}

// ------------------------------------------------------------------------------------------
// trees.cpp, restated on flat node arrays (ids, not pointers)
// ------------------------------------------------------------------------------------------
EqfSolver::ETree *EqfSolver::createNewTree(int8_t type, int rNode)      // :183-203
{
    ETree *t = new ETree();
    std::vector<ETree *> &set = (type == T_I) ? setBtI : setBtII;
    t->id = (int) set.size(); t->type = type; t->root = -1; t->upToDate = false; t->extraArc = -1;
    set.push_back(t);
    if (rNode >= 0) {
        t->root = rNode;
        t->nodesInTree.push_back(rNode);
        memberOfTreeID[rNode] = t->id; memberOfTreeType[rNode] = t->type;
    }
    return t;
}

void EqfSolver::setTreeIDAndType(ETree *t, int newID, int8_t newType)   // :218-229
{
    t->id = newID; t->type = newType;
    for (size_t k = 0; k < t->nodesInTree.size(); ++k) {
        const int i = t->nodesInTree[k];
        memberOfTreeID[i] = newID; memberOfTreeType[i] = newType;
    }
}

void EqfSolver::removeTree(ETree *t)                               // :232-256: the last tree of the type takes the freed id
{
    std::vector<ETree *> &set = (t->type == T_I) ? setBtI : setBtII;
    const int oldID = t->id;
    if ((size_t) oldID < set.size() - 1) {
        ETree *last = set.back();
        setTreeIDAndType(last, oldID, t->type);
        set[(size_t) oldID] = last;
    }
    set.pop_back();
}

void EqfSolver::convertTItoTII(ETree *t)                           // :259-272 (upToDate is left as it is)
{
    const int newID = (int) setBtII.size();
    removeTree(t);
    setTreeIDAndType(t, newID, T_II);
    t->extraArc = -1;
    setBtII.push_back(t);
}

void EqfSolver::convertTIItoTI(ETree *t, int32_t av)               // :275-288
{
    const int newID = (int) setBtI.size();
    removeTree(t);
    setTreeIDAndType(t, newID, T_I);
    t->extraArc = av;
    t->potDirty = true;                                 // (theta comes from the new extra arc even when the rooting stays)
    setBtI.push_back(t);
}

int EqfSolver::insertIntoTrees(int32_t av)                         // :294-363
{
    const int u = tail[av], v = head[av];
    const int tUID = memberOfTreeID[u], tVID = memberOfTreeID[v];
    const int8_t tUType = memberOfTreeType[u], tVType = memberOfTreeType[v];
    if (tUID != -1 && tVID != -1) {
        if (tUID == tVID && tUType == tVType) {
            if (tUType == T_I) { gne_set_error("Error: Insertion adds cycle to type I tree"); return GNE_ERR_TREE; }
            convertTIItoTI(setBtII[(size_t) tUID], av);
        } else if (tUType == T_II && tVType == T_II) {
            mergeTxAndTII(setBtII[(size_t) tUID], setBtII[(size_t) tVID], av);
        } else if (tUType == T_I && tVType == T_II) {
            mergeTxAndTII(setBtI[(size_t) tUID], setBtII[(size_t) tVID], av);
        } else if (tUType == T_II && tVType == T_I) {
            mergeTxAndTII(setBtI[(size_t) tVID], setBtII[(size_t) tUID], av);
        } else {
            gne_set_error("Error: Insertion merges two type I trees"); return GNE_ERR_TREE;
        }
    } else {
        ETree *tUV;
        if (tUID == -1 && tVID == -1) tUV = createNewTree(u == v ? T_I : T_II, u);
        else if (tUID != -1) tUV = (tUType == T_I) ? setBtI[(size_t) tUID] : setBtII[(size_t) tUID];
        else tUV = (tVType == T_I) ? setBtI[(size_t) tVID] : setBtII[(size_t) tVID];
        expandTreeUsingArc(tUV, av);
    }
    return GNE_OK;
}

void EqfSolver::mergeTxAndTII(ETree *tX, ETree *tII, int32_t av)   // :366-392
{
    removeTree(tII);                                    // (may hand tII's id to tX when tX is the last type II tree)
    setTreeIDAndType(tII, tX->id, tX->type);
    tX->nodesInTree.insert(tX->nodesInTree.end(), tII->nodesInTree.begin(), tII->nodesInTree.end());
    for (size_t k = 0; k < tII->arcsInTree.size(); ++k) {
        posInTree[(size_t) tII->arcsInTree[k]] = (int32_t) tX->arcsInTree.size();
        tX->arcsInTree.push_back(tII->arcsInTree[k]);
    }
    posInTree[(size_t) av] = (int32_t) tX->arcsInTree.size();
    tX->arcsInTree.push_back(av);
    incidentArcs[tail[av]].push_back(av);
    incidentArcs[head[av]].push_back(av);
    tX->upToDate = false;
    tX->potDirty = true;
    delete tII;
}

void EqfSolver::expandTreeUsingArc(ETree *t, int32_t av)           // :395-428 (the id alone is compared, as there)
{
    const int u = tail[av], v = head[av];
    if (memberOfTreeID[u] != t->id) {
        t->nodesInTree.push_back(u);
        memberOfTreeID[u] = t->id; memberOfTreeType[u] = t->type;
    } else if (memberOfTreeID[v] != t->id) {
        t->nodesInTree.push_back(v);
        memberOfTreeID[v] = t->id; memberOfTreeType[v] = t->type;
    }
    if (u != v) {
        posInTree[(size_t) av] = (int32_t) t->arcsInTree.size();
        t->arcsInTree.push_back(av);
        incidentArcs[u].push_back(av);
        incidentArcs[v].push_back(av);
    } else {
        t->extraArc = av;
    }
    t->upToDate = false;
    t->potDirty = true;
}

int EqfSolver::removeFromTrees(int32_t av)                         // :450-494
{
    const int u = tail[av], v = head[av];
    const int tUID = memberOfTreeID[u], tVID = memberOfTreeID[v];
    const int8_t tUType = memberOfTreeType[u], tVType = memberOfTreeType[v];
    if (tUID != tVID || tUType != tVType) { gne_set_error("Error: Attempting to remove arc spanning two trees"); return GNE_ERR_TREE; }
    ETree *tUV = (tUType == T_I) ? setBtI[(size_t) tUID] : setBtII[(size_t) tUID];
    if (tUV->type == T_I && av == tUV->extraArc) {
        convertTItoTII(tUV);
    } else {
        removeTree(tUV);
        splitTree(tUV, av);
        if (tUV->type == T_I) GNE_TRY(insertIntoTrees(tUV->extraArc));
        delete tUV;
    }
    return GNE_OK;
}

void EqfSolver::splitTree(ETree *tX, int32_t removedArc)           // :497-566: the DFS that defines the parts' node and arc order
{
    ETree *tTop = createNewTree(T_II, -1);
    ETree *tBottom = createNewTree(T_II, -1);
    ETree *tCur = tTop;
    std::stack<int> curNodeStack, prevNodeStack;
    curNodeStack.push(tX->root);
    prevNodeStack.push(-1);
    while (!curNodeStack.empty()) {
        int curNode = curNodeStack.top(); curNodeStack.pop();
        if (curNode < 0) {                              // the bottom part is done
            tCur = tTop;
            if (curNodeStack.empty()) break;
            curNode = curNodeStack.top(); curNodeStack.pop();
        }
        const int prevNode = prevNodeStack.top(); prevNodeStack.pop();
        memberOfTreeID[curNode] = tCur->id; memberOfTreeType[curNode] = tCur->type;
        if (tCur->nodesInTree.empty()) tCur->root = curNode;
        tCur->nodesInTree.push_back(curNode);
        std::vector<int32_t> &inc = incidentArcs[curNode];
        int removedArcInd = -1;
        for (int k = 0; k < (int) inc.size(); ++k) {
            const int32_t nextArc = inc[(size_t) k];
            if (nextArc == removedArc) { removedArcInd = k; continue; }
            const int nextNode = (tail[nextArc] == curNode) ? head[nextArc] : tail[nextArc];
            if (nextNode != prevNode) {
                posInTree[(size_t) nextArc] = (int32_t) tCur->arcsInTree.size();
                tCur->arcsInTree.push_back(nextArc);
                curNodeStack.push(nextNode);
                prevNodeStack.push(curNode);
            }
        }
        if (removedArcInd >= 0) {
            inc[(size_t) removedArcInd] = inc.back();
            inc.pop_back();
            if (prevNode > -2) {                        // first end of the removed arc: its other end starts the bottom part
                const int nextNode = (tail[removedArc] == curNode) ? head[removedArc] : tail[removedArc];
                curNodeStack.push(-1);
                curNodeStack.push(nextNode);
                prevNodeStack.push(-2);
                tCur = tBottom;
            }
        }
    }
}

void EqfSolver::updateTrees()                                      // :572-591
{
    for (size_t k = 0; k < setBtI.size(); ++k) {
        ETree *t = setBtI[k];
        if (!t->upToDate) { t->root = tail[t->extraArc]; initializeTreeIndices(t); t->upToDate = true; t->potDirty = true; }
    }
    for (size_t k = 0; k < setBtII.size(); ++k) {
        ETree *t = setBtII[k];
        if (!t->upToDate) { initializeTreeIndices(t); t->upToDate = true; }
    }
}

void EqfSolver::initializeTreeIndices(ETree *t)                    // :597-644
{
    std::stack<int> curNodeStack;
    std::stack<int32_t> predArcStack;
    std::vector<int32_t> &threads = t->threads;
    threads.clear();
    curNodeStack.push(t->root);
    while (!curNodeStack.empty()) {
        const int curNode = curNodeStack.top(); curNodeStack.pop();
        int predNode = -1;
        if (!threads.empty()) {
            const int32_t pa = predArcStack.top(); predArcStack.pop();
            predNode = (curNode == tail[pa]) ? head[pa] : tail[pa];
            depth[curNode] = depth[predNode] + 1;
            predArc[curNode] = pa;
            thread[threads.back()] = curNode;
        } else {
            depth[curNode] = 0;
            predArc[curNode] = -1;
        }
        pred[curNode] = predNode;
        threadPos[curNode] = (int32_t) threads.size();
        threads.push_back(curNode);
        const std::vector<int32_t> &inc = incidentArcs[curNode];
        for (int k = (int) inc.size() - 1; k >= 0; --k) {
            const int32_t av = inc[(size_t) k];
            const int u = tail[av], v = head[av];
            if (u == predNode || v == predNode) continue;
            curNodeStack.push(curNode == u ? v : u);
            predArcStack.push(av);
        }
    }
    thread[threads.back()] = t->root;
}

// ------------------------------------------------------------------------------------------
// basis sets: gnePivot.cpp:487-600; setNBarc is mirrored on the device (position order)
// ------------------------------------------------------------------------------------------
int EqfSolver::insertIntoB(int32_t v)
{
    if (isArc(v)) {
        setBarc.push_back(v); setLoc[v] = B_var; setInd[v] = (int32_t) setBarc.size() - 1;
        GNE_TRY(insertIntoTrees(v));
    } else {
        setBeqf.push_back(v); setLoc[v] = B_var; setInd[v] = (int32_t) setBeqf.size() - 1;
    }
    return GNE_OK;
}

int EqfSolver::removeFromB(int32_t v)
{
    std::vector<int32_t> &set = isArc(v) ? setBarc : setBeqf;
    const int32_t prevInd = setInd[v];
    const int32_t mv = set.back();
    set[(size_t) prevInd] = mv; setInd[mv] = prevInd; set.pop_back();
    if (isArc(v)) GNE_TRY(removeFromTrees(v));
    setLoc[v] = NIL; setInd[v] = -1;
    return GNE_OK;
}

int EqfSolver::insertIntoNB(int32_t v, int8_t loc)
{
    if (isArc(v)) {
        setNBarc.push_back(v); setLoc[v] = loc; setInd[v] = (int32_t) setNBarc.size() - 1;
        if (costsOnDevice) {
            GNE_TRY(gne_nb_write(dev, (int64_t) setNBarc.size() - 1, tail[v], head[v], cost[v], mult[v], loc));
            GNE_TRY(gne_nb_set_count(dev, (int64_t) setNBarc.size()));
        }
    } else {
        setNBeqf.push_back(v); setLoc[v] = loc; setInd[v] = (int32_t) setNBeqf.size() - 1;
    }
    return GNE_OK;
}

int EqfSolver::removeFromNB(int32_t v)
{
    std::vector<int32_t> &set = isArc(v) ? setNBarc : setNBeqf;
    const int32_t prevInd = setInd[v];
    const int32_t mv = set.back();
    set[(size_t) prevInd] = mv; setInd[mv] = prevInd; set.pop_back();
    setLoc[v] = NIL; setInd[v] = -1;
    if (isArc(v) && costsOnDevice) {
        if (mv != v) GNE_TRY(gne_nb_write(dev, prevInd, tail[mv], head[mv], cost[mv], mult[mv], setLoc[mv]));
        GNE_TRY(gne_nb_set_count(dev, (int64_t) setNBarc.size()));
    }
    return GNE_OK;
}

// ------------------------------------------------------------------------------------------
// the s x s systems: GSL's gsl_linalg_LU_decomp / gsl_linalg_LU_solve (linalg/lu.c of GSL 1.x - 2.5), row-major
// ------------------------------------------------------------------------------------------
void EqfSolver::lu_decomp(std::vector<double> &a, int s, std::vector<int> &perm)
{
    perm.resize((size_t) s);
    for (int i = 0; i < s; ++i) perm[(size_t) i] = i;
    for (int j = 0; j + 1 < s; ++j) {
        double big = fabs(a[(size_t) j * s + j]);       // pivot: the FIRST row of largest magnitude in column j
        int piv = j;
        for (int i = j + 1; i < s; ++i) {
            const double aij = fabs(a[(size_t) i * s + j]);
            if (aij > big) { big = aij; piv = i; }
        }
        if (piv != j) {
            for (int k = 0; k < s; ++k) std::swap(a[(size_t) j * s + k], a[(size_t) piv * s + k]);
            std::swap(perm[(size_t) j], perm[(size_t) piv]);
        }
        const double ajj = a[(size_t) j * s + j];
        if (ajj != 0.0) {
            for (int i = j + 1; i < s; ++i) {
                const double aij = a[(size_t) i * s + j] / ajj;
                a[(size_t) i * s + j] = aij;
                for (int k = j + 1; k < s; ++k) a[(size_t) i * s + k] = a[(size_t) i * s + k] - aij * a[(size_t) j * s + k];
            }
        }
    }
}

void EqfSolver::lu_solve(const std::vector<double> &lu, int s, const std::vector<int> &perm, const std::vector<double> &b,
                         std::vector<double> &x)
{
    x.resize((size_t) s);
    for (int i = 0; i < s; ++i) x[(size_t) i] = b[(size_t) perm[(size_t) i]];       // gsl_permute_vector
    for (int i = 1; i < s; ++i) {                       // dtrsv Lower / NoTrans / Unit (gslcblas source_trsv_r.h)
        double tmp = x[(size_t) i];
        for (int j = 0; j < i; ++j) tmp -= lu[(size_t) i * s + j] * x[(size_t) j];
        x[(size_t) i] = tmp;
    }
    if (s > 0) {                                        // dtrsv Upper / NoTrans / NonUnit
        x[(size_t) s - 1] = x[(size_t) s - 1] / lu[(size_t) (s - 1) * s + (s - 1)];
        for (int i = s - 2; i >= 0; --i) {
            double tmp = x[(size_t) i];
            for (int j = i + 1; j < s; ++j) tmp -= lu[(size_t) i * s + j] * x[(size_t) j];
            x[(size_t) i] = tmp / lu[(size_t) i * s + i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// gnePotential.cpp
// ------------------------------------------------------------------------------------------
int EqfSolver::computePotentials()                                 // :28-61
{
    // The reference recomputes every tree every iteration.  A Type I tree's potentials depend only on its own arcs, rooting,
    // extra arc and costs: untouched since the last call (potDirty) they would come out bit for bit the same, so the tree is
    // skipped.  Type II trees are coupled through the s x s system and are always recomputed.
    upNodes.clear();
    for (size_t k = 0; k < setBtI.size(); ++k) {
        ETree *t = setBtI[k];
        if (!t->potDirty) continue;
        computePotentialsInTermsOfRoot(t);
        finalizePotentialsTypeI(t);
        t->potDirty = false;
        upNodes.insert(upNodes.end(), t->nodesInTree.begin(), t->nodesInTree.end());
    }
    for (size_t k = 0; k < setBtII.size(); ++k) {
        computePotentialsInTermsOfRoot(setBtII[k]);
        upNodes.insert(upNodes.end(), setBtII[k]->nodesInTree.begin(), setBtII[k]->nodesInTree.end());
    }
    if (!setBeqf.empty()) computePotentialsWithEqFlowSets();
    if (hostOnly) return GNE_OK;
    // the device prices against these; only the rewritten ones travel (all of them when that is most of the vector)
    if (!potsOnDevice || upNodes.size() * 2 > (size_t) numNodes) {
        GNE_TRY(gne_pot_set_all(dev, potential.data()));
        potsOnDevice = true;
    } else if (!upNodes.empty()) {
        upPi.resize(upNodes.size());
        for (size_t q = 0; q < upNodes.size(); ++q) upPi[q] = potential[(size_t) upNodes[q]];
        GNE_TRY(gne_pot_set(dev, (int64_t) upNodes.size(), upNodes.data(), upPi.data()));
    }
    return GNE_OK;
}

void EqfSolver::computePotentialsInTermsOfRoot(ETree *t)           // :64-110
{
    const int thetaID = (t->type == T_I) ? -1 : t->id;
    const std::vector<int32_t> &th = t->threads;
    const int h = th[0];
    potA0[h] = 0.0; potA1[h] = 1.0; potTheta[h] = thetaID;
    for (size_t q = 1; q < th.size(); ++q) {
        const int j = th[q];
        const int i = pred[j];
        const int32_t pa = predArc[j];
        const double c = cost[pa], mu = mult[pa];
        if (tail[pa] == i) {
            potA0[j] = (potA0[i] - c) / mu;
            potA1[j] = potA1[i] / mu;
        } else {
            potA0[j] = (mu * potA0[i]) + c;
            potA1[j] = mu * potA1[i];
        }
        potTheta[j] = thetaID;
    }
}

void EqfSolver::finalizePotentialsTypeI(ETree *t)                  // :113-138
{
    const int32_t ex = t->extraArc;
    const int alpha = tail[ex], beta = head[ex];
    const double cAB = cost[ex], muAB = mult[ex];
    const double thetaValue = (cAB - potA0[alpha] + muAB * potA0[beta]) / (potA1[alpha] - muAB * potA1[beta]);
    for (size_t k = 0; k < t->nodesInTree.size(); ++k) {
        const int i = t->nodesInTree[k];
        const double was = potential[i];
        potential[i] = potA0[i] + potA1[i] * thetaValue;
        nonFinitePots += (int) !std::isfinite(potential[i]) - (int) !std::isfinite(was);
        potA0[i] = 0.0; potA1[i] = 0.0; potTheta[i] = -2;
    }
}

void EqfSolver::computePotentialsWithEqFlowSets()                  // :142-214
{
    const int s = (int) setBeqf.size();
    std::vector<double> mat((size_t) s * s, 0.0), rhs((size_t) s, 0.0), x, rowVals((size_t) s);
    std::vector<int> perm;
    // The reference's row loop is dense over the nodes (:159-168).  A node outside the set contributes dri = 0: `rhs -= 0 * pi`
    // and `row += 0 * a1` change nothing (rhs and row never hold -0.0: they start at +0.0 or a cost, x - x rounds to +0.0) -
    // unless the potential term is not finite (0 * inf = NaN).  So the sparse rows are walked when every term is finite, the
    // dense loop otherwise (the reference does reach NaN states; the fixtures keep such runs).
    bool finite = true;
    for (int i = 0; i < numNodes && finite; ++i)
        finite = (memberOfTreeType[i] == T_I) ? std::isfinite(potential[i]) : (std::isfinite(potA0[i]) && std::isfinite(potA1[i]));
    for (int b = 0; b < s; ++b) {
        const int32_t v = setBeqf[(size_t) b];
        const int r = (int) (v - A);
        double rhsVal = cost[v];
        std::fill(rowVals.begin(), rowVals.end(), 0.0);
        auto term = [&](int i, double dri) {
            if (memberOfTreeType[i] == T_I) {
                rhsVal -= dri * potential[i];
            } else {
                rhsVal -= dri * potA0[i];
                rowVals[(size_t) potTheta[i]] += dri * potA1[i];
            }
        };
        if (finite) {
            for (int64_t q = csrStart[(size_t) r]; q < csrStart[(size_t) r + 1]; ++q) term(csrNode[(size_t) q], csrVal[(size_t) q]);
        } else {
            const double *dr = &nodeVals[(size_t) r * numNodes];
            for (int i = 0; i < numNodes; ++i) term(i, dr[i]);
        }
        rhs[(size_t) b] = rhsVal;
        for (int k = 0; k < s; ++k) mat[(size_t) b * s + k] = rowVals[(size_t) k];
    }
    lu_decomp(mat, s, perm);
    lu_solve(mat, s, perm, rhs, x);
    for (size_t t = 0; t < setBtII.size(); ++t) {       // (:203-210 walks all nodes and skips the Type I ones)
        const std::vector<int32_t> &nodes = setBtII[t]->nodesInTree;
        for (size_t k = 0; k < nodes.size(); ++k) {
            const int i = nodes[k];
            const double was = potential[i];
            potential[i] = potA0[i] + potA1[i] * x[(size_t) potTheta[i]];
            nonFinitePots += (int) !std::isfinite(potential[i]) - (int) !std::isfinite(was);
            potA0[i] = 0.0; potA1[i] = 0.0; potTheta[i] = -2;
        }
    }
}

// ------------------------------------------------------------------------------------------
// genneteq.h:334-403
// ------------------------------------------------------------------------------------------
double EqfSolver::computeDeltaVar(int32_t v) const                 // :356-403 (initialType is SELFLOOPS: no artificial-cost branch)
{
    const double xVar = flow[v], yVar = deltaFlow[v];
    if (yVar > TOL) {
        if (fabs(UB[v] - xVar) < TOL) return 0.0;
        return (UB[v] - xVar) / yVar;
    } else if (yVar < (-1.0 * TOL)) {
        if (fabs(xVar) < TOL) return 0.0;
        return xVar / (-1.0 * yVar);
    }
    return DBL_MAX;
}

void EqfSolver::checkVarForBlocking(int32_t v)                     // :334-354
{
    const double deltaVar = computeDeltaVar(v);
    if (delta > deltaVar + TOL) blockingVars.clear();
    if (delta >= deltaVar - TOL) { delta = deltaVar; blockingVars.push_back(v); }
}

// ------------------------------------------------------------------------------------------
// gneFlow.cpp
// ------------------------------------------------------------------------------------------
void EqfSolver::computeFlows(int updateOption)                     // :30-56
{
    flowCost = 0.0;
    if (!setBeqf.empty()) {
        setSupplyWithThetas();
        computeBoundedFlows(true);
        computeFlowsWithEqFlowSets();
    } else {
        setSupplyWithEqs();
        computeBoundedFlows(false);
    }
    for (size_t k = 0; k < setBtI.size(); ++k) computeFlowsTypeI(setBtI[k]);
    setFlowValues(updateOption);
}

void EqfSolver::setFlowValues(int updateOption)                    // :59-150 (the error report is the reference's printf only)
{
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<int32_t> &set = pass == 0 ? setBarc : setBeqf;
        for (size_t k = 0; k < set.size(); ++k) {
            const int32_t v = set[k];
            if (updateOption == 0 || updateOption > 1) flow[v] = deltaFlow[v];
            if (updateOption == 0) { deltaFlow[v] = 0.0; flowCost += flow[v] * cost[v]; }
            else { flowCost += flow[v] * cost[v]; deltaFlow[v] = 0.0; }
        }
    }
}

// (computeFlowsPivotingTI, gneFlow.cpp:155-185, lives in pivotTI: the supply injection and the two trees' sweeps)

void EqfSolver::computeFlowsPivotingTII(int32_t var)               // :188-209
{
    const double t0 = wallNow();
    if (!setBeqf.empty()) {
        setSupplyWithThetas(var);
        computeFlowsWithEqFlowSets();
    } else {
        setSupplyWithEqs(var);
    }
    const double t1 = wallNow();
    for (size_t k = 0; k < setBtI.size(); ++k) computeFlowsTypeI(setBtI[k]);
    secsTIIsets += t1 - t0; secsTIItrees += wallNow() - t1;
}

void EqfSolver::computeFlowsWithEqFlowSets()                       // :212-256
{
    const int s = (int) setBeqf.size();
    for (int b = 0; b < s; ++b) {                       // (:216-221 is dense over the nodes: x - 0.0 is x, so only the entries)
        const int r = (int) (setBeqf[(size_t) b] - A);
        for (int64_t q = csrStart[(size_t) r]; q < csrStart[(size_t) r + 1]; ++q)
            swt[(size_t) csrNode[(size_t) q] * S + b + 1] -= csrVal[(size_t) q];
    }
    std::vector<double> mat((size_t) s * s, 0.0), rhs((size_t) s, 0.0), x;
    std::vector<int> perm;
    for (size_t t = 0; t < setBtII.size(); ++t) {
        const int h = computeFlowsTypeII(setBtII[t]);
        const double *rootSupply = &swt[(size_t) h * S];
        rhs[t] = (-1.0) * rootSupply[0];
        for (int b = 0; b < s; ++b) mat[t * s + b] = rootSupply[b + 1];
    }
    lu_decomp(mat, s, perm);
    lu_solve(mat, s, perm, rhs, x);
    computeNumericFlowsAndSupply(x);
}

void EqfSolver::computeNumericFlowsAndSupply(const std::vector<double> &x)     // :259-306
{
    const int s = (int) setBeqf.size();
    for (size_t t = 0; t < setBtII.size(); ++t) {
        const std::vector<int32_t> &arcs = setBtII[t]->arcsInTree;
        for (size_t k = 0; k < arcs.size(); ++k) {
            const int32_t av = arcs[k];
            const int j = (predArc[tail[av]] == av) ? tail[av] : head[av];         // the arc's flow vector sits in its lower end's row
            const double *avFlow = &fwt[(size_t) j * S];
            double actualFlow = avFlow[0];
            for (int b = 0; b < s; ++b) actualFlow += avFlow[b + 1] * x[(size_t) b];
            deltaFlow[av] = actualFlow;
            checkVarForBlocking(av);
        }
    }
    for (int b = 0; b < s; ++b) {
        const int32_t v = setBeqf[(size_t) b];
        deltaFlow[v] = x[(size_t) b];
        checkVarForBlocking(v);
    }
    for (int i = 0; i < numNodes; ++i) {
        const double *sup = &swt[(size_t) i * S];
        double actualSupply = sup[0];
        for (int b = 0; b < s; ++b) actualSupply += sup[b + 1] * x[(size_t) b];
        supA0[i] = actualSupply;
        if (memberOfTreeType[i] == T_II) supA0[i] = 0.0;
        supA1[i] = 0.0;
    }
}

void EqfSolver::computeBoundedFlows(bool withEqFlowSets)           // :312-368
{
    for (size_t k = 0; k < setNBarc.size(); ++k) {
        const int32_t a = setNBarc[k];
        if (setLoc[a] == L_var) { flow[a] = 0.0; continue; }
        flow[a] = UB[a];
        flowCost += flow[a] * cost[a];
        if (withEqFlowSets) {
            swt[(size_t) tail[a] * S] -= flow[a];
            swt[(size_t) head[a] * S] += flow[a] * mult[a];
        } else {
            supA0[tail[a]] -= flow[a];
            supA0[head[a]] += flow[a] * mult[a];
        }
    }
    for (size_t k = 0; k < setNBeqf.size(); ++k) {
        const int32_t v = setNBeqf[k];
        if (setLoc[v] == L_var) { flow[v] = 0.0; continue; }
        flow[v] = UB[v];
        flowCost += flow[v] * cost[v];
        const double *dr = &nodeVals[(size_t) (v - A) * numNodes];
        if (withEqFlowSets) for (int i = 0; i < numNodes; ++i) swt[(size_t) i * S] -= flow[v] * dr[i];
        else for (int i = 0; i < numNodes; ++i) supA0[i] -= flow[v] * dr[i];
    }
}

void EqfSolver::setSupplyWithThetas()                              // :374-383
{
    S = (int) setBeqf.size() + 1;
    swt.assign((size_t) numNodes * S, 0.0);
    fwt.resize((size_t) numNodes * S);                  // (every row read later is written first, by computeFlowsTypeII)
    for (int i = 0; i < numNodes; ++i) swt[(size_t) i * S] = supply[i];
}

void EqfSolver::setSupplyWithEqs()                                 // :386-394
{
    for (int i = 0; i < numNodes; ++i) { supA0[i] = supply[i]; supA1[i] = 0.0; }
}

void EqfSolver::setSupplyWithThetas(int32_t var)                   // :397-438
{
    S = (int) setBeqf.size() + 1;
    swt.assign((size_t) numNodes * S, 0.0);
    fwt.resize((size_t) numNodes * S);
    if (isArc(var)) {
        const int u = tail[var], v = head[var];
        if (setLoc[var] == L_var) {
            if (u == v) swt[(size_t) u * S] = -1.0 * (1.0 - mult[var]);
            else { swt[(size_t) u * S] = -1.0; swt[(size_t) v * S] = mult[var]; }
        } else {
            if (u == v) swt[(size_t) u * S] = 1.0 - mult[var];
            else { swt[(size_t) u * S] = 1.0; swt[(size_t) v * S] = -1.0 * mult[var]; }
        }
    } else {
        const double *dr = &nodeVals[(size_t) (var - A) * numNodes];
        if (setLoc[var] == L_var) for (int i = 0; i < numNodes; ++i) swt[(size_t) i * S] = -1.0 * dr[i];
        else for (int i = 0; i < numNodes; ++i) swt[(size_t) i * S] = dr[i];
    }
}

void EqfSolver::setSupplyWithEqs(int32_t var)                      // :441-474
{
    if (isArc(var)) {
        const int u = tail[var], v = head[var];
        if (setLoc[var] == L_var) {
            if (u == v) supA0[u] = -1.0 * (1.0 - mult[var]);
            else { supA0[u] = -1.0; supA0[v] = mult[var]; }
        } else {
            if (u == v) supA0[u] = 1.0 - mult[var];
            else { supA0[u] = 1.0; supA0[v] = -1.0 * mult[var]; }
        }
    } else {
        const double *dr = &nodeVals[(size_t) (var - A) * numNodes];
        if (setLoc[var] == L_var) for (int i = 0; i < numNodes; ++i) supA0[i] = -1.0 * dr[i];
        else for (int i = 0; i < numNodes; ++i) supA0[i] = dr[i];
    }
}

void EqfSolver::computeFlowsTypeI(ETree *t)                        // :480-560
{
    const int32_t ex = t->extraArc;
    const int alpha = tail[ex], beta = head[ex];
    const double muAB = mult[ex];
    flowA0[ex] = 0.0; flowA1[ex] = 1.0;
    supA1[alpha] -= 1.0;
    supA1[beta] += muAB;
    const std::vector<int32_t> &th = t->threads;
    const int h = t->root;
    for (size_t q = th.size(); q-- > 1;) {              // reverse thread order, the root (th[0]) excluded
        const int j = th[q];
        const int32_t av = predArc[j];
        if (tail[av] == j) {
            const int i = head[av];
            const double muJI = mult[av];
            flowA0[av] = supA0[j]; flowA1[av] = supA1[j];
            supA0[i] += supA0[j] * muJI;
            supA1[i] += supA1[j] * muJI;
        } else {
            const int i = tail[av];
            const double muIJ = mult[av];
            flowA0[av] = supA0[j] / (-1.0 * muIJ);
            flowA1[av] = supA1[j] / (-1.0 * muIJ);
            supA0[i] += supA0[j] / muIJ;
            supA1[i] += supA1[j] / muIJ;
        }
        supA0[j] = 0.0; supA1[j] = 0.0;
    }
    const double theta = (-1.0 * supA0[h]) / supA1[h];
    supA0[h] = 0.0; supA1[h] = 0.0;
    for (size_t k = 0; k < t->arcsInTree.size(); ++k) {
        const int32_t av = t->arcsInTree[k];
        deltaFlow[av] = flowA0[av] + flowA1[av] * theta;
        flowA0[av] = 0.0; flowA1[av] = 0.0;
        checkVarForBlocking(av);
    }
    deltaFlow[ex] = flowA0[ex] + flowA1[ex] * theta;
    flowA0[ex] = 0.0; flowA1[ex] = 0.0;
    checkVarForBlocking(ex);
}

// computeFlowsTypeI for a PIVOT between Type I trees, where every supply is zero except at the entering arc's ends (`sources`)
// and - through theta - at the extra arc's ends: e(j) and x(predArc(j)) are then exact zeros off the root paths of those <= 4
// nodes, a zero flow change neither blocks (computeDeltaVar: max) nor moves a flow, so only the path nodes are swept - in reverse
// thread order, which is the order the reference adds the children's contributions in - and only the path arcs are ratio-tested,
// in arcsInTree order.  Same values as the full sweep; only the sign of exact zeros can differ.  Returns false (nothing touched)
// when the tree's indices do not allow it; a theta that is not finite (0 * inf would make every arc NaN) falls back to the
// reference's full final loop.
bool EqfSolver::computeFlowsTypeISparse(ETree *t, const int32_t *sources, int numSources, std::vector<int32_t> &arcsOut)
{
    arcsOut.clear();
    if (!sparseFlows || !t->upToDate || t->threads.empty() || t->threads[0] != t->root) return false;
    const int32_t ex = t->extraArc;
    const int alpha = tail[ex], beta = head[ex];
    const double muAB = mult[ex];
    const int h = t->root;
    // the active nodes: the union of the root paths
    if (++stampNow == INT32_MAX) { std::fill(nodeStamp.begin(), nodeStamp.end(), 0); stampNow = 1; }
    activeNodes.clear();
    auto climb = [&](int j) {
        while (j != -1 && nodeStamp[(size_t) j] != stampNow) { nodeStamp[(size_t) j] = stampNow; activeNodes.push_back(j); j = pred[j]; }
    };
    for (int k = 0; k < numSources; ++k) climb(sources[k]);
    climb(alpha); climb(beta);
    if (nodeStamp[(size_t) h] != stampNow) return false;            // (cannot happen in a consistent tree)
    flowA0[ex] = 0.0; flowA1[ex] = 1.0;
    supA1[alpha] -= 1.0;
    supA1[beta] += muAB;
    std::sort(activeNodes.begin(), activeNodes.end(), [&](int32_t a, int32_t b) { return threadPos[(size_t) a] > threadPos[(size_t) b]; });
    for (size_t q = 0; q < activeNodes.size(); ++q) {
        const int j = activeNodes[q];
        if (j == h) continue;
        const int32_t av = predArc[j];
        arcsOut.push_back(av);
        if (tail[av] == j) {
            const int i = head[av];
            const double muJI = mult[av];
            flowA0[av] = supA0[j]; flowA1[av] = supA1[j];
            supA0[i] += supA0[j] * muJI;
            supA1[i] += supA1[j] * muJI;
        } else {
            const int i = tail[av];
            const double muIJ = mult[av];
            flowA0[av] = supA0[j] / (-1.0 * muIJ);
            flowA1[av] = supA1[j] / (-1.0 * muIJ);
            supA0[i] += supA0[j] / muIJ;
            supA1[i] += supA1[j] / muIJ;
        }
        supA0[j] = 0.0; supA1[j] = 0.0;
    }
    const double theta = (-1.0 * supA0[h]) / supA1[h];
    supA0[h] = 0.0; supA1[h] = 0.0;
    if (!std::isfinite(theta)) {
        // every arc of the tree turns NaN in the reference (0 + 0 * theta): its full final loop, and a full apply afterwards
        arcsOut.clear(); arcsOut.push_back(-1);
        for (size_t k = 0; k < t->arcsInTree.size(); ++k) {
            const int32_t av = t->arcsInTree[k];
            deltaFlow[av] = flowA0[av] + flowA1[av] * theta;
            flowA0[av] = 0.0; flowA1[av] = 0.0;
            checkVarForBlocking(av);
        }
    } else {
        std::sort(arcsOut.begin(), arcsOut.end(), [&](int32_t a, int32_t b) { return posInTree[(size_t) a] < posInTree[(size_t) b]; });
        for (size_t k = 0; k < arcsOut.size(); ++k) {
            const int32_t av = arcsOut[k];
            deltaFlow[av] = flowA0[av] + flowA1[av] * theta;
            flowA0[av] = 0.0; flowA1[av] = 0.0;
            checkVarForBlocking(av);
        }
    }
    deltaFlow[ex] = flowA0[ex] + flowA1[ex] * theta;
    flowA0[ex] = 0.0; flowA1[ex] = 0.0;
    checkVarForBlocking(ex);
    return true;
}

int EqfSolver::computeFlowsTypeII(ETree *t)                        // :563-592
{
    const std::vector<int32_t> &th = t->threads;
    const int h = t->root;
    for (size_t q = th.size(); q-- > 1;) {
        const int j = th[q];
        const int32_t av = predArc[j];
        double *ej = &swt[(size_t) j * S];
        double *fj = &fwt[(size_t) j * S];
        if (tail[av] == j) {
            double *ei = &swt[(size_t) head[av] * S];
            const double muJI = mult[av];
            for (int k = 0; k < S; ++k) {
                const double e_j_k = ej[k];
                fj[k] = e_j_k;
                ei[k] += e_j_k * muJI;
            }
        } else {
            double *ei = &swt[(size_t) tail[av] * S];
            const double muIJ = mult[av];
            for (int k = 0; k < S; ++k) {
                const double e_j_k = ej[k];
                fj[k] = (-1.0 * e_j_k) / muIJ;
                ei[k] += e_j_k / muIJ;
            }
        }
    }
    return h;
}

// ------------------------------------------------------------------------------------------
// gnePivot.cpp
// ------------------------------------------------------------------------------------------
int EqfSolver::pivot()                                             // :28-60
{
    const double t0 = wallNow();
    if (isArc(eVar) && memberOfTreeType[tail[eVar]] == T_I && memberOfTreeType[head[eVar]] == T_I) {
        pivotTI(eVar);
        secsFlowsTI += wallNow() - t0;
    } else {
        pivotTII();
        ++pivotsTII;
        secsFlowsTII += wallNow() - t0;
    }
    if (lVar == -1) return GNE_OK;
    const double t1 = wallNow();
    GNE_TRY(updateBasis());
    secsBasis += wallNow() - t1;
    checkForCycling();
    return status;
}

void EqfSolver::pivotTI(int32_t eav)                               // :63-119
{
    ETree *tU = setBtI[(size_t) memberOfTreeID[tail[eav]]];
    ETree *tV = setBtI[(size_t) memberOfTreeID[head[eav]]];
    delta = DBL_MAX;
    // computeFlowsPivotingTI (gneFlow.cpp:155-185): the supply injection, then the two trees - path-sparse where possible
    if (setLoc[eav] == L_var) {
        if (tail[eav] == head[eav]) supA0[tail[eav]] = -1.0 * (1.0 - mult[eav]);
        else { supA0[tail[eav]] = -1.0; supA0[head[eav]] = mult[eav]; }
    } else {
        if (tail[eav] == head[eav]) supA0[tail[eav]] = (1.0 - mult[eav]);
        else { supA0[tail[eav]] = 1.0; supA0[head[eav]] = -1.0 * mult[eav]; }
    }
    const bool same = (tU->id == tV->id);
    const int32_t ends[2] = { tail[eav], head[eav] };
    bool sparseU, sparseV = true;
    if (same) {
        sparseU = computeFlowsTypeISparse(tU, ends, 2, activeArcsU);
        if (!sparseU) computeFlowsTypeI(tU);
    } else {
        sparseU = computeFlowsTypeISparse(tU, ends, 1, activeArcsU);
        if (!sparseU) computeFlowsTypeI(tU);
        sparseV = computeFlowsTypeISparse(tV, ends + 1, 1, activeArcsV);
        if (!sparseV) computeFlowsTypeI(tV);
    }
    identifyLeavingVar();
    if (lVar == -1) return;
    for (int side = 0; side < 2; ++side) {
        ETree *t = side == 0 ? tU : tV;
        if (side == 1 && same) break;
        const bool sparse = side == 0 ? sparseU : sparseV;
        const std::vector<int32_t> &act = side == 0 ? activeArcsU : activeArcsV;
        // (an arc off the root paths has deltaFlow = +-0: `flow += delta * 0` and `flowCost += cost * (delta * 0)` move nothing)
        const bool full = !sparse || (!act.empty() && act[0] == -1) || !std::isfinite(delta);
        const std::vector<int32_t> &arcs = full ? t->arcsInTree : act;
        for (size_t k = 0; k < arcs.size(); ++k) {
            const int32_t av = arcs[k];
            flow[av] += (delta * deltaFlow[av]);
            flowCost += cost[av] * (delta * deltaFlow[av]);
            deltaFlow[av] = 0.0;
        }
        const int32_t ex = t->extraArc;
        flow[ex] += (delta * deltaFlow[ex]);
        flowCost += cost[ex] * (delta * deltaFlow[ex]);
        deltaFlow[ex] = 0.0;
    }
    if (setLoc[eav] == L_var) { flow[eav] = delta; flowCost += cost[eav] * delta; }
    else { flow[eav] = UB[eav] - delta; flowCost -= cost[eav] * delta; }
}

void EqfSolver::pivotTII()                                         // :122-160
{
    delta = DBL_MAX;
    computeFlowsPivotingTII(eVar);
    identifyLeavingVar();
    if (lVar == -1) return;
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<int32_t> &set = pass == 0 ? setBarc : setBeqf;
        for (size_t k = 0; k < set.size(); ++k) {
            const int32_t v = set[k];
            flow[v] += (delta * deltaFlow[v]);
            flowCost += cost[v] * (delta * deltaFlow[v]);
            deltaFlow[v] = 0.0;
        }
    }
    if (setLoc[eVar] == L_var) { flow[eVar] = delta; flowCost += cost[eVar] * delta; }
    else { flow[eVar] = UB[eVar] - delta; flowCost -= cost[eVar] * delta; }
}

void EqfSolver::identifyLeavingVar()                               // :166-209 (an empty list ends the solve as unbounded instead of indexing it)
{
    if (UB[eVar] <= delta + TOL) {
        delta = UB[eVar];
        lVar = eVar;
    } else {
        if (blockingVars.empty()) { isUnbounded = true; lVar = -1; return; }
        lVar = blockingVars[(size_t) (drand48() * blockingVars.size())];
    }
    if (delta <= TOL) { ++totalDegenPivots; ++consecutiveDegenPivots; }
    else consecutiveDegenPivots = 0;
}

int EqfSolver::updateBasis()                                       // :254-309
{
    if (eVar == lVar) {
        const int8_t was = setLoc[eVar];
        GNE_TRY(removeFromNB(eVar));
        GNE_TRY(insertIntoNB(eVar, was == L_var ? U_var : L_var));
    } else {
        GNE_TRY(removeFromNB(eVar));
        GNE_TRY(removeFromB(lVar));
        GNE_TRY(insertIntoB(eVar));
        if (flow[lVar] > (0.5 * UB[lVar])) GNE_TRY(insertIntoNB(lVar, U_var));
        else GNE_TRY(insertIntoNB(lVar, L_var));
        updateTrees();
    }
    if (setLoc[lVar] == L_var) {
        if (fabs(flow[lVar] - 0.0) > TOL) flow[lVar] = 0.0;
    } else {
        if (fabs(flow[lVar] - UB[lVar]) > TOL) flow[lVar] = UB[lVar];
    }
    return GNE_OK;
}

void EqfSolver::checkForCycling()                                  // :312-356
{
    if (cycleList.size() >= 10) cycleList.pop_back();
    cycleList.push_front(std::make_pair(eVar, lVar));
    cyclingDetected = false;
    std::list<std::pair<int32_t, int32_t> >::iterator it = cycleList.begin();
    for (++it; it != cycleList.end(); ++it) {
        if (cycleList.front().first == it->first && cycleList.front().second == it->second) { cyclingDetected = true; break; }
    }
    if (cyclingDetected) {
        ++totalCyclesDetected;
        ++consecutiveCycles;
        if (consecutiveCycles > 5) { gne_set_error("Stuck in cycling loop.  Retry with different pivots."); status = GNE_ERR_CYCLING; }
    } else {
        consecutiveCycles = 0;
    }
    lastEVarInd = eVar;
}

// ------------------------------------------------------------------------------------------
// gnePivotRules.cpp: the arc loops run on the device, the set loops continue their state on the host
// ------------------------------------------------------------------------------------------
bool EqfSolver::violates(int32_t v, double rc) const               // :406-407 and its copies
{
    return (setLoc[v] == L_var && rc < (-1.0 * TOL)) || (setLoc[v] == U_var && rc > TOL);
}

double EqfSolver::hostRedCost(int32_t v) const                     // genneteq.h:312-330 against the host's copy of the potentials
{
    if (isArc(v)) return cost[v] - potential[tail[v]] + mult[v] * potential[head[v]];
    const double *dr = &nodeVals[(size_t) (v - A) * numNodes];
    double rc = cost[v];
    for (int i = 0; i < numNodes; ++i) rc -= (dr[i] * potential[i]);
    return rc;
}

// computeEqfRedCost of every nonbasic set, setNBeqf order.  Up to 128 nonbasic sets: the rows are resident, the call travels as
// kernel arguments and the answer comes back through mapped memory (launch now, fetch later - the arc pass goes in between);
// more: the rows travel with the call (gne_price_eqf).
int EqfSolver::priceSetsLaunch()
{
    const int q = (int) setNBeqf.size();
    eqfRc.resize((size_t) q);
    setsLaunched = false;
    if (q == 0 || !setsResident || q > 128) return GNE_OK;
    std::vector<int32_t> ids((size_t) q);
    std::vector<double> c((size_t) q);
    for (int k = 0; k < q; ++k) { ids[(size_t) k] = (int32_t) (setNBeqf[(size_t) k] - A); c[(size_t) k] = cost[setNBeqf[(size_t) k]]; }
    GNE_TRY(gne_price_eqf_launch(dev, q, ids.data(), c.data(), nonFinitePots));
    setsLaunched = true;
    return GNE_OK;
}

int EqfSolver::priceSetsFetch()
{
    const int q = (int) setNBeqf.size();
    if (q == 0) return GNE_OK;
    if (setsLaunched) { setsLaunched = false; return gne_price_eqf_fetch(dev, q, eqfRc.data()); }
    std::vector<int64_t> rowStart((size_t) q + 1, 0);
    std::vector<int32_t> node;
    std::vector<double> val, c((size_t) q);
    for (int k = 0; k < q; ++k) {
        const int r = (int) (setNBeqf[(size_t) k] - A);
        node.insert(node.end(), csrNode.begin() + csrStart[(size_t) r], csrNode.begin() + csrStart[(size_t) r + 1]);
        val.insert(val.end(), csrVal.begin() + csrStart[(size_t) r], csrVal.begin() + csrStart[(size_t) r + 1]);
        rowStart[(size_t) k + 1] = (int64_t) node.size();
        c[(size_t) k] = cost[setNBeqf[(size_t) k]];
    }
    return gne_price_eqf(dev, q, rowStart.data(), node.data(), val.data(), c.data(), eqfRc.data());
}

int EqfSolver::priceSets()
{
    GNE_TRY(priceSetsLaunch());
    return priceSetsFetch();
}

int EqfSolver::fetchViolators()                                    // computeReducedCosts(false), gnePotential.cpp:217-246
{
    isOptimal = true;
    int64_t cnt = 0;
    if (listPos.size() < 1024) { listPos.resize(1024); listViol.resize(1024); }
    int rc = gne_price_violators(dev, 0.0, &cnt, listPos.data(), listViol.data(), (int64_t) listPos.size());
    if (rc == GNE_ERR_CAPACITY) {
        listPos.resize((size_t) cnt + 1024); listViol.resize((size_t) cnt + 1024);
        rc = gne_price_violators(dev, 0.0, &cnt, listPos.data(), listViol.data(), (int64_t) listPos.size());
    }
    if (rc) return rc;
    arcsPriced += (double) setNBarc.size();
    offendingVars.resize((size_t) cnt); offendingViol.resize((size_t) cnt);
    for (int64_t i = 0; i < cnt; ++i) { offendingVars[(size_t) i] = setNBarc[(size_t) listPos[(size_t) i]]; offendingViol[(size_t) i] = listViol[(size_t) i]; }
    GNE_TRY(priceSets());
    for (size_t k = 0; k < setNBeqf.size(); ++k) {
        if (violates(setNBeqf[k], eqfRc[k])) { offendingVars.push_back(setNBeqf[k]); offendingViol.push_back(fabs(eqfRc[k])); }
    }
    if (!offendingVars.empty()) isOptimal = false;
    return GNE_OK;
}

// order 0: arcs then sets (getVarWithLargestViolation :133-157), 1: arcs first (:160-180), 2: sets first (:183-203)
int EqfSolver::getVarWithLargestViolation(int order, bool useWeighted, int32_t *out)
{
    isOptimal = true;
    double L = -1.0;
    int64_t arcCount = 0;
    std::vector<int32_t> sets;                          // the set members of offendingVars (always behind the arc members)
    bool arcsPricedNow = false;
    // the set loop of computeEqfsWithLargestViolation (:415-429) continuing checkVarForOffending's state (:431-452)
    auto setPass = [&](bool weighted) {
        for (size_t k = 0; k < setNBeqf.size(); ++k) {
            const int32_t v = setNBeqf[k];
            const double rc = eqfRc[k];
            if (!violates(v, rc)) continue;
            isOptimal = false;
            const double frc = weighted ? fabs(rc) / numOrigArcs[(size_t) (v - A)] : fabs(rc);
            if (L < frc - TOL) { arcCount = 0; sets.clear(); }
            if (L <= frc + TOL) { L = frc; sets.push_back(v); }
        }
    };
    auto arcPass = [&]() -> int {                       // only ever entered with L == -1 and an empty list
        int any = 0;
        GNE_TRY(gne_price_lv_begin(dev, &L, &arcCount, &any));
        arcsPriced += (double) setNBarc.size();
        arcsPricedNow = true;
        if (any) isOptimal = false;
        return GNE_OK;
    };
    // (the sets are priced before the arc pass is started: nothing else touches the handle between the pass and the
    //  gne_price_lv_at that resolves the picked member of a list the device kept)
    if (order == 2) {
        GNE_TRY(priceSets());
        setPass(false);
        lastViolation = L;
        if (sets.empty()) { GNE_TRY(arcPass()); lastViolation = L; }
    } else if (order == 0) {
        GNE_TRY(priceSetsLaunch());                     // queued ahead of the arc pass: the wait for the pass covers both
        GNE_TRY(arcPass());
        GNE_TRY(priceSetsFetch());
        setPass(useWeighted);
        lastViolation = L;
    } else {
        GNE_TRY(arcPass());
        lastViolation = L;
        if (arcCount == 0) { GNE_TRY(priceSets()); setPass(false); lastViolation = L; }
    }
    const size_t total = (size_t) arcCount + sets.size();
    if (total == 0) { *out = -1; return GNE_OK; }
    const size_t idx = (size_t) (drand48() * total);
    if (idx < (size_t) arcCount) {
        if (!arcsPricedNow) { gne_set_error("EqfSolver: arc member without an arc pass"); return GNE_ERR_ARG; }
        int64_t pos = -1;
        GNE_TRY(gne_price_lv_at(dev, (int64_t) idx, &pos));
        *out = setNBarc[(size_t) pos];
    } else {
        *out = sets[idx - (size_t) arcCount];
    }
    return GNE_OK;
}

int EqfSolver::getFirstVarWithViolation(int32_t *out)              // :206-228
{
    isOptimal = true;
    int64_t pos = -1;
    GNE_TRY(gne_price_first(dev, &pos));
    arcsPriced += (double) (pos < 0 ? (int64_t) setNBarc.size() : pos + 1);
    if (pos >= 0) { isOptimal = false; *out = setNBarc[(size_t) pos]; return GNE_OK; }
    GNE_TRY(priceSets());
    for (size_t k = 0; k < setNBeqf.size(); ++k) {
        if (violates(setNBeqf[k], eqfRc[k])) { isOptimal = false; *out = setNBeqf[k]; return GNE_OK; }
    }
    *out = -1;
    return GNE_OK;
}

int EqfSolver::getRandomVarWithViolation(int32_t *out)             // :273-281
{
    GNE_TRY(fetchViolators());
    if (offendingVars.empty()) { *out = -1; return GNE_OK; }
    *out = offendingVars[(size_t) (drand48() * offendingVars.size())];
    return GNE_OK;
}

int EqfSolver::getRandomVarWeightedByViolation(int32_t *out)       // :284-304
{
    GNE_TRY(fetchViolators());
    *out = -1;
    if (!offendingVars.empty()) {
        double sum = 0.0;
        for (size_t i = 0; i < offendingVars.size(); ++i) sum += offendingViol[i];
        const double level = drand48() * sum;
        double run = 0.0;
        for (size_t i = 0; i < offendingVars.size(); ++i) {
            run += offendingViol[i];
            if (level <= run) { *out = offendingVars[i]; break; }
        }
    }
    return GNE_OK;
}

int EqfSolver::getRandomVarWithLargeViolation(bool useWeighted, int32_t *out)  // :307-362
{
    GNE_TRY(fetchViolators());                          // arcs in position order, then the sets
    std::priority_queue<VarWithViol, std::vector<VarWithViol>, violcomp> q;
    for (size_t i = 0; i < offendingVars.size(); ++i) {
        const int32_t v = offendingVars[i];
        VarWithViol vv; vv.var = v;
        vv.violation = (!isArc(v) && useWeighted) ? -1.0 * offendingViol[i] / numOrigArcs[(size_t) (v - A)] : -1.0 * offendingViol[i];
        q.push(vv);
        while ((int) q.size() > numLargestViolVarsToTrack) q.pop();
    }
    std::vector<int32_t> fin;
    while (!q.empty()) { fin.push_back(q.top().var); q.pop(); }
    if (fin.empty()) { *out = -1; return GNE_OK; }
    *out = fin[(size_t) (drand48() * fin.size())];
    return GNE_OK;
}

int EqfSolver::getVarBySimAnnealing(bool arcsFirst, int32_t *out)  // :364-396
{
    if (drand48() < lastViolation / baselineViolation) GNE_TRY(getVarWithLargestViolation(arcsFirst ? 1 : 2, false, out));
    else GNE_TRY(getVarWithLargestViolation(0, true, out));
    if (arcsFirst && lastViolation * 100 < baselineViolation && baselineViolation > 5000) baselineViolation = lastViolation;
    return GNE_OK;
}

int EqfSolver::getNextCandidate(bool useWeighted, int32_t *out)    // :456-535
{
    isOptimal = true;
    if (itersSinceUpdate < 0 || itersSinceUpdate > majorIterationLength) {
        while (!candidateList.empty()) candidateList.pop();        // updateCandidateList
        GNE_TRY(fetchViolators());
        for (size_t i = 0; i < offendingVars.size(); ++i) {
            const int32_t v = offendingVars[i];
            VarWithViol vv; vv.var = v;
            vv.violation = (!isArc(v) && useWeighted) ? offendingViol[i] / numOrigArcs[(size_t) (v - A)] : offendingViol[i];
            candidateList.push(vv);
        }
        itersSinceUpdate = 0;
    } else {
        ++itersSinceUpdate;
    }
    int32_t found = -1;
    while (!candidateList.empty()) {
        const int32_t c = candidateList.top().var;
        candidateList.pop();
        const bool atLB = (setLoc[c] == L_var);
        const double rc = hostRedCost(c);              // one variable against the host's copy of the potentials
        if ((atLB && rc < (-1.0 * TOL)) || (!atLB && rc > TOL)) { isOptimal = false; found = c; break; }
    }
    if (found == -1 && itersSinceUpdate > 0) {
        itersSinceUpdate = -1;
        return getNextCandidate(false, out);            // (the reference's retry drops useWeighted, :494)
    }
    *out = found;
    return GNE_OK;
}

// updateCandidateList2 (:613-677): arcs and sets are scanned from their rolling cursors, interleaved by one drand48 draw per
// step while both sides last, until clMaxSize (= n) violators are listed.  The device supplies which positions violate
// (every arc violator in position order with |rc|, the sets' reduced costs); the host replays the walk - no reduced cost is
// computed here - and jumps over the draw-free stretches.
int EqfSolver::updateCandidateList2()
{
    candidateList2.clear(); candidateViol2.clear();
    const int64_t nA = (int64_t) setNBarc.size(), nE = (int64_t) setNBeqf.size();
    const size_t clMaxSize = (size_t) numNodes;
    const double arcProb = ((double) nA) / ((double) (nA + nE));
    if (nbArcPivotInd >= nA) nbArcPivotInd = 0;
    if (nbEqfPivotInd >= nE) nbEqfPivotInd = 0;
    const int64_t startA = nbArcPivotInd, startE = nbEqfPivotInd;
    bool exA = (nA == 0), exE = (nE == 0);
    // arc violators in position order
    int64_t cnt = 0;
    if (nA > 0) {
        if (listPos.size() < 1024) { listPos.resize(1024); listViol.resize(1024); }
        int rc = gne_price_violators(dev, 0.0, &cnt, listPos.data(), listViol.data(), (int64_t) listPos.size());
        if (rc == GNE_ERR_CAPACITY) {
            listPos.resize((size_t) cnt + 1024); listViol.resize((size_t) cnt + 1024);
            rc = gne_price_violators(dev, 0.0, &cnt, listPos.data(), listViol.data(), (int64_t) listPos.size());
        }
        if (rc) return rc;
    }
    if (nE > 0) GNE_TRY(priceSets());
    // the next violator at or behind the cursor, cyclically; `left` = violators the cursor has not passed yet
    int64_t vNext = cnt ? (int64_t) (std::lower_bound(listPos.begin(), listPos.begin() + cnt, startA) - listPos.begin()) % cnt : 0;
    int64_t left = cnt, scanned = 0;
    auto tryToAddArc = [&]() {                          // :680-689 + the cursor step of its callers
        if (left > 0 && listPos[(size_t) vNext] == nbArcPivotInd) {
            isOptimal = false;
            candidateList2.push_back(setNBarc[(size_t) nbArcPivotInd]); candidateViol2.push_back(listViol[(size_t) vNext]);
            vNext = (vNext + 1) % cnt; --left;
        }
        nbArcPivotInd = (nbArcPivotInd + 1) % nA;
        exA = (nbArcPivotInd == startA);
        ++scanned;
    };
    auto tryToAddEqf = [&]() {                          // :692-701
        const int32_t v = setNBeqf[(size_t) nbEqfPivotInd];
        const double rc = eqfRc[(size_t) nbEqfPivotInd];
        if (violates(v, rc)) { isOptimal = false; candidateList2.push_back(v); candidateViol2.push_back(fabs(rc)); }
        nbEqfPivotInd = (nbEqfPivotInd + 1) % nE;
        exE = (nbEqfPivotInd == startE);
    };
    for (int i = 0; i < 5 && !exA; ++i) tryToAddArc();                 // numArcsToAlwaysCheck
    for (int i = 0; i < 1 && !exE; ++i) tryToAddEqf();                 // numEqfsToAlwaysCheck
    while (candidateList2.size() < clMaxSize && !exA && !exE) {
        if (drand48() < arcProb) tryToAddArc();
        else tryToAddEqf();
    }
    if (exA) while (candidateList2.size() < clMaxSize && !exE) tryToAddEqf();
    if (exE && !exA && candidateList2.size() < clMaxSize) {
        // no draws from here on: the arcs up to the violator that fills the list, or all that are left
        const int64_t want = (int64_t) (clMaxSize - candidateList2.size());
        const int64_t take = std::min(want, left);
        int64_t lastPos = -1;
        for (int64_t k = 0; k < take; ++k) {
            lastPos = listPos[(size_t) vNext];
            isOptimal = false;
            candidateList2.push_back(setNBarc[(size_t) lastPos]); candidateViol2.push_back(listViol[(size_t) vNext]);
            vNext = (vNext + 1) % cnt; --left;
        }
        int64_t end;                                    // the cursor afterwards
        if (take == want) end = (lastPos + 1) % nA;     // the list is full right behind that violator
        else end = startA;                              // ran once round
        scanned += (end - nbArcPivotInd + nA) % nA + ((end == nbArcPivotInd) ? nA : 0);
        nbArcPivotInd = end;
    }
    arcsPriced += (double) scanned;
    return GNE_OK;
}

int EqfSolver::getNextCandidate2(bool useWeighted, int32_t *out)   // :550-610
{
    isOptimal = true;
    if (itersSinceUpdate < 0 || itersSinceUpdate > majorIterationLength) {
        GNE_TRY(updateCandidateList2());
        itersSinceUpdate = 0;
    } else {
        ++itersSinceUpdate;
    }
    int32_t ev = -1;
    int64_t evInd = -1;
    double largest = -1.0;
    size_t k = 0;
    while (k < candidateList2.size()) {
        const int32_t c = candidateList2[k];
        const bool atLB = (setLoc[c] == L_var);
        bool viol;
        double a;
        if (itersSinceUpdate > 0) {
            const double rc = hostRedCost(c);           // one list member against the host's copy of the potentials
            viol = (atLB && rc < (-1.0 * TOL)) || (!atLB && rc > TOL);
            a = fabs(rc);
        } else {                                        // "reduced cost is up to date": the value of the list pass
            viol = true;
            a = candidateViol2[k];
        }
        if (viol) {
            isOptimal = false;
            if (!isArc(c) && useWeighted) a = a / numOrigArcs[(size_t) (c - A)];
            if (largest < a) { largest = a; ev = c; evInd = (int64_t) k; }
            ++k;
        } else {
            candidateList2[k] = candidateList2.back(); candidateList2.pop_back();
            candidateViol2[k] = candidateViol2.back(); candidateViol2.pop_back();
        }
    }
    if (ev == -1 && itersSinceUpdate > 0) {
        itersSinceUpdate = -1;
        return getNextCandidate2(false, out);           // (the reference's retry drops useWeighted, :588)
    }
    if (evInd >= 0) {
        candidateList2[(size_t) evInd] = candidateList2.back(); candidateList2.pop_back();
        candidateViol2[(size_t) evInd] = candidateViol2.back(); candidateViol2.pop_back();
    }
    *out = ev;
    return GNE_OK;
}

int EqfSolver::quickSelect(bool useWeighted, int32_t *out)         // :710-764
{
    isOptimal = true;
    int32_t bestVar = -1;
    double largestViol = 0.0;
    if (nbArcPivotInd >= (int64_t) setNBarc.size()) nbArcPivotInd = 0;
    if (nbEqfPivotInd >= (int64_t) setNBeqf.size()) nbEqfPivotInd = 0;
    const int64_t nbArcPivotStart = nbArcPivotInd, nbEqfPivotStart = nbEqfPivotInd;
    int eqfViolations = 0;
    bool exhaustedEqfs = setNBeqf.empty();
    if (!exhaustedEqfs) GNE_TRY(priceSets());
    while (eqfViolations < 1 && !exhaustedEqfs) {
        const int32_t v = setNBeqf[(size_t) nbEqfPivotInd];
        const double rc = eqfRc[(size_t) nbEqfPivotInd];
        if (violates(v, rc)) {
            isOptimal = false;
            ++eqfViolations;
            const double frc = useWeighted ? fabs(rc) / numOrigArcs[(size_t) (v - A)] : fabs(rc);
            if (frc > largestViol) { largestViol = frc; bestVar = v; }
        }
        nbEqfPivotInd = (nbEqfPivotInd + 1) % (int64_t) setNBeqf.size();
        exhaustedEqfs = (nbEqfPivotInd == nbEqfPivotStart);
    }
    if (!setNBarc.empty()) {
        // the arc loop: first strictly-largest |rc| among the scanned arcs; it replaces the set only when strictly larger
        int64_t bestPos = -1, newCursor = nbArcPivotStart, scanned = 0, violations = 0;
        double bestViol = 0.0;
        GNE_TRY(gne_price_qs(dev, nbArcPivotStart, (int64_t) numNodes, &bestPos, &bestViol, &newCursor, &scanned, &violations));
        arcsPriced += (double) scanned;
        if (violations > 0) isOptimal = false;
        if (bestPos >= 0 && bestViol > largestViol) { largestViol = bestViol; bestVar = setNBarc[(size_t) bestPos]; }
        nbArcPivotInd = newCursor;
    }
    if ((loopIter % majorIterationLength) > 0) { nbArcPivotInd = nbArcPivotStart; nbEqfPivotInd = nbEqfPivotStart; }
    *out = bestVar;
    return GNE_OK;
}

int EqfSolver::selectEnteringVariable(int32_t *out)                // :36-127
{
    if (cyclingDetected) return getRandomVarWithLargeViolation(false, out);
    const int rule = presolve ? phaseIpivot : phaseIIpivot;
    switch (rule) {
      case GNE_LV: return getVarWithLargestViolation(0, false, out);
      case GNE_LVW: return getVarWithLargestViolation(0, true, out);
      case GNE_LVA: return getVarWithLargestViolation(1, false, out);
      case GNE_LVE: return getVarWithLargestViolation(2, false, out);
      case GNE_FV: return getFirstVarWithViolation(out);
      case GNE_RV: return getRandomVarWithViolation(out);
      case GNE_RWV: return getRandomVarWeightedByViolation(out);
      case GNE_RLV: return getRandomVarWithLargeViolation(false, out);
      case GNE_RLVW: return getRandomVarWithLargeViolation(true, out);
      case GNE_CL: return getNextCandidate(false, out);
      case GNE_CLW: return getNextCandidate(true, out);
      case GNE_SAA: return getVarBySimAnnealing(true, out);
      case GNE_SAE: return getVarBySimAnnealing(false, out);
      case GNE_CL2: return getNextCandidate2(false, out);
      case GNE_CLW2: return getNextCandidate2(true, out);
      case GNE_QS: return quickSelect(false, out);
      case GNE_QSW: return quickSelect(true, out);
    }
    gne_set_error("pivot rule %d is not available with equal-flow sets (SD)", rule);
    return GNE_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------
// genneteq.cpp
// ------------------------------------------------------------------------------------------
void EqfSolver::initializeStats(bool usePresolve)                  // :123-167
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
    nbArcPivotInd = 0; nbEqfPivotInd = 0;
    baselineViolation = lastViolation;
}

int EqfSolver::setVariableCosts()                                  // :304-323 + the device's cost column
{
    if (presolve) for (int64_t i = 0; i < V; ++i) cost[(size_t) i] = isArtificial[(size_t) i] ? 1.0 : 0.0;
    else for (int64_t i = 0; i < V; ++i) cost[(size_t) i] = actualCost[(size_t) i];
    for (size_t k = 0; k < setBtI.size(); ++k) setBtI[k]->potDirty = true;       // every potential depends on the costs
    if (hostOnly) return GNE_OK;
    const size_t N = setNBarc.size();
    std::vector<int32_t> t(N), h(N);
    std::vector<double> c(N), mu(N);
    std::vector<int8_t> st(N);
    for (size_t p = 0; p < N; ++p) { const int32_t a = setNBarc[p]; t[p] = tail[a]; h[p] = head[a]; c[p] = cost[a]; mu[p] = mult[a]; st[p] = setLoc[a]; }
    GNE_TRY(gne_nb_reset(dev, (int64_t) N, t.data(), h.data(), c.data(), mu.data(), st.data()));
    costsOnDevice = true;
    return GNE_OK;
}

void EqfSolver::checkFeasibility()                                 // :227-288 (useArtificialCosts is empty with SELFLOOPS)
{
    isInfeasible = false;
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<int32_t> &set = pass == 0 ? setBarc : setBeqf;
        for (size_t k = 0; k < set.size(); ++k) {
            const int32_t v = set[k];
            if ((isArtificial[v] && flow[v] > TOL) || (flow[v] < 0.0 - TOL) || (flow[v] > UB[v] + TOL)) { isInfeasible = true; return; }
        }
    }
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<int32_t> &set = pass == 0 ? setNBarc : setNBeqf;
        for (size_t k = 0; k < set.size(); ++k) if (isArtificial[set[k]] && setLoc[set[k]] == U_var) { isInfeasible = true; return; }
    }
}

int EqfSolver::verifyOptimality()                                  // gnePotential.cpp:249-288 (its printf report aside)
{
    GNE_TRY(computePotentials());
    return fetchViolators();
}

int EqfSolver::replayTrace(int draws1, int draws2, int64_t p1, int64_t total, const int32_t *e, const int32_t *l, int64_t *mismatch)
{
    if (!hostOnly) { gne_set_error("replayTrace needs a host-only solver"); return GNE_ERR_ARG; }
    *mismatch = -1;
    int64_t k = 0;
    for (int phase = 0; phase < 2; ++phase) {
        const bool pre = (phase == 0);
        const int draws = pre ? draws1 : draws2;
        const bool pIwasInfeasible = pre ? false : isInfeasible;
        initializeStats(pre);
        GNE_TRY(setVariableCosts());
        if (!presolve && pIwasInfeasible) { isInfeasible = true; break; }
        computeFlows(0);
        const int64_t end = pre ? p1 : total;
        for (; k < end; ++k) {
            if ((loopIter % 1000) == 0) computeFlows(2);
            { const double q0 = wallNow(); GNE_TRY(computePotentials()); secsPotentials += wallNow() - q0; }
            if (cyclingDetected) (void) drand48();
            else for (int d = 0; d < draws; ++d) (void) drand48();
            eVar = e[k];
            if (eVar < 0 || eVar >= V || (setLoc[eVar] != L_var && setLoc[eVar] != U_var)) { *mismatch = k; return GNE_OK; }
            GNE_TRY(pivot());
            if (lVar != l[k]) { *mismatch = k; return GNE_OK; }
            ++loopIter; ++pivots;
        }
        if (!cyclingDetected && draws == 2) (void) drand48();
        computeFlows(0);
        if (presolve) checkFeasibility();
        GNE_TRY(computePotentials());
    }
    if (getenv("GNE_HOST_PROFILE"))
        fprintf(stderr, "eqf host replay: %lld pivots (%lld through pivotTII): potentials %.3f s, flows of Type I pivots %.3f s, of Type II pivots %.3f s, "
                "basis + trees %.3f s; inside the Type II pivots: sets / Type II trees / systems %.3f s, Type I trees %.3f s\n", (long long) pivots, (long long) pivotsTII, secsPotentials, secsFlowsTI, secsFlowsTII, secsBasis, secsTIIsets, secsTIItrees);
    return GNE_OK;
}

int EqfSolver::solve(bool usePresolve, int64_t maxPivots, int64_t *iters, int *finished)   // :60-117
{
    if (hostOnly) { gne_set_error("solve needs a device: genneteq_b200 has no CPU fallback"); return GNE_ERR_CUDA; }
    *finished = 1; *iters = 0;
    const bool pIwasInfeasible = usePresolve ? false : isInfeasible;
    initializeStats(usePresolve);
    GNE_TRY(setVariableCosts());
    if (!presolve && pIwasInfeasible) { isInfeasible = true; return GNE_OK; }
    const double t0 = wallNow();
    computeFlows(0);
    int64_t budget = maxPivots;
    // GNE_EQF_DEBUG=k: a line per stage from pivot k on (stderr; which call a stuck solve sits in)
    const char *dbgEnv = getenv("GNE_EQF_DEBUG");
    const long dbgFrom = dbgEnv ? atol(dbgEnv) : -1;
    while (!isOptimal && !isUnbounded) {
        if (maxPivots >= 0 && budget == 0) { *finished = 0; break; }
        if ((loopIter % 1000) == 0) computeFlows(2);
        const bool dbg = dbgFrom >= 0 && loopIter >= dbgFrom;
        if (dbg) fprintf(stderr, "[eqf %d] potentials (s = %zu, trees %zu + %zu)\n", loopIter, setBeqf.size(), setBtI.size(), setBtII.size());
        const double a = wallNow();
        GNE_TRY(computePotentials());
        const double b = wallNow();
        secsPotentials += b - a;
        if (dbg) fprintf(stderr, "[eqf %d] pricing (rule %d, cycling %d)\n", loopIter, presolve ? phaseIpivot : phaseIIpivot, (int) cyclingDetected);
        GNE_TRY(selectEnteringVariable(&eVar));
        const double c = wallNow();
        secsPricing += c - b;
        if (dbg) fprintf(stderr, "[eqf %d] entering %d, optimal %d, objective %.17g\n", loopIter, eVar, (int) isOptimal, flowCost);
        if (eVar != -1) {
            GNE_TRY(pivot());
            secsPivot += wallNow() - c;
            if (lVar != -1) {
                if (traceLen < traceCap) { traceE[traceLen] = eVar; traceL[traceLen] = lVar; }
                ++traceLen;
            }
            ++loopIter; ++pivots;
            if (presolve) ++phaseIiters; else ++phaseIIiters;
            if (budget > 0) --budget;
        }
    }
    if (*finished) {
        computeFlows(0);
        if (presolve) checkFeasibility();
        GNE_TRY(verifyOptimality());
    }
    secsTotal += wallNow() - t0;
    *iters = loopIter;
    return GNE_OK;
}

} // namespace gneb200
