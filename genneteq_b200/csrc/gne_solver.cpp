// genneteq_b200 — host mirror of GenNetEq driving the device operators; see gne_solver.h.
// Built with -ffp-contract=off: the reference's arithmetic has no fused multiply-add.
#include "gne_solver.h"
#include "gne_internal.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <time.h>

namespace gneb200 {

static const double TOL = GNE_ROUNDOFF_TOLERANCE;       // main.h:36-43
enum { B_var = 0, L_var = 1, U_var = 2, NIL = 3 };       // variables.h:23

static inline double wallNow()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1.0e-9 * (double) ts.tv_nsec;
}

#define GNE_TRY(call) do { int rc_ = (call); if (rc_ != GNE_OK) return rc_; } while (0)

GenNetEq::GenNetEq() {}

GenNetEq::~GenNetEq()
{
    if (dev) gne_dev_destroy(dev);
}

// glibc drand48 (never seeded by the reference => X0 = 0): X' = (0x5DEECE66D X + 0xB) mod 2^48
double GenNetEq::drand48()
{
    rand48State = (0x5DEECE66DULL * rand48State + 0xBULL) & ((1ULL << 48) - 1ULL);
    return ldexp((double) rand48State, -48);
}

// ------------------------------------------------------------------------------------------
// gneInit.cpp:25-186 (initialType is forced to SELFLOOPS there, :31)
// ------------------------------------------------------------------------------------------
int GenNetEq::initialize(Graph *g, int /*initType: ignored by the reference too*/, int cudaDev)
{
    if (g->getNumEqualFlowSets() != 0) { gne_set_error("equal-flow sets are out of scope"); return GNE_ERR_UNSUPPORTED; }
    cudaDevice = cudaDev;
    numNodes = g->getNumNodes();
    bigM = g->getBigM();
    m = g->getNumArcs();
    M = m + numNodes;
    tail.resize(M); head.resize(M); mult.resize(M); UB.resize(M); actualCost.resize(M); isArtificial.assign(M, 0);
    for (int64_t a = 0; a < m; ++a) {                   // gneInit.cpp:82-107
        tail[a] = g->tail[a]; head[a] = g->head[a]; UB[a] = g->cap[a]; actualCost[a] = g->cost[a]; mult[a] = g->mult[a];
    }
    supply.resize(numNodes);
    for (int i = 0; i < numNodes; ++i) {                // gneInit.cpp:110-127 artificial self-loops
        supply[i] = g->supply(i);
        const int64_t a = m + i;
        tail[a] = i; head[a] = i; UB[a] = DBL_MAX; actualCost[a] = bigM;
        mult[a] = (supply[i] >= 0.0) ? 0.5 : 2.0;
        isArtificial[a] = 1;
    }
    cost.assign(M, 0.0); flow.assign(M, 0.0); deltaFlow.assign(M, 0.0); flowA0.assign(M, 0.0); flowA1.assign(M, 0.0);
    setLoc.assign(M, NIL); setInd.assign(M, -1);
    supA0.assign(numNodes, 0.0); supA1.assign(numNodes, 0.0);
    potA0.assign(numNodes, 0.0); potA1.assign(numNodes, 0.0); potSlot.assign(numNodes, -1);
    thetaOfSlot.clear();
    forest.init(numNodes, M, tail.data(), head.data());
    nodeMark.assign(numNodes, 0);
    setBarc.clear(); setNBarc.clear(); uPos.clear(); lDirty.clear(); flowStale.clear();
    setBarc.reserve(numNodes); setNBarc.reserve(m + 1);
    for (int64_t a = 0; a < M; ++a) {                   // formInitialBFS, gneInit.cpp:174-185
        if (isArtificial[a]) GNE_TRY(insertIntoB((int32_t) a));
        else GNE_TRY(insertIntoNB((int32_t) a, L_var));
    }
    forest.updateTrees();
    phaseIiters = 0; phaseIIiters = 0;
    return GNE_OK;
}

void GenNetEq::shardRange(int64_t m, int rank, int nranks, int64_t *lo, int64_t *hi)
{
    const int64_t cap = m + 16;
    int64_t per = (m + nranks - 1) / nranks;
    per = ((per + 2047) / 2048) * 2048;
    if (per == 0) per = 2048;
    *lo = (int64_t) rank * per;
    if (*lo >= cap) { *lo = *hi = ((cap + 2047) / 2048) * 2048; return; }      // empty shard, tile-aligned, beyond the list
    *hi = (rank == nranks - 1) ? cap : std::min<int64_t>(*lo + per, cap);
}

int GenNetEq::setComm(const uint8_t id[128], int rank, int nranks, gne_allgather_fn hostGather, void *hostCtx)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) { gne_set_error("setComm: bad rank %d of %d", rank, nranks); return GNE_ERR_ARG; }
    if (dev) { gne_set_error("setComm must precede the first solve"); return GNE_ERR_ARG; }
    commRank = rank; commRanks = nranks;
    // contiguous position ranges, boundaries on tile multiples (2048); a rank whose range would start at or
    // beyond the end of the list owns an EMPTY shard (lists shorter than 2048 x ranks): ranges never overlap
    const int64_t cap = m + 16;
    int64_t lo = 0, hi = 0;
    shardRange(m, rank, nranks, &lo, &hi);
    GNE_TRY(gne_dev_create(&dev, cudaDevice, numNodes + 16, cap, lo, hi));
    if (nranks > 1) {
        if (hostGather) GNE_TRY(gne_comm_init_host(dev, rank, nranks, hostGather, hostCtx));
        else GNE_TRY(gne_comm_init(dev, id, rank, nranks));
    }
    return GNE_OK;
}

// ------------------------------------------------------------------------------------------
// basis sets: gnePivot.cpp:487-600, mirrored on the device
// ------------------------------------------------------------------------------------------
int GenNetEq::writeNbRecord(int64_t pos, int32_t v)
{
    if (!dev || !costsOnDevice) return GNE_OK;
    return gne_nb_write(dev, pos, tail[v], head[v], cost[v], mult[v], setLoc[v]);
}

int GenNetEq::insertIntoB(int32_t v)
{
    setBarc.push_back(v); setLoc[v] = B_var; setInd[v] = (int32_t) setBarc.size() - 1;
    int rc = forest.insertArc(v);
    if (rc) { gne_set_error("basis forest: insertion failed (%d): %s", rc, rc == -2 ? "adds cycle to type I tree" : "merges two type I trees"); return GNE_ERR_TREE; }
    return GNE_OK;
}

int GenNetEq::removeFromB(int32_t v)
{
    const int32_t prevInd = setInd[v];
    const int32_t mv = setBarc.back();
    setBarc[prevInd] = mv; setInd[mv] = prevInd; setBarc.pop_back();
    int rc = forest.removeArc(v);
    if (rc) { gne_set_error("basis forest: removal failed (%d)", rc); return GNE_ERR_TREE; }
    setLoc[v] = NIL; setInd[v] = -1;
    return GNE_OK;
}

int GenNetEq::insertIntoNB(int32_t v, int8_t loc)
{
    setNBarc.push_back(v); setLoc[v] = loc; setInd[v] = (int32_t) setNBarc.size() - 1;
    if (loc == U_var) { uPos.insert(setInd[v]); markFlowStale(forest.treeOf(tail[v])); markFlowStale(forest.treeOf(head[v])); }
    if (dev && costsOnDevice) {
        GNE_TRY(writeNbRecord((int64_t) setNBarc.size() - 1, v));
        GNE_TRY(gne_nb_set_count(dev, (int64_t) setNBarc.size()));
    }
    return GNE_OK;
}

int GenNetEq::removeFromNB(int32_t v)
{
    const int32_t prevInd = setInd[v];
    const int32_t mv = setNBarc.back();
    if (setLoc[v] == U_var) { uPos.erase(prevInd); markFlowStale(forest.treeOf(tail[v])); markFlowStale(forest.treeOf(head[v])); }
    if (mv != v && setLoc[mv] == U_var) {
        // a U arc changes position: the order of the supply additions at its end nodes may change with it
        uPos.erase((int32_t) setNBarc.size() - 1); uPos.insert(prevInd);
        markFlowStale(forest.treeOf(tail[mv])); markFlowStale(forest.treeOf(head[mv]));
    }
    setNBarc[prevInd] = mv; setInd[mv] = prevInd; setNBarc.pop_back();
    setLoc[v] = NIL; setInd[v] = -1;
    if (dev && costsOnDevice) {
        if (mv != v) GNE_TRY(writeNbRecord(prevInd, mv));
        GNE_TRY(gne_nb_set_count(dev, (int64_t) setNBarc.size()));
    }
    return GNE_OK;
}

// ------------------------------------------------------------------------------------------
// genneteq.h inline helpers
// ------------------------------------------------------------------------------------------
double GenNetEq::hostPotential(int32_t node) const
{
    const double p = potA1[node] * thetaOfSlot[potSlot[node]];     // gnePotential.cpp:130, unfused
    return potA0[node] + p;
}

double GenNetEq::computeArcRedCost(int32_t a)                     // genneteq.h:312-318
{
    const double t = mult[a] * hostPotential(head[a]);
    return (cost[a] - hostPotential(tail[a])) + t;
}

double GenNetEq::computeDeltaVar(int32_t v) const                 // genneteq.h:364-403, initialType == SELFLOOPS
{
    const double x = flow[v], y = deltaFlow[v];
    if (y > TOL) {
        if (fabs(UB[v] - x) < TOL) return 0.0;
        return (UB[v] - x) / y;
    } else if (y < (-1.0 * TOL)) {
        if (fabs(x) < TOL) return 0.0;
        return x / (-1.0 * y);
    }
    return DBL_MAX;
}

void GenNetEq::checkVarForBlocking(int32_t v)                     // genneteq.h:334-360
{
    const double dv = computeDeltaVar(v);
    if (delta > dv + TOL) blockingVars.clear();
    if (delta >= dv - TOL) { delta = dv; blockingVars.push_back(v); }
}

// ------------------------------------------------------------------------------------------
// gneFlow.cpp
// ------------------------------------------------------------------------------------------
void GenNetEq::computeFlowsTypeI(int32_t slot)                    // gneFlow.cpp:480-560
{
    forest.ensureThreads(slot);
    forest.ensureArcs(slot);                            // (the full sweep walks arcs(tree) in order)
    const Tree &t = forest.tree(slot);
    const int32_t ex = t.extraArc;
    const int32_t alpha = tail[ex], beta = head[ex];
    const double muAB = mult[ex];
    flowA0[ex] = 0.0; flowA1[ex] = 1.0;
    supA1[alpha] -= 1.0;
    supA1[beta] += muAB;
    const int32_t h = t.root;
    for (int64_t k = (int64_t) t.threads.size() - 1; k >= 0; --k) {
        const int32_t j = t.threads[k];
        if (j == h) break;
        const int32_t a = forest.predArcOf(j);
        const double mu = mult[a];
        if (tail[a] == j) {
            const int32_t i = head[a];
            flowA0[a] = supA0[j]; flowA1[a] = supA1[j];
            supA0[i] += supA0[j] * mu;
            supA1[i] += supA1[j] * mu;
        } else {
            const int32_t i = tail[a];
            flowA0[a] = supA0[j] / (-1.0 * mu);
            flowA1[a] = supA1[j] / (-1.0 * mu);
            supA0[i] += supA0[j] / mu;
            supA1[i] += supA1[j] / mu;
        }
        supA0[j] = 0.0; supA1[j] = 0.0;
    }
    const double theta = (-1.0 * supA0[h]) / supA1[h];
    supA0[h] = 0.0; supA1[h] = 0.0;
    for (size_t k = 0; k < t.arcs.size(); ++k) {
        const int32_t a = t.arcs[k];
        const double p = flowA1[a] * theta;
        deltaFlow[a] = flowA0[a] + p;
        flowA0[a] = 0.0; flowA1[a] = 0.0;
        checkVarForBlocking(a);
    }
    const double p = flowA1[ex] * theta;
    deltaFlow[ex] = flowA0[ex] + p;
    flowA0[ex] = 0.0; flowA1[ex] = 0.0;
    checkVarForBlocking(ex);
}

void GenNetEq::computeFlows(int updateOption)                     // gneFlow.cpp:30-56
{
    static const bool selfCheck = getenv("GNE_FLOW_CHECK") != nullptr;
    if (selfCheck && !inFlowCheck) {
        // debugging aid: the skipping version against the reference's full walk (every tree, every U arc) - bit for bit
        inFlowCheck = true;
        const std::vector<double> flowBefore(flow);
        const std::vector<int32_t> lDirtyBefore(lDirty);
        computeFlows(updateOption);
        const std::vector<double> flowSkipping(flow);
        const double costSkipping = flowCost;
        flow = flowBefore; lDirty = lDirtyBefore;
        flowStale.assign(flowStale.size(), 1);
        computeFlows(updateOption);
        inFlowCheck = false;
        if (memcmp(flowSkipping.data(), flow.data(), sizeof(double) * flow.size()) != 0 || memcmp(&costSkipping, &flowCost, sizeof(double)) != 0) {
            fprintf(stderr, "GNE_FLOW_CHECK: computeFlows with skipped trees differs from the full walk (iteration %lld)\n", (long long) loopIter);
            abort();
        }
        return;
    }
    flowCost = 0.0;
    const int ns = forest.numSlots();
    if ((int) flowStale.size() < ns) flowStale.resize((size_t) ns, 1);
    // Trees that no pivot, structure change or U-arc move touched since the last call would be recomputed from the
    // same supplies, in the same order, to the same bits: they are skipped (their nodes' supplies stay zero).
    for (int i = 0; i < numNodes; ++i) {                                               // :387-395
        if (flowStale[forest.treeOf(i)]) { supA0[i] = supply[i]; supA1[i] = 0.0; }
        else { supA0[i] = 0.0; supA1[i] = 0.0; }
    }
    // :343-353 walks setNBarc in position order: L arcs get flow 0, U arcs flow UB and their supply terms.  Only
    // the recorded L arcs can hold anything but +0.0, and the U arcs are kept in position order (uPos).
    for (size_t k = 0; k < lDirty.size(); ++k) if (setLoc[lDirty[k]] == L_var) flow[lDirty[k]] = 0.0;
    lDirty.clear();
    for (std::set<int32_t>::const_iterator it = uPos.begin(); it != uPos.end(); ++it) {
        const int32_t a = setNBarc[*it];
        flow[a] = UB[a];
        flowCost += flow[a] * cost[a];
        if (flowStale[forest.treeOf(tail[a])]) supA0[tail[a]] -= flow[a];
        if (flowStale[forest.treeOf(head[a])]) supA0[head[a]] += flow[a] * mult[a];
    }
    // one pass per (stale) type I tree; the trees are independent, so any order gives the same values
    for (int s = 0; s < ns; ++s) if (forest.tree(s).type == TREE_I && flowStale[s]) computeFlowsTypeI(s);
    for (size_t k = 0; k < setBarc.size(); ++k) {                                      // setFlowValues :59-149
        const int32_t a = setBarc[k];
        if (flowStale[forest.treeOf(tail[a])] && (updateOption == 0 || updateOption > 1)) flow[a] = deltaFlow[a];
        flowCost += flow[a] * cost[a];
        deltaFlow[a] = 0.0;
    }
    for (int s = 0; s < ns; ++s) flowStale[s] = 0;
}

void GenNetEq::computeFlowsPivotingTI(int32_t eav, int32_t tU, int32_t tV)           // gneFlow.cpp:155-185
{
    if (setLoc[eav] == L_var) {
        if (tail[eav] == head[eav]) supA0[tail[eav]] = -1.0 * (1.0 - mult[eav]);
        else { supA0[tail[eav]] = -1.0; supA0[head[eav]] = mult[eav]; }
    } else {
        if (tail[eav] == head[eav]) supA0[tail[eav]] = (1.0 - mult[eav]);
        else { supA0[tail[eav]] = 1.0; supA0[head[eav]] = -1.0 * mult[eav]; }
    }
    computeFlowsTypeI(tU);
    if (tU != tV) computeFlowsTypeI(tV);
}

// computeFlowsTypeI for a pivot, restricted to where it is non-zero.  During a pivot every supply is
// zero except at the entering arc's ends and at the extra arc's ends (theta coefficient), so e(j)
// and x(predArc(j)) are +-0 off the root paths of those (at most four) nodes; +-0 neither changes a
// sum nor blocks (genneteq.h:399-402).  Nodes on the paths are processed in reverse thread order
// (the order of the reference's sweep, hence the same order of additions at every junction), the
// arcs with a flow change are ratio-tested in arcsInTree order.  Values are identical to the full
// sweep; only the sign of exact zeros can differ.
void GenNetEq::computeFlowsSparse(int32_t slot, const int32_t *starts, int nStarts, std::vector<int32_t> &activeArcs)
{
    const Tree &t = forest.tree(slot);
    const int32_t ex = t.extraArc;
    const int32_t alpha = tail[ex], beta = head[ex];
    const double muAB = mult[ex];
    flowA0[ex] = 0.0; flowA1[ex] = 1.0;
    supA1[alpha] -= 1.0;
    supA1[beta] += muAB;
    const int32_t h = t.root;
    activeNodes.clear();
    int32_t st[6];
    int ns = 0;
    for (int k = 0; k < nStarts; ++k) st[ns++] = starts[k];
    st[ns++] = alpha; st[ns++] = beta;
    for (int k = 0; k < ns; ++k) {
        int32_t v = st[k];
        while (v >= 0 && !nodeMark[v]) { activeNodes.push_back(v); nodeMark[v] = (int32_t) activeNodes.size(); v = forest.predOf(v); }
    }
    // Reverse thread order restricted to the active nodes = post-order of the tree they form, the
    // children of a junction taken from the last incident entry to the first (the thread visits a
    // node's children in incident order, trees.cpp:626-640) - no thread positions needed.
    const size_t na = activeNodes.size();
    actHead.assign(na, -1); actNext.assign(na, -1);
    for (size_t q = 0; q < na; ++q) {
        const int32_t j = activeNodes[q];
        if (j == h) continue;
        const int32_t pq = nodeMark[forest.predOf(j)] - 1;
        actNext[q] = actHead[pq]; actHead[pq] = (int32_t) q;
    }
    actOrder.clear(); actStack.clear();
    actStack.push_back(nodeMark[h] - 1);
    while (!actStack.empty()) {
        const int32_t top = actStack.back(); actStack.pop_back();
        if (top < 0) { actOrder.push_back(activeNodes[-top - 1]); continue; }
        actStack.push_back(-top - 1);                   // emit after the children
        actKids.clear();
        for (int32_t c = actHead[top]; c >= 0; c = actNext[c]) actKids.push_back(c);
        if (actKids.size() > 1) {
            const int32_t p = activeNodes[top];
            const Forest &fr = forest;
            const std::vector<int32_t> &an = activeNodes;
            std::sort(actKids.begin(), actKids.end(), [&fr, &an, p](int32_t x, int32_t y) {
                return fr.incidentIndex(p, fr.predArcOf(an[x])) < fr.incidentIndex(p, fr.predArcOf(an[y])); });
        }
        // the child with the highest incident index is processed first: push it last
        for (size_t c = 0; c < actKids.size(); ++c) actStack.push_back(actKids[c]);
    }
    // arcsInTree order of the active arcs when the tree's order is implicit (Forest::split): the split DFS restricted to
    // the active nodes - a node's arcs in the snapshot order of its incident list (`ord`), its children entered from the
    // last to the first - then the arcs appended since the last split (tail), by their explicit positions
    const bool implicitOrder = t.implicitOrder;
    if (implicitOrder) {
        orderedArcs.clear(); activeKeys.clear();
        actStack.clear();
        actStack.push_back(nodeMark[h] - 1);
        while (!actStack.empty()) {
            const int32_t top = actStack.back(); actStack.pop_back();
            const int32_t p = activeNodes[top];
            ordKids.clear();
            for (int32_t c = actHead[top]; c >= 0; c = actNext[c]) {
                const int32_t a = forest.predArcOf(activeNodes[c]);
                const Inc *e = forest.entryOf(p, a);
                if (e->pos >= POS_TAIL_BASE || e->pos == POS_TAIL_OTHER) {
                    // an appended subtree: all its arcs sit in the tail with explicit positions
                    actKids.clear(); actKids.push_back(c);
                    while (!actKids.empty()) {
                        const int32_t x = actKids.back(); actKids.pop_back();
                        const int32_t ax = forest.predArcOf(activeNodes[x]);
                        activeKeys.push_back(std::make_pair(forest.posOfArc(ax, tail[ax], head[ax]), ax));
                        for (int32_t y = actHead[x]; y >= 0; y = actNext[y]) actKids.push_back(y);
                    }
                } else {
                    ordKids.push_back(std::make_pair(e->ord, c));
                }
            }
            if (ordKids.size() > 1) std::sort(ordKids.begin(), ordKids.end());
            for (size_t c = 0; c < ordKids.size(); ++c) orderedArcs.push_back(forest.predArcOf(activeNodes[ordKids[c].second]));
            for (size_t c = 0; c < ordKids.size(); ++c) actStack.push_back(ordKids[c].second);     // the last entry is entered first
        }
        std::sort(activeKeys.begin(), activeKeys.end());
        for (size_t k = 0; k < activeKeys.size(); ++k) orderedArcs.push_back(activeKeys[k].second);
    }
    for (size_t q = 0; q < na; ++q) nodeMark[activeNodes[q]] = 0;
    activeNodes.swap(actOrder);
    activeArcs.clear();
    for (size_t q = 0; q < activeNodes.size(); ++q) {
        const int32_t j = activeNodes[q];
        if (j == h) continue;
        const int32_t a = forest.predArcOf(j);
        const double mu = mult[a];
        if (tail[a] == j) {
            const int32_t i = head[a];
            flowA0[a] = supA0[j]; flowA1[a] = supA1[j];
            supA0[i] += supA0[j] * mu;
            supA1[i] += supA1[j] * mu;
        } else {
            const int32_t i = tail[a];
            flowA0[a] = supA0[j] / (-1.0 * mu);
            flowA1[a] = supA1[j] / (-1.0 * mu);
            supA0[i] += supA0[j] / mu;
            supA1[i] += supA1[j] / mu;
        }
        supA0[j] = 0.0; supA1[j] = 0.0;
        activeArcs.push_back(a);
    }
    const double theta = (-1.0 * supA0[h]) / supA1[h];
    supA0[h] = 0.0; supA1[h] = 0.0;
    // ratio test in arcsInTree order: sort the few active arcs by their index in arcs(tree)
    if (implicitOrder) {
        activeArcs = orderedArcs;
    } else {
        activeKeys.clear();
        for (size_t k = 0; k < activeArcs.size(); ++k) {
            const int32_t a = activeArcs[k];
            activeKeys.push_back(std::make_pair(forest.posOfArc(a, tail[a], head[a]), a));
        }
        std::sort(activeKeys.begin(), activeKeys.end());
        for (size_t k = 0; k < activeArcs.size(); ++k) activeArcs[k] = activeKeys[k].second;
    }
    for (size_t k = 0; k < activeArcs.size(); ++k) {
        const int32_t a = activeArcs[k];
        const double p = flowA1[a] * theta;
        deltaFlow[a] = flowA0[a] + p;
        flowA0[a] = 0.0; flowA1[a] = 0.0;
        checkVarForBlocking(a);
    }
    const double p = flowA1[ex] * theta;
    deltaFlow[ex] = flowA0[ex] + p;
    flowA0[ex] = 0.0; flowA1[ex] = 0.0;
    checkVarForBlocking(ex);
}

void GenNetEq::applyFlowSparse(int32_t slot, const std::vector<int32_t> &activeArcs)   // gnePivot.cpp:79-91 on the changed arcs
{
    for (size_t k = 0; k < activeArcs.size(); ++k) {
        const int32_t a = activeArcs[k];
        const double d = delta * deltaFlow[a];
        flow[a] += d;
        flowCost += cost[a] * d;
        deltaFlow[a] = 0.0;
    }
    const int32_t ex = forest.tree(slot).extraArc;
    const double d = delta * deltaFlow[ex];
    flow[ex] += d;
    flowCost += cost[ex] * d;
    deltaFlow[ex] = 0.0;
}

// ------------------------------------------------------------------------------------------
// gnePotential.cpp: the root-relative sweep on the host for the trees a pivot touched, the
// finalize on the device.  Untouched trees keep (a0, a1, theta), hence pi, bit for bit.
// ------------------------------------------------------------------------------------------
// One node of computePotentialsInTermsOfRoot (gnePotential.cpp:89-108): (a0, a1) of j from its
// predecessor; records the node for upload when (a0, a1, slot) differ from what the device holds.
inline bool GenNetEq::relabelNode(int32_t j, int32_t slot)
{
    const int32_t i = forest.predOf(j);
    double a0, a1;
    if (i < 0) {                                        // root: pi_h = theta  (:78-82)
        a0 = 0.0; a1 = 1.0;
    } else {
        const int32_t a = forest.predArcOf(j);
        const double c0 = cost[a], mu = mult[a];
        if (tail[a] == i) {                             // pi(j) = (pi(i) - c_ij) / mu_ij
            a0 = (potA0[i] - c0) / mu;
            a1 = potA1[i] / mu;
        } else {                                        // pi(j) = mu_ji pi(i) + c_ji
            const double p = mu * potA0[i];
            a0 = p + c0;
            a1 = mu * potA1[i];
        }
    }
    if (memcmp(&potA0[j], &a0, 8) != 0 || memcmp(&potA1[j], &a1, 8) != 0 || potSlot[j] != slot) {
        potA0[j] = a0; potA1[j] = a1; potSlot[j] = slot;
        upNode.push_back(j); upA0.push_back(a0); upA1.push_back(a1); upSlot.push_back(slot);
        return true;
    }
    return false;
}

int GenNetEq::computePotentials()
{
    upNode.clear(); upSlot.clear(); upA0.clear(); upA1.clear(); upThetaSlot.clear(); upTheta.clear(); upFinal.clear(); upFinalSlots.clear();
    bool finalizeAll = false;
    if ((int) thetaOfSlot.size() < forest.numSlots()) thetaOfSlot.resize(forest.numSlots(), 0.0);
    if (allPotDirty) {
        // costs changed (phase start): every tree is swept in thread order
        for (size_t c = 0; c < forest.changed.size(); ++c) {
            if (forest.tree(forest.changed[c]).type == TREE_FREE) continue;
            forest.ensureThreads(forest.changed[c]);
            const Tree &t = forest.tree(forest.changed[c]);
            for (size_t q = 0; q < t.threads.size(); ++q) relabelNode(t.threads[q], forest.changed[c]);
        }
        finalizeAll = true;
    } else {
        // only the nodes whose root path or tree changed in the last re-threading, parents first
        for (size_t q = 0; q < forest.dirtyNodes.size(); ++q) {
            const int32_t j = forest.dirtyNodes[q];
            if (relabelNode(j, forest.treeOf(j))) upFinal.push_back(j);
        }
    }
    forest.clearDirty();
    for (size_t c = 0; c < forest.changed.size(); ++c) {
        const int32_t slot = forest.changed[c];
        forest.changedFlag[slot] = 0;
        const Tree &t = forest.tree(slot);
        if (t.type == TREE_FREE) continue;
        if (t.type != TREE_I) { gne_set_error("type II tree at computePotentials (equal-flow sets are out of scope)"); return GNE_ERR_UNSUPPORTED; }
        // theta, gnePotential.cpp:116-124
        const int32_t ex = t.extraArc;
        const int32_t alpha = tail[ex], beta = head[ex];
        const double pb0 = mult[ex] * potA0[beta];
        const double pb1 = mult[ex] * potA1[beta];
        const double theta = ((cost[ex] - potA0[alpha]) + pb0) / (potA1[alpha] - pb1);
        if (allPotDirty || memcmp(&thetaOfSlot[slot], &theta, 8) != 0) {
            thetaOfSlot[slot] = theta;
            upThetaSlot.push_back(slot); upTheta.push_back(theta);
            // every node of the tree gets a new pi (theta multiplies a1 != 0): a short node list rides with the update,
            // a big tree is finalized on the device by slot (no O(|tree|) list is built or sent)
            if (!finalizeAll) {
                if (t.nodes.size() <= 64 && upFinal.size() + t.nodes.size() <= 160) upFinal.insert(upFinal.end(), t.nodes.begin(), t.nodes.end());
                else upFinalSlots.push_back(slot);
            }
        }
    }
    forest.changed.clear();
    allPotDirty = false;
    if (!dev) return GNE_OK;                            // host-only replay
    // one call: queued setNBarc writes + dirty records + thetas + finalize (gnePotential.cpp:127-135)
    // short finalize lists travel with their values: pi = a0 + a1 * theta by the same unfused operations the device would
    // use (hostPotential), so the relabel folded into the next pricing pass is stores only
    upPi.clear();
    if (!finalizeAll && upFinal.size() <= 160) for (size_t i = 0; i < upFinal.size(); ++i) upPi.push_back(hostPotential(upFinal[i]));
    GNE_TRY(gne_pot_update_fused_pi(dev, (int64_t) upNode.size(), upNode.data(), upA0.data(), upA1.data(), upSlot.data(),
                                    (int64_t) upThetaSlot.size(), upThetaSlot.data(), upTheta.data(),
                                    finalizeAll ? -1 : (int64_t) upFinal.size(), upFinal.data(),
                                    (!finalizeAll && upFinal.size() <= 160) ? upPi.data() : nullptr));
    if (!finalizeAll && !upFinalSlots.empty()) GNE_TRY(gne_pot_finalize_slots(dev, (int) upFinalSlots.size(), upFinalSlots.data()));
    return GNE_OK;
}

int GenNetEq::fetchViolators(double threshold)
{
    int64_t cnt = 0;
    if (listPos.size() < 1024) { listPos.resize(1024); listViol.resize(1024); }
    int rc = gne_price_violators(dev, threshold, &cnt, listPos.data(), listViol.data(), (int64_t) listPos.size());
    if (rc == GNE_ERR_CAPACITY) {
        listPos.resize((size_t) cnt + 1024); listViol.resize((size_t) cnt + 1024);
        rc = gne_price_violators(dev, threshold, &cnt, listPos.data(), listViol.data(), (int64_t) listPos.size());
    }
    if (rc) return rc;
    arcsPriced += (double) setNBarc.size();
    float ms = 0; gne_dev_last_kernel_ms(dev, &ms); kernelMs += ms;
    offendingVars.resize((size_t) cnt); offendingViol.resize((size_t) cnt);
    for (int64_t i = 0; i < cnt; ++i) { offendingVars[i] = setNBarc[listPos[i]]; offendingViol[i] = listViol[i]; }
    return GNE_OK;
}

int GenNetEq::computeReducedCosts(bool /*includeBasicVars: only feeds prints in the reference*/)   // gnePotential.cpp:217-258
{
    isOptimal = true;
    GNE_TRY(fetchViolators(0.0));
    if (!offendingVars.empty()) isOptimal = false;
    return GNE_OK;
}

int GenNetEq::verifyOptimality()                                  // gnePotential.cpp:261-287
{
    GNE_TRY(computePotentials());
    return computeReducedCosts(true);
}

// ------------------------------------------------------------------------------------------
// gnePivotRules.cpp
// ------------------------------------------------------------------------------------------
int GenNetEq::getVarWithLargestViolation(int32_t *out)            // :133-157 (LVA :160-180, LVE :183-203 coincide at p = 0)
{
    isOptimal = true;
    offendingVars.clear();
    GNE_TRY(gne_price_lv_launch(dev));
    return completeLargestViolation(out);
}

// second half of getVarWithLargestViolation: the pass has been launched (gne_price_lv_launch)
int GenNetEq::completeLargestViolation(int32_t *out)
{
    double largest = -1.0;
    int64_t cnt = 0;
    int any = 0;
    // the device leaves offendingVars where it is (it can be hundreds of thousands long in Phase I);
    // only its length and the member drand48 picks are needed here
    GNE_TRY(gne_price_lv_complete(dev, &largest, &cnt, &any));
    arcsPriced += (double) setNBarc.size();
    float ms = 0; gne_dev_last_kernel_ms(dev, &ms); kernelMs += ms;
    if (any) isOptimal = false;
    lastViolation = largest;
    if (cnt == 0) { *out = -1; return GNE_OK; }
    const int64_t idx = (int64_t) (size_t) (drand48() * (size_t) cnt);       // selectRandomOffendingVar == true (:152-153)
    int64_t pos = -1;
    GNE_TRY(gne_price_lv_at(dev, idx, &pos));
    *out = setNBarc[pos];
    return GNE_OK;
}

int GenNetEq::getFirstVarWithViolation(int32_t *out)              // :206-228
{
    isOptimal = true;
    int64_t pos = -1;
    GNE_TRY(gne_price_first(dev, &pos));
    arcsPriced += (double) (pos < 0 ? (int64_t) setNBarc.size() : pos + 1);
    float ms = 0; gne_dev_last_kernel_ms(dev, &ms); kernelMs += ms;
    if (pos < 0) { *out = -1; return GNE_OK; }
    isOptimal = false;
    *out = setNBarc[pos];
    return GNE_OK;
}

int GenNetEq::getRandomVarWithViolation(int32_t *out)             // :273-281
{
    GNE_TRY(computeReducedCosts(false));
    if (offendingVars.empty()) { *out = -1; return GNE_OK; }
    *out = offendingVars[(size_t) (drand48() * offendingVars.size())];
    return GNE_OK;
}

int GenNetEq::getRandomVarWeightedByViolation(int32_t *out)       // :284-304
{
    GNE_TRY(computeReducedCosts(false));
    *out = -1;
    if (!offendingVars.empty()) {
        double sum = 0.0;
        for (size_t i = 0; i < offendingVars.size(); ++i) sum += offendingViol[i];     // fabs(var->redCost)
        const double level = drand48() * sum;
        double run = 0.0;
        for (size_t i = 0; i < offendingVars.size(); ++i) {
            run += offendingViol[i];
            if (level <= run) { *out = offendingVars[i]; break; }
        }
    }
    return GNE_OK;
}

int GenNetEq::getRandomVarWithLargeViolation(int32_t *out)        // :307-362
{
    isOptimal = true;
    GNE_TRY(fetchViolators(0.0));
    std::priority_queue<VarWithViol, std::vector<VarWithViol>, violcomp> q;
    for (size_t i = 0; i < offendingVars.size(); ++i) {             // position order, as the reference pushes
        isOptimal = false;
        VarWithViol vv; vv.var = offendingVars[i]; vv.violation = -1.0 * offendingViol[i];
        q.push(vv);
        while ((int) q.size() > numLargestViolVarsToTrack) q.pop();
    }
    std::vector<int32_t> fin;
    while (!q.empty()) { fin.push_back(q.top().var); q.pop(); }
    if (fin.empty()) { *out = -1; return GNE_OK; }
    *out = fin[(size_t) (drand48() * fin.size())];
    return GNE_OK;
}

int GenNetEq::getVarBySimAnnealing(bool arcsFirst, int32_t *out)  // :364-396; both branches are LV at p = 0
{
    (void) drand48();
    GNE_TRY(getVarWithLargestViolation(out));
    if (arcsFirst && lastViolation * 100 < baselineViolation && baselineViolation > 5000) baselineViolation = lastViolation;
    return GNE_OK;
}

int GenNetEq::updateCandidateList()                               // :504-535
{
    while (!candidateList.empty()) candidateList.pop();
    GNE_TRY(fetchViolators(0.0));
    for (size_t i = 0; i < offendingVars.size(); ++i) {
        isOptimal = false;
        VarWithViol vv; vv.var = offendingVars[i]; vv.violation = offendingViol[i];
        candidateList.push(vv);
    }
    return GNE_OK;
}

int GenNetEq::getNextCandidate(int32_t *out)                      // :457-501
{
    isOptimal = true;
    if (itersSinceUpdate < 0 || itersSinceUpdate > majorIterationLength) {
        GNE_TRY(updateCandidateList());
        itersSinceUpdate = 0;
    } else {
        ++itersSinceUpdate;
    }
    int32_t ev = -1;
    while (!candidateList.empty()) {
        const int32_t c = candidateList.top().var;
        candidateList.pop();
        const bool atLB = (setLoc[c] == L_var);
        const double rc = computeArcRedCost(c);
        if ((atLB && rc < (-1.0 * TOL)) || (!atLB && rc > TOL)) { isOptimal = false; ev = c; break; }
    }
    if (ev == -1 && itersSinceUpdate > 0) {
        itersSinceUpdate = -1;
        return getNextCandidate(out);
    }
    *out = ev;
    return GNE_OK;
}

int GenNetEq::updateCandidateList2()                              // :613-677 with setNBeqf empty (tryToAddArc :679-689)
{
    candidateList2.clear(); candidateViol2.clear();
    const int64_t numNB = (int64_t) setNBarc.size();
    if (nbArcPivotInd >= numNB) nbArcPivotInd = 0;
    const size_t cap = (size_t) clMaxSize + (size_t) numArcsToAlwaysCheck + 16;
    if (listPos.size() < cap) { listPos.resize(cap); listViol.resize(cap); }
    int64_t cnt = 0, nc = nbArcPivotInd, sc = 0;
    // exhaustedEqfs is true from the start, so the drand48 arc/eqf interleave (:643-653) never runs
    GNE_TRY(gne_price_scan_list(dev, nbArcPivotInd, (int64_t) clMaxSize, (int64_t) numArcsToAlwaysCheck, &cnt,
                                listPos.data(), listViol.data(), (int64_t) listPos.size(), &nc, &sc));
    arcsPriced += (double) sc;
    if (cnt > 0) isOptimal = false;
    for (int64_t i = 0; i < cnt; ++i) { candidateList2.push_back(setNBarc[listPos[i]]); candidateViol2.push_back(listViol[i]); }
    nbArcPivotInd = nc;                                 // rolling cursor: never restored by this rule
    return GNE_OK;
}

int GenNetEq::getNextCandidate2(int32_t *out)                     // :550-610
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
        double arc;
        if (itersSinceUpdate > 0) {
            const double rc = computeArcRedCost(c);     // re-priced with the current potentials
            viol = (atLB && rc < (-1.0 * TOL)) || (!atLB && rc > TOL);
            arc = fabs(rc);
        } else {                                        // "reduced cost is up to date": the value the list pass stored
            viol = true;
            arc = candidateViol2[k];
        }
        if (viol) {
            isOptimal = false;
            if (largest < arc) { largest = arc; ev = c; evInd = (int64_t) k; }
            ++k;
        } else {
            candidateList2[k] = candidateList2.back(); candidateList2.pop_back();
            candidateViol2[k] = candidateViol2.back(); candidateViol2.pop_back();
        }
    }
    if (ev == -1 && itersSinceUpdate > 0) {
        itersSinceUpdate = -1;
        return getNextCandidate2(out);
    }
    if (evInd >= 0) {
        candidateList2[evInd] = candidateList2.back(); candidateList2.pop_back();
        candidateViol2[evInd] = candidateViol2.back(); candidateViol2.pop_back();
    }
    *out = ev;
    return GNE_OK;
}

int GenNetEq::quickSelect(int32_t *out)                           // :710-764, arc loop
{
    isOptimal = true;
    if (nbArcPivotInd >= (int64_t) setNBarc.size()) nbArcPivotInd = 0;
    const int64_t start = nbArcPivotInd;
    int64_t bestPos = -1, newCursor = start, scanned = 0, violations = 0;
    double bestViol = 0.0;
    GNE_TRY(gne_price_qs(dev, start, (int64_t) numNodes, &bestPos, &bestViol, &newCursor, &scanned, &violations));
    arcsPriced += (double) scanned;
    float ms = 0; gne_dev_last_kernel_ms(dev, &ms); kernelMs += ms;
    if (violations > 0) isOptimal = false;
    nbArcPivotInd = newCursor;
    if ((loopIter % majorIterationLength) > 0) nbArcPivotInd = start;
    *out = (bestPos >= 0) ? setNBarc[bestPos] : -1;
    return GNE_OK;
}

int GenNetEq::selectEnteringVariable(int32_t *out)                // :36-127
{
    if (cyclingDetected) return getRandomVarWithLargeViolation(out);      // useRandomJumpWhenCycling == true
    const int rule = presolve ? phaseIpivot : phaseIIpivot;
    switch (rule) {
      case GNE_LV: case GNE_LVW: case GNE_LVA: case GNE_LVE: return getVarWithLargestViolation(out);
      case GNE_FV: return getFirstVarWithViolation(out);
      case GNE_RV: return getRandomVarWithViolation(out);
      case GNE_RWV: return getRandomVarWeightedByViolation(out);
      case GNE_RLV: case GNE_RLVW: return getRandomVarWithLargeViolation(out);
      case GNE_CL: case GNE_CLW: return getNextCandidate(out);
      case GNE_SAA: return getVarBySimAnnealing(true, out);
      case GNE_SAE: return getVarBySimAnnealing(false, out);
      case GNE_CL2: case GNE_CLW2: return getNextCandidate2(out);
      case GNE_QS: case GNE_QSW: return quickSelect(out);
    }
    gne_set_error("pivot rule %d is not supported (SD relies on the stale computeMinDelta, gnePivot.cpp:406)", rule);
    return GNE_ERR_UNSUPPORTED;
}

int GenNetEq::identifyEnteringVariable() { return selectEnteringVariable(&eVar); }   // :28-33

// ------------------------------------------------------------------------------------------
// gnePivot.cpp
// ------------------------------------------------------------------------------------------
void GenNetEq::identifyLeavingVar()                               // :166-209
{
    if (UB[eVar] <= delta + TOL) {
        delta = UB[eVar];
        lVar = eVar;
    } else {
        if (blockingVars.empty()) { isUnbounded = true; lVar = -1; return; }
        lVar = blockingVars[(size_t) (drand48() * blockingVars.size())];   // selectRandomBlockingVar == true
    }
    if (delta <= TOL) { ++totalDegenPivots; ++consecutiveDegenPivots; }
    else consecutiveDegenPivots = 0;
}

void GenNetEq::applyTreeFlow(int32_t slot)                        // :79-91
{
    forest.ensureArcs(slot);
    const Tree &t = forest.tree(slot);
    for (size_t k = 0; k < t.arcs.size(); ++k) {
        const int32_t a = t.arcs[k];
        const double d = delta * deltaFlow[a];
        flow[a] += d;
        flowCost += cost[a] * d;
        deltaFlow[a] = 0.0;
    }
    const int32_t ex = t.extraArc;
    const double d = delta * deltaFlow[ex];
    flow[ex] += d;
    flowCost += cost[ex] * d;
    deltaFlow[ex] = 0.0;
}

int GenNetEq::pivotTI(int32_t eav)                                // :63-119
{
    const int32_t tU = forest.treeOf(tail[eav]);
    const int32_t tV = forest.treeOf(head[eav]);
    delta = DBL_MAX;
    markFlowStale(tU); markFlowStale(tV);                           // their flows change below
    const double f0 = wallNow();
    if (!sparseFlows) {
        computeFlowsPivotingTI(eav, tU, tV);
        identifyLeavingVar();
        if (lVar == -1) return GNE_OK;
        applyTreeFlow(tU);
        if (tU != tV) applyTreeFlow(tV);
    } else {
        // supply injection of computeFlowsPivotingTI (gneFlow.cpp:160-174), then the sparse sweeps
        if (setLoc[eav] == L_var) {
            if (tail[eav] == head[eav]) supA0[tail[eav]] = -1.0 * (1.0 - mult[eav]);
            else { supA0[tail[eav]] = -1.0; supA0[head[eav]] = mult[eav]; }
        } else {
            if (tail[eav] == head[eav]) supA0[tail[eav]] = (1.0 - mult[eav]);
            else { supA0[tail[eav]] = 1.0; supA0[head[eav]] = -1.0 * mult[eav]; }
        }
        int32_t starts[2] = { tail[eav], head[eav] };
        if (tU == tV) {
            computeFlowsSparse(tU, starts, 2, activeArcsU);
        } else {
            computeFlowsSparse(tU, starts, 1, activeArcsU);
            computeFlowsSparse(tV, starts + 1, 1, activeArcsV);
        }
        identifyLeavingVar();
        if (lVar == -1) return GNE_OK;
        applyFlowSparse(tU, activeArcsU);
        if (tU != tV) applyFlowSparse(tV, activeArcsV);
    }
    if (setLoc[eav] == L_var) { flow[eav] = delta; flowCost += cost[eav] * delta; }
    else { flow[eav] = UB[eav] - delta; flowCost -= cost[eav] * delta; }
    secsFlows += wallNow() - f0;
    return GNE_OK;
}

int GenNetEq::updateBasis()                                       // :254-309
{
    if (eVar == lVar) {
        const int8_t was = setLoc[eVar];
        GNE_TRY(removeFromNB(eVar));
        GNE_TRY(insertIntoNB(eVar, was == L_var ? U_var : L_var));
    } else {
        GNE_TRY(removeFromNB(eVar));
        const double t0 = wallNow();
        if (forest.rehang(lVar, eVar)) {
            // the forest did removeArc(lVar) + insertArc(eVar) in one step; the set bookkeeping of removeFromB / insertIntoB
            const int32_t prevInd = setInd[lVar];
            const int32_t mv = setBarc.back();
            setBarc[prevInd] = mv; setInd[mv] = prevInd; setBarc.pop_back();
            setLoc[lVar] = NIL; setInd[lVar] = -1;
            setBarc.push_back(eVar); setLoc[eVar] = B_var; setInd[eVar] = (int32_t) setBarc.size() - 1;
        } else {
            GNE_TRY(removeFromB(lVar));
            GNE_TRY(insertIntoB(eVar));
        }
        secsTrees += wallNow() - t0;
        if (flow[lVar] > (0.5 * UB[lVar])) GNE_TRY(insertIntoNB(lVar, U_var));
        else GNE_TRY(insertIntoNB(lVar, L_var));
        const double t1 = wallNow();
        forest.updateTrees();
        secsTrees += wallNow() - t1;
    }
    if (setLoc[lVar] == L_var) {                                   // bound snap :294-306
        if (fabs(flow[lVar] - 0.0) > TOL) flow[lVar] = 0.0;
        if (flow[lVar] != 0.0 || std::signbit(flow[lVar])) lDirty.push_back(lVar);      // the next computeFlows resets it to +0.0
    } else {
        if (fabs(flow[lVar] - UB[lVar]) > TOL) flow[lVar] = UB[lVar];
    }
    return GNE_OK;
}

void GenNetEq::checkForCycling()                                  // :312-356
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

int GenNetEq::pivot()                                             // :28-60
{
    if (forest.typeOf(tail[eVar]) == TREE_I && forest.typeOf(head[eVar]) == TREE_I) {
        GNE_TRY(pivotTI(eVar));
    } else {
        gne_set_error("pivotTII reached (equal-flow sets are out of scope)");
        return GNE_ERR_UNSUPPORTED;
    }
    if (lVar == -1) return GNE_OK;
    GNE_TRY(updateBasis());
    for (size_t c = 0; c < forest.changed.size(); ++c) markFlowStale(forest.changed[c]);     // split parts, merged trees
    markFlowStale(forest.treeOf(tail[eVar])); markFlowStale(forest.treeOf(head[eVar]));
    markFlowStale(forest.treeOf(tail[lVar])); markFlowStale(forest.treeOf(head[lVar]));
    checkForCycling();
    return status;
}

// ------------------------------------------------------------------------------------------
// genneteq.cpp
// ------------------------------------------------------------------------------------------
void GenNetEq::initializeStats(bool usePresolve)                  // :123-167
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
    clMaxSize = numNodes;
    nbArcPivotInd = 0;
    numArcsToAlwaysCheck = 5;
    baselineViolation = lastViolation;
}

int GenNetEq::setVariableCosts()                                  // :304-323 + re-upload of the cost column
{
    if (presolve) for (int64_t i = 0; i < M; ++i) cost[i] = isArtificial[i] ? 1.0 : 0.0;
    else for (int64_t i = 0; i < M; ++i) cost[i] = actualCost[i];
    // every tree's potentials depend on the costs
    for (int s = 0; s < forest.numSlots(); ++s) if (forest.tree(s).type != TREE_FREE) forest.markChanged(s);
    allPotDirty = true;
    if (hostOnly) return GNE_OK;
    if (!dev) {
        uint8_t none[128] = { 0 };
        GNE_TRY(setComm(none, 0, 1));
    }
    const size_t N = setNBarc.size();
    if (!costsOnDevice) {
        std::vector<int32_t> t(N), h(N);
        std::vector<double> c(N), mu(N);
        std::vector<int8_t> st(N);
        for (size_t p = 0; p < N; ++p) { const int32_t a = setNBarc[p]; t[p] = tail[a]; h[p] = head[a]; c[p] = cost[a]; mu[p] = mult[a]; st[p] = setLoc[a]; }
        GNE_TRY(gne_nb_reset(dev, (int64_t) N, t.data(), h.data(), c.data(), mu.data(), st.data()));
        costsOnDevice = true;
    } else {
        std::vector<double> c(N);
        for (size_t p = 0; p < N; ++p) c[p] = cost[setNBarc[p]];
        GNE_TRY(gne_dev_flush(dev));
        GNE_TRY(gne_nb_set_costs(dev, (int64_t) N, c.data()));
    }
    return GNE_OK;
}

// Host-only replay of a recorded pivot sequence (no device, no pricing): the entering arc of each
// pivot comes from the trace, flows / ratio test / leaving arc / basis update run as in solve(),
// and the leaving arc is compared with the trace.  `draws` = drand48 calls the pricing rule makes
// per successful call (LV 1, QS/CL/FV 0, SAA/SAE 2, RV/RWV/RLV 1); the cycling override draws once.
int GenNetEq::replayTrace(int draws, int64_t p1, int64_t total, const int32_t *e, const int32_t *l, int64_t *mismatch)
{
    hostOnly = true;
    *mismatch = -1;
    int64_t k = 0;
    for (int phase = 0; phase < 2; ++phase) {
        const bool pre = (phase == 0);
        const bool pIwasInfeasible = pre ? false : isInfeasible;
        initializeStats(pre);
        GNE_TRY(setVariableCosts());
        if (!presolve && pIwasInfeasible) { isInfeasible = true; break; }
        computeFlows(0);
        const int64_t end = pre ? p1 : total;
        for (; k < end; ++k) {
            if ((loopIter % 1000) == 0) { const double q0 = wallNow(); computeFlows(2); secsTotal += wallNow() - q0; }
            { const double q0 = wallNow(); GNE_TRY(computePotentials()); secsPotentials += wallNow() - q0; }
            if (cyclingDetected) (void) drand48();
            else for (int d = 0; d < draws; ++d) (void) drand48();
            eVar = e[k];
            { const double q0 = wallNow(); GNE_TRY(pivot()); secsPivot += wallNow() - q0; }
            if (lVar != l[k]) { *mismatch = k; return GNE_OK; }
            ++loopIter; ++pivots;
        }
        // the terminating pricing call of the phase: the simulated-annealing rules draw before they
        // price (gnePivotRules.cpp:369,386), every other rule draws only when it found a violator
        if (!cyclingDetected && draws == 2) (void) drand48();
        computeFlows(0);
        if (presolve) checkFeasibility();
        GNE_TRY(computePotentials());
    }
    if (getenv("GNE_HOST_PROFILE"))
        fprintf(stderr, "host replay: pivot %.3f s, potentials %.3f s, periodic full flows %.3f s; flows %.3f s, trees %.3f s (split %.3f, reindex %.3f, merge %.3f)\n",
                secsPivot, secsPotentials, secsTotal, secsFlows, secsTrees,
                forest.secsSplit, forest.secsReindex, forest.secsMerge),
        fprintf(stderr, "  nodes visited: split %lld (%lld subtree cuts, %lld re-hangs, %lld full DFS), reindex %lld in %lld calls, order materialised %lld\n", (long long) forest.nodesSplit,
                (long long) forest.callsSplitFast, (long long) forest.callsRehang, (long long) forest.callsSplitFull, (long long) forest.nodesReindex, (long long) forest.callsReindex, (long long) forest.nodesMaterialized);
    return GNE_OK;
}

void GenNetEq::checkFeasibility()                                 // :227-288
{
    isInfeasible = false;
    for (size_t k = 0; k < setBarc.size(); ++k) {
        const int32_t a = setBarc[k];
        if ((isArtificial[a] && flow[a] > TOL) || flow[a] < 0.0 - TOL || flow[a] > UB[a] + TOL) { isInfeasible = true; return; }
    }
    for (std::set<int32_t>::const_iterator it = uPos.begin(); it != uPos.end(); ++it)       // nonbasic artificial at its upper bound
        if (isArtificial[setNBarc[*it]]) { isInfeasible = true; return; }
}

// GenNetEq::solve (genneteq.cpp:60-117) as a state machine, so that one host thread can drive several solvers:
//   solveBegin   everything before the loop;
//   solveStep    advances by one stage and never blocks on the device unless `block`: PREP (loop head: window / progress
//                hooks, periodic flows, computePotentials, then either the LV pass is LAUNCHED or - any other rule - the
//                entering variable is selected synchronously), WAIT (the launched pass: complete it when its result is
//                there), then the pivot.  Returns 1 while there is more to do (2: nothing could be done, the pass is still
//                running), 0 when the loop has ended, -(gne_status) on error;
//   solveEnd     everything after the loop.
// solve() = solveBegin + blocking solveStep until 0 + solveEnd.
int GenNetEq::solveBegin(bool usePresolve, int64_t maxPivots)
{
    const bool pIwasInfeasible = usePresolve ? false : isInfeasible;
    initializeStats(usePresolve);
    GNE_TRY(setVariableCosts());
    stFinished = 1;
    stSkip = false;
    stStage = 0;
    if (!presolve && pIwasInfeasible) { isInfeasible = true; stSkip = true; return GNE_OK; }
    stT0 = wallNow();
    computeFlows(0);
    stMaxPivots = maxPivots;
    stBudget = maxPivots;
    callPivots = 0;
    return GNE_OK;
}

#define STEP_TRY(call) do { int rc_ = (call); if (rc_ != GNE_OK) return -rc_; } while (0)
int GenNetEq::solveStep(bool block)
{
    if (stSkip) return 0;
    if (stStage == 0) {
        if (!(!isOptimal && !isUnbounded)) return 0;
        if (winStart >= 0 && callPivots == winStart && winT0 == 0.0) { gne_dev_sync(dev); getCounters(winC0); winT0 = wallNow(); }
        if (winStart >= 0 && callPivots == winStart + winLen && winT1 == 0.0) { gne_dev_sync(dev); winT1 = wallNow(); getCounters(winC1); }
        if (progEvery > 0 && progBuf && callPivots % progEvery == 0 && progRows < progCap) {
            double *row = progBuf + 14 * progRows++;
            row[0] = wallNow();
            getCounters(row + 1);
            size_t big = 0;
            for (int sl = 0; sl < forest.numSlots(); ++sl) if (forest.tree(sl).type != TREE_FREE) big = std::max(big, forest.tree(sl).nodes.size());
            row[13] = (double) big;
        }
        if (stMaxPivots >= 0 && stBudget == 0) { stFinished = 0; return 0; }
        if ((loopIter % 1000) == 0) computeFlows(2);
        const double a = wallNow();
        STEP_TRY(computePotentials());
        stB = wallNow();
        secsPotentials += stB - a;
        const int rule = presolve ? phaseIpivot : phaseIIpivot;
        const bool lvRule = rule == GNE_LV || rule == GNE_LVW || rule == GNE_LVA || rule == GNE_LVE;
        if (lvRule && !cyclingDetected) {
            // getVarWithLargestViolation, first half: the pass is on its way
            isOptimal = true;
            offendingVars.clear();
            // (a solver of a batch parks the pass in its thread's launch group: one launch for several instances)
            STEP_TRY(lvGroup ? gne_lv_group_add(lvGroup, dev) : gne_price_lv_launch(dev));
            stStage = 1;
            if (!block) return 1;
        } else {
            STEP_TRY(identifyEnteringVariable());
            stStage = 2;
        }
    }
    if (stStage == 1) {
        if (!block && !gne_price_lv_ready(dev)) return 2;       // still waiting: nothing was done
        if (block && lvGroup) STEP_TRY(gne_lv_group_launch(lvGroup));     // (nothing may stay parked while this thread waits)
        const double c0 = wallNow();
        STEP_TRY(completeLargestViolation(&eVar));
        secsComplete += wallNow() - c0;
        stStage = 2;
    }
    // stage 2: the pivot
    const double c = wallNow();
    secsPricing += c - stB;
    if (eVar != -1) {
        STEP_TRY(pivot());
        secsPivot += wallNow() - c;
        if (lVar != -1) {
            if (traceLen < traceCap) { traceE[traceLen] = eVar; traceL[traceLen] = lVar; }
            ++traceLen;
        }
        ++loopIter; ++pivots; ++callPivots;
        if (presolve) ++phaseIiters; else ++phaseIIiters;
        if (stBudget > 0) --stBudget;
    }
    stStage = 0;
    return 1;
}

#undef STEP_TRY

int GenNetEq::solveEnd(int64_t *iters, int *finished)
{
    *finished = stFinished;
    *iters = 0;
    if (stSkip) return GNE_OK;
    if (stFinished) {
        computeFlows(0);
        if (presolve) checkFeasibility();
        GNE_TRY(verifyOptimality());
    }
    secsTotal += wallNow() - stT0;
    *iters = loopIter;
    return GNE_OK;
}

int GenNetEq::solve(bool usePresolve, int64_t maxPivots, int64_t *iters, int *finished)   // :60-117
{
    *finished = 1; *iters = 0;
    GNE_TRY(solveBegin(usePresolve, maxPivots));
    while (true) {
        const int r = solveStep(true);
        if (r < 0) return -r;
        if (r == 0) break;
    }
    return solveEnd(iters, finished);
}

void GenNetEq::getFlows(double *out) const { memcpy(out, flow.data(), sizeof(double) * (size_t) M); }

int GenNetEq::getPotentials(double *out)
{
    // the device holds the authoritative pi; slots 0..numNodes-1 of its node arrays are ours
    std::vector<double> tmp((size_t) numNodes + 16);
    GNE_TRY(gne_pot_get_all(dev, tmp.data()));
    memcpy(out, tmp.data(), sizeof(double) * (size_t) numNodes);
    return GNE_OK;
}

void GenNetEq::getStates(int32_t *out) const { for (int64_t i = 0; i < M; ++i) out[i] = setLoc[i]; }

void GenNetEq::getWindow(double *o) const
{
    o[0] = (winT1 > 0.0 && winT0 > 0.0) ? winT1 - winT0 : -1.0;
    for (int i = 0; i < 12; ++i) { o[1 + i] = winC0[i]; o[13 + i] = winC1[i]; }
    o[25] = winT0; o[26] = winT1;                       // CLOCK_MONOTONIC seconds (comparable across threads)
}

void GenNetEq::getCounters(double *o) const
{
    int64_t h2d = 0, d2h = 0;
    if (dev) gne_dev_traffic(dev, &h2d, &d2h);
    o[0] = secsPricing; o[1] = secsPotentials; o[2] = secsPivot; o[3] = secsTotal; o[4] = arcsPriced; o[5] = kernelMs;
    o[6] = (double) h2d; o[7] = (double) d2h; o[8] = dev ? (double) gne_dev_launch_count(dev) : 0.0; o[9] = (double) pivots;
    o[10] = secsTrees; o[11] = secsFlows;
}

} // namespace gneb200
