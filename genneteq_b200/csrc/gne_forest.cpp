// genneteq_b200 — host basis forest; see gne_forest.h.
#include "gne_forest.h"

#include <algorithm>

namespace gneb200 {

void Forest::init(int nNodes, int64_t numVars, const int32_t *tailIn, const int32_t *headIn)
{
    n = nNodes; tail = tailIn; head = headIn;
    treeSlot.assign(n, -1); pred.assign(n, -1); predArc.assign(n, -1); depth.assign(n, -1); thread.assign(n, -1);
    threadPos.assign(n, -1); nodeFlag.assign(n, 1);
    arcPos.assign((size_t) numVars, -1);
    trees.clear(); freeSlots.clear(); stale.clear(); changed.clear(); changedFlag.clear(); dirtyNodes.clear();
    trees.reserve((size_t) n + 8);
    pool.clear(); pool.reserve((size_t) 4 * n + 64);
    incStart.assign(n, 0); incLen.assign(n, 0); incCap.assign(n, 0);
}

void Forest::markChanged(int32_t slot)
{
    if ((size_t) slot >= changedFlag.size()) changedFlag.resize((size_t) slot + 1, 0);
    if (!changedFlag[slot]) { changedFlag[slot] = 1; changed.push_back(slot); }
}

void Forest::touch(int32_t slot)
{
    Tree &t = trees[slot];
    t.upToDate = false;
    if (!t.queued) { t.queued = true; stale.push_back(slot); }
    markChanged(slot);
}

// Tree::createNewTree (trees.cpp:183-205)
int32_t Forest::newTree(int8_t type, int32_t rootNode)
{
    int32_t slot;
    if (!freeSlots.empty()) { slot = freeSlots.back(); freeSlots.pop_back(); }
    else { slot = (int32_t) trees.size(); trees.emplace_back(); }
    Tree &t = trees[slot];
    t.type = type; t.root = -1; t.extraArc = -1;
    t.nodes.clear(); t.arcs.clear(); t.threads.clear();
    touch(slot);
    if (rootNode >= 0) {
        t.root = rootNode;
        t.nodes.push_back(rootNode);
        setSlot(rootNode, slot);
    }
    return slot;
}

void Forest::freeTree(int32_t slot)
{
    Tree &t = trees[slot];
    t.type = TREE_FREE; t.root = -1; t.extraArc = -1;
    t.nodes.clear(); t.arcs.clear(); t.threads.clear();
    if (t.nodes.capacity() > 4096) {                    // give big buffers back, keep small ones for reuse
        std::vector<int32_t>().swap(t.nodes);
        std::vector<int32_t>().swap(t.arcs);
        std::vector<int32_t>().swap(t.threads);
    }
    freeSlots.push_back(slot);
}

// Tree::setTreeIDAndType (trees.cpp:220-230): membership only; slots never get renumbered
void Forest::relabel(int32_t slot, int32_t newSlot)
{
    const std::vector<int32_t> &nd = trees[slot].nodes;
    for (size_t k = 0; k < nd.size(); ++k) setSlot(nd[k], newSlot);
}

void Forest::incPush(int32_t node, int32_t arc, int32_t nbr)
{
    if (incLen[node] == incCap[node]) {
        const int32_t nc = std::max(4, 2 * incCap[node]);
        const int64_t ns = (int64_t) pool.size();
        pool.resize(pool.size() + (size_t) nc);
        for (int32_t k = 0; k < incLen[node]; ++k) pool[ns + k] = pool[incStart[node] + k];
        incStart[node] = ns; incCap[node] = nc;
    }
    Inc e; e.arc = arc; e.nbr = nbr;
    pool[incStart[node] + incLen[node]++] = e;
}

void Forest::compactPool()
{
    std::vector<Inc> np;
    np.reserve((size_t) 4 * n + 64);
    for (int i = 0; i < n; ++i) {
        if (incLen[i] == 0) { incStart[i] = 0; incCap[i] = 0; continue; }
        int32_t c = 4;
        while (c < incLen[i]) c *= 2;
        const int64_t ns = (int64_t) np.size();
        np.resize(np.size() + (size_t) c);
        for (int32_t k = 0; k < incLen[i]; ++k) np[ns + k] = pool[incStart[i] + k];
        incStart[i] = ns; incCap[i] = c;
    }
    pool.swap(np);
}

// Tree::mergeTxAndTII (trees.cpp:365-395)
void Forest::mergeInto(int32_t tX, int32_t tII, int32_t av)
{
    relabel(tII, tX);
    Tree &x = trees[tX];
    Tree &y = trees[tII];
    x.nodes.insert(x.nodes.end(), y.nodes.begin(), y.nodes.end());
    for (size_t k = 0; k < y.arcs.size(); ++k) { arcPos[y.arcs[k]] = (int32_t) x.arcs.size(); x.arcs.push_back(y.arcs[k]); }
    arcPos[av] = (int32_t) x.arcs.size();
    x.arcs.push_back(av);
    incPush(tail[av], av, head[av]);
    incPush(head[av], av, tail[av]);
    touch(tX);
    freeTree(tII);
}

// Tree::expandTreeUsingArc (trees.cpp:398-437)
void Forest::expandWithArc(int32_t t, int32_t av)
{
    const int32_t u = tail[av], v = head[av];
    Tree &x = trees[t];
    if (treeSlot[u] != t) { x.nodes.push_back(u); setSlot(u, t); }
    else if (treeSlot[v] != t) { x.nodes.push_back(v); setSlot(v, t); }
    if (u != v) {
        arcPos[av] = (int32_t) x.arcs.size();
        x.arcs.push_back(av);
        incPush(u, av, v);
        incPush(v, av, u);
    } else {
        x.extraArc = av;
    }
    touch(t);
}

// Tree::insertIntoTrees (trees.cpp:290-362)
int Forest::insertArc(int32_t av)
{
    const int32_t u = tail[av], v = head[av];
    const int32_t tU = treeSlot[u], tV = treeSlot[v];
    if (tU >= 0 && tV >= 0) {
        if (tU == tV) {
            if (trees[tU].type == TREE_I) return -2;     // "Insertion adds cycle to type I tree"
            trees[tU].type = TREE_I;                     // convertTIItoTI: threads stay valid
            trees[tU].extraArc = av;
            markChanged(tU);
        } else {
            const int8_t aU = trees[tU].type, aV = trees[tV].type;
            if (aU == TREE_II && aV == TREE_II) mergeInto(tU, tV, av);
            else if (aU == TREE_I && aV == TREE_II) mergeInto(tU, tV, av);
            else if (aU == TREE_II && aV == TREE_I) mergeInto(tV, tU, av);
            else return -3;                              // "Insertion merges two type I trees"
        }
    } else {
        int32_t t;
        if (tU < 0 && tV < 0) t = newTree(u == v ? TREE_I : TREE_II, u);
        else t = (tU >= 0) ? tU : tV;
        expandWithArc(t, av);
    }
    return 0;
}

// Tree::splitTree (trees.cpp:502-566): DFS from the old root over the incident lists; the part
// reached through the removed arc becomes `bottom`, the rest `top`.  The discovery order IS the
// new arcsInTree / nodesInTree order.  `top` keeps the old tree's slot (the reference creates two
// fresh trees; tree ids have no observable effect), so its nodes keep their slot.
int32_t Forest::split(int32_t tOld, int32_t removedArc)
{
    const int32_t bottom = newTree(TREE_II, -1);         // may grow `trees`: take references afterwards
    const int32_t top = tOld;
    const int32_t oldRoot = trees[top].root;
    {
        Tree &o = trees[top];
        oldNodes.clear(); oldArcs.clear();
        oldNodes.swap(o.nodes); oldArcs.swap(o.arcs);    // recycle capacity; contents are rebuilt below
        o.type = TREE_II; o.extraArc = -1; o.root = -1;
        touch(top);
    }
    int32_t cur = top;
    std::vector<int32_t> &curStack = stackA, &prevStack = stackB;
    curStack.clear(); prevStack.clear();
    curStack.push_back(oldRoot);
    prevStack.push_back(-1);
    while (!curStack.empty()) {
        int32_t node = curStack.back(); curStack.pop_back();
        if (node < 0) {
            cur = top;
            if (curStack.empty()) break;
            node = curStack.back(); curStack.pop_back();
        }
        const int32_t prev = prevStack.back(); prevStack.pop_back();
        Tree &tc = trees[cur];
        setSlot(node, cur);
        if (tc.nodes.empty()) tc.root = node;
        tc.nodes.push_back(node);
        int32_t removedInd = -1, removedNbr = -1;
        const int64_t base = incStart[node];
        const int32_t len = incLen[node];
        for (int32_t k = 0; k < len; ++k) {
            const Inc e = pool[base + k];
            if (e.arc == removedArc) { removedInd = k; removedNbr = e.nbr; continue; }
            if (e.nbr != prev) {
                arcPos[e.arc] = (int32_t) tc.arcs.size();
                tc.arcs.push_back(e.arc);
                curStack.push_back(e.nbr);
                prevStack.push_back(node);
            }
        }
        if (removedInd >= 0) {
            pool[base + removedInd] = pool[base + len - 1];
            --incLen[node];
            if (prev > -2) {
                curStack.push_back(-1);
                curStack.push_back(removedNbr);
                prevStack.push_back(-2);
                cur = bottom;
            }
        }
    }
    if (trees[bottom].nodes.empty()) { freeTree(bottom); return -1; }
    return bottom;
}

// Tree::removeFromTrees (trees.cpp:456-499)
int Forest::removeArc(int32_t av)
{
    const int32_t u = tail[av], v = head[av];
    if (treeSlot[u] != treeSlot[v] || treeSlot[u] < 0) return -4;   // "remove arc spanning two trees"
    const int32_t t = treeSlot[u];
    if (trees[t].type == TREE_I && av == trees[t].extraArc) {
        trees[t].type = TREE_II;                         // convertTItoTII: threads stay valid
        trees[t].extraArc = -1;
        markChanged(t);
    } else {
        const int8_t oldType = trees[t].type;
        const int32_t oldExtra = trees[t].extraArc;
        split(t, av);
        if (oldType == TREE_I) { const int rc = insertArc(oldExtra); if (rc) return rc; }
    }
    return 0;
}

// Tree::initializeTreeIndices (trees.cpp:597-644) + the dirty-node list
void Forest::reindex(int32_t slot)
{
    Tree &t = trees[slot];
    std::vector<int32_t> &nodeStack = stackA, &arcStack = stackB, &predStack = stackC;
    nodeStack.clear(); arcStack.clear(); predStack.clear();
    t.threads.clear();
    t.threads.reserve(t.nodes.size());
    nodeStack.push_back(t.root);
    arcStack.push_back(-1);
    predStack.push_back(-1);
    int32_t last = -1;
    while (!nodeStack.empty()) {
        const int32_t cur = nodeStack.back(); nodeStack.pop_back();
        const int32_t pa = arcStack.back(); arcStack.pop_back();
        const int32_t predNode = predStack.back(); predStack.pop_back();
        const bool dirty = (nodeFlag[cur] & 1) || pred[cur] != predNode || predArc[cur] != pa ||
                           (predNode >= 0 && (nodeFlag[predNode] & 2));
        nodeFlag[cur] = dirty ? 2 : 0;
        if (dirty) dirtyNodes.push_back(cur);
        depth[cur] = (predNode >= 0) ? depth[predNode] + 1 : 0;
        predArc[cur] = pa;
        pred[cur] = predNode;
        if (last >= 0) thread[last] = cur;
        last = cur;
        threadPos[cur] = (int32_t) t.threads.size();
        t.threads.push_back(cur);
        const int64_t base = incStart[cur];
        for (int32_t k = incLen[cur] - 1; k >= 0; --k) {
            const Inc e = pool[base + k];
            if (e.nbr == predNode) continue;
            nodeStack.push_back(e.nbr);
            arcStack.push_back(e.arc);
            predStack.push_back(cur);
        }
    }
    thread[last] = t.root;
}

// Tree::updateTrees (trees.cpp:572-591), restricted to the trees touched since the last call
void Forest::updateTrees()
{
    if (pool.size() > (size_t) 12 * n + 4096) compactPool();
    for (size_t k = 0; k < stale.size(); ++k) {
        const int32_t slot = stale[k];
        Tree &t = trees[slot];
        t.queued = false;
        if (t.type == TREE_FREE || t.upToDate) continue;
        if (t.type == TREE_I) t.root = tail[t.extraArc];  // keep the root on the cycle (:578)
        reindex(slot);
        t.upToDate = true;
    }
    stale.clear();
}

} // namespace gneb200
