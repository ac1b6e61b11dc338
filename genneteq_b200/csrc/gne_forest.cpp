// genneteq_b200 — host basis forest; see gne_forest.h.
#include "gne_forest.h"

#include <algorithm>
#include <stdlib.h>
#include <time.h>

namespace gneb200 {

static inline double nowS() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double) ts.tv_sec + 1.0e-9 * (double) ts.tv_nsec; }

void Forest::init(int nNodes, int64_t numVars, const int32_t *tailIn, const int32_t *headIn)
{
    n = nNodes; tail = tailIn; head = headIn;
    (void) numVars;
    { const char *ff = getenv("GNE_FOREST_FAST"); fastSplit = !(ff && ff[0] == '0'); }
    node.assign((size_t) n, NodeRec());
    trees.clear(); freeSlots.clear(); stale.clear(); changed.clear(); changedFlag.clear(); dirtyNodes.clear();
    trees.reserve((size_t) n + 8);
    pool.clear(); pool.reserve((size_t) 4 * n + 64);
    // (incident list position / length / capacity live in the node record: one cache line per node in a DFS)
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
    t.nodes.clear(); t.arcs.clear(); t.threads.clear(); t.pending.clear();
    t.tailArcs.clear(); t.ordFix.clear(); t.implicitOrder = false;
    t.rooted = false; t.threadsValid = false;
    touch(slot);
    if (rootNode >= 0) {
        t.root = rootNode;
        pushNode(slot, rootNode);
        setSlot(rootNode, slot);
    }
    return slot;
}

void Forest::freeTree(int32_t slot)
{
    Tree &t = trees[slot];
    t.type = TREE_FREE; t.root = -1; t.extraArc = -1;
    t.nodes.clear(); t.arcs.clear(); t.threads.clear();
    t.tailArcs.clear(); t.ordFix.clear(); t.implicitOrder = false;
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

void Forest::incPush(int32_t v, int32_t arc, int32_t nbr, int32_t pos)
{
    if (node[v].incLen == node[v].incCap) {
        const int32_t nc = std::max(4, 2 * node[v].incCap);
        const int64_t ns = (int64_t) pool.size();
        pool.resize(pool.size() + (size_t) nc);
        for (int32_t k = 0; k < node[v].incLen; ++k) pool[ns + k] = pool[node[v].incStart + k];
        node[v].incStart = ns; node[v].incCap = nc;
    }
    Inc e; e.arc = arc; e.nbr = nbr; e.pos = pos; e.ord = node[v].incLen;
    pool[node[v].incStart + node[v].incLen++] = e;
}

int32_t Forest::posOfArc(int32_t a, int32_t u, int32_t v) const
{
    for (int side = 0; side < 2; ++side) {
        const int32_t w = side ? v : u;
        const int64_t base = node[w].incStart;
        for (int32_t k = 0; k < node[w].incLen; ++k) if (pool[base + k].arc == a && pool[base + k].pos >= 0) return pool[base + k].pos;
    }
    return -1;
}

void Forest::compactPool()
{
    std::vector<Inc> np;
    np.reserve((size_t) 4 * n + 64);
    for (int i = 0; i < n; ++i) {
        if (node[i].incLen == 0) { node[i].incStart = 0; node[i].incCap = 0; continue; }
        int32_t c = 4;
        while (c < node[i].incLen) c *= 2;
        const int64_t ns = (int64_t) np.size();
        np.resize(np.size() + (size_t) c);
        for (int32_t k = 0; k < node[i].incLen; ++k) np[ns + k] = pool[node[i].incStart + k];
        node[i].incStart = ns; node[i].incCap = c;
    }
    pool.swap(np);
}

const Inc *Forest::entryOf(int32_t v, int32_t a) const
{
    const int64_t base = node[v].incStart;
    for (int32_t k = 0; k < node[v].incLen; ++k) if (pool[base + k].arc == a) return &pool[base + k];
    return nullptr;
}

void Forest::applyOrdFix(Tree &t)
{
    for (size_t q = 0; q < t.ordFix.size(); ++q) {
        const int32_t v = t.ordFix[q];
        const int64_t base = node[v].incStart;
        for (int32_t k = 0; k < node[v].incLen; ++k) pool[base + k].ord = k;
        node[v].flag &= (uint8_t) ~4;
    }
    t.ordFix.clear();
}

// Materialise arcs(tree) of a tree whose order is implicit: the split DFS (see split / splitFull) from the root over the
// arcs that were in the tree at its last split - children in the incident order of THAT moment (`ord`), appended subtrees
// (tail arcs) not entered - followed by the tail arcs in the order they were appended.  O(|tree|).
void Forest::ensureArcs(int32_t slot)
{
    Tree &t = trees[slot];
    if (!t.implicitOrder) return;
    t.arcs.clear();
    std::vector<int32_t> &st = stackA;
    std::vector<std::pair<int32_t, int32_t> > &srt = sortBuf;
    st.clear();
    st.push_back(t.root);
    while (!st.empty()) {
        const int32_t nd = st.back(); st.pop_back();
        const int32_t prev = (nd == t.root) ? -1 : node[nd].pred;
        const int64_t base = node[nd].incStart;
        const int32_t len = node[nd].incLen;
        if (node[nd].flag & 4) {                         // entries out of snapshot order: visit them by `ord`
            srt.clear();
            for (int32_t k = 0; k < len; ++k) srt.push_back(std::make_pair(pool[base + k].ord, k));
            std::sort(srt.begin(), srt.end());
        }
        for (int32_t q = 0; q < len; ++q) {
            Inc &e = pool[base + ((node[nd].flag & 4) ? srt[q].second : q)];
            if (e.pos >= POS_TAIL_BASE || e.pos == POS_TAIL_OTHER) continue;      // appended since the last split
            if (e.nbr == prev) { e.pos = POS_NONE; continue; }
            e.pos = (int32_t) t.arcs.size();
            t.arcs.push_back(e.arc);
            st.push_back(e.nbr);
        }
    }
    const int32_t nBase = (int32_t) t.arcs.size();
    for (size_t i = 0; i < t.tailArcs.size(); ++i) {
        const int32_t a = t.tailArcs[i];
        for (int side = 0; side < 2; ++side) {
            const int32_t w = side ? head[a] : tail[a];
            const int64_t base = node[w].incStart;
            for (int32_t k = 0; k < node[w].incLen; ++k) {
                Inc &e = pool[base + k];
                if (e.arc != a) continue;
                if (e.pos >= POS_TAIL_BASE) e.pos = nBase + (e.pos - POS_TAIL_BASE);
                else if (e.pos == POS_TAIL_OTHER) e.pos = POS_NONE;
            }
        }
        t.arcs.push_back(a);
    }
    t.tailArcs.clear();
    t.implicitOrder = false;
    applyOrdFix(t);
    nodesMaterialized += (int64_t) t.nodes.size();
}

// Tree::mergeTxAndTII (trees.cpp:365-395): arcs(x) = arcs(x) ++ arcs(y) ++ [av]
void Forest::mergeInto(int32_t tX, int32_t tII, int32_t av)
{
    const double t0 = nowS();
    if (trees[tX].rooted) {                              // remember where the new subtree hangs
        Attach at;
        at.arc = av;
        if (node[tail[av]].slot == tX) { at.parent = tail[av]; at.child = head[av]; }
        else { at.parent = head[av]; at.child = tail[av]; }
        trees[tX].pending.push_back(at);
    }
    ensureArcs(tII);                                     // y's order explicit (a cut subtree already is: split wrote it)
    Tree &x = trees[tX];
    Tree &y = trees[tII];
    // one pass over y's nodes: membership (Tree::setTreeIDAndType, trees.cpp:220-230), nodes(x), and the positions of y's
    // arcs, which keep their relative order behind x's - in x's tail (explicit positions POS_TAIL_BASE + index) when
    // x's order is implicit, else shifted behind arcs(x)
    const bool tailMode = x.implicitOrder;
    const int32_t shift = tailMode ? POS_TAIL_BASE + (int32_t) x.tailArcs.size() : (int32_t) x.arcs.size();
    for (size_t q = 0; q < y.nodes.size(); ++q) {
        const int32_t w = y.nodes[q];
        setSlot(w, tX);
        pushNode(tX, w);
        const int64_t base = node[w].incStart;
        if (tailMode) { for (int32_t k = 0; k < node[w].incLen; ++k) { Inc &e = pool[base + k]; e.pos = (e.pos >= 0) ? e.pos + shift : (int32_t) POS_TAIL_OTHER; } }
        else if (shift) { for (int32_t k = 0; k < node[w].incLen; ++k) if (pool[base + k].pos >= 0) pool[base + k].pos += shift; }
    }
    if (tailMode) {
        x.tailArcs.insert(x.tailArcs.end(), y.arcs.begin(), y.arcs.end());
        const int32_t pav = POS_TAIL_BASE + (int32_t) x.tailArcs.size();
        x.tailArcs.push_back(av);
        incPush(tail[av], av, head[av], pav);
        incPush(head[av], av, tail[av], POS_TAIL_OTHER);
    } else {
        x.arcs.insert(x.arcs.end(), y.arcs.begin(), y.arcs.end());
        const int32_t pav = (int32_t) x.arcs.size();
        x.arcs.push_back(av);
        incPush(tail[av], av, head[av], pav);
        incPush(head[av], av, tail[av], POS_NONE);
    }
    x.ordFix.insert(x.ordFix.end(), y.ordFix.begin(), y.ordFix.end());
    touch(tX);
    freeTree(tII);
    secsMerge += nowS() - t0;
}

// Tree::expandTreeUsingArc (trees.cpp:398-437)
void Forest::expandWithArc(int32_t t, int32_t av)
{
    const int32_t u = tail[av], v = head[av];
    ensureArcs(t);
    Tree &x = trees[t];
    if (node[u].slot != t) { pushNode(t, u); setSlot(u, t); }
    else if (node[v].slot != t) { pushNode(t, v); setSlot(v, t); }
    if (u != v) {
        const int32_t pav = (int32_t) x.arcs.size();
        x.arcs.push_back(av);
        incPush(u, av, v, pav);
        incPush(v, av, u, -1);
    } else {
        x.extraArc = av;
    }
    x.rooted = false;                                    // grown arc by arc: re-thread from the root
    touch(t);
}

// Tree::insertIntoTrees (trees.cpp:290-362)
int Forest::insertArc(int32_t av)
{
    const int32_t u = tail[av], v = head[av];
    const int32_t tU = node[u].slot, tV = node[v].slot;
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
int32_t Forest::splitFull(int32_t tOld, int32_t removedArc)
{
    const double t0 = nowS();
    { Tree &o = trees[tOld]; applyOrdFix(o); o.tailArcs.clear(); o.implicitOrder = false; }   // a new explicit order is written below
    const int32_t bottom = newTree(TREE_II, -1);         // may grow `trees`: take references afterwards
    const int32_t top = tOld;
    const int32_t oldRoot = trees[top].root;
    {
        Tree &o = trees[top];
        oldNodes.clear(); oldArcs.clear();
        oldNodes.swap(o.nodes); oldArcs.swap(o.arcs);    // recycle capacity; contents are rebuilt below
        o.type = TREE_II; o.extraArc = -1; o.root = -1;
        o.threadsValid = false;                          // a block of the thread disappears; the rooting of the rest stays
        touch(top);
    }
    int32_t cur = top;
    std::vector<int32_t> &curStack = stackA, &prevStack = stackB;
    curStack.clear(); prevStack.clear();
    curStack.push_back(oldRoot);
    prevStack.push_back(-1);
    while (!curStack.empty()) {
        int32_t nd = curStack.back(); curStack.pop_back();
        if (nd < 0) {
            cur = top;
            if (curStack.empty()) break;
            nd = curStack.back(); curStack.pop_back();
        }
        const int32_t prev = prevStack.back(); prevStack.pop_back();
        Tree &tc = trees[cur];
        setSlot(nd, cur);
        if (tc.nodes.empty()) tc.root = nd;
        pushNode(cur, nd);
        int32_t removedInd = -1, removedNbr = -1;
        const int64_t base = node[nd].incStart;
        const int32_t len = node[nd].incLen;
        for (int32_t k = 0; k < len; ++k) {
            Inc &e = pool[base + k];
            if (e.arc == removedArc) { removedInd = k; removedNbr = e.nbr; continue; }
            if (e.nbr != prev) {
                e.pos = (int32_t) tc.arcs.size();       // the arc's index in arcs(tree) lives with its discoverer
                tc.arcs.push_back(e.arc);
                curStack.push_back(e.nbr);
                prevStack.push_back(nd);
            } else {
                e.pos = POS_NONE;
            }
        }
        if (removedInd >= 0) {
            pool[base + removedInd] = pool[base + len - 1];
            pool[base + removedInd].ord = removedInd;
            --node[nd].incLen;
            if (prev > -2) {
                curStack.push_back(-1);
                curStack.push_back(removedNbr);
                prevStack.push_back(-2);
                cur = bottom;
            }
        }
    }
    secsSplit += nowS() - t0;
    nodesSplit += (int64_t) (trees[top].nodes.size() + trees[bottom].nodes.size()); ++callsSplitFull;
    if (trees[bottom].nodes.empty()) { freeTree(bottom); return -1; }
    return bottom;
}

// Tree::splitTree without visiting the part that keeps the root.  What the reference's DFS (splitFull) produces for a tree
// whose rooting is valid is known without running it:
//   * bottom = the subtree under the removed arc (pred pointers), its root the arc's child end; top = the rest, same root;
//   * arcs(top) / arcs(bottom) = the DFS order from those roots over the incident lists as they stand NOW - kept implicit:
//     an arc's rank is (pop rank of the parent end, index in its list), and two parent ends compare at their lowest common
//     ancestor by the list index of the branches (the later entry is popped first).  The lists only change afterwards by
//     appends (merges) and by the swap-with-last removal done here, whose moved entries keep their old index in `ord`
//     until the next split of the tree (Tree::ordFix).
// Cost: O(|bottom|) + the two incident lists, instead of O(|tree|).  The arcs merged into the tree since its last split
// (its tail) become part of the implicit order here, as the reference's DFS would re-sort them.
int32_t Forest::split(int32_t top, int32_t removedArc)
{
    {
        const Tree &o = trees[top];
        if (!fastSplit || !o.rooted || !o.pending.empty() || o.root < 0) return splitFull(top, removedArc);
    }
    const double t0 = nowS();
    const int32_t u = tail[removedArc], v = head[removedArc];
    int32_t c;                                           // child end of the removed arc
    if (node[v].predArc == removedArc && node[v].pred == u) c = v;
    else if (node[u].predArc == removedArc && node[u].pred == v) c = u;
    else return splitFull(top, removedArc);
    const int32_t par = (c == v) ? u : v;
    const int32_t bottom = newTree(TREE_II, -1);         // may grow `trees`: take references afterwards
    Tree &o = trees[top];
    Tree &bt = trees[bottom];
    // the new snapshot: tail arcs and lagging list indices are absorbed into "the lists as they stand now"
    for (size_t i = 0; i < o.tailArcs.size(); ++i) {
        const int32_t a = o.tailArcs[i];
        for (int side = 0; side < 2; ++side) {
            const int32_t w = side ? head[a] : tail[a];
            const int64_t base = node[w].incStart;
            for (int32_t k = 0; k < node[w].incLen; ++k) if (pool[base + k].arc == a) pool[base + k].pos = POS_NONE;
        }
    }
    o.tailArcs.clear();
    applyOrdFix(o);
    o.arcs.clear();
    o.implicitOrder = true;
    o.type = TREE_II; o.extraArc = -1;
    o.threadsValid = false;                              // a block of the thread disappears; the rooting of the rest stays
    touch(top);
    // bottom: the subtree of c, visited in the split DFS order from c over the lists as they stand (the removed arc
    // skipped), so that arcs(bottom) and the positions in its incident entries are written in the same pass: bottom is
    // almost always merged into another tree right away, which needs its order explicit
    std::vector<int32_t> &st = stackA;
    st.clear();
    st.push_back(c);
    int64_t visited = 0;
    while (!st.empty()) {
        const int32_t nd = st.back(); st.pop_back();
        ++visited;
        {   // leave nodes(top) by swap-with-last
            const int32_t idx = node[nd].nodeIdx, last = o.nodes.back();
            o.nodes[idx] = last; node[last].nodeIdx = idx; o.nodes.pop_back();
        }
        setSlot(nd, bottom);
        pushNode(bottom, nd);
        const int32_t prev = node[nd].pred;
        const int64_t base = node[nd].incStart;
        for (int32_t k = 0; k < node[nd].incLen; ++k) {
            Inc &e = pool[base + k];
            if (e.nbr == prev) { e.pos = POS_NONE; continue; }            // (for c: the removed arc's entry, nbr == par)
            e.pos = (int32_t) bt.arcs.size();
            bt.arcs.push_back(e.arc);
            st.push_back(e.nbr);
        }
    }
    bt.root = c; bt.implicitOrder = false; bt.rooted = false;
    node[c].pred = -1; node[c].predArc = -1; node[c].flag |= 1;
    // remove the arc from the two incident lists (swap with last).  In top the moved entry keeps its snapshot index
    // (`ord`) until the next split; bottom's order is explicit, so there the index is simply brought up to date
    for (int side = 0; side < 2; ++side) {
        const int32_t w = side ? c : par;
        const int64_t base = node[w].incStart;
        const int32_t len = node[w].incLen;
        for (int32_t k = 0; k < len; ++k) {
            if (pool[base + k].arc != removedArc) continue;
            if (k != len - 1) {
                pool[base + k] = pool[base + len - 1];
                if (side) pool[base + k].ord = k;
                else if (!(node[w].flag & 4)) { node[w].flag |= 4; o.ordFix.push_back(w); }
            }
            --node[w].incLen;
            break;
        }
    }
    secsSplit += nowS() - t0;
    nodesSplit += visited; ++callsSplitFast;
    return bottom;
}

bool Forest::rehang(int32_t out, int32_t in)
{
    if (!fastSplit) return false;
    const int32_t u = tail[out], v = head[out];
    const int32_t t = node[u].slot;
    if (t < 0 || node[v].slot != t || node[tail[in]].slot != t || node[head[in]].slot != t) return false;
    {
        const Tree &o = trees[t];
        if (o.type != TREE_I || out == o.extraArc || tail[in] == head[in] || !o.rooted || !o.pending.empty() || o.root < 0) return false;
    }
    int32_t c;                                           // child end of the leaving arc
    if (node[v].predArc == out && node[v].pred == u) c = v;
    else if (node[u].predArc == out && node[u].pred == v) c = u;
    else return false;
    const int32_t par = (c == v) ? u : v;
    // x is in the subtree of c  <=>  climbing from x to c's depth ends at c
    const int32_t dc = node[c].depth;
    auto below = [&](int32_t x) { while (node[x].depth > dc) x = node[x].pred; return x == c; };
    const int32_t ex = trees[t].extraArc;
    if (below(head[ex]) || below(tail[ex])) return false;                 // the leaving arc is on the cycle: the root may change
    const bool tb = below(tail[in]), hb = below(head[in]);
    if (tb == hb) return false;                                           // the entering arc does not reconnect the two parts
    const double t0 = nowS();
    Tree &o = trees[t];
    // (1) what split() does to the part that keeps the root: a new snapshot of the implicit order
    for (size_t i = 0; i < o.tailArcs.size(); ++i) {
        const int32_t a = o.tailArcs[i];
        for (int side = 0; side < 2; ++side) {
            const int32_t w = side ? head[a] : tail[a];
            const int64_t base = node[w].incStart;
            for (int32_t k = 0; k < node[w].incLen; ++k) if (pool[base + k].arc == a) pool[base + k].pos = POS_NONE;
        }
    }
    o.tailArcs.clear();
    applyOrdFix(o);
    if (!o.implicitOrder) { o.arcs.clear(); o.implicitOrder = true; }
    o.threadsValid = false;
    // (2) arcs(bottom) in the split DFS order from c (lists as they stand, the leaving arc skipped), written straight
    //     into the tail: arcs(t) = arcs(top) ++ arcs(bottom) ++ [in], as mergeInto would leave them
    std::vector<int32_t> &st = stackA;
    st.clear();
    st.push_back(c);
    int64_t visited = 0;
    while (!st.empty()) {
        const int32_t nd = st.back(); st.pop_back();
        ++visited;
        const int32_t prev = node[nd].pred;
        const int64_t base = node[nd].incStart;
        const int32_t len = node[nd].incLen;
        for (int32_t k = 0; k < len; ++k) {
            Inc &e = pool[base + k];
            if (e.nbr == prev) { e.pos = (e.arc == out) ? (int32_t) POS_NONE : (int32_t) POS_TAIL_OTHER; continue; }   // (no parallel tree arcs)
            e.pos = POS_TAIL_BASE + (int32_t) o.tailArcs.size();
            o.tailArcs.push_back(e.arc);
            st.push_back(e.nbr);
        }
    }
    // (3) the leaving arc leaves the two incident lists (swap with last); moved entries keep their snapshot index
    for (int side = 0; side < 2; ++side) {
        const int32_t w = side ? c : par;
        const int64_t base = node[w].incStart;
        const int32_t len = node[w].incLen;
        for (int32_t k = 0; k < len; ++k) {
            if (pool[base + k].arc != out) continue;
            if (k != len - 1) {
                pool[base + k] = pool[base + len - 1];
                if (!(node[w].flag & 4)) { node[w].flag |= 4; o.ordFix.push_back(w); }
            }
            --node[w].incLen;
            break;
        }
    }
    // (4) the entering arc: appended to both lists and to the tail, the subtree re-rooted from its end by updateTrees
    Attach at;
    at.arc = in;
    if (tb) { at.parent = head[in]; at.child = tail[in]; } else { at.parent = tail[in]; at.child = head[in]; }
    o.pending.push_back(at);
    const int32_t pav = POS_TAIL_BASE + (int32_t) o.tailArcs.size();
    o.tailArcs.push_back(in);
    incPush(tail[in], in, head[in], pav);
    incPush(head[in], in, tail[in], POS_TAIL_OTHER);
    touch(t);
    secsSplit += nowS() - t0;
    nodesSplit += visited; ++callsRehang;
    return true;
}

// Tree::removeFromTrees (trees.cpp:456-499)
int Forest::removeArc(int32_t av)
{
    const int32_t u = tail[av], v = head[av];
    if (node[u].slot != node[v].slot || node[u].slot < 0) return -4;   // "remove arc spanning two trees"
    const int32_t t = node[u].slot;
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
    const double t0 = nowS();
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
        NodeRec &r = node[cur];
        const bool dirty = (r.flag & 1) || r.pred != predNode || r.predArc != pa || (predNode >= 0 && (node[predNode].flag & 2));
        r.flag = (uint8_t) ((r.flag & 4) | (dirty ? 2 : 0));
        if (dirty) dirtyNodes.push_back(cur);
        r.depth = (predNode >= 0) ? node[predNode].depth + 1 : 0;
        r.predArc = pa;
        r.pred = predNode;
        if (last >= 0) node[last].thread = cur;
        last = cur;
        r.threadPos = (int32_t) t.threads.size();
        t.threads.push_back(cur);
        const int64_t base = node[cur].incStart;
        for (int32_t k = node[cur].incLen - 1; k >= 0; --k) {
            const Inc e = pool[base + k];
            if (e.nbr == predNode) continue;
            nodeStack.push_back(e.nbr);
            arcStack.push_back(e.arc);
            predStack.push_back(cur);
        }
    }
    node[last].thread = t.root;
    t.rooted = true; t.threadsValid = true; t.pending.clear();
    secsReindex += nowS() - t0;
    nodesReindex += (int64_t) t.threads.size(); ++callsReindex;
}

// A subtree hung under at.parent through at.arc: pred / predArc / depth of its nodes only.  The
// rest of the tree keeps its rooting, so nothing else has to be visited.
void Forest::rethreadSubtree(int32_t slot, const Attach &at)
{
    const double t0 = nowS();
    std::vector<int32_t> &nodeStack = stackA, &arcStack = stackB, &predStack = stackC;
    nodeStack.clear(); arcStack.clear(); predStack.clear();
    nodeStack.push_back(at.child);
    arcStack.push_back(at.arc);
    predStack.push_back(at.parent);
    int64_t visited = 0;
    while (!nodeStack.empty()) {
        const int32_t cur = nodeStack.back(); nodeStack.pop_back();
        const int32_t pa = arcStack.back(); arcStack.pop_back();
        const int32_t predNode = predStack.back(); predStack.pop_back();
        NodeRec &r = node[cur];
        const bool dirty = (r.flag & 1) || r.pred != predNode || r.predArc != pa || (node[predNode].flag & 2);
        r.flag = (uint8_t) ((r.flag & 4) | (dirty ? 2 : 0));
        if (dirty) dirtyNodes.push_back(cur);
        r.depth = node[predNode].depth + 1;
        r.predArc = pa;
        r.pred = predNode;
        ++visited;
        const int64_t base = node[cur].incStart;
        for (int32_t k = node[cur].incLen - 1; k >= 0; --k) {
            const Inc e = pool[base + k];
            if (e.nbr == predNode) continue;
            nodeStack.push_back(e.nbr);
            arcStack.push_back(e.arc);
            predStack.push_back(cur);
        }
    }
    (void) slot;
    secsReindex += nowS() - t0;
    nodesReindex += visited; ++callsReindex;
}

// preorder thread (initializeTreeIndices' visiting order) rebuilt from the rooting, when somebody needs it
void Forest::ensureThreads(int32_t slot)
{
    Tree &t = trees[slot];
    if (t.threadsValid || t.type == TREE_FREE) return;
    std::vector<int32_t> &nodeStack = stackA;
    nodeStack.clear();
    t.threads.clear();
    t.threads.reserve(t.nodes.size());
    nodeStack.push_back(t.root);
    int32_t last = -1;
    while (!nodeStack.empty()) {
        const int32_t cur = nodeStack.back(); nodeStack.pop_back();
        NodeRec &r = node[cur];
        if (last >= 0) node[last].thread = cur;
        last = cur;
        r.threadPos = (int32_t) t.threads.size();
        t.threads.push_back(cur);
        const int64_t base = node[cur].incStart;
        for (int32_t k = node[cur].incLen - 1; k >= 0; --k) if (pool[base + k].nbr != r.pred) nodeStack.push_back(pool[base + k].nbr);
    }
    node[last].thread = t.root;
    t.threadsValid = true;
}

int32_t Forest::incidentIndex(int32_t v, int32_t a) const
{
    const int64_t base = node[v].incStart;
    for (int32_t k = 0; k < node[v].incLen; ++k) if (pool[base + k].arc == a) return k;
    return -1;
}

// Tree::updateTrees (trees.cpp:572-591), restricted to the trees touched since the last call.  A
// tree whose root stays (the extra arc did not change) keeps the rooting of everything that was in
// it: removing a subtree changes no other node's predecessor, and a subtree hung in through a
// new arc only needs its own nodes re-rooted.  Only a changed root forces the full DFS.
void Forest::updateTrees()
{
    if (pool.size() > (size_t) 12 * n + 4096) compactPool();
    for (size_t k = 0; k < stale.size(); ++k) {
        const int32_t slot = stale[k];
        Tree &t = trees[slot];
        t.queued = false;
        if (t.type == TREE_FREE || t.upToDate) continue;
        const int32_t newRoot = (t.type == TREE_I) ? tail[t.extraArc] : t.root;   // keep the root on the cycle (:578)
        if (!t.rooted || newRoot != t.root) {
            ensureArcs(slot);                            // an implicit order refers to the old rooting: write it down first
            t.root = newRoot;
            reindex(slot);
        } else {
            for (size_t q = 0; q < t.pending.size(); ++q) rethreadSubtree(slot, t.pending[q]);
            t.pending.clear();
            t.threadsValid = false;
        }
        t.upToDate = true;
    }
    stale.clear();
}

} // namespace gneb200
