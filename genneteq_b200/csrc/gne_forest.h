// genneteq_b200 — host basis forest (the reference's L1 layer, trees.h / trees.cpp, re-designed
// on flat arrays).  The forest defines every order the simplex tie-breaks depend on:
//   * arcs(t)    order of Tree::arcsInTree   -> order of blockingVars (genneteq.h:334-360)
//   * threads(t) preorder of initializeTreeIndices (trees.cpp:597-644) -> order in which
//                children's supplies are summed in computeFlowsTypeI (gneFlow.cpp:498-531)
//   * incident lists (Node::incidentArcs)    -> both of the above after the next split
// so the sequence of primitive operations follows trees.cpp exactly, while the bookkeeping that
// has no observable effect does not: trees live in stable slots (no id renumbering on removal,
// trees.cpp:233-256; the part of a split tree that keeps the old root keeps the slot), only trees
// touched since the last update are re-threaded (trees.cpp:572-591 scans every tree), incident
// lists sit in one pool and carry the neighbour node, so a DFS never touches the arc arrays.
// On top of the reference's fields it records what makes the per-pivot work sparse:
//   * posOfArc(a)   index of a in arcs(tree)          -> ratio test over the few arcs with flow change
//   * threadPos     index of a node in threads(tree)  -> flow sweep over the three root paths only
//   * dirtyNodes    nodes whose root path or tree changed, in thread order -> potential sweep
// Per-node fields live in one 32-byte record and the arc position rides in the incident entry of
// the node the arc was discovered from, so a DFS touches about two cache lines per node.
#pragma once
#include <stdint.h>
#include <utility>
#include <vector>

namespace gneb200 {

enum { TREE_FREE = 0, TREE_I = 1, TREE_II = 2 };        // variables.h:25 tree_type

struct Attach { int32_t arc, parent, child; };          // a subtree hung under `parent` through `arc` since the last re-thread

struct Tree {
    int8_t type = TREE_FREE;
    bool upToDate = false;
    bool queued = false;                                // in Forest::stale
    bool rooted = false;                                // pred/predArc/depth valid w.r.t. root (except `pending` subtrees)
    bool threadsValid = false;                          // `threads` (and Node::thread / threadPos) match the structure
    std::vector<Attach> pending;
    int32_t root = -1;
    int32_t extraArc = -1;
    std::vector<int32_t> nodes;                         // Tree::nodesInTree (as a set: its order has no observable effect)
    std::vector<int32_t> arcs;                          // Tree::arcsInTree - materialised only while !implicitOrder
    std::vector<int32_t> threads;                       // Tree::threads (rebuilt lazily: Forest::ensureThreads)
    // Implicit arcsInTree order (see Forest::split): the order is "the split DFS from `root` over the incident lists as they
    // stood at the last split" for the arcs that were in the tree then, followed by `tailArcs` (arcs appended by merges
    // since, with explicit positions).  Nothing of size |tree| is touched when a subtree is cut off.
    bool implicitOrder = false;
    std::vector<int32_t> tailArcs;
    std::vector<int32_t> ordFix;                        // nodes whose incident entries' `ord` lag behind their physical index
};

// incident arc, the node at its other end, and (in the entry of the node the arc was discovered
// from / appended at; -1 in the other one) its index in arcs(tree)
// `ord`: the entry's index in its node's list at the tree's last split (the order that split's DFS used); equals the
// physical index except at the nodes listed in Tree::ordFix.
struct Inc { int32_t arc, nbr, pos, ord; };
enum { POS_NONE = -1, POS_TAIL_OTHER = -3, POS_TAIL_BASE = 1 << 30 };   // pos codes: not the holder / other end of a tail arc / holder of tail arc i

struct alignas(64) NodeRec {                            // Node's tree fields (variables.h:63-76) + bookkeeping: one cache line
    int32_t slot = -1;                                  // memberOfTreeID (stable slot)
    int32_t pred = -1, predArc = -1, depth = -1, thread = -1;
    int32_t threadPos = -1;
    int32_t nodeIdx = -1;                               // index in nodes(tree)
    int64_t incStart = 0;                               // this node's incident entries: pool[incStart .. incStart + incLen)
    int32_t incLen = 0, incCap = 0;
    uint8_t flag = 1;                                   // bit0: slot changed since last re-thread, bit1: dirty in this pass, bit2: `ord` of the entries lags (Tree::ordFix)
};

class Forest {
  public:
    void init(int n, int64_t numVars, const int32_t *tail, const int32_t *head);

    // Tree::insertIntoTrees / removeFromTrees / updateTrees (trees.h:41-44).  Return 0 or a
    // negative code where the reference prints and exit(-1)s.
    int insertArc(int32_t av);
    int removeArc(int32_t av);
    // removeArc(out) followed by insertArc(in) in one step, for the common pivot inside one type I tree: `out` is a tree arc
    // off the cycle and `in` hangs the subtree it cuts off back into the rest.  Same resulting lists, orders and rooting
    // as the two calls (split + re-insertion of the extra arc + merge), without moving the subtree's nodes to another
    // tree slot and back.  Returns false (nothing done) when the pivot is not of that shape.
    bool rehang(int32_t out, int32_t in);
    void updateTrees();
    // The preorder thread of a tree is only needed by the full flow recompute, the full potential
    // sweep and the accessors used in tests; pivots that keep a tree's root only re-thread the
    // subtrees that were hung into it, so the thread vector is rebuilt on demand.
    void ensureThreads(int32_t slot);
    // index of the incident entry of arc a in node v's list (order of v's children in the thread)
    int32_t incidentIndex(int32_t v, int32_t a) const;

    int32_t treeOf(int32_t v) const { return node[v].slot; }
    int32_t predOf(int32_t v) const { return node[v].pred; }
    int32_t predArcOf(int32_t v) const { return node[v].predArc; }
    int32_t threadPosOf(int32_t v) const { return node[v].threadPos; }
    const Tree &tree(int32_t slot) const { return trees[slot]; }
    int8_t typeOf(int32_t v) const { return node[v].slot < 0 ? (int8_t) TREE_FREE : trees[node[v].slot].type; }
    // index of basic arc a (ends u, v) in arcs(tree): scans the two short incident lists (explicit order, or a tail arc)
    int32_t posOfArc(int32_t a, int32_t u, int32_t v) const;
    // materialise arcs(tree) (and the positions in the incident entries) of a tree whose order is implicit
    void ensureArcs(int32_t slot);
    // the incident entry of arc a at node v (nullptr if absent)
    const Inc *entryOf(int32_t v, int32_t a) const;
    bool fastSplit = true;                              // GNE_FOREST_FAST=0: always the reference's full-DFS split
    int numSlots() const { return (int) trees.size(); }

    // slots whose structure or extra arc changed since the caller last cleared this list
    std::vector<int32_t> changed;
    std::vector<uint8_t> changedFlag;
    void markChanged(int32_t slot);
    // nodes whose (pred, predArc), an ancestor's, or whose slot changed in the last re-threading,
    // grouped by tree and in thread order (parents before children); cleared by the consumer
    std::vector<int32_t> dirtyNodes;
    void clearDirty() { for (size_t k = 0; k < dirtyNodes.size(); ++k) node[dirtyNodes[k]].flag &= (uint8_t) ~2; dirtyNodes.clear(); }

    std::vector<NodeRec> node;
    double secsSplit = 0, secsReindex = 0, secsMerge = 0;
    int64_t nodesSplit = 0, nodesReindex = 0, callsReindex = 0, nodesMaterialized = 0, callsSplitFast = 0, callsSplitFull = 0, callsRehang = 0;   // wall-clock split of the forest work

  private:
    int n = 0;
    const int32_t *tail = nullptr, *head = nullptr;
    std::vector<Tree> trees;
    std::vector<int32_t> freeSlots;
    std::vector<int32_t> stale;                         // slots with upToDate == false

    // incident lists (Node::incidentArcs): push_back / swap-with-last removal, pooled
    std::vector<Inc> pool;
    void incPush(int32_t v, int32_t arc, int32_t nbr, int32_t pos);
    void compactPool();

    std::vector<int32_t> stackA, stackB, stackC;        // DFS scratch
    std::vector<std::pair<int32_t, int32_t> > sortBuf;
    std::vector<int32_t> oldNodes, oldArcs;             // the vectors of a tree being split

    int32_t newTree(int8_t type, int32_t rootNode);
    void freeTree(int32_t slot);
    void relabel(int32_t slot, int32_t newSlot);
    void touch(int32_t slot);                           // upToDate = false
    void setSlot(int32_t v, int32_t slot) { NodeRec &r = node[v]; if (r.slot != slot) { r.slot = slot; r.flag |= 1; } }
    void mergeInto(int32_t tX, int32_t tII, int32_t av);
    void expandWithArc(int32_t t, int32_t av);
    int32_t split(int32_t t, int32_t removedArc);       // returns the slot of the bottom part; t keeps the top
    int32_t splitFull(int32_t t, int32_t removedArc);   // the reference's DFS over the whole tree (trees.cpp:502-566)
    void pushNode(int32_t slot, int32_t v) { node[v].nodeIdx = (int32_t) trees[slot].nodes.size(); trees[slot].nodes.push_back(v); }
    void applyOrdFix(Tree &t);
    void reindex(int32_t slot);
    void rethreadSubtree(int32_t slot, const Attach &at);
};

} // namespace gneb200
