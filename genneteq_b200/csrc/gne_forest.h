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
//   * arcPos[a]     index of a in arcs(tree)          -> ratio test over the few arcs with flow change
//   * threadPos[v]  index of v in threads(tree)       -> flow sweep over the three root paths only
//   * dirtyNodes    nodes whose root path or tree changed, in thread order -> potential sweep
#pragma once
#include <stdint.h>
#include <vector>

namespace gneb200 {

enum { TREE_FREE = 0, TREE_I = 1, TREE_II = 2 };        // variables.h:25 tree_type

struct Tree {
    int8_t type = TREE_FREE;
    bool upToDate = false;
    bool queued = false;                                // in Forest::stale
    int32_t root = -1;
    int32_t extraArc = -1;
    std::vector<int32_t> nodes;                         // Tree::nodesInTree
    std::vector<int32_t> arcs;                          // Tree::arcsInTree
    std::vector<int32_t> threads;                       // Tree::threads
};

struct Inc { int32_t arc, nbr; };                       // incident arc + the node at its other end

class Forest {
  public:
    void init(int n, int64_t numVars, const int32_t *tail, const int32_t *head);

    // Tree::insertIntoTrees / removeFromTrees / updateTrees (trees.h:41-44).  Return 0 or a
    // negative code where the reference prints and exit(-1)s.
    int insertArc(int32_t av);
    int removeArc(int32_t av);
    void updateTrees();

    int32_t treeOf(int32_t node) const { return treeSlot[node]; }
    const Tree &tree(int32_t slot) const { return trees[slot]; }
    int8_t typeOf(int32_t node) const { return treeSlot[node] < 0 ? (int8_t) TREE_FREE : trees[treeSlot[node]].type; }
    int numSlots() const { return (int) trees.size(); }

    // slots whose structure or extra arc changed since the caller last cleared this list
    std::vector<int32_t> changed;
    std::vector<uint8_t> changedFlag;
    void markChanged(int32_t slot);
    // nodes whose (pred, predArc), an ancestor's, or whose slot changed in the last re-threading,
    // grouped by tree and in thread order (parents before children); cleared by the consumer
    std::vector<int32_t> dirtyNodes;

    // per node
    std::vector<int32_t> treeSlot, pred, predArc, depth, thread, threadPos;
    // per arc variable
    std::vector<int32_t> arcPos;

  private:
    int n = 0;
    const int32_t *tail = nullptr, *head = nullptr;
    std::vector<Tree> trees;
    std::vector<int32_t> freeSlots;
    std::vector<int32_t> stale;                         // slots with upToDate == false
    std::vector<uint8_t> nodeFlag;                      // bit0: slot changed since last re-thread, bit1: dirty in this pass

    // incident lists (Node::incidentArcs): push_back / swap-with-last removal, pooled
    std::vector<Inc> pool;
    std::vector<int64_t> incStart;
    std::vector<int32_t> incLen, incCap;
    void incPush(int32_t node, int32_t arc, int32_t nbr);
    void compactPool();

    std::vector<int32_t> stackA, stackB, stackC;        // DFS scratch
    std::vector<int32_t> oldNodes, oldArcs;             // the vectors of a tree being split

    int32_t newTree(int8_t type, int32_t rootNode);
    void freeTree(int32_t slot);
    void relabel(int32_t slot, int32_t newSlot);
    void touch(int32_t slot);                           // upToDate = false
    void setSlot(int32_t node, int32_t slot) { if (treeSlot[node] != slot) { treeSlot[node] = slot; nodeFlag[node] |= 1; } }
    void mergeInto(int32_t tX, int32_t tII, int32_t av);
    void expandWithArc(int32_t t, int32_t av);
    int32_t split(int32_t t, int32_t removedArc);       // returns the slot of the bottom part; t keeps the top
    void reindex(int32_t slot);
};

} // namespace gneb200
