// genneteq_b200 — layer 2 of the C ABI (solver mirror), the sparse Graph loader for the
// reference's text format, and the synthetic-instance generator used by the benchmarks.
#include "gne_solver.h"
#include "gne_internal.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <sys/prctl.h>
#include <time.h>
#include <vector>

using namespace gneb200;

// No exception may cross the C ABI (std::length_error / bad_alloc from absurd sizes would call std::terminate in a
// C caller): every entry point that allocates runs inside guarded().
template <typename F>
static int guarded(const char *where, F f)
{
    try { return f(); }
    catch (const std::bad_alloc &) { gne_set_error("%s: out of memory (sizes too large?)", where); return GNE_ERR_ARG; }
    catch (const std::exception &e) { gne_set_error("%s: %s", where, e.what()); return GNE_ERR_ARG; }
    catch (...) { gne_set_error("%s: unknown exception", where); return GNE_ERR_ARG; }
}

// ------------------------------------------------------------------------------------------
// Graph: graph.cpp:31-53, 248-296, 358-447 read sparsely
// ------------------------------------------------------------------------------------------
int Graph::initializeFromArrays(int n, int64_t m, const int32_t *t, const int32_t *h, const double *ub,
                                const double *c, const double *mu, const double *sup, double bigMIn)
{
    if (n <= 0 || m < 0) { gne_set_error("Graph: bad sizes"); return GNE_ERR_ARG; }
    numNodes = n; numEqualFlowSets = 0;
    tail.clear(); head.clear(); cap.clear(); cost.clear(); mult.clear(); eqfOf.clear(); eqfNodeVals.clear();
    double maxCost = 0.0, maxCap = 0.0;                 // graph.cpp:36-37, 412-417
    for (int64_t a = 0; a < m; ++a) {
        if (t[a] < 0 || t[a] >= n || h[a] < 0 || h[a] >= n) { gne_set_error("Graph: arc %lld has a node out of range", (long long) a); return GNE_ERR_ARG; }
        if (a > 0 && (t[a] < t[a - 1] || (t[a] == t[a - 1] && h[a] <= h[a - 1]))) {
            gne_set_error("Graph: arcs must be sorted by (tail, head) without duplicates (arc %lld)", (long long) a);
            return GNE_ERR_ARG;
        }
        if (maxCost < c[a]) maxCost = c[a];
        if (maxCap < ub[a]) maxCap = ub[a];
        if (ub[a] <= 0.0) continue;                     // "arc doesn't exist", gneInit.cpp:84
        tail.push_back(t[a]); head.push_back(h[a]); cap.push_back(ub[a]); cost.push_back(c[a]); mult.push_back(mu[a]);
    }
    supplies.assign(sup, sup + n);
    const int numArcs = (int) m;
    bigM = (bigMIn >= 0.0) ? bigMIn : numArcs * maxCost * maxCap;      // graph.cpp:43
    return GNE_OK;
}

int Graph::initializeFromArraysEqf(int n, int64_t m, const int32_t *t, const int32_t *h, const double *ub, const double *c,
                                   const double *mu, const int32_t *eqf, int numSets, const double *sup, double bigMIn)
{
    if (numSets < 0 || (numSets > 0 && !eqf)) { gne_set_error("Graph: bad equal-flow arguments"); return GNE_ERR_ARG; }
    if (numSets > 0 && (double) numSets * (double) n > 4.0e9) { gne_set_error("Graph: %d sets x %d nodes is beyond the dense rows this path keeps", numSets, n); return GNE_ERR_ARG; }
    int rc = initializeFromArrays(n, m, t, h, ub, c, mu, sup, bigMIn);
    if (rc) return rc;
    numEqualFlowSets = numSets;
    eqfOf.clear(); eqfNodeVals.assign((size_t) numSets * (size_t) n, 0.0);
    for (int64_t a = 0; a < m; ++a) {
        const int r = numSets > 0 ? eqf[a] - 1 : -1;    // EQF_IND_OFFSET, graph.h:18
        if (r >= numSets) { gne_set_error("Graph: arc %lld names equal-flow set %d of %d", (long long) a, r + 1, numSets); return GNE_ERR_ARG; }
        if (r >= 0) {                                   // graph.cpp:403-410, in the order the arcs are given
            eqfNodeVals[(size_t) r * n + t[a]] += 1.0;
            eqfNodeVals[(size_t) r * n + h[a]] += -1.0 * mu[a];
        }
        if (ub[a] <= 0.0) continue;
        eqfOf.push_back(r < 0 ? -1 : r);
    }
    return GNE_OK;
}

void Graph::toEqfInstance(EqfInstance &out) const
{
    out.n = numNodes; out.p = numEqualFlowSets; out.bigM = bigM;
    out.tail.clear(); out.head.clear(); out.cap.clear(); out.cost.clear(); out.mult.clear();
    out.supply = supplies;
    out.numOrigArcs.assign((size_t) numEqualFlowSets, 0);
    out.setUB.assign((size_t) numEqualFlowSets, std::numeric_limits<double>::max());       // gneInit.cpp:69
    out.setCost.assign((size_t) numEqualFlowSets, 0.0);
    for (size_t a = 0; a < tail.size(); ++a) {          // (tail, head) order, arcs with capacity > 0 only: gneInit.cpp:82-105
        const int r = eqfOf.empty() ? -1 : eqfOf[a];
        if (r < 0) {
            out.tail.push_back(tail[a]); out.head.push_back(head[a]); out.cap.push_back(cap[a]); out.cost.push_back(cost[a]); out.mult.push_back(mult[a]);
        } else {
            out.setCost[(size_t) r] += cost[a];
            out.numOrigArcs[(size_t) r] += 1;
            if (out.setUB[(size_t) r] > cap[a]) out.setUB[(size_t) r] = cap[a];
        }
    }
    out.nodeVals = eqfNodeVals;
}

int Graph::initialize(Config *conf)
{
    FILE *f = fopen(conf->inFile, "r");
    if (!f) { gne_set_error("Unable to open file: %s", conf->inFile); return GNE_ERR_ARG; }
    int n = 0, numArcs = 0, numEq = 0;
    long arcsRead = 0;
    bool haveP = false;
    double maxCost = 0.0, maxCap = 0.0;
    struct Rec { double ub, cost, mult; int eq; };
    std::vector<double> nodeVals;                       // d_r(i), increments in FILE order (graph.cpp:403-410)
    std::map<std::pair<int32_t, int32_t>, Rec> arcs;    // a later line overwrites (dense matrices do, graph.cpp:406-410)
    std::vector<double> sup;
    char *line = nullptr;
    size_t lcap = 0;
    int rc = GNE_OK;
    while (getline(&line, &lcap, f) > 0) {
        if (line[0] == '\n' || line[0] == '\0') continue;
        if (line[0] == 'p') {                           // processProblemLine, graph.cpp:358-372
            char tmp[64];
            if (sscanf(line, "p %63s %d %d %d", tmp, &n, &numArcs, &numEq) != 4) { gne_set_error("Improperly formatted problem line"); rc = GNE_ERR_ARG; break; }
            if (n <= 0 || numArcs < 0 || n > (1 << 30) || numEq < 0) { gne_set_error("problem line: bad sizes (n = %d, arcs = %d, sets = %d)", n, numArcs, numEq); rc = GNE_ERR_ARG; break; }
            if ((double) numEq * (double) n > 4.0e9) { gne_set_error("problem line: %d sets x %d nodes is beyond the dense rows this path keeps", numEq, n); rc = GNE_ERR_ARG; break; }
            sup.assign(n, 0.0);
            nodeVals.assign((size_t) numEq * (size_t) n, 0.0);
            haveP = true;
        } else if (line[0] == 'a' || line[0] == 'e') {  // processArcLine, graph.cpp:382-419 ('e' lines fail its sscanf)
            int n1, n2, eq;
            double lb, ub, c, mu;
            ++arcsRead;
            if (line[0] == 'e' || sscanf(line, "a %d %d %lf %lf %lf %lf %d", &n1, &n2, &lb, &ub, &c, &mu, &eq) != 7) {
                gne_set_error("Improperly formatted arc line"); rc = GNE_ERR_ARG; break;
            }
            if (!haveP || n1 < 1 || n1 > n || n2 < 1 || n2 > n) { gne_set_error("arc line names a node out of range"); rc = GNE_ERR_ARG; break; }
            if (eq - 1 >= numEq) { gne_set_error("arc line names equal-flow set %d of %d", eq, numEq); rc = GNE_ERR_ARG; break; }
            Rec r; r.ub = ub; r.cost = c; r.mult = mu;  // the lower bound is parsed and dropped (graph.cpp:386-410)
            r.eq = eq - 1 >= 0 ? eq - 1 : -1;           // EQF_IND_OFFSET
            if (r.eq >= 0) {
                nodeVals[(size_t) r.eq * n + (n1 - 1)] += 1.0;
                nodeVals[(size_t) r.eq * n + (n2 - 1)] += -1.0 * mu;
            }
            arcs[std::make_pair((int32_t) (n1 - 1), (int32_t) (n2 - 1))] = r;
            if (maxCost < c) maxCost = c;
            if (maxCap < ub) maxCap = ub;
        } else if (line[0] == 'n') {                    // processNodeLine, graph.cpp:430-447
            int id;
            double s;
            if (sscanf(line, "n %d %lf", &id, &s) != 2) { gne_set_error("Improperly formatted node line"); rc = GNE_ERR_ARG; break; }
            if (!haveP || id < 1 || id > n) { gne_set_error("node line out of range"); rc = GNE_ERR_ARG; break; }
            sup[id - 1] = s;
        }                                               // anything else is a comment
    }
    free(line);
    fclose(f);
    if (rc) return rc;
    if (!haveP) { gne_set_error("no problem line"); return GNE_ERR_ARG; }
    if (arcsRead != numArcs) { gne_set_error("Number of arcs specified does not match the number of arcs in the file"); return GNE_ERR_ARG; }
    numNodes = n; numEqualFlowSets = numEq;
    tail.clear(); head.clear(); cap.clear(); cost.clear(); mult.clear(); eqfOf.clear();
    for (std::map<std::pair<int32_t, int32_t>, Rec>::const_iterator it = arcs.begin(); it != arcs.end(); ++it) {
        if (it->second.ub <= 0.0) continue;
        tail.push_back(it->first.first); head.push_back(it->first.second);
        cap.push_back(it->second.ub); cost.push_back(it->second.cost); mult.push_back(it->second.mult);
        eqfOf.push_back(it->second.eq);
    }
    eqfNodeVals.swap(nodeVals);
    supplies = sup;
    bigM = numArcs * maxCost * maxCap;                  // graph.cpp:43
    return GNE_OK;
}

// ------------------------------------------------------------------------------------------
// instance files: the reference's text format, written the way
// Graph::exportColFile does (graph.cpp:94-121: comment line, `p efpgn n m 0`, `a i j 0.0 UB cost mult 0`
// with %lf = 6 decimals, one `n` line per node), and a binary structure-of-arrays form for the sizes
// a text file is impractical for (header + the six columns, arcs sorted by (tail, head)).
// ------------------------------------------------------------------------------------------
extern "C" int gne_write_col(const char *path, int n, int64_t m, const int32_t *tail, const int32_t *head, const double *ub,
                             const double *cost, const double *mult, const double *supply)
{
    FILE *f = fopen(path, "w");
    if (!f) { gne_set_error("Unable to open file %s", path); return GNE_ERR_ARG; }
    fprintf(f, "c   n1  n2      LB  UB  COST    MULT    EQFLW\n");
    fprintf(f, "p efpgn %d %lld %d\n", n, (long long) m, 0);
    for (int64_t a = 0; a < m; ++a) fprintf(f, "a %d %d %lf %lf %lf %lf %d\n", tail[a] + 1, head[a] + 1, 0.0, ub[a], cost[a], mult[a], 0);
    for (int i = 0; i < n; ++i) fprintf(f, "n %d %lf\n", i + 1, supply[i]);
    fclose(f);
    return GNE_OK;
}

static const char GNE_BIN_MAGIC[8] = { 'G', 'N', 'E', 'S', 'O', 'A', '1', 0 };

extern "C" int gne_write_bin(const char *path, int n, int64_t m, const int32_t *tail, const int32_t *head, const double *ub,
                             const double *cost, const double *mult, const double *supply)
{
    FILE *f = fopen(path, "wb");
    if (!f) { gne_set_error("Unable to open file %s", path); return GNE_ERR_ARG; }
    int64_t hdr[2] = { (int64_t) n, m };
    bool ok = fwrite(GNE_BIN_MAGIC, 1, 8, f) == 8 && fwrite(hdr, 8, 2, f) == 2;
    ok = ok && fwrite(tail, 4, (size_t) m, f) == (size_t) m && fwrite(head, 4, (size_t) m, f) == (size_t) m;
    ok = ok && fwrite(ub, 8, (size_t) m, f) == (size_t) m && fwrite(cost, 8, (size_t) m, f) == (size_t) m;
    ok = ok && fwrite(mult, 8, (size_t) m, f) == (size_t) m && fwrite(supply, 8, (size_t) n, f) == (size_t) n;
    fclose(f);
    if (!ok) { gne_set_error("short write to %s", path); return GNE_ERR_ARG; }
    return GNE_OK;
}

extern "C" int gne_read_bin_header(const char *path, int *n, int64_t *m)
{
    FILE *f = fopen(path, "rb");
    if (!f) { gne_set_error("Unable to open file: %s", path); return GNE_ERR_ARG; }
    char magic[8];
    int64_t hdr[2];
    const bool ok = fread(magic, 1, 8, f) == 8 && !memcmp(magic, GNE_BIN_MAGIC, 8) && fread(hdr, 8, 2, f) == 2;
    int64_t fileBytes = -1;
    if (fseek(f, 0, SEEK_END) == 0) fileBytes = (int64_t) ftell(f);
    fclose(f);
    if (!ok) { gne_set_error("%s is not a genneteq_b200 binary instance", path); return GNE_ERR_ARG; }
    if (hdr[0] <= 0 || hdr[0] > (1 << 30) || hdr[1] < 0 || fileBytes != 24 + 32 * hdr[1] + 8 * hdr[0]) {
        gne_set_error("%s: header sizes (n = %lld, m = %lld) do not match the file length", path, (long long) hdr[0], (long long) hdr[1]);
        return GNE_ERR_ARG;
    }
    *n = (int) hdr[0]; *m = hdr[1];
    return GNE_OK;
}

extern "C" int gne_read_bin(const char *path, int n, int64_t m, int32_t *tail, int32_t *head, double *ub, double *cost,
                            double *mult, double *supply)
{
    int n2; int64_t m2;
    int rc = gne_read_bin_header(path, &n2, &m2); if (rc) return rc;
    if (n2 != n || m2 != m) { gne_set_error("gne_read_bin: sizes differ from the header"); return GNE_ERR_ARG; }
    FILE *f = fopen(path, "rb");
    if (!f) { gne_set_error("Unable to open file: %s", path); return GNE_ERR_ARG; }
    fseek(f, 24, SEEK_SET);
    bool ok = fread(tail, 4, (size_t) m, f) == (size_t) m && fread(head, 4, (size_t) m, f) == (size_t) m;
    ok = ok && fread(ub, 8, (size_t) m, f) == (size_t) m && fread(cost, 8, (size_t) m, f) == (size_t) m;
    ok = ok && fread(mult, 8, (size_t) m, f) == (size_t) m && fread(supply, 8, (size_t) n, f) == (size_t) n;
    fclose(f);
    if (!ok) { gne_set_error("short read from %s", path); return GNE_ERR_ARG; }
    return GNE_OK;
}

// the text format parsed sparsely into caller-sized arrays (two calls: sizes, then data)
extern "C" int gne_read_col(const char *path, int *n, int64_t *m, int32_t *tail, int32_t *head, double *ub, double *cost,
                            double *mult, double *supply, double *bigM)
{
    return guarded("gne_read_col", [&]() -> int {
    Config conf;
    conf.inFile = path;
    Graph g;
    int rc = g.initialize(&conf);
    if (rc) return rc;
    if (g.getNumEqualFlowSets() > 0) { gne_set_error("gne_read_col: the file has equal-flow sets (use gne_read_col_eqf)"); return GNE_ERR_UNSUPPORTED; }
    const int64_t cnt = g.getNumArcs();
    if (tail) {
        if (*m < cnt || *n < g.getNumNodes()) { gne_set_error("gne_read_col: arrays too small"); return GNE_ERR_CAPACITY; }
        for (int64_t a = 0; a < cnt; ++a) { tail[a] = g.tail[a]; head[a] = g.head[a]; ub[a] = g.cap[a]; cost[a] = g.cost[a]; mult[a] = g.mult[a]; }
        for (int i = 0; i < g.getNumNodes(); ++i) supply[i] = g.supply(i);
    }
    *n = g.getNumNodes(); *m = cnt;
    if (bigM) *bigM = g.getBigM();
    return GNE_OK;
    });
}

extern "C" int gne_read_col_eqf(const char *path, int *n, int64_t *m, int *numSets, int32_t *tail, int32_t *head, double *ub,
                                double *cost, double *mult, int32_t *eqf, double *supply, double *bigM)
{
    return guarded("gne_read_col_eqf", [&]() -> int {
    if (!n || !m || !numSets) { gne_set_error("gne_read_col_eqf: bad argument"); return GNE_ERR_ARG; }
    Config conf;
    conf.inFile = path;
    Graph g;
    int rc = g.initialize(&conf);
    if (rc) return rc;
    const int64_t cnt = g.getNumArcs();
    if (tail) {
        if (*m < cnt || *n < g.getNumNodes()) { gne_set_error("gne_read_col_eqf: arrays too small"); return GNE_ERR_CAPACITY; }
        for (int64_t a = 0; a < cnt; ++a) {
            tail[a] = g.tail[a]; head[a] = g.head[a]; ub[a] = g.cap[a]; cost[a] = g.cost[a]; mult[a] = g.mult[a];
            if (eqf) eqf[a] = g.eqfOf.empty() ? 0 : g.eqfOf[(size_t) a] + 1;
        }
        for (int i = 0; i < g.getNumNodes(); ++i) supply[i] = g.supply(i);
    }
    *n = g.getNumNodes(); *m = cnt; *numSets = g.getNumEqualFlowSets();
    if (bigM) *bigM = g.getBigM();
    return GNE_OK;
    });
}

// ------------------------------------------------------------------------------------------
// layer 2 C ABI
// ------------------------------------------------------------------------------------------
struct gne_solver {
    GenNetEq solver;
    EqfSolver *eq = nullptr;                            // instances with equal-flow sets run in this engine instead (gne_eqf.h)
    ~gne_solver() { delete eq; }
};
#define NO_EQF(s, what) do { if ((s)->eq) { gne_set_error(what ": not available for instances with equal-flow sets"); return GNE_ERR_UNSUPPORTED; } } while (0)

static int make_solver(gne_solver **out, Graph &g, int initType, int cudaDevice)
{
    gne_solver *s = new gne_solver();
    int rc;
    if (g.getNumEqualFlowSets() > 0) {
        EqfInstance inst;
        g.toEqfInstance(inst);
        s->eq = new EqfSolver();
        rc = s->eq->initialize(inst, cudaDevice);
    } else {
        rc = s->solver.initialize(&g, initType, cudaDevice);
    }
    if (rc) { delete s; return rc; }
    *out = s;
    return GNE_OK;
}

extern "C" int gne_solver_create(gne_solver **out, int cudaDevice, int n, int64_t m, const int32_t *tail, const int32_t *head,
                                 const double *ub, const double *cost, const double *mult, const double *supply, double bigM)
{
    return guarded("gne_solver_create", [&]() -> int {
    if (!out) return GNE_ERR_ARG;
    if (gne_cuda_device_count() <= 0) { gne_set_error("no CUDA device: genneteq_b200 has no CPU fallback"); return GNE_ERR_CUDA; }
    Graph g;
    int rc = g.initializeFromArrays(n, m, tail, head, ub, cost, mult, supply, bigM);
    if (rc) return rc;
    return make_solver(out, g, 0, cudaDevice);
    });
}

extern "C" int gne_solver_create_eqf(gne_solver **out, int cudaDevice, int n, int64_t m, const int32_t *tail, const int32_t *head,
                                     const double *ub, const double *cost, const double *mult, const int32_t *eqf, int numSets,
                                     const double *supply, double bigM)
{
    return guarded("gne_solver_create_eqf", [&]() -> int {
    if (!out) return GNE_ERR_ARG;
    if (gne_cuda_device_count() <= 0) { gne_set_error("no CUDA device: genneteq_b200 has no CPU fallback"); return GNE_ERR_CUDA; }
    Graph g;
    int rc = g.initializeFromArraysEqf(n, m, tail, head, ub, cost, mult, eqf, numSets, supply, bigM);
    if (rc) return rc;
    return make_solver(out, g, 0, cudaDevice);
    });
}

extern "C" int gne_solver_create_from_col(gne_solver **out, int cudaDevice, const char *path)
{
    return guarded("gne_solver_create_from_col", [&]() -> int {
    if (!out || !path) return GNE_ERR_ARG;
    if (gne_cuda_device_count() <= 0) { gne_set_error("no CUDA device: genneteq_b200 has no CPU fallback"); return GNE_ERR_CUDA; }
    Config conf;
    conf.inFile = path;
    Graph g;
    int rc = g.initialize(&conf);
    if (rc) return rc;
    return make_solver(out, g, conf.initType, cudaDevice);
    });
}

extern "C" int gne_shard_range(int64_t m, int rank, int nranks, int64_t *lo, int64_t *hi)
{
    if (m < 0 || nranks < 1 || rank < 0 || rank >= nranks || !lo || !hi) { gne_set_error("gne_shard_range: bad argument"); return GNE_ERR_ARG; }
    GenNetEq::shardRange(m, rank, nranks, lo, hi);
    return GNE_OK;
}

extern "C" int gne_solver_destroy(gne_solver *s) { delete s; return GNE_OK; }
extern "C" int gne_solver_set_rules(gne_solver *s, int p1, int p2)
{
    if (s->eq) { s->eq->phaseIpivot = p1; s->eq->phaseIIpivot = p2; return GNE_OK; }
    s->solver.phaseIpivot = p1; s->solver.phaseIIpivot = p2; return GNE_OK;
}
// One host thread per group of solvers; a thread keeps stepping the solvers of its group round-robin and never
// blocks on the device while another of its solvers has host work to do.  With perLaunch > 0 the LV passes a thread has
// prepared during one sweep over its solvers leave in launch groups (gne_lv_group: one grid for up to perLaunch
// instances) instead of one launch each: the per-instance launch cost is what limits a batch of small instances.
// alignPivot >= 0: no solver starts its pivot number alignPivot before every solver of the batch has got there (or ended):
// the instances then run that part of their solves side by side whatever their first pivots cost (bench.py's window).
static int solve_batch_impl(gne_solver **s, int B, int presolve, int64_t maxPivots, int threads, int perLaunch,
                            int64_t *iters, int *finished, int64_t *groupLaunches, int64_t alignPivot = -1)
{
    if (B <= 0 || !s || !iters || !finished || perLaunch < 0) { gne_set_error("gne_solver_solve_batch: bad argument"); return GNE_ERR_ARG; }
    for (int i = 0; i < B; ++i) NO_EQF(s[i], "gne_solver_solve_batch");
    int T = threads > 0 ? threads : (int) std::thread::hardware_concurrency();
    if (T <= 0) T = 1;
    if (T > B) T = B;
    std::vector<int> status((size_t) B, GNE_OK);
    std::vector<std::string> errors((size_t) B);
    std::vector<int64_t> launchesOf((size_t) T, 0);
    std::atomic<int> arrived(0);                        // solvers that have reached alignPivot (or ended before it)
    std::vector<char> hasArrived((size_t) B, 0);
    auto work = [&](int w) {
        std::vector<int> mine;
        for (int i = w; i < B; i += T) mine.push_back(i);
        std::vector<char> live(mine.size(), 1);
        size_t alive = 0;
        gne_lv_group *group = nullptr;
        for (size_t k = 0; k < mine.size(); ++k) {
            const int i = mine[k];
            iters[i] = 0; finished[i] = 1;
            status[i] = s[i]->solver.solveBegin(presolve != 0, maxPivots);
            if (status[i] != GNE_OK) { errors[i] = gne_last_error(); live[k] = 0; } else ++alive;
            if (!live[k] && alignPivot >= 0) { hasArrived[(size_t) i] = 1; arrived.fetch_add(1); }
        }
        // (the device handles exist once solveBegin has run: now the thread's launch group can be made on their device)
        if (perLaunch > 0) {
            for (size_t k = 0; k < mine.size() && !group; ++k) {
                gne_dev *d0 = live[k] ? s[mine[k]]->solver.device() : nullptr;
                if (d0 && gne_lv_group_create(&group, gne_dev_cuda_device(d0), perLaunch) != GNE_OK) {
                    const std::string why = gne_last_error();
                    for (size_t q = 0; q < mine.size(); ++q) if (live[q]) { status[mine[q]] = GNE_ERR_CUDA; errors[mine[q]] = why; live[q] = 0; }
                    alive = 0; group = nullptr;
                    break;
                }
            }
            for (size_t k = 0; k < mine.size(); ++k) s[mine[k]]->solver.lvGroup = group;
        }
        prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0);   // (the default 50 us slack would stretch the short sleeps below)
        bool groupFailed = false;
        // GNE_BATCH_STATS=1: where this thread's time goes (stderr, once per thread at the end)
        static const bool wantStats = getenv("GNE_BATCH_STATS") != nullptr;
        auto nowNs = []() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (int64_t) t.tv_sec * 1000000000ll + t.tv_nsec; };
        int64_t nsStep1 = 0, nsStep2 = 0, nsLaunch = 0, nsSleep = 0, nsIdleSweep = 0, nStep1 = 0, nStep2 = 0, nLaunch = 0, nSleep = 0, nIdle = 0;
        const int64_t tStart = wantStats ? nowNs() : 0;
        while (alive > 0) {
            bool progressed = false;
            const int64_t tSweep = wantStats ? nowNs() : 0;
            for (size_t k = 0; k < mine.size(); ++k) {
                if (!live[k]) continue;
                const int i = mine[k];
                if (alignPivot >= 0 && s[i]->solver.atLoopHead() && s[i]->solver.pivotsThisCall() == alignPivot) {
                    if (!hasArrived[(size_t) i]) { hasArrived[(size_t) i] = 1; arrived.fetch_add(1); }
                    if (arrived.load() < B) continue;   // (held at the line until the whole batch stands there)
                }
                // a solver whose result has arrived is taken through its pivot AND the preparation of its next pass
                // (two steps) before the next solver is looked at; with a single live solver left there is nothing to
                // interleave with: block in the step
                const int64_t t0 = wantStats ? nowNs() : 0;
                int r = s[i]->solver.solveStep(alive == 1);
                if (r == 2) continue;
                progressed = true;
                const int64_t t1 = wantStats ? nowNs() : 0;
                // (the pivot is done: the preparation of the next pass follows at once - unless that pass is the one behind the
                //  starting line and the rest of the batch is not there yet; the check at the top of the sweep lets it go later)
                if (r == 1 && alignPivot >= 0 && s[i]->solver.atLoopHead() && s[i]->solver.pivotsThisCall() == alignPivot) {
                    if (!hasArrived[(size_t) i]) { hasArrived[(size_t) i] = 1; arrived.fetch_add(1); }
                    if (arrived.load() < B) continue;
                }
                if (r == 1) r = s[i]->solver.solveStep(alive == 1);
                if (wantStats) { const int64_t t2 = nowNs(); nsStep1 += t1 - t0; nsStep2 += t2 - t1; ++nStep1; ++nStep2; }
                if (r == 1 || r == 2) continue;
                if (r < 0) { status[i] = -r; errors[i] = gne_last_error(); }
                else { status[i] = s[i]->solver.solveEnd(&iters[i], &finished[i]); if (status[i] != GNE_OK) errors[i] = gne_last_error(); }
                live[k] = 0; --alive;
                if (alignPivot >= 0 && !hasArrived[(size_t) i]) { hasArrived[(size_t) i] = 1; arrived.fetch_add(1); }
            }
            // whatever has been parked and not sent yet goes out now
            const int64_t tL = wantStats ? nowNs() : 0;
            if (wantStats && !progressed) { nsIdleSweep += tL - tSweep; ++nIdle; }
            // (only when the sweep found nothing else to do: while solvers keep coming in, their passes fill the group - a
            //  launch of two instances occupies the device about as long as a launch of eight)
            if (group && gne_lv_group_pending(group) > 0 && !progressed) {
                if (gne_lv_group_launch(group) != GNE_OK) { groupFailed = true; break; }
                if (wantStats) { nsLaunch += nowNs() - tL; ++nLaunch; }
            }
            // every pass of this thread's solvers is still running: leave the core to the other threads for a moment
            // (a thread that spins through its quantum is descheduled for milliseconds as soon as anything else
            // wants a core, and its instances stall with it)
            if (!progressed) {
                const int64_t tS = wantStats ? nowNs() : 0;
                struct timespec ts = { 0, 2000 }; nanosleep(&ts, nullptr);
                if (wantStats) { nsSleep += nowNs() - tS; ++nSleep; }
            }
        }
        if (wantStats) {
            double sc = 0, sp = 0, so = 0;
            for (size_t k = 0; k < mine.size(); ++k) { sc += s[mine[k]]->solver.secsComplete; sp += s[mine[k]]->solver.pivotSeconds(); so += s[mine[k]]->solver.potentialSeconds(); }
            fprintf(stderr, "[batch thread %d] inside its solvers: complete %.3f ms, pivot %.3f ms, potentials %.3f ms\n", w, 1e3 * sc, 1e3 * sp, 1e3 * so);
        }
        if (wantStats)
            fprintf(stderr, "[batch thread %d] %zu solvers, %.3f ms total | first steps (complete+pivot) %lld x %.2f us | second steps (prep+park) %lld x %.2f us | "
                    "explicit launches %lld x %.2f us | idle sweeps %lld x %.2f us | sleeps %lld x %.2f us\n", w, mine.size(), 1e-6 * (nowNs() - tStart),
                    (long long) nStep1, nStep1 ? 1e-3 * nsStep1 / nStep1 : 0.0, (long long) nStep2, nStep2 ? 1e-3 * nsStep2 / nStep2 : 0.0,
                    (long long) nLaunch, nLaunch ? 1e-3 * nsLaunch / nLaunch : 0.0, (long long) nIdle, nIdle ? 1e-3 * nsIdleSweep / nIdle : 0.0,
                    (long long) nSleep, nSleep ? 1e-3 * nsSleep / nSleep : 0.0);
        if (groupFailed) {
            const std::string why = gne_last_error();
            for (size_t k = 0; k < mine.size(); ++k) if (live[k]) { status[mine[k]] = GNE_ERR_CUDA; errors[mine[k]] = why; }
        }
        for (size_t k = 0; k < mine.size(); ++k) s[mine[k]]->solver.lvGroup = nullptr;
        if (group) { launchesOf[(size_t) w] = gne_lv_group_launch_count(group); gne_lv_group_destroy(group); }
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < T; ++w) pool.emplace_back(work, w);
    work(0);
    for (size_t k = 0; k < pool.size(); ++k) pool[k].join();
    if (groupLaunches) { *groupLaunches = 0; for (int w = 0; w < T; ++w) *groupLaunches += launchesOf[(size_t) w]; }
    for (int i = 0; i < B; ++i) if (status[i] != GNE_OK) { gne_set_error("solver %d of the batch: %s", i, errors[i].c_str()); return status[i]; }
    return GNE_OK;
}
extern "C" int gne_solver_solve_batch(gne_solver **s, int B, int presolve, int64_t maxPivots, int threads, int64_t *iters, int *finished) {
    return guarded("gne_solver_solve_batch", [&]() -> int { return solve_batch_impl(s, B, presolve, maxPivots, threads, 0, iters, finished, nullptr); });
}
extern "C" int gne_solver_solve_batch_grouped(gne_solver **s, int B, int presolve, int64_t maxPivots, int threads, int perLaunch,
                                              int64_t *iters, int *finished, int64_t *groupLaunches) {
    return guarded("gne_solver_solve_batch_grouped", [&]() -> int { return solve_batch_impl(s, B, presolve, maxPivots, threads, perLaunch, iters, finished, groupLaunches); });
}
extern "C" int gne_solver_solve_batch_ex(gne_solver **s, int B, int presolve, int64_t maxPivots, int threads, int perLaunch, int64_t alignPivot,
                                         int64_t *iters, int *finished, int64_t *groupLaunches) {
    return guarded("gne_solver_solve_batch_ex", [&]() -> int { return solve_batch_impl(s, B, presolve, maxPivots, threads, perLaunch, iters, finished, groupLaunches, alignPivot); });
}
extern "C" int gne_solver_set_comm_host(gne_solver *s, int rank, int nranks, gne_allgather_fn allgather, void *ctx) {
    return guarded("gne_solver_set_comm_host", [&]() -> int {
        if (!allgather) { gne_set_error("gne_solver_set_comm_host: no all-gather callback"); return GNE_ERR_ARG; }
        uint8_t none[128] = { 0 };
        if (s->eq) return s->eq->setComm(none, rank, nranks, allgather, ctx);
        return s->solver.setComm(none, rank, nranks, allgather, ctx); });
}
extern "C" int gne_solver_set_comm(gne_solver *s, const uint8_t id[128], int rank, int nranks) {
    return guarded("gne_solver_set_comm", [&]() -> int { if (s->eq) return s->eq->setComm(id, rank, nranks, nullptr, nullptr); return s->solver.setComm(id, rank, nranks);     });
}
extern "C" int gne_solver_set_trace(gne_solver *s, int32_t *e, int32_t *l, int64_t cap)
{
    if (s->eq) s->eq->setTrace(e, l, cap); else s->solver.setTrace(e, l, cap);
    return GNE_OK;
}
extern "C" int64_t gne_solver_trace_len(const gne_solver *s) { return s->eq ? s->eq->getTraceLen() : s->solver.getTraceLen(); }
extern "C" int gne_solver_solve(gne_solver *s, int presolve, int64_t maxPivots, int64_t *iters, int *finished)
{
    return guarded("gne_solver_solve", [&]() -> int {
    int64_t it = 0;
    int fin = 0;
    int rc = s->eq ? s->eq->solve(presolve != 0, maxPivots, &it, &fin) : s->solver.solve(presolve != 0, maxPivots, &it, &fin);
    if (iters) *iters = it;
    if (finished) *finished = fin;
    return rc;
    });
}
extern "C" int gne_solver_num_nodes(const gne_solver *s) { return s->eq ? s->eq->getNumNodes() : s->solver.getNumNodes(); }
extern "C" int64_t gne_solver_num_vars(const gne_solver *s) { return s->eq ? s->eq->getNumVars() : s->solver.getNumVars(); }
extern "C" int gne_solver_num_sets(const gne_solver *s) { return s->eq ? s->eq->getNumSets() : 0; }
extern "C" double gne_solver_objective(const gne_solver *s) { return s->eq ? s->eq->getFlowCost() : s->solver.getFlowCost(); }
extern "C" int gne_solver_feasible(const gne_solver *s) { return (s->eq ? s->eq->getIsInfeasible() : s->solver.getIsInfeasible()) ? 0 : 1; }
extern "C" int gne_solver_get_flows(gne_solver *s, double *out) { if (s->eq) s->eq->getFlows(out); else s->solver.getFlows(out); return GNE_OK; }
extern "C" int gne_solver_get_potentials(gne_solver *s, double *out) {
    if (s->eq) { s->eq->getPotentials(out); return GNE_OK; }
    return guarded("gne_solver_get_potentials", [&]() -> int { return s->solver.getPotentials(out);     });
}
extern "C" int gne_solver_compute_potentials(gne_solver *s)
{
    NO_EQF(s, "gne_solver_compute_potentials");
    int rc = s->solver.refreshPotentials();
    if (rc) return rc;
    return gne_dev_sync(s->solver.device());
}
extern "C" int gne_solver_get_states(gne_solver *s, int32_t *out) { if (s->eq) s->eq->getStates(out); else s->solver.getStates(out); return GNE_OK; }
extern "C" int gne_solver_get_forest(gne_solver *s, int32_t *tree, int32_t *pred, int32_t *predArc, int32_t *depth, int32_t *thread)
{
    NO_EQF(s, "gne_solver_get_forest");
    Forest &f = s->solver.getForestMutable();
    const int n = s->solver.getNumNodes();
    for (int sl = 0; sl < f.numSlots(); ++sl) f.ensureThreads(sl);
    for (int i = 0; i < n; ++i) { tree[i] = f.node[i].slot; pred[i] = f.node[i].pred; predArc[i] = f.node[i].predArc; depth[i] = f.node[i].depth; thread[i] = f.node[i].thread; }
    return GNE_OK;
}
extern "C" int gne_solver_get_counters(const gne_solver *s, double *out12) { if (s->eq) s->eq->getCounters(out12); else s->solver.getCounters(out12); return GNE_OK; }
extern "C" gne_dev *gne_solver_device(gne_solver *s) { return s->eq ? s->eq->device() : s->solver.device(); }
extern "C" int gne_solver_set_window(gne_solver *s, int64_t startPivot, int64_t numPivots) { NO_EQF(s, "gne_solver_set_window"); s->solver.setWindow(startPivot, numPivots); return GNE_OK; }
extern "C" int gne_solver_set_progress(gne_solver *s, int64_t every, double *buf, int64_t capRows) { NO_EQF(s, "gne_solver_set_progress"); s->solver.setProgress(every, buf, capRows); return GNE_OK; }
extern "C" int64_t gne_solver_progress_rows(const gne_solver *s) { return s->solver.getProgressRows(); }
extern "C" int gne_solver_get_window(const gne_solver *s, double *out27) { s->solver.getWindow(out27); return GNE_OK; }

// ------------------------------------------------------------------------------------------
// host-only forest replay (no device): the basis exchanges of a given pivot trace
// ------------------------------------------------------------------------------------------
extern "C" int gne_forest_replay(int n, int64_t m, const int32_t *tailIn, const int32_t *headIn, const double *supply,
                                 int64_t pivots, const int32_t *e, const int32_t *l,
                                 int32_t *tree, int32_t *pred, int32_t *predArc, int32_t *depth, int32_t *thread)
{
    return guarded("gne_forest_replay", [&]() -> int {
    (void) supply;
    const int64_t M = m + n;
    std::vector<int32_t> tail((size_t) M), head((size_t) M);
    for (int64_t a = 0; a < m; ++a) { tail[a] = tailIn[a]; head[a] = headIn[a]; }
    for (int i = 0; i < n; ++i) { tail[m + i] = i; head[m + i] = i; }
    Forest f;
    f.init(n, M, tail.data(), head.data());
    std::vector<uint8_t> basic((size_t) M, 0);
    for (int i = 0; i < n; ++i) { if (f.insertArc((int32_t) (m + i))) return GNE_ERR_TREE; basic[m + i] = 1; }
    f.updateTrees();
    for (int64_t k = 0; k < pivots; ++k) {
        if (e[k] == l[k]) continue;                     // bound flip: the forest is untouched (gnePivot.cpp:256-267)
        if (!basic[l[k]] || basic[e[k]]) { gne_set_error("gne_forest_replay: pivot %lld is not a valid exchange", (long long) k); return GNE_ERR_ARG; }
        int rc = f.removeArc(l[k]);
        if (!rc) rc = f.insertArc(e[k]);
        if (rc) { gne_set_error("gne_forest_replay: forest error %d at pivot %lld", rc, (long long) k); return GNE_ERR_TREE; }
        basic[l[k]] = 0; basic[e[k]] = 1;
        f.updateTrees();
    }
    for (int sl = 0; sl < f.numSlots(); ++sl) f.ensureThreads(sl);
    for (int i = 0; i < n; ++i) { tree[i] = f.node[i].slot; pred[i] = f.node[i].pred; predArc[i] = f.node[i].predArc; depth[i] = f.node[i].depth; thread[i] = f.node[i].thread; }
    return GNE_OK;
    });
}

// host-only replay of a whole recorded solve: flows, ratio test, leaving arc, basis update and
// the root-relative potential sweep run exactly as in gne_solver_solve, without any device
extern "C" int gne_host_replay(int n, int64_t m, const int32_t *tail, const int32_t *head, const double *ub, const double *cost,
                               const double *mult, const double *supply, double bigM, int drawsPerPricing, int sparseFlows,
                               int64_t phase1Pivots, int64_t totalPivots, const int32_t *e, const int32_t *l,
                               int64_t *mismatch, double *flowsOut, double *potentialsOut, double *objective)
{
    return guarded("gne_host_replay", [&]() -> int {
    Graph g;
    int rc = g.initializeFromArrays(n, m, tail, head, ub, cost, mult, supply, bigM);
    if (rc) return rc;
    GenNetEq s;
    rc = s.initialize(&g, 0, 0);
    if (rc) return rc;
    s.setSparseFlows(sparseFlows != 0);
    rc = s.replayTrace(drawsPerPricing, phase1Pivots, totalPivots, e, l, mismatch);
    if (rc) return rc;
    if (flowsOut) s.getFlows(flowsOut);
    if (potentialsOut) s.getHostPotentials(potentialsOut);
    if (objective) *objective = s.getFlowCost();
    return GNE_OK;
    });
}

extern "C" int gne_host_lu_solve(int s, const double *a, const double *b, double *x, double *lu, int32_t *perm)
{
    return guarded("gne_host_lu_solve", [&]() -> int {
    if (s <= 0 || !a || !b || !x) { gne_set_error("gne_host_lu_solve: bad argument"); return GNE_ERR_ARG; }
    std::vector<double> m(a, a + (size_t) s * s), rhs(b, b + s), sol;
    std::vector<int> p;
    EqfSolver::luSolve(m, s, rhs, p, sol);
    for (int i = 0; i < s; ++i) { x[i] = sol[(size_t) i]; if (perm) perm[i] = p[(size_t) i]; }
    if (lu) for (size_t i = 0; i < (size_t) s * s; ++i) lu[i] = m[i];
    return GNE_OK;
    });
}

extern "C" int gne_host_replay_eqf(int n, int64_t m, const int32_t *tail, const int32_t *head, const double *ub, const double *cost,
                                   const double *mult, const int32_t *eqf, int numSets, const double *supply, double bigM,
                                   int drawsPhaseI, int drawsPhaseII, int64_t phase1Pivots, int64_t totalPivots, const int32_t *e, const int32_t *l,
                                   int64_t *mismatch, double *flowsOut, double *potentialsOut, double *objective)
{
    return guarded("gne_host_replay_eqf", [&]() -> int {
    Graph g;
    int rc = g.initializeFromArraysEqf(n, m, tail, head, ub, cost, mult, eqf, numSets, supply, bigM);
    if (rc) return rc;
    if (numSets <= 0) { gne_set_error("gne_host_replay_eqf: no equal-flow sets (use gne_host_replay)"); return GNE_ERR_ARG; }
    EqfInstance inst;
    g.toEqfInstance(inst);
    EqfSolver s;
    rc = s.initialize(inst, 0, true);
    if (rc) return rc;
    rc = s.replayTrace(drawsPhaseI, drawsPhaseII, phase1Pivots, totalPivots, e, l, mismatch);
    if (rc) return rc;
    if (flowsOut) s.getFlows(flowsOut);
    if (potentialsOut) s.getPotentials(potentialsOut);
    if (objective) *objective = s.getFlowCost();
    return GNE_OK;
    });
}

extern "C" int gne_host_replay_eqf_col(const char *path, int drawsPhaseI, int drawsPhaseII, int64_t phase1Pivots, int64_t totalPivots,
                                       const int32_t *e, const int32_t *l, int64_t *mismatch, int64_t numVars, double *flowsOut,
                                       double *potentialsOut, double *objective)
{
    return guarded("gne_host_replay_eqf_col", [&]() -> int {
    if (!path || !mismatch) { gne_set_error("gne_host_replay_eqf_col: bad argument"); return GNE_ERR_ARG; }
    Config conf;
    conf.inFile = path;
    Graph g;
    int rc = g.initialize(&conf);
    if (rc) return rc;
    if (g.getNumEqualFlowSets() <= 0) { gne_set_error("gne_host_replay_eqf_col: the file has no equal-flow sets"); return GNE_ERR_ARG; }
    EqfInstance inst;
    g.toEqfInstance(inst);
    EqfSolver s;
    rc = s.initialize(inst, 0, true);
    if (rc) return rc;
    if (flowsOut && numVars != s.getNumVars()) { gne_set_error("gne_host_replay_eqf_col: %lld variables, buffer for %lld", (long long) s.getNumVars(), (long long) numVars); return GNE_ERR_CAPACITY; }
    rc = s.replayTrace(drawsPhaseI, drawsPhaseII, phase1Pivots, totalPivots, e, l, mismatch);
    if (rc) return rc;
    if (flowsOut) s.getFlows(flowsOut);
    if (potentialsOut) s.getPotentials(potentialsOut);
    if (objective) *objective = s.getFlowCost();
    return GNE_OK;
    });
}

// ------------------------------------------------------------------------------------------
// synthetic instances (gne_generate.cpp; the same file also builds the reference arm's stand-alone generator library)
// ------------------------------------------------------------------------------------------
extern "C" int gne_generate_instance_impl(int n, int64_t m, uint64_t seed, int mode,
                                          int32_t *tail, int32_t *head, double *ub, double *cost, double *mult, double *supply);

extern "C" int gne_generate_instance(int n, int64_t m, uint64_t seed, int mode,
                                     int32_t *tail, int32_t *head, double *ub, double *cost, double *mult, double *supply)
{
    const int rc = gne_generate_instance_impl(n, m, seed, mode, tail, head, ub, cost, mult, supply);
    if (rc) gne_set_error("gne_generate_instance: need n >= 3 and n <= m <= n(n-1)");
    return rc;
}
