// genneteq_b200 — layer 1 of the C ABI: device-resident state and kernel launches.
// See include/genneteq_b200.h for the contract of every entry point.
#include "gne_kernels.cuh"
#include "gne_exchange.cuh"
#include "gne_stream.cuh"
#ifdef GNE_TUNING
#include "gne_variants.cuh"            // tuning harness: only in `make tuning` builds (libgenneteq_b200_tuning.so)
#endif
#include "gne_internal.h"
#include "../../include/genneteq_b200.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <vector>
#include <sched.h>
#include <sys/prctl.h>
#include <time.h>

using namespace gne;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void gne_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *gne_last_error(void) { return g_err; }
extern "C" const char *gne_version(void) { return "genneteq_b200 0.1 (sm_100a)"; }

extern "C" int gne_cuda_device_count(void)
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            gne_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return GNE_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

// ------------------------------------------------------------------------------------------
// NCCL through dlopen: the library loads on machines without NCCL; gne_comm_* fail cleanly.
// ------------------------------------------------------------------------------------------
namespace {
struct NcclId { char b[128]; };                         // ncclUniqueId is 128 opaque bytes, passed by value
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool nccl_load()
{
    if (g_nccl.lib) return true;
    const char *names[] = { "libnccl.so.2", "libnccl.so", nullptr };
    for (int i = 0; names[i] && !g_nccl.lib; ++i) g_nccl.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!g_nccl.lib) { gne_set_error("NCCL not found: %s", dlerror()); return false; }
    *(void **) &g_nccl.GetUniqueId = dlsym(g_nccl.lib, "ncclGetUniqueId");
    *(void **) &g_nccl.CommInitRank = dlsym(g_nccl.lib, "ncclCommInitRank");
    *(void **) &g_nccl.CommDestroy = dlsym(g_nccl.lib, "ncclCommDestroy");
    *(void **) &g_nccl.AllGather = dlsym(g_nccl.lib, "ncclAllGather");
    *(void **) &g_nccl.GetErrorString = dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather) {
        gne_set_error("NCCL symbols missing");
        return false;
    }
    return true;
}
} // namespace

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct gne_dev {
    // equal-flow rows kept on the device (gne_eqf_set_rows) and the mapped result block of gne_price_eqf_launch / _fetch
    int eqfP = 0;
    int64_t *eqfRowStart = nullptr;
    int32_t *eqfNode = nullptr;
    double *eqfVal = nullptr;
    double *eqfOut = nullptr;                      // mapped pinned: EQF_CALL_MAX values, then the sequence word
    unsigned long long eqfSeq = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    int n = 0;
    int64_t cap = 0, lo = 0, hi = 0, N = 0;        // global capacity, shard [lo, hi), global count
    int64_t localCap = 0;                          // allocated shard length (multiple of TILE)
    // nonbasic shard (structure-of-arrays, local index = pos - lo)
    int32_t *tail = nullptr, *head = nullptr;
    double *cost = nullptr, *mult = nullptr;
    int8_t *state = nullptr;
    // nodes
    double *a0 = nullptr, *a1 = nullptr, *theta = nullptr, *pi = nullptr;
    int32_t *slot = nullptr;
    // tile scratch
    int G = 0;                                     // tiles covering localCap
    LvTiles lv{};
    int32_t *tileCnt = nullptr;
    int64_t *tileOff = nullptr;
    double *tileVmin = nullptr, *tileVmax = nullptr;
    SelectResult *selRes = nullptr;                // mapped pinned
    // state of the last gne_price_lv_begin
    int lvMode = 0;                                // 0: list on the host, 1: all compacted candidates, on the device
    std::vector<int64_t> lvList;
    int64_t lvCount = 0;
    double lvFloor = 0.0;
    QsTiles qs{};
    int GQ = 0;
    unsigned long long *firstPos = nullptr;
    // mapped pinned host blocks
    LvResult *lvRes = nullptr;
    QsState *qsState = nullptr;
    int64_t *hostScalar = nullptr;                 // [0] compaction total, [1] first pos
    // staging for updates (pinned host + device), grown on demand
    void *stageHost = nullptr, *stageDev = nullptr;
    size_t stageBytes = 0;
    // compaction output (device), grown on demand
    int64_t *outPos = nullptr;
    double *outViol = nullptr;
    int64_t outCap = 0;
    double *flushBuf = nullptr;
    int64_t flushLen = 0;
    float lastMs = 0.f;
    int64_t launches = 0;
    int64_t h2dBytes = 0, d2hBytes = 0;
    int64_t qsWindow = 0;                          // adaptive window of the quickSelect scan
    int64_t qsArcsLaunched = 0;                    // arcs the quickSelect kernels were launched over (whole windows)
    NbWriteOps pending{};                          // queued gne_nb_write records (<= 8)
    bool stageInFlight = false;                    // the staging buffer feeds a kernel not yet synchronised
    LvCtas lvc{};                                  // per-CTA records of the TMA-streamed LV pass
    int streamGrid = 0;                            // 2 CTAs per SM
    int stStages = ST_STAGES, stCtasPerSm = ST_CTAS_PER_SM;     // shape of the streamed pass (GNE_ST_SHAPE="stages,ctas" for tuning runs)
    int streamGridFull = 0;                        // ... of the whole device (gne_dev_set_stream_ctas lowers streamGrid for batches)
    unsigned int *lvTicket = nullptr;              // counters of the streamed pass, one 128-byte line each (zero between passes):
    unsigned int *lvArrive = nullptr;              //   last-CTA ticket, first-CTA election, next tile, update-visible flag
    unsigned long long *lvTileCtr = nullptr, *lvUpdFlag = nullptr;
    SmallUpdate pendingSmall{};                    // argument-only update not launched yet: the next streamed LV pass carries it
    bool havePendingSmall = false;
    int streamOverflows = 0, streamHoldoff = 0;    // adaptive fallback to the per-tile kernel
    bool lastWasStream = false;
    int lvKernelMode = 0;                          // 0 auto, 1 per-tile kernel, 2 TMA-streamed kernel
    SmallUpdate lastSmall{};                       // the last argument-only update (bench replay)
    bool haveLastSmall = false;
    bool timing = false;                           // bracket pricing kernels with CUDA events (gne_dev_set_timing)
    std::vector<cudaEvent_t> benchEv;              // per-step event pairs of gne_bench_price_lv
    bool exchangeFused = false;                    // the last streamed pass carried the shard exchange itself
    bool lvLaunched = false;                       // a pass started by gne_price_lv_launch awaits gne_price_lv_complete
    cudaStream_t passStream = nullptr;             // the stream that pass runs on (the handle's own, or its launch group's)
    bool fastMath = false;                         // gne_dev_set_fast_math / GNE_FAST_MATH
    bool pdl = true;                               // programmatic dependent launch between a pivot's kernels (GNE_PDL=0 disables)
    unsigned long long lvSeq = 0;                  // sequence number of the last LV pass launched
    unsigned long long qsSeq = 0;
    QsState *qsDev = nullptr;                      // device-resident carry of the quickSelect scan
    // multi-GPU
    void *comm = nullptr;                          // NCCL communicator (gne_comm_init) ...
    gne_allgather_fn hostGather = nullptr;         // ... or the host program's all-gather (gne_comm_init_host)
    void *hostGatherCtx = nullptr;
    char *agDev = nullptr;                         // NCCL staging of small host all-gathers
    size_t agBytes = 0;
    long pollSleepNs = 0;                          // > 0: the result poll sleeps this long between looks (batches)
    long long xTimeoutCycles = 8000000000ll;       // peer wait inside the exchange kernels (GNE_EXCHANGE_TIMEOUT_S, default ~4 s)
    int rank = 0, nranks = 1;
    void *gatherSend = nullptr, *gatherRecv = nullptr, *gatherHost = nullptr;
    LvResult *lvResDev = nullptr;                  // sharded mode: the local LV result stays on the device
    struct Mailbox *mailbox = nullptr;             // this rank's receive buffer, written by the peers over NVLink
    struct Mailbox **peersDev = nullptr;           // device array [nranks] of every rank's mailbox (IPC-mapped)
    void *peerPtr[8] = { nullptr };
    bool p2p = false;
    unsigned long long xseq = 0, bseq = 0;
    unsigned long long *blobHost = nullptr, *blobDev = nullptr;
    int64_t *cntDev = nullptr;                     // [1 + nranks] counts exchange
    char *xSend = nullptr, *xRecv = nullptr;       // variable-length list exchange, grown on demand
    size_t xSendBytes = 0, xRecvBytes = 0;
};

static NbView view_of(const gne_dev *d)
{
    NbView v;
    v.tail = d->tail; v.head = d->head; v.cost = d->cost; v.mult = d->mult; v.state = d->state;
    v.lo = d->lo;
    v.end = std::min(d->N, d->hi);
    v.begin = d->lo;
    v.tile0 = 0;
    return v;
}

static inline int tiles_for(const gne_dev *d)
{
    // tiles that contain at least one priced position of this shard
    const int64_t end = std::min(d->N, d->hi);
    if (end <= d->lo) return 0;
    return (int) ((end - d->lo + TILE - 1) / TILE);
}

static int ensure_stage(gne_dev *d, size_t bytes)
{
    if (bytes <= d->stageBytes) return GNE_OK;
    size_t nb = std::max<size_t>(bytes, std::max<size_t>(2 * d->stageBytes, (size_t) 1 << 20));
    if (d->stageHost) cudaFreeHost(d->stageHost);
    if (d->stageDev) cudaFree(d->stageDev);
    d->stageHost = nullptr; d->stageDev = nullptr; d->stageBytes = 0;
    CK(cudaHostAlloc(&d->stageHost, nb, cudaHostAllocMapped));
    CK(cudaMalloc(&d->stageDev, nb));
    d->stageBytes = nb;
    return GNE_OK;
}

static int ensure_out(gne_dev *d, int64_t cap)
{
    if (cap <= d->outCap) return GNE_OK;
    int64_t nc = std::max<int64_t>(cap, std::max<int64_t>(2 * d->outCap, 1 << 16));
    if (d->outPos) cudaFree(d->outPos);
    if (d->outViol) cudaFree(d->outViol);
    d->outPos = nullptr; d->outViol = nullptr; d->outCap = 0;
    CK(cudaMalloc(&d->outPos, sizeof(int64_t) * nc));
    CK(cudaMalloc(&d->outViol, sizeof(double) * nc));
    d->outCap = nc;
    return GNE_OK;
}

static int dev_create_fill(gne_dev *d, int cudaDevice, int n, int64_t nbCapacity, int64_t shardLo, int64_t shardHi);
static int flush_pending(gne_dev *d);

// The streamed LV pass in its production shape (3 stages of 25.6 KB, 2 CTAs per SM) and the shapes kept for tuning runs.
typedef void (*StKernel)(const NbView, const double *, const LvCtas, const int64_t, const LvFused, const SmallUpdate, const NbMut, const PotMut);
static StKernel st_kernel_for(int stages, int ctasPerSm)
{
    if (stages == 3 && ctasPerSm == 2) return k_price_lv_stream<3, 2>;
    if (stages == 2 && ctasPerSm == 2) return k_price_lv_stream<2, 2>;
    if (stages == 2 && ctasPerSm == 3) return k_price_lv_stream<2, 3>;
    if (stages == 4 && ctasPerSm == 2) return k_price_lv_stream<4, 2>;
    if (stages == 3 && ctasPerSm == 1) return k_price_lv_stream<3, 1>;
    return nullptr;
}

extern "C" int gne_dev_create(gne_dev **out, int cudaDevice, int n, int64_t nbCapacity, int64_t shardLo, int64_t shardHi)
{
    if (!out || n <= 0 || nbCapacity <= 0 || shardLo < 0 || shardHi < shardLo) { gne_set_error("gne_dev_create: bad argument"); return GNE_ERR_ARG; }
    if (shardLo % TILE != 0) { gne_set_error("gne_dev_create: shardLo must be a multiple of %d", TILE); return GNE_ERR_ARG; }
    int cnt = gne_cuda_device_count();
    if (cnt <= 0) { gne_set_error("no CUDA device: genneteq_b200 has no CPU fallback"); return GNE_ERR_CUDA; }
    if (cudaDevice < 0 || cudaDevice >= cnt) { gne_set_error("cuda device %d out of range (%d present)", cudaDevice, cnt); return GNE_ERR_ARG; }
    CK(cudaSetDevice(cudaDevice));
    gne_dev *d = new gne_dev();
    const int rc = dev_create_fill(d, cudaDevice, n, nbCapacity, shardLo, shardHi);
    if (rc != GNE_OK) { gne_dev_destroy(d); return rc; }      // every member is null-initialised: destroy frees what was allocated
    *out = d;
    return GNE_OK;
}

static int dev_create_fill(gne_dev *d, int cudaDevice, int n, int64_t nbCapacity, int64_t shardLo, int64_t shardHi)
{
    d->device = cudaDevice; d->n = n; d->cap = nbCapacity;
    // a shard that starts at or beyond the capacity is EMPTY (more ranks than tiles): lo == hi, nothing is ever priced
    if (shardLo >= nbCapacity) shardHi = shardLo;
    d->lo = shardLo; d->hi = std::max(shardLo, std::min(shardHi, nbCapacity)); d->N = 0;
    d->localCap = ((d->hi - d->lo + TILE - 1) / TILE) * TILE;
    if (d->localCap == 0) d->localCap = TILE;
    d->G = (int) (d->localCap / TILE);
    CK(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&d->ev0));
    CK(cudaEventCreate(&d->ev1));
    const size_t L = (size_t) d->localCap + 4 * ST_TILE;       // padding: the streamed pass copies whole tiles
    CK(cudaMalloc(&d->tail, 4 * L)); CK(cudaMalloc(&d->head, 4 * L));
    CK(cudaMalloc(&d->cost, 8 * L)); CK(cudaMalloc(&d->mult, 8 * L)); CK(cudaMalloc(&d->state, L));
    CK(cudaMemsetAsync(d->tail, 0, 4 * L, d->stream)); CK(cudaMemsetAsync(d->head, 0, 4 * L, d->stream));
    CK(cudaMemsetAsync(d->cost, 0, 8 * L, d->stream)); CK(cudaMemsetAsync(d->mult, 0, 8 * L, d->stream));
    CK(cudaMemsetAsync(d->state, 0, L, d->stream));
    CK(cudaMalloc(&d->a0, 8 * (size_t) n)); CK(cudaMalloc(&d->a1, 8 * (size_t) n));
    CK(cudaMalloc(&d->theta, 8 * (size_t) n)); CK(cudaMalloc(&d->pi, 8 * (size_t) n));
    CK(cudaMalloc(&d->slot, 4 * (size_t) n));
    CK(cudaMemsetAsync(d->a0, 0, 8 * (size_t) n, d->stream)); CK(cudaMemsetAsync(d->a1, 0, 8 * (size_t) n, d->stream));
    CK(cudaMemsetAsync(d->theta, 0, 8 * (size_t) n, d->stream)); CK(cudaMemsetAsync(d->pi, 0, 8 * (size_t) n, d->stream));
    CK(cudaMemsetAsync(d->slot, 0, 4 * (size_t) n, d->stream));
    CK(cudaMalloc(&d->lv.tileMax, 8 * (size_t) d->G)); CK(cudaMalloc(&d->lv.tileCnt, 4 * (size_t) d->G));
    CK(cudaMalloc(&d->lv.candOff, 4 * (size_t) d->G * CAND_K)); CK(cudaMalloc(&d->lv.candViol, 8 * (size_t) d->G * CAND_K));
    CK(cudaMalloc(&d->tileCnt, 4 * (size_t) d->G)); CK(cudaMalloc(&d->tileOff, 8 * ((size_t) d->G + 1)));
    CK(cudaMalloc(&d->tileVmin, 8 * (size_t) d->G)); CK(cudaMalloc(&d->tileVmax, 8 * (size_t) d->G));
    CK(cudaHostAlloc(&d->selRes, sizeof(SelectResult), cudaHostAllocMapped));
    d->GQ = (int) ((d->localCap + QTILE - 1) / QTILE) + 1;
    CK(cudaMalloc(&d->qs.cnt, 4 * (size_t) d->GQ)); CK(cudaMalloc(&d->qs.bestV, 8 * (size_t) d->GQ)); CK(cudaMalloc(&d->qs.bestJ, 8 * (size_t) d->GQ));
    CK(cudaMalloc(&d->firstPos, 8));
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cudaDevice);
        { const char *sh = getenv("GNE_ST_SHAPE"); int a = 0, b = 0; if (sh && sscanf(sh, "%d,%d", &a, &b) == 2) { d->stStages = a; d->stCtasPerSm = b; } }
        if (!st_kernel_for(d->stStages, d->stCtasPerSm)) { gne_set_error("GNE_ST_SHAPE=%d,%d is not an instantiated shape of the streamed pass", d->stStages, d->stCtasPerSm); return GNE_ERR_ARG; }
        d->streamGrid = sms * d->stCtasPerSm;
        if (d->streamGrid > 2 * ST_THREADS) d->streamGrid = 2 * ST_THREADS;        // the reduction reads two records per thread
        d->streamGridFull = d->streamGrid;
        { const char *sg = getenv("GNE_STREAM_CTAS"); if (sg && atoi(sg) > 0) d->streamGrid = std::min(d->streamGridFull, atoi(sg)); }
        CK(cudaMalloc(&d->lvc.rec, sizeof(CtaRec) * (size_t) d->streamGrid));
        CK(cudaMalloc(&d->lvTicket, 512)); CK(cudaMemset(d->lvTicket, 0, 512));
        d->lvArrive = d->lvTicket + 32;
        d->lvTileCtr = reinterpret_cast<unsigned long long *>(d->lvTicket + 64);
        d->lvUpdFlag = reinterpret_cast<unsigned long long *>(d->lvTicket + 96);
        CK(cudaMalloc(&d->lvc.candPos, 8 * (size_t) d->streamGrid * ST_CAND)); CK(cudaMalloc(&d->lvc.candViol, 8 * (size_t) d->streamGrid * ST_CAND));
        CK(cudaMalloc(&d->lvc.candSecond, 8 * (size_t) d->streamGrid * ST_CAND));
        CK(cudaFuncSetAttribute(st_kernel_for(d->stStages, d->stCtasPerSm), cudaFuncAttributeMaxDynamicSharedMemorySize, st_smem_bytes(d->stStages)));
    }
    CK(cudaHostAlloc(&d->lvRes, sizeof(LvResult), cudaHostAllocMapped));
    CK(cudaHostAlloc(&d->qsState, sizeof(QsState), cudaHostAllocMapped));
    CK(cudaHostAlloc(&d->hostScalar, 64, cudaHostAllocMapped));
    memset(d->lvRes, 0, sizeof(LvResult));
    { const char *tm = getenv("GNE_TIMING"); d->timing = tm && tm[0] == '1'; }
    { const char *pd = getenv("GNE_PDL"); d->pdl = !(pd && pd[0] == '0'); }
    { const char *fm = getenv("GNE_FAST_MATH"); const int v = (fm && fm[0] == '1') ? 1 : 0; CK(cudaMemcpyToSymbol(c_gne_fast_math, &v, sizeof(int))); d->fastMath = v != 0; }
    { const char *ps = getenv("GNE_POLL_SLEEP_US"); if (ps && atof(ps) > 0.0) d->pollSleepNs = (long) (atof(ps) * 1000.0); }
    { const char *lk = getenv("GNE_LV_KERNEL"); if (lk) d->lvKernelMode = !strcmp(lk, "stream") ? 2 : (!strcmp(lk, "tiles") ? 1 : 0); }
    memset(d->qsState, 0, sizeof(QsState));
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    return GNE_OK;
}

extern "C" int gne_dev_destroy(gne_dev *d)
{
    if (!d) return GNE_OK;
    cudaSetDevice(d->device);
    if (d->stream) cudaStreamSynchronize(d->stream);
    for (int p = 0; p < 8; ++p) if (d->peerPtr[p] && d->peerPtr[p] != (void *) d->mailbox) cudaIpcCloseMemHandle(d->peerPtr[p]);
    if (d->mailbox) cudaFree(d->mailbox);
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    void *dev[] = { d->tail, d->head, d->cost, d->mult, d->state, d->a0, d->a1, d->theta, d->pi, d->slot,
                    d->lv.tileMax, d->lv.tileCnt, d->lv.candOff, d->lv.candViol, d->tileCnt, d->tileOff,
                    d->qs.cnt, d->qs.bestV, d->qs.bestJ, d->firstPos, d->tileVmin, d->tileVmax, d->lvc.rec, d->lvTicket, d->lvc.candPos, d->lvc.candViol, d->lvc.candSecond, d->qsDev, d->stageDev, d->outPos, d->outViol,
                    d->flushBuf, d->gatherSend, d->gatherRecv, d->cntDev, d->xSend, d->xRecv, d->lvResDev, d->peersDev, d->blobDev, d->agDev,
                    d->eqfRowStart, d->eqfNode, d->eqfVal };
    for (void *p : dev) if (p) cudaFree(p);
    void *host[] = { d->lvRes, d->qsState, d->hostScalar, d->stageHost, d->gatherHost, d->selRes, d->blobHost, d->eqfOut };
    for (void *p : host) if (p) cudaFreeHost(p);
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    if (d->ev2) cudaEventDestroy(d->ev2);
    if (d->ev3) cudaEventDestroy(d->ev3);
    for (cudaEvent_t e : d->benchEv) cudaEventDestroy(e);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
    return GNE_OK;
}

extern "C" int gne_dev_sync(gne_dev *d)
{
    CK(cudaSetDevice(d->device));
    { int rcf = flush_pending(d); if (rcf) return rcf; }
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    return GNE_OK;
}

extern "C" int gne_dev_last_kernel_ms(gne_dev *d, float *ms) { *ms = d->lastMs; return GNE_OK; }
extern "C" int gne_dev_set_timing(gne_dev *d, int on) { d->timing = on != 0; return GNE_OK; }
extern "C" int gne_dev_set_stream_ctas(gne_dev *d, int ctas)
{
    if (ctas < 0) { gne_set_error("gne_dev_set_stream_ctas: negative"); return GNE_ERR_ARG; }
    d->streamGrid = ctas == 0 ? d->streamGridFull : std::min(ctas, d->streamGridFull);
    return GNE_OK;
}
extern "C" int gne_dev_set_fast_math(gne_dev *d, int on)
{
    CK(cudaSetDevice(d->device));
    { int rcf = flush_pending(d); if (rcf) return rcf; }
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    const int v = on ? 1 : 0;
    CK(cudaMemcpyToSymbol(c_gne_fast_math, &v, sizeof(int)));
    d->fastMath = on != 0;
    return GNE_OK;
}
extern "C" int gne_dev_set_lv_kernel(gne_dev *d, int mode) { if (mode < 0 || mode > 2) return GNE_ERR_ARG; d->lvKernelMode = mode; return GNE_OK; }
extern "C" int gne_dev_last_lv_kernel(gne_dev *d) { return d->lastWasStream ? 2 : 1; }
extern "C" int64_t gne_dev_launch_count(gne_dev *d) { return d->launches; }
extern "C" int gne_dev_cuda_device(const gne_dev *d) { return d ? d->device : -1; }
void gne_dev_traffic(gne_dev *d, int64_t *h2d, int64_t *d2h) { *h2d = d->h2dBytes; *d2h = d->d2hBytes; }

// ------------------------------------------------------------------------------------------
// nonbasic set mirror
// ------------------------------------------------------------------------------------------
extern "C" int gne_nb_reset(gne_dev *d, int64_t N, const int32_t *tail, const int32_t *head, const double *cost,
                            const double *mult, const int8_t *state)
{
    if (N < 0 || N > d->cap) { gne_set_error("gne_nb_reset: N out of range"); return GNE_ERR_ARG; }
    CK(cudaSetDevice(d->device));
    { int rcf = flush_pending(d); if (rcf) return rcf; }
    const int64_t a = d->lo, b = std::min(N, d->hi);
    // the pricing kernels gather pi[tail], pi[head] unchecked: reject a bad record here (as gne_nb_write does)
    for (int64_t p = a; p < b; ++p)
        if (tail[p] < 0 || tail[p] >= d->n || head[p] < 0 || head[p] >= d->n) { gne_set_error("gne_nb_reset: position %lld names a node out of range", (long long) p); return GNE_ERR_ARG; }
    d->N = N;
    if (b > a) {
        const size_t c = (size_t) (b - a);
        CK(cudaMemcpyAsync(d->tail, tail + a, 4 * c, cudaMemcpyHostToDevice, d->stream));
        CK(cudaMemcpyAsync(d->head, head + a, 4 * c, cudaMemcpyHostToDevice, d->stream));
        CK(cudaMemcpyAsync(d->cost, cost + a, 8 * c, cudaMemcpyHostToDevice, d->stream));
        CK(cudaMemcpyAsync(d->mult, mult + a, 8 * c, cudaMemcpyHostToDevice, d->stream));
        CK(cudaMemcpyAsync(d->state, state + a, c, cudaMemcpyHostToDevice, d->stream));
        d->h2dBytes += 25 * (int64_t) c;
    }
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;          // caller may free its arrays
    return GNE_OK;
}

extern "C" int gne_nb_set_costs(gne_dev *d, int64_t N, const double *cost)
{
    if (N != d->N) { gne_set_error("gne_nb_set_costs: N mismatch"); return GNE_ERR_ARG; }
    CK(cudaSetDevice(d->device));
    { int rcf = flush_pending(d); if (rcf) return rcf; }
    const int64_t a = d->lo, b = std::min(N, d->hi);
    if (b > a) {
        CK(cudaMemcpyAsync(d->cost, cost + a, 8 * (size_t) (b - a), cudaMemcpyHostToDevice, d->stream));
        d->h2dBytes += 8 * (b - a);
    }
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    return GNE_OK;
}

// queued writes are flushed by one k_nb_write launch before the next kernel that reads the list
static int flush_writes(gne_dev *d, NbWriteOps &ops)
{
    if (ops.n == 0) return GNE_OK;
    NbMut m; m.tail = d->tail; m.head = d->head; m.cost = d->cost; m.mult = d->mult; m.state = d->state;
    k_nb_write<<<1, 32, 0, d->stream>>>(m, ops);
    CK(cudaGetLastError());
    ++d->launches;
    d->h2dBytes += 25 * ops.n;
    ops.n = 0;
    return GNE_OK;
}

static NbWriteOps &pending_of(gne_dev *d) { return d->pending; }

// Everything queued but not launched yet, in order: the argument-only update that waits for the next streamed LV pass
// (it carries the set-mirror writes that were queued before it), then the set-mirror writes queued since.
static int flush_pending(gne_dev *d)
{
    if (d->havePendingSmall) {
        NbMut mm; mm.tail = d->tail; mm.head = d->head; mm.cost = d->cost; mm.mult = d->mult; mm.state = d->state;
        k_update_small<<<1, 256, 0, d->stream>>>(mm, d->pendingSmall, d->a0, d->a1, d->slot, d->theta, d->pi);
        CK(cudaGetLastError());
        ++d->launches;
        d->havePendingSmall = false;
    }
    return flush_writes(d, pending_of(d));
}

int gne_dev_flush(gne_dev *d) { return flush_pending(d); }

extern "C" int gne_nb_write(gne_dev *d, int64_t pos, int32_t tail, int32_t head, double cost, double mult, int8_t state)
{
    if (pos < 0 || pos >= d->cap) { gne_set_error("gne_nb_write: pos out of range"); return GNE_ERR_ARG; }
    if (tail < 0 || tail >= d->n || head < 0 || head >= d->n) { gne_set_error("gne_nb_write: node out of range"); return GNE_ERR_ARG; }
    if (pos < d->lo || pos >= d->hi) return GNE_OK;     // another rank's position
    CK(cudaSetDevice(d->device));
    NbWriteOps &ops = pending_of(d);
    // the batch is applied by concurrent threads: a later write to the same position replaces
    // the queued one instead of racing with it
    int i = -1;
    for (int k = 0; k < ops.n; ++k) if (ops.local[k] == pos - d->lo) i = k;
    if (i < 0) {
        if (ops.n == 8) { int rc = flush_writes(d, ops); if (rc) return rc; }
        i = ops.n++;
    }
    ops.local[i] = pos - d->lo; ops.tail[i] = tail; ops.head[i] = head; ops.cost[i] = cost; ops.mult[i] = mult; ops.state[i] = state;
    return GNE_OK;
}

extern "C" int gne_nb_set_count(gne_dev *d, int64_t N)
{
    if (N < 0 || N > d->cap) { gne_set_error("gne_nb_set_count: N out of range"); return GNE_ERR_ARG; }
    d->N = N;
    return GNE_OK;
}

extern "C" int gne_nb_read(gne_dev *d, int64_t lo, int64_t hi, int32_t *tail, int32_t *head, double *cost, double *mult, int8_t *state)
{
    if (lo < d->lo || hi > d->hi || hi < lo) { gne_set_error("gne_nb_read: range outside the shard"); return GNE_ERR_ARG; }
    CK(cudaSetDevice(d->device));
    int rc = gne_dev_flush(d); if (rc) return rc;
    const size_t c = (size_t) (hi - lo), o = (size_t) (lo - d->lo);
    if (tail) CK(cudaMemcpyAsync(tail, d->tail + o, 4 * c, cudaMemcpyDeviceToHost, d->stream));
    if (head) CK(cudaMemcpyAsync(head, d->head + o, 4 * c, cudaMemcpyDeviceToHost, d->stream));
    if (cost) CK(cudaMemcpyAsync(cost, d->cost + o, 8 * c, cudaMemcpyDeviceToHost, d->stream));
    if (mult) CK(cudaMemcpyAsync(mult, d->mult + o, 8 * c, cudaMemcpyDeviceToHost, d->stream));
    if (state) CK(cudaMemcpyAsync(state, d->state + o, c, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    return GNE_OK;
}

// ------------------------------------------------------------------------------------------
// potentials
// ------------------------------------------------------------------------------------------
static inline int blocks_for(int64_t n, int b) { return (int) ((n + b - 1) / b); }

extern "C" int gne_pot_update_nodes(gne_dev *d, int64_t count, const int32_t *node, const double *a0, const double *a1, const int32_t *slot)
{
    if (count <= 0) return GNE_OK;
    CK(cudaSetDevice(d->device));
    { int rcf = flush_pending(d); if (rcf) return rcf; }
    const size_t bytes = (size_t) count * 24;
    int rc = ensure_stage(d, bytes); if (rc) return rc;
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;           // the staging buffer may still be in flight
    char *h = (char *) d->stageHost;
    memcpy(h, a0, 8 * (size_t) count);
    memcpy(h + 8 * (size_t) count, a1, 8 * (size_t) count);
    memcpy(h + 16 * (size_t) count, node, 4 * (size_t) count);
    memcpy(h + 20 * (size_t) count, slot, 4 * (size_t) count);
    CK(cudaMemcpyAsync(d->stageDev, h, bytes, cudaMemcpyHostToDevice, d->stream));
    char *g = (char *) d->stageDev;
    k_pot_update_nodes<<<blocks_for(count, 256), 256, 0, d->stream>>>(
        count, (const int32_t *) (g + 16 * (size_t) count), (const double *) g, (const double *) (g + 8 * (size_t) count),
        (const int32_t *) (g + 20 * (size_t) count), d->a0, d->a1, d->slot);
    CK(cudaGetLastError());
    ++d->launches;
    d->h2dBytes += (int64_t) bytes;
    d->stageInFlight = true;
    return GNE_OK;
}

extern "C" int gne_pot_update_thetas(gne_dev *d, int64_t count, const int32_t *slot, const double *theta)
{
    if (count <= 0) return GNE_OK;
    CK(cudaSetDevice(d->device));
    { int rcf = flush_pending(d); if (rcf) return rcf; }
    const size_t bytes = (size_t) count * 12;
    int rc = ensure_stage(d, bytes); if (rc) return rc;
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    char *h = (char *) d->stageHost;
    memcpy(h, theta, 8 * (size_t) count);
    memcpy(h + 8 * (size_t) count, slot, 4 * (size_t) count);
    CK(cudaMemcpyAsync(d->stageDev, h, bytes, cudaMemcpyHostToDevice, d->stream));
    char *g = (char *) d->stageDev;
    k_pot_update_thetas<<<blocks_for(count, 256), 256, 0, d->stream>>>(count, (const int32_t *) (g + 8 * (size_t) count), (const double *) g, d->theta);
    CK(cudaGetLastError());
    ++d->launches;
    d->h2dBytes += (int64_t) bytes;
    d->stageInFlight = true;
    return GNE_OK;
}

// One call per pivot from the solver: queued set writes + dirty records + thetas + finalize.
// nFinal < 0 => finalize every node.  Small updates ride in a single one-CTA kernel that reads the
// mapped pinned staging buffer directly; large ones are copied to the device first.
extern "C" int gne_pot_update_fused(gne_dev *d, int64_t nNodes, const int32_t *node, const double *a0, const double *a1,
                                    const int32_t *slot, int64_t nTheta, const int32_t *tslot, const double *theta,
                                    int64_t nFinal, const int32_t *fnode)
{
    return gne_pot_update_fused_pi(d, nNodes, node, a0, a1, slot, nTheta, tslot, theta, nFinal, fnode, nullptr);
}

extern "C" int gne_pot_update_fused_pi(gne_dev *d, int64_t nNodes, const int32_t *node, const double *a0, const double *a1,
                                       const int32_t *slot, int64_t nTheta, const int32_t *tslot, const double *theta,
                                       int64_t nFinal, const int32_t *fnode, const double *fpi)
{
    CK(cudaSetDevice(d->device));
    const bool full = nFinal < 0;
    const int64_t nf = full ? 0 : nFinal;
    const size_t bytes = (size_t) nNodes * 24 + (size_t) nTheta * 12 + (size_t) nf * 4 + 64;
    if (nNodes == 0 && nTheta == 0 && nf == 0 && !full) return GNE_OK;       // queued set-mirror writes ride with the next pass
    if (d->havePendingSmall) { int rcf = flush_pending(d); if (rcf) return rcf; }   // two updates without a pass in between: in order
    if (!full && nNodes <= SU_NODES && nTheta <= SU_THETAS && nf <= SU_FINAL) {
        // tiny update: everything as kernel arguments (no staging buffer, no host memory reads).  Not launched here:
        // the next streamed LV pass applies it itself (gne_stream.cuh); any other reader of the state launches it
        // first as k_update_small (flush_pending).
        SmallUpdate &su = d->pendingSmall;
        NbWriteOps &q = pending_of(d);
        su.ops = q;
        su.nNodes = (int) nNodes; su.nTheta = (int) nTheta; su.nFinal = (int) nf;
        su.hasPi = (fpi != nullptr && !d->fastMath) ? 1 : 0;             // (the fused mode finalizes on the device, with its FMA)
        if (su.hasPi) for (int64_t i = 0; i < nf; ++i) su.fpi[i] = fpi[i];
        for (int64_t i = 0; i < nNodes; ++i) { su.node[i] = node[i]; su.a0[i] = a0[i]; su.a1[i] = a1[i]; su.slot[i] = slot[i]; }
        for (int64_t i = 0; i < nTheta; ++i) { su.tslot[i] = tslot[i]; su.theta[i] = theta[i]; }
        for (int64_t i = 0; i < nf; ++i) su.fnode[i] = fnode[i];
        d->havePendingSmall = true;
        d->lastSmall = su; d->lastSmall.ops.n = 0; d->haveLastSmall = true;     // replayable (idempotent) by the bench helper
        d->h2dBytes += 25 * q.n + 24 * nNodes + 12 * nTheta + 4 * nf;
        q.n = 0;
        return GNE_OK;
    }
    int rc = ensure_stage(d, bytes); if (rc) return rc;
    if (d->stageInFlight) { CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false; d->stageInFlight = false; }
    char *h = (char *) d->stageHost;
    size_t o = 0;
    const size_t oA0 = o; memcpy(h + o, a0, 8 * (size_t) nNodes); o += 8 * (size_t) nNodes;
    const size_t oA1 = o; memcpy(h + o, a1, 8 * (size_t) nNodes); o += 8 * (size_t) nNodes;
    const size_t oTh = o; memcpy(h + o, theta, 8 * (size_t) nTheta); o += 8 * (size_t) nTheta;
    const size_t oNd = o; memcpy(h + o, node, 4 * (size_t) nNodes); o += 4 * (size_t) nNodes;
    const size_t oSl = o; memcpy(h + o, slot, 4 * (size_t) nNodes); o += 4 * (size_t) nNodes;
    const size_t oTs = o; memcpy(h + o, tslot, 4 * (size_t) nTheta); o += 4 * (size_t) nTheta;
    const size_t oFn = o; if (nf) memcpy(h + o, fnode, 4 * (size_t) nf); o += 4 * (size_t) nf;
    d->h2dBytes += (int64_t) o;
    d->stageInFlight = true;
    NbMut m; m.tail = d->tail; m.head = d->head; m.cost = d->cost; m.mult = d->mult; m.state = d->state;
    const bool small = !full && o <= (size_t) 96 * 1024;
    const char *src = small ? h : (const char *) d->stageDev;
    if (!small) CK(cudaMemcpyAsync(d->stageDev, h, o, cudaMemcpyHostToDevice, d->stream));
    FusedUpdate u;
    u.nNodes = nNodes; u.node = (const int32_t *) (src + oNd); u.a0 = (const double *) (src + oA0); u.a1 = (const double *) (src + oA1);
    u.slot = (const int32_t *) (src + oSl);
    u.nTheta = nTheta; u.tslot = (const int32_t *) (src + oTs); u.theta = (const double *) (src + oTh);
    u.nFinal = nf; u.fnode = (const int32_t *) (src + oFn);
    NbWriteOps &ops = pending_of(d);
    if (small) {
        d->h2dBytes += 25 * ops.n;
        k_update_fused<<<1, 1024, 0, d->stream>>>(m, ops, u, d->a0, d->a1, d->slot, d->theta, d->pi);
        ops.n = 0;
        ++d->launches;
    } else {
        rc = flush_writes(d, ops); if (rc) return rc;
        if (nNodes) { k_pot_update_nodes<<<blocks_for(nNodes, 256), 256, 0, d->stream>>>(nNodes, u.node, u.a0, u.a1, u.slot, d->a0, d->a1, d->slot); ++d->launches; }
        if (nTheta) { k_pot_update_thetas<<<blocks_for(nTheta, 256), 256, 0, d->stream>>>(nTheta, u.tslot, u.theta, d->theta); ++d->launches; }
        if (full) k_pot_finalize<<<blocks_for(d->n, 256), 256, 0, d->stream>>>(d->n, d->a0, d->a1, d->slot, d->theta, d->pi);
        else k_pot_finalize_list<<<blocks_for(nf, 256), 256, 0, d->stream>>>(nf, u.fnode, d->a0, d->a1, d->slot, d->theta, d->pi);
        ++d->launches;
    }
    CK(cudaGetLastError());
    return GNE_OK;
}

extern "C" int gne_pot_finalize(gne_dev *d)
{
    CK(cudaSetDevice(d->device));
    { int rcf = flush_pending(d); if (rcf) return rcf; }
    k_pot_finalize<<<blocks_for(d->n, 256), 256, 0, d->stream>>>(d->n, d->a0, d->a1, d->slot, d->theta, d->pi);
    CK(cudaGetLastError());
    ++d->launches;
    return GNE_OK;
}

extern "C" int gne_pot_finalize_slots(gne_dev *d, int count, const int32_t *slots)
{
    if (count <= 0) return GNE_OK;
    CK(cudaSetDevice(d->device));
    { int rcf = flush_pending(d); if (rcf) return rcf; }
    for (int k = 0; k < count; k += 8) {
        SlotSet ss; ss.n = std::min(8, count - k);
        for (int i = 0; i < 8; ++i) ss.slot[i] = i < ss.n ? slots[k + i] : -1;
        k_pot_finalize_slots<<<blocks_for(d->n, 256), 256, 0, d->stream>>>(d->n, d->a0, d->a1, d->slot, d->theta, d->pi, ss);
        CK(cudaGetLastError());
        ++d->launches;
    }
    d->h2dBytes += 4 * count;
    return GNE_OK;
}

extern "C" int gne_pot_set(gne_dev *d, int64_t count, const int32_t *node, const double *pi)
{
    if (count <= 0) return GNE_OK;
    CK(cudaSetDevice(d->device));
    { int rcf = flush_pending(d); if (rcf) return rcf; }
    const size_t bytes = (size_t) count * 12;
    int rc = ensure_stage(d, bytes); if (rc) return rc;
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    char *h = (char *) d->stageHost;
    memcpy(h, pi, 8 * (size_t) count);
    memcpy(h + 8 * (size_t) count, node, 4 * (size_t) count);
    CK(cudaMemcpyAsync(d->stageDev, h, bytes, cudaMemcpyHostToDevice, d->stream));
    char *g = (char *) d->stageDev;
    k_pot_set<<<blocks_for(count, 256), 256, 0, d->stream>>>(count, (const int32_t *) (g + 8 * (size_t) count), (const double *) g, d->pi);
    CK(cudaGetLastError());
    ++d->launches;
    d->h2dBytes += (int64_t) bytes;
    d->stageInFlight = true;
    return GNE_OK;
}

extern "C" int gne_pot_set_all(gne_dev *d, const double *pi)
{
    CK(cudaSetDevice(d->device));
    { int rcf = flush_pending(d); if (rcf) return rcf; }
    CK(cudaMemcpyAsync(d->pi, pi, 8 * (size_t) d->n, cudaMemcpyHostToDevice, d->stream));
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    d->h2dBytes += 8 * (int64_t) d->n;
    return GNE_OK;
}

extern "C" int gne_pot_get_all(gne_dev *d, double *pi)
{
    CK(cudaSetDevice(d->device));
    { int rcf = flush_pending(d); if (rcf) return rcf; }
    CK(cudaMemcpyAsync(pi, d->pi, 8 * (size_t) d->n, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    d->d2hBytes += 8 * (int64_t) d->n;
    return GNE_OK;
}

// ------------------------------------------------------------------------------------------
// pricing
// ------------------------------------------------------------------------------------------
static int sharded_gather_lists(gne_dev *d, int64_t localCount, std::vector<int64_t> &pos, std::vector<double> &viol);
static int small_allgather(gne_dev *d, const void *sendHost, void *recvHost, size_t bytes);

// ordered compaction of this shard's violators with |rc| >= threshold, in two steps: count per tile
// + offsets (+ min / max of the qualifying values in hostScalar[4..5]), then the write pass
// view restricted to global positions [rlo, rhi) of this shard; *G = tiles to launch
static NbView range_view(const gne_dev *d, int64_t rlo, int64_t rhi, int *G)
{
    NbView nb = view_of(d);
    rlo = std::max(rlo, d->lo); rhi = std::min(rhi, nb.end);
    if (rhi <= rlo) { *G = 0; return nb; }
    nb.begin = rlo; nb.end = rhi;
    nb.tile0 = (int32_t) ((rlo - d->lo) / TILE);
    *G = (int) ((rhi - d->lo + TILE - 1) / TILE) - nb.tile0;
    return nb;
}

static int viol_count_device(gne_dev *d, double threshold, int64_t *total, double *vmin, double *vmax,
                             int64_t rlo = 0, int64_t rhi = INT64_MAX)
{
    int G = 0;
    NbView nb = range_view(d, rlo, rhi, &G);
    *total = 0;
    if (vmin) { *vmin = 0.0; *vmax = 0.0; }
    if (G == 0) return GNE_OK;
    k_viol_count<<<G, TILE_THREADS, 0, d->stream>>>(nb, d->pi, threshold, d->tileCnt, d->tileVmin, d->tileVmax);
    k_scan_tiles<<<1, 1024, 0, d->stream>>>(d->tileCnt, d->tileOff, G, d->hostScalar, d->tileVmin, d->tileVmax, (double *) (d->hostScalar + 4));
    CK(cudaGetLastError());
    d->launches += 2;
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    *total = d->hostScalar[0];
    if (vmin) { *vmin = ((double *) (d->hostScalar + 4))[0]; *vmax = ((double *) (d->hostScalar + 4))[1]; }
    d->d2hBytes += 24;
    return GNE_OK;
}

static int viol_write_device(gne_dev *d, double threshold, int64_t total, int64_t rlo = 0, int64_t rhi = INT64_MAX)
{
    if (total == 0) return GNE_OK;
    int rc = ensure_out(d, total); if (rc) return rc;
    int G = 0;
    NbView nb = range_view(d, rlo, rhi, &G);
    k_viol_write<<<G, TILE_THREADS, 0, d->stream>>>(nb, d->pi, threshold, d->tileOff, d->outPos, d->outViol, d->outCap);
    CK(cudaGetLastError());
    ++d->launches;
    return GNE_OK;
}

static int viol_compact_device(gne_dev *d, double threshold, int64_t *total)
{
    int rc = viol_count_device(d, threshold, total, nullptr, nullptr); if (rc) return rc;
    return viol_write_device(d, threshold, *total);
}

// The sequential rule of checkVarForOffending (gnePivotRules.cpp:431-452) on a candidate list
// that holds, in position order, every violator with |rc| >= floorV.  Returns false when the
// replay could have been influenced by a violator below floorV (see DESIGN.md "LV tie rule"):
//   (1) the first candidate must clear whatever lower elements left behind, and
//   (2) the running value must never come within tolerance of an excluded element.
static bool replay_band_rule(const int64_t *pos, const double *viol, int64_t cnt, double floorV, bool floorIsZero,
                             std::vector<int64_t> &list, double *largest)
{
    list.clear();
    double L = -1.0;                                    // gnePivotRules.cpp:137
    double minL = DBL_MAX;
    for (int64_t i = 0; i < cnt; ++i) {
        const double v = viol[i];
        if (L < v - GNE_ROUNDOFF_TOLERANCE) list.clear();
        if (L <= v + GNE_ROUNDOFF_TOLERANCE) { L = v; list.push_back(pos[i]); if (L < minL) minL = L; }
    }
    *largest = L;
    if (floorIsZero || cnt == 0) return true;           // every violator was in the list
    // largest excluded value is < floorV; excluded element x is pushed iff L <= x + tol
    const double below = nextafter(floorV, 0.0);
    if (!(below < viol[0] - GNE_ROUNDOFF_TOLERANCE)) return false;        // (1)
    if (!(minL > below + GNE_ROUNDOFF_TOLERANCE)) return false;           // (2)
    return true;
}

// The LV pass is the streamed one-launch kernel (gne_stream.cuh) unless the per-tile kernels were asked for
// (gne_dev_set_lv_kernel / GNE_LV_KERNEL=tiles) or the handle's streamed lists keep overflowing (hold-off).
static bool use_stream_kernel(const gne_dev *d)
{
    if (d->lvKernelMode == 1) return false;
    if (d->lvKernelMode == 2) return true;
    return d->streamHoldoff == 0;
}

// Launch on the handle's stream; with `dependent` the kernel may become resident while its predecessor
// still runs (programmatic stream serialization) - it calls pdl_wait() before touching anything the
// predecessor wrote.  Used only between the kernels of one pivot.
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(gne_dev *d, bool dependent, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&... args)
{
    if (!(dependent && d->pdl)) {
        kernel<<<grid, block, smem, d->stream>>>(std::forward<Args>(args)...);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = cudaLaunchConfig_t();
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = d->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// One LV pass over this shard.  Streamed kernel: ONE launch - it applies the queued argument-only update (or `replay`,
// the bench helper's idempotent copy of the last one) and the queued set-mirror writes itself, prices, reduces in its
// last CTA and, sharded over peer mailboxes, exchanges + merges there too (result block in d->lvRes either way).
// Per-tile kernels: queued work is launched first, then tiles + finish (+ a stand-alone exchange by the caller).
// Arguments of one streamed LV pass over this shard (consumes the queued argument-only update and set-mirror writes).
static int fill_stream_pass(gne_dev *d, LvResult *target, bool withExchange, const SmallUpdate *replay, LvPass &P)
{
    SmallUpdate &su = P.u;
    const bool fromPending = !replay && d->havePendingSmall;
    if (replay) su = *replay;
    else if (fromPending) su = d->pendingSmall;
    else { su.ops.n = 0; su.nNodes = su.nTheta = su.nFinal = su.hasPi = 0; }
    NbWriteOps &q = pending_of(d);
    if (q.n > 0) {
        // set-mirror writes queued after the update: fold them in (a later write to a position replaces the earlier one)
        bool fits = true;
        NbWriteOps merged = su.ops;
        for (int k = 0; k < q.n && fits; ++k) {
            int i = -1;
            for (int j = 0; j < merged.n; ++j) if (merged.local[j] == q.local[k]) i = j;
            if (i < 0) { if (merged.n == 8) { fits = false; break; } i = merged.n++; }
            merged.local[i] = q.local[k]; merged.tail[i] = q.tail[k]; merged.head[i] = q.head[k];
            merged.cost[i] = q.cost[k]; merged.mult[i] = q.mult[k]; merged.state[i] = q.state[k];
        }
        if (fits) { su.ops = merged; d->h2dBytes += 25 * q.n; q.n = 0; }
        else {                                       // too many: launch everything queued, the pass carries nothing
            int rc = flush_pending(d); if (rc) return rc;
            if (!replay) { su.ops.n = 0; su.nNodes = su.nTheta = su.nFinal = 0; }
        }
    }
    if (fromPending) d->havePendingSmall = false;
    P.nb = view_of(d);
    P.pi = d->pi;
    P.out = d->lvc;
    P.numTiles = P.nb.end > P.nb.lo ? (P.nb.end - P.nb.lo + ST_TILE - 1) / ST_TILE : 0;
    LvFused &f = P.f;
    memset(&f, 0, sizeof(f));
    f.ticket = d->lvTicket; f.arrive = d->lvArrive; f.tileCtr = d->lvTileCtr; f.updFlag = d->lvUpdFlag;
    f.res = target; f.seq = ++d->lvSeq; f.resOnHost = (target == d->lvRes) ? 1 : 0;
    f.single = P.numTiles <= (int64_t) d->streamGrid ? 1 : 0;
    f.nranks = 1;
    if (withExchange && d->nranks > 1 && d->p2p) {
        // sharded over peer mailboxes: the last CTA exchanges the shard's record, merges the ranks' records and
        // publishes the merged block where the host polls
        f.nranks = d->nranks; f.rank = d->rank; f.peers = (Mailbox *const *) d->peersDev;
        f.hostOut = d->lvRes;
        f.xseq = ++d->xseq; f.xTimeout = d->xTimeoutCycles;
        d->exchangeFused = true;
    }
    NbMut &mm = P.nbMut; mm.tail = d->tail; mm.head = d->head; mm.cost = d->cost; mm.mult = d->mult; mm.state = d->state;
    PotMut &pm = P.pm; pm.a0 = d->a0; pm.a1 = d->a1; pm.slot = d->slot; pm.theta = d->theta; pm.pi = d->pi;
    P.grid = f.single ? (int) std::max<int64_t>(1, P.numTiles) : d->streamGrid;
    return GNE_OK;
}

static int launch_lv_pass(gne_dev *d, LvResult *target, cudaEvent_t evA, cudaEvent_t evB, bool withExchange = true,
                          const SmallUpdate *replay = nullptr)
{
    d->lastWasStream = use_stream_kernel(d);
    d->exchangeFused = false;
    d->passStream = d->stream;
    if (d->lastWasStream) {
        static thread_local LvPass P;
        int rcf = fill_stream_pass(d, target, withExchange, replay, P); if (rcf) return rcf;
        LvFused &f = P.f;
        const int grid = P.grid;
        const int64_t numTiles = P.numTiles;
        (void) numTiles;
#ifdef GNE_STAMPS
        static long long *stampsDev = nullptr, *stampsHost = nullptr;
        static long passes = 0;
        if (!stampsDev) { cudaMalloc(&stampsDev, 128); cudaHostAlloc(&stampsHost, 128, cudaHostAllocDefault); }
        if (passes > 0 && passes % 50 == 0) {
            cudaStreamSynchronize(d->stream);
            cudaMemcpy(stampsHost, stampsDev, 64, cudaMemcpyDeviceToHost);
            const long long *h = stampsHost;
            fprintf(stderr, "[stamps] grid %d tiles %lld: first CTA in -> last CTA at its first gather %lld ns | -> last CTA leaves the loop %lld | ticket %lld | records loaded %lld | merged %lld | published %lld | done %lld  (total %lld ns)\n",
                    grid, (long long) numTiles, h[1] - h[0], h[2] - h[0], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6], h[7] - h[0]);
        }
        ++passes;
        { long long init[8] = { 0x7fffffffffffffffll, 0, 0, 0, 0, 0, 0, 0 }; cudaMemcpyAsync(stampsDev, init, 64, cudaMemcpyHostToDevice, d->stream); }
        f.stamps = stampsDev;
#endif
        if (evA) CK(cudaEventRecord(evA, d->stream));
        st_kernel_for(d->stStages, d->stCtasPerSm)<<<grid, ST_BLOCK, st_smem_bytes(d->stStages), d->stream>>>(P.nb, P.pi, P.out, P.numTiles, f, P.u, P.nbMut, P.pm);
        CK(cudaGetLastError());
        ++d->launches;
        if (evB) CK(cudaEventRecord(evB, d->stream));
    } else {
        int rc = flush_pending(d); if (rc) return rc;
        const int G = tiles_for(d);
        NbView nb = view_of(d);
        if (evA) CK(cudaEventRecord(evA, d->stream));
        if (G > 0) { CK(launch_k(d, true, k_price_lv_tiles, dim3(G), dim3(TILE_THREADS), 0, nb, (const double *) d->pi, d->lv)); ++d->launches; }
        if (evB) CK(cudaEventRecord(evB, d->stream));
        CK(launch_k(d, true, k_price_lv_finish, dim3(1), dim3(FIN_THREADS), 0, d->lv, G, d->lo, target, (unsigned long long) ++d->lvSeq));
        ++d->launches;
        if (d->streamHoldoff > 0) --d->streamHoldoff;
    }
    CK(cudaGetLastError());
    return GNE_OK;
}

static int price_lv_local(gne_dev *d)
{
    // one pass over this shard: result in d->lvRes (mapped pinned) with sequence number d->lvSeq - single GPU, or
    // sharded with the exchange fused into the pass; otherwise in d->lvResDev for the stand-alone exchange
    const bool hostBlock = d->nranks == 1;
    int rc = launch_lv_pass(d, hostBlock ? d->lvRes : d->lvResDev, d->timing ? d->ev0 : nullptr, nullptr);
    if (rc) return rc;
    if (d->timing) CK(cudaEventRecord(d->ev1, d->stream));
    return GNE_OK;
}

// Wait for the result block of the pass just launched.  The reduction kernel publishes a sequence
// number in mapped pinned memory after a system-scope fence, so the host polls that word (a few
// hundred ns of latency) instead of paying a stream synchronisation; the stream is queried now and
// then so that a faulted kernel is reported instead of spinning forever.
static int wait_host_seq(gne_dev *d, const volatile unsigned long long *seq, unsigned long long expect)
{
    unsigned long spins = 0;
    while (*seq != expect) {
        // batches run one host thread per instance, possibly more threads than cores: after the first few
        // microseconds of polling give the core away between polls - by a short sleep when the handle was told
        // so (GNE_POLL_SLEEP_US: the pass of a batch instance takes >100 us, the other instances' threads need the core),
        // else by a yield
        if (++spins > 4096 && (spins & 63) == 0) {
            if (d->pollSleepNs > 0) {
                static thread_local bool slackSet = false;           // the default 50 us timer slack would triple a 20 us sleep
                if (!slackSet) { prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0); slackSet = true; }
                struct timespec ts = { 0, d->pollSleepNs };
                nanosleep(&ts, nullptr);
            }
            else sched_yield();
        }
        if ((spins & 0xfffff) == 0) {                    // now and then
            cudaError_t q = cudaStreamQuery(d->passStream ? d->passStream : d->stream);
            if (q == cudaSuccess) { if (*seq == expect) break; gne_set_error("pricing pass finished without publishing its result"); return GNE_ERR_CUDA; }
            if (q != cudaErrorNotReady) { gne_set_error("pricing pass failed: %s", cudaGetErrorString(q)); return GNE_ERR_CUDA; }
        }
    }
    d->stageInFlight = false;
    return GNE_OK;
}

// The result of the pass just launched, as the host reads it: from the fast header alone when that says everything
// (at most one candidate - the maximum itself - and no overflow code), else from the full block, which the device wrote
// and fenced before the header.
struct LvView { double maxViol; int any, overflow; int64_t count; const int64_t *pos; const double *viol; };

static int wait_lv_result(gne_dev *d, LvView *v)
{
    const volatile LvResult *r = d->lvRes;
    int rc = wait_host_seq(d, &r->hSeqA, d->lvSeq); if (rc) return rc;
    const unsigned long long want = d->lvSeq & ((1ull << 48) - 1ull);
    unsigned long spins = 0;
    while ((r->hTagB >> 16) != want) {                   // the second 16-byte chunk lands with (or right behind) the first
        if (++spins > (1ul << 28)) { gne_set_error("pricing pass published half a result header"); return GNE_ERR_CUDA; }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    const unsigned long long tag = r->hTagB;
    v->any = (int) ((tag >> 1) & 1ull); v->overflow = (int) ((tag >> 2) & 3ull);
    const int64_t hc = (int64_t) ((tag >> 4) & 4095ull);
    if (hc <= 1 && !v->overflow) {
        v->maxViol = r->hMax; v->count = hc;
        v->pos = const_cast<const int64_t *>(reinterpret_cast<const volatile int64_t *>(&r->hPos0));
        v->viol = const_cast<const double *>(&r->hMax);
    } else {
        v->maxViol = r->maxViol; v->count = r->count; v->any = r->any; v->overflow = r->overflow;
        v->pos = const_cast<const int64_t *>(r->pos); v->viol = const_cast<const double *>(r->viol);
    }
    if (d->timing) { CK(cudaEventSynchronize(d->ev1)); CK(cudaEventElapsedTime(&d->lastMs, d->ev0, d->ev1)); }
    else d->lastMs = 0.f;
    return GNE_OK;
}

// Long tie lists (hundreds of thousands of arcs share the maximal |rc| once a large zero-potential
// tree exists in Phase I) are never moved to the host: when every compacted candidate lies within
// tol/2 of the others, the sequential rule keeps all of them (no element can clear, every element
// is pushed) and leaves L = the last one's value, so only the count, the picked element and the last
// element are needed — k_viol_select resolves them from the tile offsets.
static const int64_t LV_DEVICE_LIST_MIN = 4096;

static int launch_shard_exchange(gne_dev *d);
static int sharded_band_fallback(gne_dev *d, double M, std::vector<int64_t> &list, double *largest);

// The LV pass in two halves, so that a host thread can look after another instance while the pass runs:
// gne_price_lv_launch starts it, gne_price_lv_ready says (without blocking) whether its result has arrived,
// gne_price_lv_complete waits for it and resolves the offending list.  gne_price_lv_begin = launch + complete.
extern "C" int gne_price_lv_launch(gne_dev *d)
{
    CK(cudaSetDevice(d->device));
    int rc = price_lv_local(d); if (rc) return rc;       // (queued updates ride with the pass or are launched ahead of it)
    if (d->nranks > 1 && d->p2p) { rc = launch_shard_exchange(d); if (rc) return rc; }
    d->lvLaunched = true;
    return GNE_OK;
}

extern "C" int gne_price_lv_ready(gne_dev *d)
{
    if (!d->lvLaunched) return 1;
    if (d->nranks > 1 && !d->p2p) return 1;              // (the all-gather path blocks in complete)
    const volatile LvResult *r = d->lvRes;
    return r->hSeqA == d->lvSeq && (r->hTagB >> 16) == (d->lvSeq & ((1ull << 48) - 1ull)) ? 1 : 0;
}

// ---- launch groups: the LV passes of several independent instances of one GPU in ONE launch (k_price_lv_batch) ----
// A group belongs to one host thread.  gne_lv_group_add stands for gne_price_lv_launch: the handle's pass is prepared
// (sequence number, folded relabel, counters) and parked in the group; gne_lv_group_launch sends everything parked as one
// grid, each instance on its own share of the CTAs.  ready / complete then work per handle as after a launch of its own.
struct gne_lv_group {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr;
    int maxPending = LVB_MAX;
    int64_t launches = 0, passes = 0;
    LvBatch args;
};

extern "C" int gne_lv_group_create(gne_lv_group **out, int cudaDevice, int maxPending)
{
    if (!out) { gne_set_error("gne_lv_group_create: bad argument"); return GNE_ERR_ARG; }
    const int cnt = gne_cuda_device_count();
    if (cnt <= 0) { gne_set_error("no CUDA device: genneteq_b200 has no CPU fallback"); return GNE_ERR_CUDA; }
    if (cudaDevice < 0 || cudaDevice >= cnt) { gne_set_error("cuda device %d out of range (%d present)", cudaDevice, cnt); return GNE_ERR_ARG; }
    CK(cudaSetDevice(cudaDevice));
    gne_lv_group *g = new gne_lv_group();
    g->device = cudaDevice;
    g->maxPending = (maxPending <= 0 || maxPending > LVB_MAX) ? LVB_MAX : maxPending;
    g->args.count = 0; g->args.ctasPer = 0;
    cudaError_t e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_price_lv_batch<ST_STAGES, ST_CTAS_PER_SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, st_smem_bytes(ST_STAGES));
    if (e != cudaSuccess) { gne_set_error("gne_lv_group_create: %s", cudaGetErrorString(e)); delete g; return GNE_ERR_CUDA; }
    *out = g;
    return GNE_OK;
}

extern "C" int gne_lv_group_launch(gne_lv_group *g)
{
    if (!g) return GNE_ERR_ARG;
    if (g->args.count == 0) return GNE_OK;
    CK(cudaSetDevice(g->device));
    k_price_lv_batch<ST_STAGES, ST_CTAS_PER_SM><<<g->args.count * g->args.ctasPer, ST_BLOCK, st_smem_bytes(ST_STAGES), g->stream>>>(g->args);
    g->args.count = 0; g->args.ctasPer = 0;
    CK(cudaGetLastError());
    ++g->launches;
    return GNE_OK;
}

static int lv_group_add(gne_lv_group *g, gne_dev *d, const SmallUpdate *replay);
extern "C" int gne_lv_group_add(gne_lv_group *g, gne_dev *d) { return lv_group_add(g, d, nullptr); }

static int lv_group_add(gne_lv_group *g, gne_dev *d, const SmallUpdate *replay)
{
    if (!g || !d) return GNE_ERR_ARG;
    if (d->lvLaunched) { gne_set_error("gne_lv_group_add: the handle's previous pass was not completed"); return GNE_ERR_ARG; }
    // passes the batched kernel does not cover go out as launches of their own: sharded handles, the per-tile kernel
    // (short lists, overflow hold-off), timed passes, tuning shapes, another device
    if (d->device != g->device || d->nranks > 1 || d->timing || !use_stream_kernel(d) || d->stStages != ST_STAGES || d->stCtasPerSm != ST_CTAS_PER_SM)
        return gne_price_lv_launch(d);
    CK(cudaSetDevice(d->device));
    d->lastWasStream = true;
    d->exchangeFused = false;
    LvPass &P = g->args.p[g->args.count];
    int rc = fill_stream_pass(d, d->lvRes, false, replay, P); if (rc) return rc;
    // work still queued on the handle's own stream (bulk uploads, a flushed update) precedes the pass
    if (cudaStreamQuery(d->stream) == cudaErrorNotReady) { CK(cudaEventRecord(g->ev, d->stream)); CK(cudaStreamWaitEvent(g->stream, g->ev, 0)); }
    d->passStream = g->stream;
    d->lvLaunched = true;
    ++d->launches; ++g->passes;
    g->args.ctasPer = std::max(g->args.ctasPer, P.grid);
    if (++g->args.count >= g->maxPending) return gne_lv_group_launch(g);
    return GNE_OK;
}

extern "C" int gne_lv_group_pending(const gne_lv_group *g) { return g ? g->args.count : 0; }
extern "C" int64_t gne_lv_group_launch_count(const gne_lv_group *g) { return g ? g->launches : 0; }

extern "C" int gne_lv_group_destroy(gne_lv_group *g)
{
    if (!g) return GNE_OK;
    cudaSetDevice(g->device);
    if (g->stream) { cudaStreamSynchronize(g->stream); cudaStreamDestroy(g->stream); }
    if (g->ev) cudaEventDestroy(g->ev);
    delete g;
    return GNE_OK;
}

extern "C" int gne_price_lv_begin(gne_dev *d, double *largestViolation, int64_t *count, int *any)
{
    int rc = gne_price_lv_launch(d); if (rc) return rc;
    return gne_price_lv_complete(d, largestViolation, count, any);
}

extern "C" int gne_price_lv_complete(gne_dev *d, double *largestViolation, int64_t *count, int *any)
{
    if (!d->lvLaunched) { gne_set_error("gne_price_lv_complete: no pass was launched"); return GNE_ERR_ARG; }
    d->lvLaunched = false;
    CK(cudaSetDevice(d->device));
    int rc = GNE_OK;
    d->lvMode = 0; d->lvList.clear(); d->lvCount = 0;
    double L = -1.0;
    int anyV = 0;
    if (d->nranks > 1 && !d->p2p) {
        // records through ncclAllGather / the host program's all-gather, merged on the host
        rc = gne_dev_price_lv_sharded(d, &L, d->lvList, &anyV); if (rc) return rc;
    } else {
        // single GPU, or sharded over peer mailboxes (the merged block arrives like a single GPU's)
        LvView view;
        rc = wait_lv_result(d, &view); if (rc) return rc;
        const LvView *r = &view;
        if (r->overflow == XCHG_TIMEOUT) { gne_set_error("sharded exchange timed out waiting for a peer rank"); return GNE_ERR_COMM; }
        d->d2hBytes += 24 + 16 * std::min<int64_t>(r->count, LV_LIST_CAP);
        anyV = r->any;
        bool ok = !anyV;
        if (d->lastWasStream && d->nranks == 1) {
            // a streamed pass whose per-CTA lists overflow repeatedly (values rising along the list) is
            // cheaper through the per-tile kernel: hold the streamed kernel off for a while
            if (anyV && r->overflow) { if (++d->streamOverflows >= 3) { d->streamHoldoff = 256; d->streamOverflows = 0; } }
            else d->streamOverflows = 0;
        }
        if (anyV && !r->overflow) ok = replay_band_rule(r->pos, r->viol, r->count, r->maxViol - LV_BAND, false, d->lvList, &L);
        const double M = anyV ? r->maxViol : 0.0;
        if (!ok && d->nranks > 1) { rc = sharded_band_fallback(d, M, d->lvList, &L); if (rc) return rc; ok = true; }
        // fallback: widen the band until the replay is provably exact
        double width = 16.0 * GNE_ROUNDOFF_TOLERANCE;
        while (!ok) {
            double floorV = M - width;
            bool zero = false;
            if (!(floorV > 0.0) || width > 1.0e-6) { floorV = 0.0; zero = true; }
            int64_t total = 0;
            double vmin = 0.0, vmax = 0.0;
            rc = viol_count_device(d, floorV, &total, &vmin, &vmax); if (rc) return rc;
            const double tol = GNE_ROUNDOFF_TOLERANCE;
            const double below = zero ? 0.0 : nextafter(floorV, 0.0);
            // within tol/4 of each other and small enough that rounding (ulp <= tol/16) cannot tip a comparison
            const bool tight = (vmax == vmin) || ((vmax - vmin) <= 0.25 * tol && vmax < 16.0);
            const bool sealed = zero || ((below < vmin - tol) && (vmin > below + tol));
            if (total >= LV_DEVICE_LIST_MIN && tight && sealed) {
                // every candidate stays in the list; resolve the last one (the final running value)
                k_viol_select<<<2, TILE_THREADS, 0, d->stream>>>(view_of(d), d->pi, floorV, d->tileOff, tiles_for(d), 0, d->selRes);
                CK(cudaGetLastError());
                ++d->launches;
                CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
                d->d2hBytes += sizeof(SelectResult);
                L = d->selRes->lastViol;
                d->lvMode = 1; d->lvCount = total; d->lvFloor = floorV;
                ok = true;
                break;
            }
            rc = viol_write_device(d, floorV, total); if (rc) return rc;
            std::vector<int64_t> hp((size_t) total);
            std::vector<double> hv((size_t) total);
            if (total > 0) {
                CK(cudaMemcpyAsync(hp.data(), d->outPos, 8 * (size_t) total, cudaMemcpyDeviceToHost, d->stream));
                CK(cudaMemcpyAsync(hv.data(), d->outViol, 8 * (size_t) total, cudaMemcpyDeviceToHost, d->stream));
                CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
                d->d2hBytes += 16 * total;
            }
            ok = replay_band_rule(hp.data(), hv.data(), total, floorV, zero, d->lvList, &L);
            width *= 64.0;
        }
    }
    if (d->lvMode == 0) d->lvCount = (int64_t) d->lvList.size();
    *largestViolation = L;
    *any = anyV;
    *count = d->lvCount;
    return GNE_OK;
}

extern "C" int gne_price_lv_at(gne_dev *d, int64_t index, int64_t *pos)
{
    if (index < 0 || index >= d->lvCount) { gne_set_error("gne_price_lv_at: index out of range"); return GNE_ERR_ARG; }
    if (d->lvMode == 0) { *pos = d->lvList[(size_t) index]; return GNE_OK; }
    CK(cudaSetDevice(d->device));
    k_viol_select<<<1, TILE_THREADS, 0, d->stream>>>(view_of(d), d->pi, d->lvFloor, d->tileOff, tiles_for(d), index, d->selRes);
    CK(cudaGetLastError());
    ++d->launches;
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    d->d2hBytes += 16;
    *pos = d->selRes->pos;
    return GNE_OK;
}

extern "C" int gne_price_lv(gne_dev *d, double *largestViolation, int64_t *count, int64_t *listPos, int64_t cap, int *any)
{
    int rc = gne_price_lv_begin(d, largestViolation, count, any); if (rc) return rc;
    if (*count > cap) { gne_set_error("gne_price_lv: list of %lld exceeds cap", (long long) *count); return GNE_ERR_CAPACITY; }
    if (d->lvMode == 0) {
        if (*count > 0) memcpy(listPos, d->lvList.data(), 8 * (size_t) *count);
        return GNE_OK;
    }
    // the list lives on the device (every compacted candidate is a member): materialise it
    rc = viol_write_device(d, d->lvFloor, d->lvCount); if (rc) return rc;
    CK(cudaMemcpyAsync(listPos, d->outPos, 8 * (size_t) d->lvCount, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    d->d2hBytes += 8 * d->lvCount;
    return GNE_OK;
}

extern "C" int gne_price_violators(gne_dev *d, double threshold, int64_t *count, int64_t *listPos, double *listViol, int64_t cap)
{
    CK(cudaSetDevice(d->device));
    int rc = gne_dev_flush(d); if (rc) return rc;
    CK(cudaEventRecord(d->ev0, d->stream));
    int64_t total = 0;
    rc = viol_compact_device(d, threshold, &total); if (rc) return rc;
    CK(cudaEventRecord(d->ev1, d->stream));
    if (d->nranks > 1) {
        std::vector<int64_t> gp;
        std::vector<double> gv;
        rc = sharded_gather_lists(d, total, gp, gv); if (rc) return rc;
        *count = (int64_t) gp.size();
        d->lastMs = 0.f;
        if (*count > cap) { gne_set_error("gne_price_violators: %lld exceeds cap", (long long) *count); return GNE_ERR_CAPACITY; }
        if (*count > 0) { memcpy(listPos, gp.data(), 8 * gp.size()); if (listViol) memcpy(listViol, gv.data(), 8 * gv.size()); }
        return GNE_OK;
    }
    *count = total;
    if (total > cap) { CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false; gne_set_error("gne_price_violators: %lld exceeds cap", (long long) total); return GNE_ERR_CAPACITY; }
    if (total > 0) {
        CK(cudaMemcpyAsync(listPos, d->outPos, 8 * (size_t) total, cudaMemcpyDeviceToHost, d->stream));
        if (listViol) CK(cudaMemcpyAsync(listViol, d->outViol, 8 * (size_t) total, cudaMemcpyDeviceToHost, d->stream));
        d->d2hBytes += (listViol ? 16 : 8) * total;
    }
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    CK(cudaEventElapsedTime(&d->lastMs, d->ev0, d->ev1));
    return GNE_OK;
}

static int exchange_blobs(gne_dev *d, const unsigned long long *mine, unsigned long long *all);

extern "C" int gne_price_first(gne_dev *d, int64_t *pos)
{
    CK(cudaSetDevice(d->device));
    int rc = gne_dev_flush(d); if (rc) return rc;
    const int G = tiles_for(d);
    CK(cudaEventRecord(d->ev0, d->stream));
    CK(cudaMemsetAsync(d->firstPos, 0xff, 8, d->stream));
    if (G > 0) { k_price_first<<<G, TILE_THREADS, 0, d->stream>>>(view_of(d), d->pi, d->firstPos); ++d->launches; }
    CK(cudaMemcpyAsync(&d->hostScalar[1], d->firstPos, 8, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaEventRecord(d->ev1, d->stream));
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    CK(cudaEventElapsedTime(&d->lastMs, d->ev0, d->ev1));
    d->d2hBytes += 8;
    unsigned long long v = (unsigned long long) d->hostScalar[1];
    if (d->nranks > 1) {
        // sharded list: the first violator is the lowest of the shards' first positions (one 128-byte blob per rank)
        unsigned long long mine[BLOB_WORDS] = { 0 }, all[BLOB_WORDS * MAX_RANKS];
        mine[0] = v;
        rc = exchange_blobs(d, mine, all); if (rc) return rc;
        for (int r = 0; r < d->nranks; ++r) v = std::min(v, all[(size_t) r * BLOB_WORDS]);
    }
    *pos = (v == ~0ull) ? -1 : (int64_t) v;
    return GNE_OK;
}

// Reduced cost of p equal-flow variables (computeEqfRedCost, genneteq.h:321-330) against the device's potentials.
// Rows in CSR form, entries of a row sorted by node.  The rows travel with the call (they change only when the set of
// nonbasic equal-flow variables does; p and the rows are small next to the arc list).
extern "C" int gne_price_eqf(gne_dev *d, int p, const int64_t *rowStart, const int32_t *node, const double *val,
                             const double *cost, double *rc)
{
    if (p < 0 || (p > 0 && (!rowStart || !cost || !rc))) { gne_set_error("gne_price_eqf: bad argument"); return GNE_ERR_ARG; }
    if (p == 0) return GNE_OK;
    const int64_t nnz = rowStart[p];
    if (rowStart[0] != 0 || nnz < 0) { gne_set_error("gne_price_eqf: rowStart must start at 0 and rise"); return GNE_ERR_ARG; }
    for (int k = 0; k < p; ++k) if (rowStart[k + 1] < rowStart[k]) { gne_set_error("gne_price_eqf: rowStart must rise"); return GNE_ERR_ARG; }
    for (int64_t q = 0; q < nnz; ++q) if (node[q] < 0 || node[q] >= d->n) { gne_set_error("gne_price_eqf: node out of range"); return GNE_ERR_ARG; }
    CK(cudaSetDevice(d->device));
    int rcf = gne_dev_flush(d); if (rcf) return rcf;
    const size_t bytes = 8 * (size_t) (p + 1) + 12 * (size_t) nnz + 16 * (size_t) p + 8 + 64;
    rcf = ensure_stage(d, bytes); if (rcf) return rcf;
    if (d->stageInFlight) { CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false; }
    char *h = (char *) d->stageHost;
    size_t o = 0;
    const size_t oRow = o; memcpy(h + o, rowStart, 8 * (size_t) (p + 1)); o += 8 * (size_t) (p + 1);
    const size_t oVal = o; if (nnz) memcpy(h + o, val, 8 * (size_t) nnz); o += 8 * (size_t) nnz;
    const size_t oCost = o; memcpy(h + o, cost, 8 * (size_t) p); o += 8 * (size_t) p;
    const size_t oRc = o; o += 8 * (size_t) p;
    const size_t oCnt = o; memset(h + o, 0, 8); o += 8;       // non-finite potentials (see k_price_eqf)
    const size_t oNode = o; if (nnz) memcpy(h + o, node, 4 * (size_t) nnz); o += 4 * (size_t) nnz;
    CK(cudaMemcpyAsync(d->stageDev, h, o, cudaMemcpyHostToDevice, d->stream));
    char *g = (char *) d->stageDev;
    k_count_nonfinite<<<blocks_for(d->n, 256), 256, 0, d->stream>>>(d->n, d->pi, (unsigned long long *) (g + oCnt));
    k_price_eqf<<<blocks_for(p, 128), 128, 0, d->stream>>>(p, (const int64_t *) (g + oRow), (const int32_t *) (g + oNode), (const double *) (g + oVal),
                                                          (const double *) (g + oCost), d->pi, (const unsigned long long *) (g + oCnt), (double *) (g + oRc));
    CK(cudaGetLastError());
    d->launches += 2;
    CK(cudaMemcpyAsync(rc, g + oRc, 8 * (size_t) p, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    d->h2dBytes += (int64_t) o; d->d2hBytes += 8 * p;
    return GNE_OK;
}

// The rows of ALL p equal-flow variables, once; gne_price_eqf_launch then names the ones to price.
extern "C" int gne_eqf_set_rows(gne_dev *d, int p, const int64_t *rowStart, const int32_t *node, const double *val)
{
    if (!d || p <= 0 || !rowStart || rowStart[0] != 0) { gne_set_error("gne_eqf_set_rows: bad argument"); return GNE_ERR_ARG; }
    const int64_t nnz = rowStart[p];
    for (int k = 0; k < p; ++k) if (rowStart[k + 1] < rowStart[k]) { gne_set_error("gne_eqf_set_rows: rowStart must rise"); return GNE_ERR_ARG; }
    for (int64_t q = 0; q < nnz; ++q) if (node[q] < 0 || node[q] >= d->n) { gne_set_error("gne_eqf_set_rows: node out of range"); return GNE_ERR_ARG; }
    CK(cudaSetDevice(d->device));
    CK(cudaStreamSynchronize(d->stream));
    if (d->eqfRowStart) { cudaFree(d->eqfRowStart); d->eqfRowStart = nullptr; }
    if (d->eqfNode) { cudaFree(d->eqfNode); d->eqfNode = nullptr; }
    if (d->eqfVal) { cudaFree(d->eqfVal); d->eqfVal = nullptr; }
    d->eqfP = 0;
    CK(cudaMalloc(&d->eqfRowStart, 8 * (size_t) (p + 1)));
    CK(cudaMalloc(&d->eqfNode, 4 * (size_t) std::max<int64_t>(nnz, 1)));
    CK(cudaMalloc(&d->eqfVal, 8 * (size_t) std::max<int64_t>(nnz, 1)));
    CK(cudaMemcpy(d->eqfRowStart, rowStart, 8 * (size_t) (p + 1), cudaMemcpyHostToDevice));
    if (nnz) { CK(cudaMemcpy(d->eqfNode, node, 4 * (size_t) nnz, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d->eqfVal, val, 8 * (size_t) nnz, cudaMemcpyHostToDevice)); }
    if (!d->eqfOut) { CK(cudaHostAlloc(&d->eqfOut, 8 * (EQF_CALL_MAX + 8), cudaHostAllocMapped)); memset(d->eqfOut, 0, 8 * (EQF_CALL_MAX + 8)); }
    d->eqfP = p;
    d->h2dBytes += 8 * (p + 1) + 12 * nnz;
    return GNE_OK;
}

// Reduced cost of the q listed sets against the device's potentials: one kernel, queued behind whatever the stream holds;
// gne_price_eqf_fetch waits for its sequence word (the arc pass may be launched in between).  GNE_ERR_CAPACITY: more than
// EQF_CALL_MAX sets or no resident rows - the caller uses gne_price_eqf.
extern "C" int gne_price_eqf_launch(gne_dev *d, int q, const int32_t *setIds, const double *cost, int64_t nonFinitePotentials)
{
    if (!d || q < 0 || (q > 0 && (!setIds || !cost)) || nonFinitePotentials < 0) { gne_set_error("gne_price_eqf_launch: bad argument"); return GNE_ERR_ARG; }
    if (q == 0) return GNE_OK;
    if (q > EQF_CALL_MAX || d->eqfP <= 0) { gne_set_error("gne_price_eqf_launch: %d sets (at most %d) or no resident rows", q, EQF_CALL_MAX); return GNE_ERR_CAPACITY; }
    static thread_local EqfCall c;
    c.q = q; c.nonFinite = (unsigned long long) nonFinitePotentials; c.seq = ++d->eqfSeq;
    for (int k = 0; k < q; ++k) {
        if (setIds[k] < 0 || setIds[k] >= d->eqfP) { gne_set_error("gne_price_eqf_launch: set %d out of range", setIds[k]); return GNE_ERR_ARG; }
        c.id[k] = setIds[k]; c.cost[k] = cost[k];
    }
    CK(cudaSetDevice(d->device));
    int rcf = flush_pending(d); if (rcf) return rcf;
    k_price_eqf_resident<<<1, EQF_CALL_MAX, 0, d->stream>>>(c, d->eqfRowStart, d->eqfNode, d->eqfVal, d->pi, d->eqfOut,
                                                           (volatile unsigned long long *) (d->eqfOut + EQF_CALL_MAX));
    CK(cudaGetLastError());
    ++d->launches;
    d->h2dBytes += 12 * q;
    return GNE_OK;
}

extern "C" int gne_price_eqf_fetch(gne_dev *d, int q, double *rc)
{
    if (!d || q < 0 || (q > 0 && !rc) || q > EQF_CALL_MAX) { gne_set_error("gne_price_eqf_fetch: bad argument"); return GNE_ERR_ARG; }
    if (q == 0) return GNE_OK;
    if (!d->eqfOut || d->eqfSeq == 0) { gne_set_error("gne_price_eqf_fetch: nothing was launched"); return GNE_ERR_ARG; }
    const cudaStream_t keep = d->passStream;
    d->passStream = d->stream;                           // (the kernel ran on the handle's own stream: that is the one to query)
    int rcw = wait_host_seq(d, (const volatile unsigned long long *) (d->eqfOut + EQF_CALL_MAX), d->eqfSeq);
    d->passStream = keep;
    if (rcw) return rcw;
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    for (int k = 0; k < q; ++k) rc[k] = const_cast<const volatile double *>(d->eqfOut)[k];
    d->d2hBytes += 8 * q;
    return GNE_OK;
}

extern "C" int gne_reduced_costs(gne_dev *d, int64_t lo, int64_t hi, double *rc)
{
    if (lo < d->lo || hi > std::min(d->N, d->hi) || hi < lo) { gne_set_error("gne_reduced_costs: bad range"); return GNE_ERR_ARG; }
    CK(cudaSetDevice(d->device));
    int r = gne_dev_flush(d); if (r) return r;
    if (hi == lo) return GNE_OK;
    r = ensure_out(d, hi - lo); if (r) return r;
    const int G = tiles_for(d);
    k_reduced_costs<<<G, TILE_THREADS, 0, d->stream>>>(view_of(d), d->pi, lo, hi, d->outViol);
    CK(cudaGetLastError());
    ++d->launches;
    CK(cudaMemcpyAsync(rc, d->outViol, 8 * (size_t) (hi - lo), cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    d->d2hBytes += 8 * (hi - lo);
    return GNE_OK;
}

// Windowed scan over the virtual indices j in [0, jTotal), pos = (cursor + j) mod N, until `want`
// violators were seen or jTotal is reached; result in d->qsState (mapped pinned).
static int qs_scan(gne_dev *d, int64_t cursor, int64_t N, int64_t jTotal, int64_t want, float *msOut)
{
    NbView nb = view_of(d);
    QsState *st = d->qsState;                           // host copy of the result (mapped pinned)
    st->violations = 0; st->bestV = 0.0; st->bestJ = -1; st->stopJ = -1; st->done = 0;
    if (msOut) *msOut = 0.f;
    if (jTotal <= 0 || want <= 0) { st->done = 1; return GNE_OK; }
    if (!d->qsDev) CK(cudaMalloc(&d->qsDev, sizeof(QsState)));
    if (d->qsWindow <= 0) d->qsWindow = std::max<int64_t>(4 * std::min<int64_t>(want, jTotal), 64 * QTILE);
    int64_t jLo = 0;
    int first = 1;
    while (true) {
        const int64_t jHi = std::min(jTotal, jLo + d->qsWindow);
        const int G = (int) ((jHi - jLo + QTILE - 1) / QTILE);
        if (d->timing) CK(cudaEventRecord(d->ev0, d->stream));
        d->qsArcsLaunched += jHi - jLo;
        k_qs_tiles<<<G, TILE_THREADS, 0, d->stream>>>(nb, d->pi, cursor, N, jLo, jHi, d->qs);
        k_qs_finish<<<1, TILE_THREADS, 0, d->stream>>>(nb, d->pi, cursor, N, jLo, jHi, jTotal, d->qs, G, want, d->qsDev, st, first, ++d->qsSeq);
        if (d->timing) CK(cudaEventRecord(d->ev1, d->stream));
        CK(cudaGetLastError());
        d->launches += 2;
        first = 0;
        // spin on the sequence number the finish kernel publishes (see wait_lv_result)
        { int rcw = wait_host_seq(d, &st->seq, d->qsSeq); if (rcw) return rcw; }
        if (d->timing && msOut) { float ms; CK(cudaEventSynchronize(d->ev1)); CK(cudaEventElapsedTime(&ms, d->ev0, d->ev1)); *msOut += ms; }
        d->d2hBytes += sizeof(QsState);
        if (st->done) {
            const int64_t sc = (st->stopJ >= 0) ? st->stopJ + 1 : jTotal;
            // adapt the next window to ~1.5x what this scan needed
            d->qsWindow = std::min<int64_t>(std::max<int64_t>(N, QTILE), std::max<int64_t>(64 * QTILE, ((sc * 3 / 2) / QTILE + 1) * QTILE));
            break;
        }
        jLo = jHi;
        d->qsWindow = std::min<int64_t>(std::max<int64_t>(N, QTILE), d->qsWindow * 4);
    }
    return GNE_OK;
}

static int price_qs_sharded(gne_dev *d, int64_t cursor, int64_t want, int64_t *bestPos, double *bestViol,
                            int64_t *newCursor, int64_t *scanned, int64_t *violations);

extern "C" int gne_price_qs(gne_dev *d, int64_t cursor, int64_t want, int64_t *bestPos, double *bestViol,
                            int64_t *newCursor, int64_t *scanned, int64_t *violations)
{
    CK(cudaSetDevice(d->device));
    int rc = gne_dev_flush(d); if (rc) return rc;
    const int64_t N = d->N;
    *bestPos = -1; *bestViol = 0.0; *scanned = 0; *violations = 0;
    if (cursor >= N) cursor = 0;                        // gnePivotRules.cpp:716
    *newCursor = cursor;
    if (N == 0 || want <= 0) return GNE_OK;
    if (d->nranks != 1) return price_qs_sharded(d, cursor, want, bestPos, bestViol, newCursor, scanned, violations);
    rc = qs_scan(d, cursor, N, N, want, &d->lastMs); if (rc) return rc;
    const QsState *st = d->qsState;
    *scanned = (st->stopJ >= 0) ? st->stopJ + 1 : N;
    *violations = st->violations;
    *bestViol = st->bestV;
    if (st->bestJ >= 0) { int64_t p = cursor + st->bestJ; if (p >= N) p -= N; *bestPos = p; }
    int64_t nc = cursor + *scanned;
    if (nc >= N) nc -= N;
    *newCursor = nc;                                    // == start when the scan was exhausted
    return GNE_OK;
}

// updateCandidateList2 + tryToAddArc (gnePivotRules.cpp:613-689 with no equal-flow sets): scan from
// the rolling cursor, always at least minScan arcs, until `want` violators are listed or the list
// wrapped; returns the violators in SCAN order with |rc| (the stored redCost the first selection
// pass reads, :571-578).
extern "C" int gne_price_scan_list(gne_dev *d, int64_t cursor, int64_t want, int64_t minScan, int64_t *count,
                                   int64_t *listPos, double *listViol, int64_t cap, int64_t *newCursor, int64_t *scanned)
{
    CK(cudaSetDevice(d->device));
    const int64_t N = d->N;
    *count = 0; *scanned = 0;
    if (cursor >= N) cursor = 0;                        // :622
    *newCursor = cursor;
    if (N == 0) return GNE_OK;
    int64_t bestPos, nc, sc = 0, viol = 0;
    double bestViol;
    if (want > 0) { int rc = gne_price_qs(d, cursor, want, &bestPos, &bestViol, &nc, &sc, &viol); if (rc) return rc; }
    sc = std::min(N, std::max(sc, std::min(minScan, N)));     // the first minScan arcs are always tried (:632-636)
    // violators of the scanned virtual range, in scan order: [cursor, min(N, cursor + sc)) then [0, rest)
    const int64_t hi1 = std::min(N, cursor + sc), rest = cursor + sc - hi1;
    int64_t got = 0;
    const int64_t seg[2][2] = { { cursor, hi1 }, { 0, rest } };
    for (int k = 0; k < 2; ++k) {
        if (seg[k][1] <= seg[k][0]) continue;
        int64_t total = 0;
        int rc = viol_count_device(d, 0.0, &total, nullptr, nullptr, seg[k][0], seg[k][1]); if (rc) return rc;
        if (d->nranks > 1) {
            // sharded list: every rank compacts its part of the segment; the parts concatenated in rank order are the
            // segment in position (= scan) order
            rc = viol_write_device(d, 0.0, total, seg[k][0], seg[k][1]); if (rc) return rc;
            std::vector<int64_t> gp;
            std::vector<double> gv;
            rc = sharded_gather_lists(d, total, gp, gv); if (rc) return rc;
            total = (int64_t) gp.size();
            if (got + total > cap) { *count = got + total; gne_set_error("gne_price_scan_list: cap too small"); return GNE_ERR_CAPACITY; }
            if (total > 0) { memcpy(listPos + got, gp.data(), 8 * (size_t) total); memcpy(listViol + got, gv.data(), 8 * (size_t) total); }
            got += total;
            continue;
        }
        if (got + total > cap) { *count = got + total; gne_set_error("gne_price_scan_list: cap too small"); return GNE_ERR_CAPACITY; }
        rc = viol_write_device(d, 0.0, total, seg[k][0], seg[k][1]); if (rc) return rc;
        if (total > 0) {
            CK(cudaMemcpyAsync(listPos + got, d->outPos, 8 * (size_t) total, cudaMemcpyDeviceToHost, d->stream));
            CK(cudaMemcpyAsync(listViol + got, d->outViol, 8 * (size_t) total, cudaMemcpyDeviceToHost, d->stream));
            CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
            d->d2hBytes += 16 * total;
        }
        got += total;
    }
    *count = got;
    *scanned = sc;
    int64_t c2 = cursor + sc;
    if (c2 >= N) c2 -= N;
    *newCursor = c2;
    return GNE_OK;
}

// ------------------------------------------------------------------------------------------
// bench helper
// ------------------------------------------------------------------------------------------

#ifdef GNE_TUNING
// tuning harness: the LV tile kernel and its experimental variants, each timed `reps` times
static const char *g_variantNames[] = {
    "production k_price_lv_tiles", "tiles<2 chunks, minb 5>", "tiles<2,6>", "tiles<2,8>", "tiles<1,8>", "tiles<4,4>",
    "tiles<2,6,no_allocate>", "tiles<1,8,no_allocate>", "persistent<minb 6> 148x6", "persistent<8> 148x8",
    "persistent<8,no_allocate> 148x8", "strided<4,8>", "strided<8,6>",
    "tma<256thr,4it,3st> 148x2", "tma<256thr,4it,4st> 148x2", "tma<256thr,2it,6st> 148x2", "tma<256thr,2it,4st> 148x4", "tma<256thr,2it,3st> 148x5",
    "tma<512thr,2it,4st> 148x2", "tma<128thr,4it,4st> 148x4", "tma<256thr,1it,8st> 148x4", "tma<512thr,2it,3st> 148x2", "tma<256thr,2it,8st> 148x2",
    "tma<256thr,4it,3st> 148x1", "production k_price_lv_stream", "k_price_lv_finish (per-tile records, mapped host result)",
    "(unused)", "production k_price_lv_stream, result block in device memory", "k_update_small replay" };
extern "C" int gne_bench_variants(gne_dev *d, int reps, float *msOut, int maxVariants, int *nVariants)
{
    CK(cudaSetDevice(d->device));
    int rc = gne_dev_flush(d); if (rc) return rc;
    const int NV = (int) (sizeof(g_variantNames) / sizeof(g_variantNames[0]));
    *nVariants = NV;
    if (maxVariants < NV) return GNE_ERR_CAPACITY;
    NbView nb = view_of(d);
    const int64_t len = std::min(d->N, d->hi) - d->lo;
    const int G2 = (int) ((len + 2047) / 2048), G1 = (int) ((len + 1023) / 1024), G4 = (int) ((len + 4095) / 4096);
    rc = ensure_out(d, G1 + 4096); if (rc) return rc;
    double *o = d->outViol;
    for (int v = 0; v < NV; ++v) {
        float best = 1e30f, sum = 0.f;
        for (int r = 0; r < reps + 2; ++r) {
            CK(cudaEventRecord(d->ev0, d->stream));
            switch (v) {
              case 0: k_price_lv_tiles<<<(int) ((len + TILE - 1) / TILE), TILE_THREADS, 0, d->stream>>>(nb, d->pi, d->lv); break;
              case 1: k_var_tiles<2, 5, false><<<G2, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o); break;
              case 2: k_var_tiles<2, 6, false><<<G2, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o); break;
              case 3: k_var_tiles<2, 8, false><<<G2, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o); break;
              case 4: k_var_tiles<1, 8, false><<<G1, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o); break;
              case 5: k_var_tiles<4, 4, false><<<G4, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o); break;
              case 6: k_var_tiles<2, 6, true><<<G2, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o); break;
              case 7: k_var_tiles<1, 8, true><<<G1, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o); break;
              case 8: k_var_persistent<6, false><<<148 * 6, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o, G1); break;
              case 9: k_var_persistent<8, false><<<148 * 8, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o, G1); break;
              case 10: k_var_persistent<8, true><<<148 * 8, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o, G1); break;
              case 11: k_var_strided<4, 8><<<G1, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o); break;
              case 12: k_var_strided<8, 6><<<G2, TILE_THREADS, 0, d->stream>>>(nb, d->pi, o); break;
#define GNE_TMA_CASE(ID, THR, IT, ST, MINB, CTAS)                                                                      \
              case ID: {                                                                                               \
                  const size_t sm = 128 + (size_t) ST * 25 * THR * IT;                                                 \
                  cudaFuncSetAttribute(k_var_tma<THR, IT, ST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm); \
                  k_var_tma<THR, IT, ST, MINB><<<148 * CTAS, THR, sm, d->stream>>>(nb, d->pi, o, len / (THR * IT));     \
              } break;
              GNE_TMA_CASE(13, 256, 4, 3, 2, 2)
              GNE_TMA_CASE(14, 256, 4, 4, 2, 2)
              GNE_TMA_CASE(15, 256, 2, 6, 2, 2)
              GNE_TMA_CASE(16, 256, 2, 4, 4, 4)
              GNE_TMA_CASE(17, 256, 2, 3, 5, 5)
              GNE_TMA_CASE(18, 512, 2, 4, 2, 2)
              GNE_TMA_CASE(19, 128, 4, 4, 4, 4)
              GNE_TMA_CASE(20, 256, 1, 8, 4, 4)
              GNE_TMA_CASE(21, 512, 2, 3, 2, 2)
              GNE_TMA_CASE(22, 256, 2, 8, 2, 2)
              GNE_TMA_CASE(23, 256, 4, 3, 1, 1)
#undef GNE_TMA_CASE
              case 24: case 26: case 27: {
                  // 24 / 27: the production one-launch pass (no folded update; result block on the host / on the device); 26: unused
                  if (v == 26) break;
                  if (!d->lvResDev) cudaMalloc(&d->lvResDev, sizeof(LvResult));
                  const int keep = d->lvKernelMode; d->lvKernelMode = 2;
                  launch_lv_pass(d, v == 27 ? d->lvResDev : d->lvRes, nullptr, nullptr, false);
                  d->lvKernelMode = keep;
              } break;
              case 25: k_price_lv_finish<<<1, FIN_THREADS, 0, d->stream>>>(d->lv, (int) ((len + TILE - 1) / TILE), d->lo, d->lvRes, ++d->lvSeq); break;
              case 28: { NbMut mm; mm.tail = d->tail; mm.head = d->head; mm.cost = d->cost; mm.mult = d->mult; mm.state = d->state;
                         k_update_small<<<1, 256, 0, d->stream>>>(mm, d->lastSmall, d->a0, d->a1, d->slot, d->theta, d->pi); } break;
            }
            CK(cudaEventRecord(d->ev1, d->stream));
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
            float ms; CK(cudaEventElapsedTime(&ms, d->ev0, d->ev1));
            if (r >= 2) { sum += ms; best = std::min(best, ms); }
            ++d->launches;
        }
        msOut[2 * v] = sum / reps; msOut[2 * v + 1] = best;
    }
    return GNE_OK;
}
extern "C" const char *gne_bench_variant_name(int i)
{
    const int NV = (int) (sizeof(g_variantNames) / sizeof(g_variantNames[0]));
    return (i >= 0 && i < NV) ? g_variantNames[i] : "";
}

#endif // GNE_TUNING

extern "C" int gne_bench_price_lv(gne_dev *d, int reps, int flushL2, int withFinalize, float *msStep, float *msTiles)
{
    CK(cudaSetDevice(d->device));
    int rc = gne_dev_flush(d); if (rc) return rc;
    if (flushL2 && !d->flushBuf) {
        d->flushLen = (int64_t) 40 * 1024 * 1024;       // 320 MB of doubles > 126 MB L2
        CK(cudaMalloc(&d->flushBuf, 8 * (size_t) d->flushLen));
    }
    if (!d->ev2) { CK(cudaEventCreate(&d->ev2)); CK(cudaEventCreate(&d->ev3)); }
    while ((int) d->benchEv.size() < 2 * reps) { cudaEvent_t e; CK(cudaEventCreate(&e)); d->benchEv.push_back(e); }
    // the relabel of a typical pivot of this window: the last argument-only update, replayed (idempotent).  The
    // streamed kernel applies it itself (one launch per step, as in the solver); the per-tile path launches it first.
    const bool replay = withFinalize == 2 && d->haveLastSmall;
    auto relabel = [&]() {
        if (replay) {
            if (use_stream_kernel(d)) return;
            NbMut mm; mm.tail = d->tail; mm.head = d->head; mm.cost = d->cost; mm.mult = d->mult; mm.state = d->state;
            k_update_small<<<1, 256, 0, d->stream>>>(mm, d->lastSmall, d->a0, d->a1, d->slot, d->theta, d->pi);
            ++d->launches;
        } else if (withFinalize) { k_pot_finalize<<<blocks_for(d->n, 256), 256, 0, d->stream>>>(d->n, d->a0, d->a1, d->slot, d->theta, d->pi); ++d->launches; }
    };
    // The steps are queued back to back and the host synchronises once per loop: per-step host wake-ups
    // would put their jitter into every cross-rank exchange of a sharded run (the early rank waits inside
    // its kernel for the late one).  Loop 0 brackets the pricing kernel alone, without the cross-rank exchange
    // (relabel, pass and reduction are one kernel); loop 1 is the step as the solver launches it.
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < reps; ++i) {
            if (flushL2) { k_flush_l2<<<148 * 8, 256, 0, d->stream>>>(d->flushBuf, d->flushLen); ++d->launches; }
            if (pass == 1) CK(cudaEventRecord(d->benchEv[2 * i], d->stream));
            relabel();
            const bool hostBlock = d->nranks == 1;
            rc = launch_lv_pass(d, hostBlock ? d->lvRes : d->lvResDev, pass == 0 ? d->benchEv[2 * i] : nullptr,
                                pass == 0 ? d->benchEv[2 * i + 1] : nullptr, pass == 1, replay ? &d->lastSmall : nullptr);
            if (rc) return rc;
            if (d->nranks > 1 && pass == 1) { rc = launch_shard_exchange(d); if (rc) return rc; }
            if (pass == 1) CK(cudaEventRecord(d->benchEv[2 * i + 1], d->stream));
        }
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
        for (int i = 0; i < reps; ++i) CK(cudaEventElapsedTime(pass == 0 ? &msTiles[i] : &msStep[i], d->benchEv[2 * i], d->benchEv[2 * i + 1]));
    }
    return GNE_OK;
}

// The same pass `reps` times back to back between ONE pair of events (no event, flush or host work between the launches):
// the average duration of a launch in a burst - the front end prepares launch k+1 while launch k runs, so this is the
// kernel without the per-launch event bracket of gne_bench_price_lv.  Single GPU only; L2 is not flushed (configs[2] and
// larger stream more than L2 holds).
extern "C" int gne_bench_price_lv_burst(gne_dev *d, int reps, int withFinalize, float *msTotal)
{
    if (!d || reps <= 0 || !msTotal) { gne_set_error("gne_bench_price_lv_burst: bad argument"); return GNE_ERR_ARG; }
    if (d->nranks > 1) { gne_set_error("gne_bench_price_lv_burst: single GPU only"); return GNE_ERR_ARG; }
    CK(cudaSetDevice(d->device));
    int rc = gne_dev_flush(d); if (rc) return rc;
    if (!d->ev2) { CK(cudaEventCreate(&d->ev2)); CK(cudaEventCreate(&d->ev3)); }
    const bool replay = withFinalize == 2 && d->haveLastSmall && use_stream_kernel(d);
    for (int i = -3; i < reps; ++i) {                       // three untimed launches first
        if (i == 0) CK(cudaEventRecord(d->ev2, d->stream));
        rc = launch_lv_pass(d, d->lvRes, nullptr, nullptr, true, replay ? &d->lastSmall : nullptr);
        if (rc) return rc;
    }
    CK(cudaEventRecord(d->ev3, d->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    CK(cudaEventElapsedTime(msTotal, d->ev2, d->ev3));
    return GNE_OK;
}

// The block rule's device leg: `reps` quickSelect passes as the solver issues them (window kernel + finish kernel per
// window, the cursor advancing by what each pass scanned), each pass timed with CUDA events around its kernels.
// arcsLaunched[i] = arcs the kernels of pass i were launched over (whole windows), arcsScanned[i] = what the rule consumed.
// Device-resident leg of a batch (configs[3]): `reps` rounds of LV passes over the B handles (each replaying its last
// argument-only relabel, idempotent), `perLaunch` instances per launch of the batched kernel, the launches spread
// round-robin over `nStreams` launch groups so that one launch's tail overlaps the next one's start.  *msTotal: device
// time of all rounds (events on the first stream, which waits for the others).
extern "C" int gne_bench_lv_groups(gne_dev **devs, int B, int reps, int perLaunch, int nStreams, float *msTotal, int64_t *launches)
{
    if (!devs || B <= 0 || reps <= 0 || !msTotal) { gne_set_error("gne_bench_lv_groups: bad argument"); return GNE_ERR_ARG; }
    nStreams = std::max(1, std::min(nStreams, 8));
    std::vector<gne_lv_group *> gr((size_t) nStreams, nullptr);
    int rc = GNE_OK;
    for (int i = 0; i < B && rc == GNE_OK; ++i) rc = gne_dev_flush(devs[i]);
    for (int k = 0; k < nStreams && rc == GNE_OK; ++k) rc = gne_lv_group_create(&gr[(size_t) k], devs[0]->device, perLaunch);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<cudaEvent_t> done((size_t) nStreams, nullptr);
    auto body = [&]() -> int {
        CK(cudaSetDevice(devs[0]->device));
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int k = 0; k < nStreams; ++k) CK(cudaEventCreateWithFlags(&done[(size_t) k], cudaEventDisableTiming));
        for (int i = 0; i < B; ++i) { CK(cudaStreamSynchronize(devs[i]->stream)); devs[i]->lvLaunched = false; }
        const int per = gr[0]->maxPending;
        for (int round = -1; round < reps; ++round) {              // round -1: warm-up, untimed
            if (round == 0) {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0, gr[0]->stream));
                for (int k = 1; k < nStreams; ++k) CK(cudaStreamWaitEvent(gr[(size_t) k]->stream, e0, 0));
            }
            int launchNo = 0;
            for (int i0 = 0; i0 < B; i0 += per, ++launchNo) {
                gne_lv_group *g = gr[(size_t) (launchNo % nStreams)];
                for (int i = i0; i < std::min(B, i0 + per); ++i) {
                    gne_dev *d = devs[i];
                    d->lvLaunched = false;
                    int r = lv_group_add(g, d, d->haveLastSmall ? &d->lastSmall : nullptr); if (r) return r;
                }
                int r = gne_lv_group_launch(g); if (r) return r;
            }
        }
        for (int k = 1; k < nStreams; ++k) { CK(cudaEventRecord(done[(size_t) k], gr[(size_t) k]->stream)); CK(cudaStreamWaitEvent(gr[0]->stream, done[(size_t) k], 0)); }
        CK(cudaEventRecord(e1, gr[0]->stream));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(msTotal, e0, e1));
        for (int i = 0; i < B; ++i) devs[i]->lvLaunched = false;
        if (launches) { *launches = 0; for (int k = 0; k < nStreams; ++k) *launches += gr[(size_t) k]->launches; }
        return GNE_OK;
    };
    if (rc == GNE_OK) rc = body();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    for (int k = 0; k < nStreams; ++k) { if (done[(size_t) k]) cudaEventDestroy(done[(size_t) k]); gne_lv_group_destroy(gr[(size_t) k]); }
    return rc;
}

extern "C" int gne_bench_price_qs(gne_dev *d, int reps, int flushL2, int64_t want, float *msStep, int64_t *arcsLaunched, int64_t *arcsScanned)
{
    CK(cudaSetDevice(d->device));
    if (d->nranks != 1) { gne_set_error("gne_bench_price_qs: single GPU only"); return GNE_ERR_UNSUPPORTED; }
    int rc = gne_dev_flush(d); if (rc) return rc;
    if (flushL2 && !d->flushBuf) {
        d->flushLen = (int64_t) 40 * 1024 * 1024;
        CK(cudaMalloc(&d->flushBuf, 8 * (size_t) d->flushLen));
    }
    const bool keep = d->timing;
    d->timing = true;
    const int64_t N = d->N;
    int64_t cursor = 0;
    for (int i = 0; i < reps; ++i) {
        if (flushL2) { k_flush_l2<<<148 * 8, 256, 0, d->stream>>>(d->flushBuf, d->flushLen); ++d->launches; }
        const int64_t before = d->qsArcsLaunched;
        float ms = 0.f;
        rc = qs_scan(d, cursor, N, N, want, &ms); if (rc) { d->timing = keep; return rc; }
        msStep[i] = ms;
        arcsLaunched[i] = d->qsArcsLaunched - before;
        const int64_t sc = (d->qsState->stopJ >= 0) ? d->qsState->stopJ + 1 : N;
        arcsScanned[i] = sc;
        cursor += sc; if (cursor >= N) cursor -= N;
    }
    d->timing = keep;
    return GNE_OK;
}

// ------------------------------------------------------------------------------------------
// multi-GPU: every rank prices its shard, the fixed-size LV records are all-gathered over
// NCCL (NVLink), and every rank runs the same merge => identical entering arc everywhere.
// ------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ void pack_record(const LvResult *res, ShardRecord *rec, int t)
{
    if (t == 0) {
        rec->maxViol = res->maxViol; rec->any = res->any; rec->count = res->count;
        rec->overflow = (res->overflow || res->count > SHARD_LIST) ? 1 : 0;
    }
    if (t < SHARD_LIST && t < res->count) { rec->pos[t] = res->pos[t]; rec->viol[t] = res->viol[t]; }
}
__global__ void k_pack_record(const LvResult *res, ShardRecord *rec) { pack_record(res, rec, threadIdx.x); }

// Stand-alone exchange step (after the per-tile reduction kernel; the TMA-streamed pass runs the same
// exchange from its own last CTA): pack the shard's record, then exchange_record (gne_exchange.cuh).
__global__ void __launch_bounds__(256)
k_lv_exchange(const LvResult *res, Mailbox *const *peers, int rank, int nranks, unsigned long long xseq, LvResult *hostOut,
              unsigned long long outSeq, long long timeoutCycles)
{
    __shared__ ShardRecord rec;
    __shared__ int sTimeout;
    __shared__ int64_t mPos[MAX_RANKS * SHARD_LIST];
    __shared__ double mViol[MAX_RANKS * SHARD_LIST];
    pdl_wait();                                         // the shard's reduction has published its result block
    pack_record(res, &rec, threadIdx.x);
    xchg_sync();
    exchange_merge(&rec, peers, rank, nranks, xseq, hostOut, outSeq, &sTimeout, timeoutCycles, mPos, mViol);
}
// all-gather of one 128-byte blob per rank through the peer mailboxes (same protocol as k_lv_exchange)
__global__ void __launch_bounds__(64)
k_blob_exchange(const unsigned long long *src, Mailbox *const *peers, int rank, int nranks, unsigned long long seq, unsigned long long *hostOut,
                unsigned long long *hostSeq, long long timeoutCycles)
{
    const int t = threadIdx.x, b = (int) (seq & 1ull);
    __shared__ unsigned long long mine[BLOB_WORDS];
    __shared__ int sTimeout;
    if (t < BLOB_WORDS) mine[t] = src[t];
    if (t == 0) sTimeout = 0;
    __syncthreads();
    for (int p = 0; p < nranks; ++p) if (t < BLOB_WORDS) peers[p]->blob[b][rank][t] = mine[t];
    __threadfence_system();
    __syncthreads();
    if (t < nranks) *reinterpret_cast<volatile unsigned long long *>(&peers[t]->bflag[b][rank]) = seq;
    Mailbox *self = peers[rank];
    if (t < nranks) {
        const long long t0 = clock64();
        while (true) {
            const unsigned long long v = *reinterpret_cast<volatile unsigned long long *>(&self->bflag[b][t]);
            if (v == seq) break;
            if (v == XCHG_POISON || clock64() - t0 > timeoutCycles) { sTimeout = 1; break; }
        }
    }
    __syncthreads();
    if (sTimeout && t < nranks) {
        *reinterpret_cast<volatile unsigned long long *>(&peers[t]->bflag[0][rank]) = XCHG_POISON;
        *reinterpret_cast<volatile unsigned long long *>(&peers[t]->bflag[1][rank]) = XCHG_POISON;
    }
    __threadfence_system();
    for (int r = 0; r < nranks; ++r)
        if (t < BLOB_WORDS) hostOut[r * BLOB_WORDS + t] = *reinterpret_cast<const volatile unsigned long long *>(&self->blob[b][r][t]);
    __syncthreads();
    if (t == 0) hostSeq[1] = sTimeout ? 1ull : 0ull;
    __threadfence_system();                             // every writer fences its own stores, then the word the host polls
    __syncthreads();
    if (t == 0) *reinterpret_cast<volatile unsigned long long *>(hostSeq) = seq;
}
} // namespace

// one 128-byte blob per rank -> all ranks (host buffers); mailbox path or NCCL
static int exchange_blobs(gne_dev *d, const unsigned long long *mine, unsigned long long *all)
{
    const int R = d->nranks;
    if (!d->blobHost) {
        CK(cudaHostAlloc(&d->blobHost, 8 * BLOB_WORDS * (size_t) (R + 2), cudaHostAllocMapped));      // mine, R gathered, the sequence word
        memset(d->blobHost, 0, 8 * BLOB_WORDS * (size_t) (R + 2));
        CK(cudaMalloc(&d->blobDev, 8 * BLOB_WORDS * (size_t) (R + 1)));
    }
    memcpy(d->blobHost, mine, 8 * BLOB_WORDS);
    if (d->p2p) {
        unsigned long long *hostSeq = d->blobHost + BLOB_WORDS * (size_t) (R + 1);
        k_blob_exchange<<<1, 64, 0, d->stream>>>(d->blobHost, d->peersDev, d->rank, R, ++d->bseq, d->blobHost + BLOB_WORDS, hostSeq, d->xTimeoutCycles);
        CK(cudaGetLastError());
        ++d->launches;
        { int rcw = wait_host_seq(d, hostSeq, d->bseq); if (rcw) return rcw; }
        if (hostSeq[1] != 0ull) { gne_set_error("sharded exchange timed out waiting for a peer rank"); return GNE_ERR_COMM; }
    } else {
        int rcg = small_allgather(d, d->blobHost, d->blobHost + BLOB_WORDS, 8 * BLOB_WORDS); if (rcg) return rcg;
    }
    memcpy(all, d->blobHost + BLOB_WORDS, 8 * BLOB_WORDS * (size_t) R);
    d->d2hBytes += 8 * BLOB_WORDS * R;
    return GNE_OK;
}

// quickSelect over a sharded list: every rank scans its part after the cursor
// (earlier in scan order) and its part before it, the (count, first-max) pairs are exchanged, every
// rank finds the segment in which the want-th violator falls, the owner rescans that segment with
// the remaining budget, and a second exchange publishes the stop position and the partial maximum.
static int price_qs_sharded(gne_dev *d, int64_t cursor, int64_t want, int64_t *bestPos, double *bestViol,
                            int64_t *newCursor, int64_t *scanned, int64_t *violations)
{
    const int R = d->nranks;
    const int64_t N = d->N;
    const int64_t lo = d->lo, end = std::max(lo, std::min(N, d->hi));
    struct Seg { int64_t a, b, cnt, pos; double v; };
    union Blob { unsigned long long w[BLOB_WORDS]; struct { int64_t a1, b1, c1, p1, a2, b2, c2, p2; double v1, v2; int64_t stop, pp; double pv; } f; };
    Blob mine; memset(&mine, 0, sizeof(mine));
    Seg local[2] = { { std::max(lo, cursor), end, 0, -1, 0.0 }, { lo, std::min(end, cursor), 0, -1, 0.0 } };
    float msTot = 0.f;
    for (int k = 0; k < 2; ++k) {
        Seg &sg = local[k];
        if (sg.b <= sg.a) { sg.a = sg.b = 0; continue; }
        float ms = 0.f;
        int rc = qs_scan(d, sg.a, N, sg.b - sg.a, INT64_MAX / 4, &ms); if (rc) return rc;
        msTot += ms;
        sg.cnt = d->qsState->violations; sg.v = d->qsState->bestV;
        sg.pos = d->qsState->bestJ >= 0 ? sg.a + d->qsState->bestJ : -1;
    }
    mine.f.a1 = local[0].a; mine.f.b1 = local[0].b; mine.f.c1 = local[0].cnt; mine.f.p1 = local[0].pos; mine.f.v1 = local[0].v;
    mine.f.a2 = local[1].a; mine.f.b2 = local[1].b; mine.f.c2 = local[1].cnt; mine.f.p2 = local[1].pos; mine.f.v2 = local[1].v;
    std::vector<Blob> all((size_t) R);
    int rc = exchange_blobs(d, mine.w, all[0].w); if (rc) return rc;
    // scan order: segments after the cursor by rank, then segments before it by rank
    std::vector<std::pair<int, int> > order;            // (rank, which)
    for (int r = 0; r < R; ++r) if (all[r].f.b1 > all[r].f.a1) order.push_back(std::make_pair(r, 0));
    for (int r = 0; r < R; ++r) if (all[r].f.b2 > all[r].f.a2) order.push_back(std::make_pair(r, 1));
    int64_t cum = 0;
    int cross = -1;
    for (size_t k = 0; k < order.size(); ++k) {
        const Blob &bl = all[order[k].first];
        const int64_t c = order[k].second ? bl.f.c2 : bl.f.c1;
        if (cum < want && cum + c >= want) { cross = (int) k; break; }
        cum += c;
    }
    double bv = 0.0;                                    // largestViol starts at 0.0, strictly-greater replaces
    int64_t bp = -1;
    const size_t upto = cross < 0 ? order.size() : (size_t) cross;
    for (size_t k = 0; k < upto; ++k) {
        const Blob &bl = all[order[k].first];
        const double v = order[k].second ? bl.f.v2 : bl.f.v1;
        const int64_t p = order[k].second ? bl.f.p2 : bl.f.p1;
        if (p >= 0 && v > bv) { bv = v; bp = p; }
    }
    if (cross < 0) {
        *violations = cum; *scanned = N; *newCursor = cursor;
        *bestViol = bv; *bestPos = bp;
        d->lastMs = msTot;
        return GNE_OK;
    }
    // the owner of the crossing segment rescans it with the remaining budget
    Blob second; memset(&second, 0, sizeof(second));
    second.f.stop = -1; second.f.pp = -1;
    if (order[cross].first == d->rank) {
        const Seg &sg = local[order[cross].second];
        float ms = 0.f;
        rc = qs_scan(d, sg.a, N, sg.b - sg.a, want - cum, &ms); if (rc) return rc;
        msTot += ms;
        second.f.stop = sg.a + d->qsState->stopJ;
        second.f.pp = d->qsState->bestJ >= 0 ? sg.a + d->qsState->bestJ : -1;
        second.f.pv = d->qsState->bestV;
    }
    rc = exchange_blobs(d, second.w, all[0].w); if (rc) return rc;
    const Blob &ow = all[order[cross].first];
    if (ow.f.pp >= 0 && ow.f.pv > bv) { bv = ow.f.pv; bp = ow.f.pp; }
    const int64_t stop = ow.f.stop;
    *violations = want;
    *scanned = (stop >= cursor ? stop - cursor : stop + N - cursor) + 1;
    *newCursor = (stop + 1 >= N) ? stop + 1 - N : stop + 1;
    *bestViol = bv; *bestPos = bp;
    d->lastMs = msTot;
    return GNE_OK;
}

extern "C" int gne_comm_unique_id(uint8_t id[128])
{
    if (!nccl_load()) return GNE_ERR_COMM;
    int r = g_nccl.GetUniqueId(id);
    if (r != 0) { gne_set_error("ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return GNE_ERR_COMM; }
    return GNE_OK;
}

// All-gather of a few host bytes per rank over whichever channel the handle was given: the host program's callback,
// or NCCL staged through a small device buffer.  Bootstrap and fallback paths only - never the per-pivot mailbox path.
static int small_allgather(gne_dev *d, const void *sendHost, void *recvHost, size_t bytes)
{
    if (d->hostGather) {
        if (d->hostGather(d->hostGatherCtx, sendHost, recvHost, (int64_t) bytes) != 0) { gne_set_error("the host all-gather callback failed"); return GNE_ERR_COMM; }
        return GNE_OK;
    }
    const size_t need = bytes * (size_t) (d->nranks + 1);
    if (need > d->agBytes) {
        if (d->agDev) cudaFree(d->agDev);
        d->agDev = nullptr; d->agBytes = 0;
        CK(cudaMalloc(&d->agDev, need));
        d->agBytes = need;
    }
    CK(cudaMemcpyAsync(d->agDev, sendHost, bytes, cudaMemcpyHostToDevice, d->stream));
    const int r = g_nccl.AllGather(d->agDev, d->agDev + bytes, bytes, /*ncclChar*/ 0, d->comm, d->stream);
    if (r != 0) { gne_set_error("ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return GNE_ERR_COMM; }
    CK(cudaMemcpyAsync(recvHost, d->agDev + bytes, bytes * (size_t) d->nranks, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    return GNE_OK;
}

// After d->comm or d->hostGather is in place: buffers of the exchange, the peer mailboxes (CUDA IPC handles
// all-gathered and mapped), and the verdict every rank must share - mailboxes only if EVERY rank mapped them all.
static int comm_setup(gne_dev *d, int rank, int nranks)
{
    d->rank = rank; d->nranks = nranks;
    { const char *ts = getenv("GNE_EXCHANGE_TIMEOUT_S"); if (ts && atof(ts) > 0.0) d->xTimeoutCycles = (long long) (atof(ts) * 2.0e9); }
    CK(cudaMalloc(&d->gatherSend, sizeof(ShardRecord)));
    CK(cudaMalloc(&d->gatherRecv, sizeof(ShardRecord) * (size_t) nranks));
    CK(cudaHostAlloc(&d->gatherHost, sizeof(ShardRecord) * (size_t) (nranks + 1) + 64, cudaHostAllocMapped));   // [nranks records][seq word ...][own record (host path)]
    memset(d->gatherHost, 0, sizeof(ShardRecord) * (size_t) (nranks + 1) + 64);
    CK(cudaMalloc(&d->lvResDev, sizeof(LvResult)));
    CK(cudaMemset(d->lvResDev, 0, sizeof(LvResult)));
    CK(cudaMalloc(&d->mailbox, sizeof(Mailbox)));
    CK(cudaMemset(d->mailbox, 0, sizeof(Mailbox)));
    const char *force = getenv("GNE_EXCHANGE");
    int okLocal = (force && (!strcmp(force, "nccl") || !strcmp(force, "host"))) ? 0 : 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (okLocal && cudaIpcGetMemHandle(&mine, d->mailbox) != cudaSuccess) { cudaGetLastError(); okLocal = 0; memset(&mine, 0, sizeof(mine)); }
    std::vector<cudaIpcMemHandle_t> handles((size_t) nranks);
    int rc = small_allgather(d, &mine, handles.data(), sizeof(cudaIpcMemHandle_t)); if (rc) return rc;
    for (int p = 0; p < nranks && okLocal; ++p) {
        if (p == rank) { d->peerPtr[p] = d->mailbox; continue; }
        if (cudaIpcOpenMemHandle(&d->peerPtr[p], handles[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); d->peerPtr[p] = nullptr; okLocal = 0; }
    }
    // every rank must take the same path: all-gather the per-rank verdicts
    unsigned char flagByte[8] = { (unsigned char) okLocal }, verdicts[8 * MAX_RANKS] = { 0 };
    rc = small_allgather(d, flagByte, verdicts, 8); if (rc) return rc;
    d->p2p = true;
    for (int p = 0; p < nranks; ++p) if (!verdicts[8 * p]) d->p2p = false;
    if (d->p2p) {
        CK(cudaMalloc(&d->peersDev, sizeof(Mailbox *) * MAX_RANKS));
        CK(cudaMemcpy(d->peersDev, d->peerPtr, sizeof(void *) * (size_t) nranks, cudaMemcpyHostToDevice));
    }
    return GNE_OK;
}

extern "C" int gne_comm_init(gne_dev *d, const uint8_t id[128], int rank, int nranks)
{
    if (nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) { gne_set_error("gne_comm_init: rank %d of %d (at most %d ranks)", rank, nranks, MAX_RANKS); return GNE_ERR_ARG; }
    if (!nccl_load()) return GNE_ERR_COMM;
    CK(cudaSetDevice(d->device));
    NcclId uid;
    memcpy(uid.b, id, 128);
    int r = g_nccl.CommInitRank(&d->comm, nranks, uid, rank);
    if (r != 0) { gne_set_error("ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return GNE_ERR_COMM; }
    return comm_setup(d, rank, nranks);
}

extern "C" int gne_comm_init_host(gne_dev *d, int rank, int nranks, gne_allgather_fn allgather, void *ctx)
{
    if (nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks || !allgather) { gne_set_error("gne_comm_init_host: rank %d of %d (at most %d ranks), callback %p", rank, nranks, MAX_RANKS, (void *) allgather); return GNE_ERR_ARG; }
    CK(cudaSetDevice(d->device));
    d->hostGather = allgather; d->hostGatherCtx = ctx;
    return comm_setup(d, rank, nranks);
}

extern "C" int gne_comm_mode(gne_dev *d) { return d->nranks <= 1 ? 0 : (d->p2p ? 2 : (d->hostGather ? 3 : 1)); }

extern "C" int gne_comm_destroy(gne_dev *d)
{
    if (d->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(d->comm); d->comm = nullptr; }
    d->hostGather = nullptr; d->hostGatherCtx = nullptr;
    d->nranks = 1; d->rank = 0;
    return GNE_OK;
}

static int launch_shard_exchange(gne_dev *d)
{
    if (d->exchangeFused) return GNE_OK;                // the pricing kernel's last CTA has already done it
    if (d->p2p) {
        // pack + NVLink peer stores + flag wait + merge + merged block to mapped pinned host: one kernel
        CK(launch_k(d, true, k_lv_exchange, dim3(1), dim3(256), 0, (const LvResult *) d->lvResDev, (Mailbox *const *) d->peersDev, d->rank, d->nranks,
                    (unsigned long long) ++d->xseq, d->lvRes, d->lvSeq, d->xTimeoutCycles));
        ++d->launches;
        return GNE_OK;
    }
    // fallback: pack this shard's record, all-gather over NCCL (or through the host program), records to pinned host
    k_pack_record<<<1, 64, 0, d->stream>>>(d->lvResDev, (ShardRecord *) d->gatherSend);
    ++d->launches;
    if (d->hostGather) {
        ShardRecord *own = (ShardRecord *) ((char *) d->gatherHost + sizeof(ShardRecord) * (size_t) d->nranks + 64);
        CK(cudaMemcpyAsync(own, d->gatherSend, sizeof(ShardRecord), cudaMemcpyDeviceToHost, d->stream));
        CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
        return small_allgather(d, own, d->gatherHost, sizeof(ShardRecord));
    }
    int r = g_nccl.AllGather(d->gatherSend, d->gatherRecv, sizeof(ShardRecord), /*ncclChar*/ 0, d->comm, d->stream);
    if (r != 0) { gne_set_error("ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return GNE_ERR_COMM; }
    CK(cudaMemcpyAsync(d->gatherHost, d->gatherRecv, sizeof(ShardRecord) * (size_t) d->nranks, cudaMemcpyDeviceToHost, d->stream));
    return GNE_OK;
}

// All-gather of the ranks' variable-length (pos, viol) lists that viol_compact_device left in
// d->outPos / d->outViol; result concatenated in rank order == global position order.  Every rank
// runs the same sequence of collectives, so a fallback can never leave a peer waiting.
static int sharded_gather_lists(gne_dev *d, int64_t localCount, std::vector<int64_t> &pos, std::vector<double> &viol)
{
    const int R = d->nranks;
    std::vector<int64_t> counts((size_t) R);
    int rc = small_allgather(d, &localCount, counts.data(), 8); if (rc) return rc;
    int64_t maxC = 0, total = 0;
    for (int k = 0; k < R; ++k) { maxC = std::max(maxC, counts[k]); total += counts[k]; }
    pos.clear(); viol.clear();
    if (maxC == 0) return GNE_OK;
    const size_t per = (size_t) maxC * 16;
    std::vector<char> host(per * R);
    if (d->hostGather) {
        // through the host program: own list to the host, all-gather there
        std::vector<char> mine(per, 0);
        if (localCount > 0) {
            CK(cudaMemcpyAsync(mine.data(), d->outPos, 8 * (size_t) localCount, cudaMemcpyDeviceToHost, d->stream));
            CK(cudaMemcpyAsync(mine.data() + 8 * (size_t) maxC, d->outViol, 8 * (size_t) localCount, cudaMemcpyDeviceToHost, d->stream));
            CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
        }
        rc = small_allgather(d, mine.data(), host.data(), per); if (rc) return rc;
    } else {
        if (per > d->xSendBytes) { if (d->xSend) cudaFree(d->xSend); d->xSend = nullptr; d->xSendBytes = 0; CK(cudaMalloc(&d->xSend, per)); d->xSendBytes = per; }
        if (per * R > d->xRecvBytes) { if (d->xRecv) cudaFree(d->xRecv); d->xRecv = nullptr; d->xRecvBytes = 0; CK(cudaMalloc(&d->xRecv, per * R)); d->xRecvBytes = per * R; }
        if (localCount > 0) {
            CK(cudaMemcpyAsync(d->xSend, d->outPos, 8 * (size_t) localCount, cudaMemcpyDeviceToDevice, d->stream));
            CK(cudaMemcpyAsync(d->xSend + 8 * (size_t) maxC, d->outViol, 8 * (size_t) localCount, cudaMemcpyDeviceToDevice, d->stream));
        }
        const int r = g_nccl.AllGather(d->xSend, d->xRecv, per, /*ncclChar*/ 0, d->comm, d->stream);
        if (r != 0) { gne_set_error("ncclAllGather(lists): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return GNE_ERR_COMM; }
        CK(cudaMemcpyAsync(host.data(), d->xRecv, per * R, cudaMemcpyDeviceToHost, d->stream));
        CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    }
    d->d2hBytes += (int64_t) (per * R);
    pos.reserve((size_t) total); viol.reserve((size_t) total);
    for (int k = 0; k < R; ++k) {
        const int64_t *p = (const int64_t *) (host.data() + per * k);
        const double *v = (const double *) (host.data() + per * k + 8 * (size_t) maxC);
        for (int64_t i = 0; i < counts[k]; ++i) { pos.push_back(p[i]); viol.push_back(v[i]); }
    }
    return GNE_OK;
}

// Fallback of a sharded pass (long tie lists, overflowed records): every rank compacts its shard down to a widening band
// below the global maximum M, the lists are all-gathered, and the sequential rule is replayed on the union.  Every rank
// takes the same branch (same merged block), so no rank can be left waiting.
static int sharded_band_fallback(gne_dev *d, double M, std::vector<int64_t> &list, double *largest)
{
    double width = 16.0 * GNE_ROUNDOFF_TOLERANCE;
    std::vector<int64_t> gp;
    std::vector<double> gv;
    while (true) {
        double floorV = M - width;
        bool zero = false;
        if (!(floorV > 0.0) || width > 1.0e-6) { floorV = 0.0; zero = true; }
        int64_t localCount = 0;
        int rc = viol_compact_device(d, floorV, &localCount); if (rc) return rc;
        rc = sharded_gather_lists(d, localCount, gp, gv); if (rc) return rc;
        if (replay_band_rule(gp.data(), gv.data(), (int64_t) gp.size(), floorV, zero, list, largest)) break;
        width *= 64.0;
    }
    return GNE_OK;
}

// Sharded pass WITHOUT peer mailboxes (GNE_EXCHANGE=nccl / host, or IPC mapping refused): the per-shard records are
// all-gathered (ncclAllGather or the host program's callback) and merged on the host by every rank.
int gne_dev_price_lv_sharded(gne_dev *d, double *largest, std::vector<int64_t> &list, int *any)
{
    // price_lv_local has been launched by the caller
    int rcx = launch_shard_exchange(d); if (rcx) return rcx;
    CK(cudaStreamSynchronize(d->stream)); d->stageInFlight = false;
    d->d2hBytes += sizeof(ShardRecord) * (int64_t) d->nranks;
    const ShardRecord *rec = (const ShardRecord *) d->gatherHost;
    // scratch kept across pivots (no allocation on the per-pivot path)
    static thread_local std::vector<double> maxima, viol;
    static thread_local std::vector<int> anys;
    static thread_local std::vector<int64_t> counts, pos, outPos;
    maxima.resize(d->nranks); anys.resize(d->nranks); counts.resize(d->nranks);
    pos.resize((size_t) d->nranks * SHARD_LIST); viol.resize((size_t) d->nranks * SHARD_LIST); outPos.resize((size_t) d->nranks * SHARD_LIST);
    for (int k = 0; k < d->nranks; ++k) {
        maxima[k] = rec[k].maxViol; anys[k] = rec[k].any;
        counts[k] = rec[k].overflow ? -1 : rec[k].count;
        const int64_t c = std::min<int64_t>(std::max<int64_t>(rec[k].count, 0), SHARD_LIST);
        for (int64_t i = 0; i < c; ++i) { pos[(size_t) k * SHARD_LIST + i] = rec[k].pos[i]; viol[(size_t) k * SHARD_LIST + i] = rec[k].viol[i]; }
    }
    int64_t cnt = 0;
    int need = 0;
    int rc = gne_merge_lv_shards(d->nranks, maxima.data(), anys.data(), counts.data(), pos.data(), viol.data(), SHARD_LIST,
                                 largest, &cnt, outPos.data(), (int64_t) outPos.size(), any, &need);
    if (rc) return rc;
    if (!need) { list.assign(outPos.begin(), outPos.begin() + cnt); return GNE_OK; }
    double M = -1.0;
    for (int k = 0; k < d->nranks; ++k) if (anys[k] && maxima[k] > M) M = maxima[k];
    return sharded_band_fallback(d, M, list, largest);
}

// Host-only: merge per-shard records.  counts[r] < 0 marks an overflowed shard.
extern "C" int gne_merge_lv_shards(int nranks, const double *maxima, const int *any, const int64_t *counts,
                                   const int64_t *listPos, const double *listViol, int64_t stride,
                                   double *largestViolation, int64_t *count, int64_t *outPos, int64_t cap, int *outAny, int *needFallback)
{
    *needFallback = 0; *count = 0; *outAny = 0; *largestViolation = -1.0;
    double M = -1.0;
    for (int r = 0; r < nranks; ++r) if (any[r]) { *outAny = 1; if (maxima[r] > M) M = maxima[r]; }
    if (!*outAny) return GNE_OK;
    const double floorV = M - LV_BAND;
    std::vector<int64_t> p;
    std::vector<double> v;
    int64_t lastPos = -1;
    for (int r = 0; r < nranks; ++r) {                  // shards are contiguous position ranges in rank order
        if (!any[r] || maxima[r] < floorV) continue;
        if (counts[r] < 0) { *needFallback = 1; return GNE_OK; }
        for (int64_t i = 0; i < counts[r]; ++i) {
            const double x = listViol[(size_t) r * stride + i];
            const int64_t q = listPos[(size_t) r * stride + i];
            if (q <= lastPos) { gne_set_error("gne_merge_lv_shards: positions not increasing (rank %d, entry %lld): overlapping shards?", r, (long long) i); return GNE_ERR_ARG; }
            lastPos = q;
            if (x >= floorV) { p.push_back(q); v.push_back(x); }
        }
    }
    std::vector<int64_t> list;
    double L;
    if (!replay_band_rule(p.data(), v.data(), (int64_t) p.size(), floorV, false, list, &L)) { *needFallback = 1; return GNE_OK; }
    if ((int64_t) list.size() > cap) { gne_set_error("gne_merge_lv_shards: cap too small"); return GNE_ERR_CAPACITY; }
    *largestViolation = L;
    *count = (int64_t) list.size();
    for (size_t i = 0; i < list.size(); ++i) outPos[i] = list[i];
    return GNE_OK;
}
