// genneteq_b200 — the LV (Dantzig) pricing pass for large shards: persistent CTAs, the five arc
// columns streamed into shared memory by the TMA engine (cp.async.bulk + mbarrier, 3 stages), and
// only the pi gathers on the LSU / L1TEX path.
//
// Why: the per-tile kernel (k_price_lv_tiles) is bound by L1TEX wavefronts with a dependent
// chain index load -> gather.  Staging the columns through the async proxy takes the streaming
// loads off that chain: with 3 stages in flight a CTA always finds its indices in shared memory
// and issues its gathers at once.  Measured on B200 (profiles/r1_variants_c3.txt): 108.3 us per
// 20M-arc pass with 16 warps/SM against 113.5 us for the per-tile kernel with 64 warps/SM;
// 2 stages or >150 KB of staging (which starves the L1 the gathers live in) are 50 % slower.
//
// Reduction: each thread keeps its largest and second-largest violation in registers; once per CTA
// the threads within the band of the CTA maximum are recorded (CtaRec + candidate arrays).  The last
// CTA to finish (ticket counter) reduces the per-CTA records to the exact global maximum and the
// in-band list in position order, publishes the result block, and - in a sharded run - exchanges the
// shard's record with the peer GPUs from the same kernel (gne_exchange.cuh): one launch per pass.
// A violator can never be skipped wrongly; a list that overflows is reported and the caller falls
// back to the ordered compaction.
#pragma once
#include "gne_kernels.cuh"
#include "gne_exchange.cuh"

namespace gne {

constexpr int ST_THREADS = 256;
constexpr int ST_ITEMS = 4;
constexpr int ST_STAGES = 3;
constexpr int ST_TILE = ST_THREADS * ST_ITEMS;          // 1024 arcs per stage
constexpr int ST_STAGE_BYTES = 25 * ST_TILE;            // 25 600 B
constexpr int ST_CTAS_PER_SM = 2;
constexpr int ST_CAND = 96;                             // in-band candidates kept per CTA
constexpr int ST_SMEM = 128 + ST_STAGES * ST_STAGE_BYTES;

__device__ __forceinline__ uint32_t st_smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(st_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void st_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(st_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(st_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(st_smem_u32(dst)), "l"(src), "r"(bytes), "r"(st_smem_u32(bar)) : "memory");
}

struct CtaRec {            // 64-byte record of one CTA; a lone in-band candidate travels in the record itself
    double max;            // -1 when the CTA saw no violator
    int32_t cnt;           // threads whose largest violation lies within the band of `max` (may exceed ST_CAND: overflow)
    int32_t pad;
    int64_t pos0;          // cnt == 1: that thread's position, largest and second-largest violation
    double viol0;
    double second0;
    double pad2[3];
};
struct LvCtas {
    CtaRec *rec;           // [G]
    int64_t *candPos;      // [G * ST_CAND] global positions, unordered           (read only when cnt > 1)
    double *candViol;      // [G * ST_CAND] the thread's largest violation
    double *candSecond;    // [G * ST_CAND] the same thread's second largest (in band => the thread may hide more)
};
// Where the last CTA of the pass leaves the result, and (sharded runs) the exchange that follows it.
struct LvFused {
    unsigned int *ticket;          // device counter, zero between passes
    LvResult *res;                 // result block: mapped pinned host memory (single GPU) or device memory (sharded)
    unsigned long long seq;        // published in res->seq last
    int resOnHost;                 // system-scope fences only when the host polls res
    int nranks, rank;              // nranks > 1 and peers != nullptr: exchange_record() after the reduction
    Mailbox *const *peers;
    ShardRecord *hostOut;
    unsigned long long *hostSeq;
    unsigned long long xseq;
    long long xTimeout;            // cycles a rank waits for its peers before failing the pass
    long long *stamps;             // debug (GNE_STAMPS builds): clock64 stamps of the last CTA
};

// Reduction over the per-CTA records by ONE CTA of ST_THREADS threads (the last one of the pass): exact
// global maximum, then every recorded candidate within the band of it, sorted by position (the CTAs
// interleave tiles, so their lists are merged by rank), published to f.res; in a sharded run the same
// CTA then exchanges the shard's record with the peers.  `scratch` is the (drained) stage memory.
// Records are read once, two per thread, straight from L2: the common case (one candidate) costs one
// dependent load level before the result is written.
__device__ __noinline__ void lv_reduce_records(const LvCtas &in, int G, const LvFused &f, unsigned char *scratch)
{
    constexpr int NT = ST_THREADS, NW = ST_THREADS / 32, LIST = 1024, PER = 2;
    double *lViol = reinterpret_cast<double *>(scratch);
    int64_t *lPos = reinterpret_cast<int64_t *>(scratch + 8 * LIST);
    int *lRank = reinterpret_cast<int *>(scratch + 16 * LIST);
    int *qG = reinterpret_cast<int *>(scratch + 20 * LIST);                 // [G] records with several candidates
    ShardRecord *xrec = reinterpret_cast<ShardRecord *>(scratch + 24 * LIST);
    __shared__ double sMaxW[NW];
    __shared__ int sCount, sOvf, sQ, sTimeout;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) { sCount = 0; sOvf = 0; sQ = 0; if (f.ticket) *f.ticket = 0u; }      // re-arm the ticket for the next pass
    double rmax[PER], rviol[PER], rsec[PER];
    int64_t rpos[PER];
    int rcnt[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int g = t + k * NT;
        rmax[k] = -1.0; rcnt[k] = 0; rpos[k] = 0; rviol[k] = 0.0; rsec[k] = 0.0;
        if (g < G) {
            const CtaRec *r = in.rec + g;
            rmax[k] = __ldcg(&r->max); rcnt[k] = __ldcg(&r->cnt); rpos[k] = __ldcg(&r->pos0);
            rviol[k] = __ldcg(&r->viol0); rsec[k] = __ldcg(&r->second0);
        }
    }
#ifdef GNE_STAMPS
    if (t == 0 && f.stamps) f.stamps[2] = clock64();
#endif
    double m = fmax(rmax[0], rmax[1]);
#ifdef GNE_STAMPS
    if (t == 0 && f.stamps) f.stamps[3] = clock64();
#endif
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) sMaxW[w] = m;
    __syncthreads();
    double M = sMaxW[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) M = fmax(M, sMaxW[i]);
    const bool any = M > 0.0;
    const double thr = __dsub_rn(M, LV_BAND);
    if (any) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int g = t + k * NT;
            if (g >= G || rmax[k] < thr) continue;
            if (rcnt[k] == 1) {
                if (rsec[k] >= thr) sOvf = 1;                             // that thread may hide further members
                const int slot = atomicAdd(&sCount, 1);
                if (slot < LIST) { lPos[slot] = rpos[k]; lViol[slot] = rviol[k]; }
            } else if (rcnt[k] > ST_CAND) sOvf = 1;
            else qG[atomicAdd(&sQ, 1)] = g;
        }
    }
    __syncthreads();
    const int nq = sQ;
    for (int qi = w; qi < nq; qi += NW) {                                 // one warp per multi-candidate record
        const int g = qG[qi];
        const int c = __ldcg(&in.rec[g].cnt);
        for (int i = lane; i < c; i += 32) {
            const double vv = __ldcg(in.candViol + (int64_t) g * ST_CAND + i);
            if (vv >= thr) {
                if (__ldcg(in.candSecond + (int64_t) g * ST_CAND + i) >= thr) sOvf = 1;
                const int slot = atomicAdd(&sCount, 1);
                if (slot < LIST) { lPos[slot] = __ldcg(in.candPos + (int64_t) g * ST_CAND + i); lViol[slot] = vv; }
            }
        }
    }
    if (nq > 0) __syncthreads();
    const int total = sCount;
    const bool bad = any && (sOvf || total > LIST);                       // longer lists go through the ordered compaction
    if (any && !bad) {
        for (int i = t; i < total; i += NT) {                             // rank sort by position (positions are distinct)
            const int64_t mine = lPos[i];
            int rank = 0;
            for (int j = 0; j < total; ++j) rank += (lPos[j] < mine);
            lRank[rank] = i;
        }
    }
    __syncthreads();
#ifdef GNE_STAMPS
    if (t == 0 && f.stamps) f.stamps[4] = clock64();
#endif
    LvResult *res = f.res;
    // one warp publishes: entries, fence, header, fence, sequence number (towards the host the fences are PCIe
    // round trips, so as few threads as possible execute them)
    if (w == 0) {
        // entries and header go out together; every writing lane fences its own stores (one round trip towards
        // the host, all lanes in parallel), then the sequence number
        const int nOut = (any && !bad) ? total : 0;
        for (int r = lane; r < nOut; r += 32) { res->pos[r] = lPos[lRank[r]]; res->viol[r] = lViol[lRank[r]]; }
        if (lane == 0) { res->maxViol = any ? M : -1.0; res->any = any ? 1 : 0; res->overflow = bad ? 1 : 0; res->count = nOut; }
        if (lane == 0 || lane < nOut) { if (f.resOnHost) __threadfence_system(); else __threadfence(); }
        __syncwarp();
#ifdef GNE_STAMPS
        if (lane == 0 && f.stamps) f.stamps[5] = clock64();
#endif
        if (lane == 0) *reinterpret_cast<volatile unsigned long long *>(&res->seq) = f.seq;
    }
    if (f.nranks > 1 && f.peers != nullptr) {
        // sharded run: this CTA also carries the shard's record to the peers (and theirs to the host)
        const int nOut = (any && !bad) ? total : 0;
        if (t == 0) {
            xrec->maxViol = any ? M : -1.0; xrec->any = any ? 1 : 0; xrec->count = nOut;
            xrec->overflow = (bad || nOut > SHARD_LIST) ? 1 : 0;
        }
        if (t < SHARD_LIST && t < nOut) { xrec->pos[t] = lPos[lRank[t]]; xrec->viol[t] = lViol[lRank[t]]; }
        __syncthreads();
        exchange_record(xrec, f.peers, f.rank, f.nranks, f.xseq, f.hostOut, f.hostSeq, &sTimeout, f.xTimeout);
    }
}

// Per tile the critical path is: wait for the stage, read the indices from shared memory, gather,
// compare - nothing else.  Each thread keeps its largest violation (first position on ties) and its
// second largest in registers; shared memory and atomics are touched once, after the last tile.
// A thread can therefore contribute one candidate; if its second value is also within the band of
// the global maximum the pass reports overflow (it may hide a third) and the caller falls back to
// the ordered compaction - with ~75 k threads that happens only for long tie lists.
__global__ void __launch_bounds__(ST_THREADS, ST_CTAS_PER_SM)
k_price_lv_stream(NbView nb, const double *__restrict__ pi, LvCtas out, int64_t numTiles, LvFused f)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);                 // [ST_STAGES]
    unsigned char *stage0 = smem + 128;
    const int t = threadIdx.x;
    if (t == 0) {
        for (int s = 0; s < ST_STAGES; ++s) st_mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_launch_dependents();
    __syncthreads();
    pdl_wait();                                         // the relabel kernel's potentials and list edits are complete
    auto issue = [&](int s, int64_t tile) {
        unsigned char *b = stage0 + (size_t) s * ST_STAGE_BYTES;
        const int64_t l0 = tile * ST_TILE;
        st_mbar_expect_tx(&full[s], (uint32_t) ST_STAGE_BYTES);
        st_bulk_g2s(b, nb.tail + l0, 4 * ST_TILE, &full[s]);
        st_bulk_g2s(b + 4 * ST_TILE, nb.head + l0, 4 * ST_TILE, &full[s]);
        st_bulk_g2s(b + 8 * ST_TILE, nb.cost + l0, 8 * ST_TILE, &full[s]);
        st_bulk_g2s(b + 16 * ST_TILE, nb.mult + l0, 8 * ST_TILE, &full[s]);
        st_bulk_g2s(b + 24 * ST_TILE, nb.state + l0, ST_TILE, &full[s]);
    };
    if (t == 0) {
        for (int s = 0; s < ST_STAGES; ++s) {
            const int64_t tile = blockIdx.x + (int64_t) s * gridDim.x;
            if (tile < numTiles) issue(s, tile);
        }
    }
    double v1 = -1.0, v2 = -1.0;                        // largest (first position on ties) and second largest
    int64_t p1 = 0;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x, ++it) {
        const int s = it % ST_STAGES;
        st_mbar_wait(&full[s], (uint32_t) ((it / ST_STAGES) & 1));
        const unsigned char *sb = stage0 + (size_t) s * ST_STAGE_BYTES;
        const int32_t *sTail = reinterpret_cast<const int32_t *>(sb);
        const int32_t *sHead = reinterpret_cast<const int32_t *>(sb + 4 * ST_TILE);
        const double *sCost = reinterpret_cast<const double *>(sb + 8 * ST_TILE);
        const double *sMult = reinterpret_cast<const double *>(sb + 16 * ST_TILE);
        const int8_t *sState = reinterpret_cast<const int8_t *>(sb + 24 * ST_TILE);
        double pt[ST_ITEMS], ph[ST_ITEMS];
#pragma unroll
        for (int k = 0; k < ST_ITEMS; ++k) { const int o = t + k * ST_THREADS; pt[k] = ld_pi(pi, sTail[o]); ph[k] = ld_pi(pi, sHead[o]); }
#pragma unroll
        for (int k = 0; k < ST_ITEMS; ++k) {
            const int o = t + k * ST_THREADS;
            const int64_t g = nb.lo + tile * ST_TILE + o;
            const double v = (g < nb.end && g >= nb.begin) ? violation(red_cost(sCost[o], sMult[o], pt[k], ph[k]), (int) sState[o]) : -1.0;
            if (v > v1) { v2 = v1; v1 = v; p1 = g; }      // positions rise within a thread: the first maximum stays
            else if (v > v2) v2 = v;
        }
        __syncthreads();                                // everyone is done with stage s
        if (t == 0) {
            const int64_t nt = tile + (int64_t) ST_STAGES * gridDim.x;
            if (nt < numTiles) issue(s, nt);
        }
    }
#ifdef GNE_STAMPS
    const long long st0 = clock64();
#endif
    // once per CTA: maximum, then the threads within the band of it
    __shared__ int sCnt, sLast;
    __shared__ int64_t sPos[ST_CAND];
    __shared__ double sViol[ST_CAND], sSecond[ST_CAND];
    if (t == 0) sCnt = 0;
    Best b; b.v = v1; b.p = p1;
    b = block_best(b);                                  // two barriers inside: sCnt = 0 is visible afterwards
    const double Mc = b.v;
    if (Mc > 0.0) {
        const double thr = __dsub_rn(Mc, LV_BAND);
        if (v1 >= thr) {
            const int slot = atomicAdd(&sCnt, 1);
            if (slot < ST_CAND) { sPos[slot] = p1; sViol[slot] = v1; sSecond[slot] = v2; }
        }
    }
    __syncthreads();
    const int cnt = sCnt;
    if (cnt > 1) {                                      // rare: several threads of this CTA within the band
        for (int i = t; i < min(cnt, ST_CAND); i += ST_THREADS) {
            out.candPos[(int64_t) blockIdx.x * ST_CAND + i] = sPos[i];
            out.candViol[(int64_t) blockIdx.x * ST_CAND + i] = sViol[i];
            out.candSecond[(int64_t) blockIdx.x * ST_CAND + i] = sSecond[i];
        }
        if (f.ticket != nullptr) __threadfence();
        __syncthreads();
    }
    // ---- sharded runs: the last CTA to get here reduces the per-CTA records and runs the exchange (one launch
    // per pass); single GPU: the reduction is its own small kernel behind a programmatic launch boundary, which
    // costs the same latency as the fence + ticket here and keeps this kernel pure streaming
    if (t == 0) {
        CtaRec *r = out.rec + blockIdx.x;
        r->max = Mc > 0.0 ? Mc : -1.0; r->cnt = cnt; r->pad = 0;
        r->pos0 = sPos[0]; r->viol0 = sViol[0]; r->second0 = sSecond[0];       // meaningful when cnt == 1
        sLast = 0;
        if (f.ticket != nullptr) {                      // (no ticket: k_price_lv_reduce follows as its own launch)
            __threadfence();
            sLast = (atomicAdd(f.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
        }
    }
    __syncthreads();
    if (!sLast) return;
#ifdef GNE_STAMPS
    if (t == 0 && f.stamps) { f.stamps[0] = st0; f.stamps[1] = clock64(); }
#endif
    __threadfence();
    lv_reduce_records(out, (int) gridDim.x, f, stage0);
#ifdef GNE_STAMPS
    if (t == 0 && f.stamps) f.stamps[7] = clock64();
#endif
}

// The same reduction as its own launch (single-GPU passes): one CTA, resident early through PDL.
__global__ void __launch_bounds__(ST_THREADS)
k_price_lv_reduce(LvCtas in, int G, LvFused f)
{
    __shared__ __align__(16) unsigned char scratch[25 * 1024];
    pdl_launch_dependents();
    pdl_wait();
    lv_reduce_records(in, G, f, scratch);
}

} // namespace gne
