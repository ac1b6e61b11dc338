// genneteq_b200 — the LV (Dantzig) pricing pass, ONE launch per pivot: persistent warp-specialised CTAs.
//
//   producer warp   one elected thread takes the next tile from a device-wide counter (dynamic scheduling: no CTA
//                   is left with one tile more than its neighbours) and streams the tile's five arc columns into
//                   shared memory with the TMA engine (cp.async.bulk + mbarrier complete_tx, 3 stages);
//   consumer warps  (8 x 32 threads) wait for a stage, read their indices from shared memory, gather pi[tail] /
//                   pi[head] through the LSU (the only L1TEX traffic left), compute the violation and release the
//                   stage through an "empty" mbarrier - no block-wide barrier per tile, warps drift freely.
//
// What used to be three launches per pivot is folded into this kernel:
//   * the relabel of the last pivot (k_update_small: <= 8 set-mirror records, <= 48 node records, <= 16 thetas,
//     <= 160 potentials to finalize, all as kernel arguments) is applied by the first CTA to arrive while every
//     CTA's TMA prologue is already in flight; the others wait for its release flag only before their first
//     gather, and patch the (at most 8) rewritten arc records into their registers where a tile holds one;
//   * the reduction over the per-CTA records is run by the last CTA to finish (ticket counter), which also resets
//     the counters for the next pass, publishes the result block to mapped pinned host memory and - in a sharded
//     run - exchanges the shard's record with the peer GPUs and merges theirs (gne_exchange.cuh).
//
// Reduction: each thread keeps its largest and second-largest violation in registers; once per CTA the threads
// within the band of the CTA maximum are recorded (CtaRec + candidate arrays).  Lists no longer than the grid
// (one tile per CTA: "single" mode) keep all four values of a thread, so every in-band candidate is recorded
// exactly.  A violator can never be skipped wrongly; a list that overflows is reported and the caller falls back
// to the ordered compaction.
#pragma once
#include "gne_kernels.cuh"
#include "gne_exchange.cuh"

namespace gne {

constexpr int ST_THREADS = 256;                         // consumer threads
constexpr int ST_PRODUCER = 32;                         // + one producer warp
constexpr int ST_UPDATER = 32;                          // + one warp that applies the folded relabel (in the elected CTA) and leaves
constexpr int ST_BLOCK = ST_THREADS + ST_PRODUCER + ST_UPDATER;
constexpr int ST_ITEMS = 4;
constexpr int ST_STAGES = 3;                            // production shape: 3 stages x 2 CTAs per SM (see st_kernel_for)
constexpr int ST_TILE = ST_THREADS * ST_ITEMS;          // 1024 arcs per stage
constexpr int ST_STAGE_BYTES = 25 * ST_TILE;            // 25 600 B
constexpr int ST_CTAS_PER_SM = 2;
constexpr int ST_CAND = 96;                             // in-band candidates kept per CTA
constexpr int ST_HEADER = 256;                          // full[], empty[] mbarriers, tile index per stage (up to 4 stages)
constexpr int ST_SMEM = ST_HEADER + ST_STAGES * ST_STAGE_BYTES;
constexpr int st_smem_bytes(int stages) { return ST_HEADER + stages * ST_STAGE_BYTES; }

__device__ __forceinline__ uint32_t st_smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(st_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void st_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(st_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(st_smem_u32(bar)) : "memory");
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
// barrier over the 256 consumer threads only (the producer warp never joins it)
__device__ __forceinline__ void st_sync_consumers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ unsigned long long st_ld_acquire(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_st_release(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ long long st_globaltimer()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct CtaRec {            // 64-byte record of one CTA; a lone in-band candidate travels in the record itself
    double max;            // -1 when the CTA saw no violator
    int32_t cnt;           // candidates within the band of `max` (may exceed ST_CAND: overflow)
    int32_t pad;
    int64_t pos0;          // cnt == 1: that candidate's position, its violation and the owning thread's second largest
    double viol0;
    double second0;
    double pad2[3];
};
struct LvCtas {
    CtaRec *rec;           // [G]
    int64_t *candPos;      // [G * ST_CAND] global positions, unordered           (read only when cnt > 1)
    double *candViol;      // [G * ST_CAND] the candidate's violation
    double *candSecond;    // [G * ST_CAND] the owning thread's second largest (in band => the thread may hide more); -1 in single mode
};
// Counters of the pass (device memory, zero between passes: the last CTA re-arms them) and where its result goes.
struct LvFused {
    unsigned int *ticket;          // last-CTA election
    unsigned int *arrive;          // first-CTA election (who applies the folded update)
    unsigned long long *tileCtr;   // next tile to stream
    unsigned long long *updFlag;   // == seq once the folded update of this pass is visible device-wide
    LvResult *res;                 // result block: mapped pinned host memory (single GPU) or device memory
    unsigned long long seq;        // published in res->seq last
    int resOnHost;                 // system-scope fences only when the host polls res
    int single;                    // one tile per CTA (grid == number of tiles): exact candidates from registers
    int nranks, rank;              // nranks > 1 and peers != nullptr: exchange + merge after the reduction ...
    Mailbox *const *peers;
    LvResult *hostOut;             // ... and the MERGED block goes here (mapped pinned host memory), with seq
    unsigned long long xseq;       // sequence number of the mailbox flags
    long long xTimeout;            // cycles a rank waits for its peers before failing the pass
    long long *stamps;             // debug (GNE_STAMPS builds): clock64 stamps of the last CTA
};
struct PotMut { double *a0; double *a1; int32_t *slot; double *theta; double *pi; };

// Reduction over the per-CTA records by the 256 consumer threads of ONE CTA (the last one of the pass): exact
// global maximum, then every recorded candidate within the band of it, sorted by position (the CTAs interleave
// tiles, so their lists are merged by rank), published to f.res; in a sharded run the same CTA instead exchanges
// the shard's record with the peers, merges the ranks' records and publishes the merged block for the host.
// `scratch` is the (drained) stage memory.  Records are read once, two per thread, straight from L2.
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
    if (t == 0) { sCount = 0; sOvf = 0; sQ = 0; }
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
    if (t == 0 && f.stamps) f.stamps[4] = st_globaltimer();              // per-CTA records loaded
#endif
    double m = fmax(rmax[0], rmax[1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) sMaxW[w] = m;
    st_sync_consumers();
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
    st_sync_consumers();
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
    if (nq > 0) st_sync_consumers();
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
    st_sync_consumers();
#ifdef GNE_STAMPS
    if (t == 0 && f.stamps) f.stamps[5] = st_globaltimer();              // merged and sorted
#endif
    const int nOut = (any && !bad) ? total : 0;
    if (!(f.nranks > 1 && f.peers != nullptr)) {
        // one warp publishes.  At most one candidate (the maximum itself) and no overflow: everything the host needs is
        // in the fast header - two 16-byte stores, no fence.  Otherwise the full block first, every writing lane fencing
        // its own stores (towards the host a fence is a PCIe round trip: all lanes pay it in parallel, once), then the header.
        if (w == 0) {
            LvResult *res = f.res;
            const bool full = nOut > 1 || bad || !f.resOnHost;
            if (full) {
                for (int r = lane; r < nOut; r += 32) { res->pos[r] = lPos[lRank[r]]; res->viol[r] = lViol[lRank[r]]; }
                if (lane == 0) { res->maxViol = any ? M : -1.0; res->any = any ? 1 : 0; res->overflow = bad ? 1 : 0; res->count = nOut; }
                if (lane == 0 || lane < nOut) { if (f.resOnHost) __threadfence_system(); else __threadfence(); }
                __syncwarp();
            }
#ifdef GNE_STAMPS
            if (lane == 0 && f.stamps) f.stamps[6] = st_globaltimer();      // full block (if any) written and fenced
#endif
            if (lane == 0) {
                if (!f.resOnHost) *reinterpret_cast<volatile unsigned long long *>(&res->seq) = f.seq;
                lv_fast_header(res, any ? M : -1.0, nOut > 0 ? lPos[lRank[0]] : 0, f.seq, nOut, any ? 1 : 0, bad ? 1 : 0);
            }
        }
        return;
    }
    // sharded run: this CTA carries the shard's record to the peers, merges theirs and publishes the merged block
    if (t == 0) {
        xrec->maxViol = any ? M : -1.0; xrec->any = any ? 1 : 0; xrec->count = nOut;
        xrec->overflow = (bad || nOut > SHARD_LIST) ? 1 : 0;
    }
    if (t < SHARD_LIST && t < nOut) { xrec->pos[t] = lPos[lRank[t]]; xrec->viol[t] = lViol[lRank[t]]; }
    st_sync_consumers();
    exchange_merge(xrec, f.peers, f.rank, f.nranks, f.xseq, f.hostOut, f.seq, &sTimeout, f.xTimeout,
                   reinterpret_cast<int64_t *>(scratch), reinterpret_cast<double *>(scratch + 8 * LIST));
}

// The folded relabel of the last pivot (what k_update_small does as its own launch), by ONE warp - the updater warp of the
// first CTA whose updater gets there; releases f.updFlag = f.seq when everything is visible device-wide.  The consumer
// warps never wait for the election: they only look at the flag before their first gather.
__device__ __forceinline__ void st_apply_update_warp(const SmallUpdate &u, const NbMut &nb, const PotMut &pm, const LvFused &f, int lane)
{
    for (int i = lane; i < u.ops.n; i += 32) {
        const int64_t l = u.ops.local[i];
        nb.tail[l] = u.ops.tail[i]; nb.head[l] = u.ops.head[i];
        nb.cost[l] = u.ops.cost[i]; nb.mult[l] = u.ops.mult[i]; nb.state[l] = u.ops.state[i];
    }
    for (int i = lane; i < u.nNodes; i += 32) { const int v = u.node[i]; pm.a0[v] = u.a0[i]; pm.a1[v] = u.a1[i]; pm.slot[v] = u.slot[i]; }
    for (int i = lane; i < u.nTheta; i += 32) pm.theta[u.tslot[i]] = u.theta[i];
    if (u.hasPi) {                                       // the host sent the finalized values: plain stores, no loads
        for (int i = lane; i < u.nFinal; i += 32) pm.pi[u.fnode[i]] = u.fpi[i];
    } else {
        __syncwarp();                                    // the records and thetas above are visible to the whole warp
        for (int i = lane; i < u.nFinal; i += 32) {
            const int v = u.fnode[i];
            pm.pi[v] = pot_value(pm.a0[v], pm.a1[v], pm.theta[pm.slot[v]]);
        }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) st_st_release(f.updFlag, f.seq);
}

// One CTA of a pass: CTA `blk` of the `grd` CTAs that price this list (a launch of its own: block index and grid size of
// that launch; a batched launch: the instance's share of the grid).  All arguments live in kernel-parameter space.
template <int STAGES>
__device__ __forceinline__ void lv_stream_cta(const NbView &nb, const double *pi, const LvCtas &out, const int64_t numTiles,
                                              const LvFused &f, const SmallUpdate &u, const NbMut &nbMut, const PotMut &pm,
                                              const int blk, const int grd, unsigned char *smem)
{
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);                 // [STAGES]
    uint64_t *empty = full + 4;                                  // [STAGES]
    volatile int64_t *sTile = reinterpret_cast<volatile int64_t *>(smem + 128);   // [STAGES] tile in each stage, -1: none left
    unsigned char *stage0 = smem + ST_HEADER;
    const int t = threadIdx.x;
#ifdef GNE_STAMPS
    if (t == 0 && f.stamps) atomicMin(reinterpret_cast<unsigned long long *>(f.stamps + 0), (unsigned long long) st_globaltimer());   // first CTA in
#endif
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) { st_mbar_init(&full[s], 1); st_mbar_init(&empty[s], ST_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t >= ST_THREADS) {
        // ---------------- producer warp: one thread streams tiles until the counter runs out
        if (t == ST_THREADS) {
            for (int it = 0;; ++it) {
                const int s = it % STAGES;
                if (it >= STAGES) st_mbar_wait(&empty[s], (uint32_t) (((it / STAGES) - 1) & 1));
                // the first STAGES tiles of a CTA are its own (no round trip to the counter before the first load is
                // issued); from then on the next tile comes from the device-wide counter
                const int64_t tile = f.single ? (it == 0 ? (int64_t) blk : numTiles)
                                   : (it < STAGES ? (int64_t) blk + (int64_t) it * grd
                                                  : (int64_t) STAGES * grd + (int64_t) atomicAdd(f.tileCtr, 1ull));
                if (tile >= numTiles) { sTile[s] = -1; st_mbar_arrive(&full[s]); break; }
                sTile[s] = tile;
                unsigned char *b = stage0 + (size_t) s * ST_STAGE_BYTES;
                const int64_t l0 = tile * ST_TILE;
                st_mbar_expect_tx(&full[s], (uint32_t) ST_STAGE_BYTES);
                st_bulk_g2s(b, nb.tail + l0, 4 * ST_TILE, &full[s]);
                st_bulk_g2s(b + 4 * ST_TILE, nb.head + l0, 4 * ST_TILE, &full[s]);
                st_bulk_g2s(b + 8 * ST_TILE, nb.cost + l0, 8 * ST_TILE, &full[s]);
                st_bulk_g2s(b + 16 * ST_TILE, nb.mult + l0, 8 * ST_TILE, &full[s]);
                st_bulk_g2s(b + 24 * ST_TILE, nb.state + l0, ST_TILE, &full[s]);
            }
        }
        if (t >= ST_THREADS + ST_PRODUCER) {
            // ---------------- updater warp: the first one to arrive (device-wide) applies the folded relabel
            const int lane = t & 31;
            if ((u.ops.n | u.nNodes | u.nTheta | u.nFinal) != 0) {
                unsigned int first = 0u;
                if (lane == 0) first = (atomicAdd(f.arrive, 1u) == 0u) ? 1u : 0u;
                first = __shfl_sync(0xffffffffu, first, 0);
                if (first) st_apply_update_warp(u, nbMut, pm, f, lane);
            }
        }
        return;
    }
    // ---------------- consumers
    const int lane = t & 31;
    const bool hasUpd = (u.ops.n | u.nNodes | u.nTheta | u.nFinal) != 0;
    double v1 = -1.0, v2 = -1.0;                        // largest (first position on ties) and second largest
    int64_t p1 = 0;
    double vv[ST_ITEMS];                                // single mode: the thread's four values
    int64_t tileOfVv = 0;
#pragma unroll
    for (int k = 0; k < ST_ITEMS; ++k) vv[k] = -1.0;
    for (int it = 0;; ++it) {
        const int s = it % STAGES;
        st_mbar_wait(&full[s], (uint32_t) ((it / STAGES) & 1));
        const int64_t tile = sTile[s];
        if (tile < 0) break;
        const unsigned char *sb = stage0 + (size_t) s * ST_STAGE_BYTES;
        const int32_t *sTail = reinterpret_cast<const int32_t *>(sb);
        const int32_t *sHead = reinterpret_cast<const int32_t *>(sb + 4 * ST_TILE);
        const double *sCost = reinterpret_cast<const double *>(sb + 8 * ST_TILE);
        const double *sMult = reinterpret_cast<const double *>(sb + 16 * ST_TILE);
        const int8_t *sState = reinterpret_cast<const int8_t *>(sb + 24 * ST_TILE);
        int tl[ST_ITEMS], hd[ST_ITEMS];
#pragma unroll
        for (int k = 0; k < ST_ITEMS; ++k) { const int o = t + k * ST_THREADS; tl[k] = sTail[o]; hd[k] = sHead[o]; }
        // a tile that holds one of the rewritten arc records (at most 8 per pivot, usually 2-3 in the whole list):
        // the stream may have fetched it before or after the update wrote it - take the record from the arguments
        bool patched = false;
        for (int i = 0; i < u.ops.n; ++i) patched |= (u.ops.local[i] / ST_TILE == tile);
        double cs[ST_ITEMS], mu[ST_ITEMS];
        int st[ST_ITEMS];
        if (patched) {
#pragma unroll
            for (int k = 0; k < ST_ITEMS; ++k) {
                const int o = t + k * ST_THREADS;
                cs[k] = sCost[o]; mu[k] = sMult[o]; st[k] = (int) sState[o];
                for (int i = 0; i < u.ops.n; ++i)
                    if (u.ops.local[i] == tile * ST_TILE + o) { tl[k] = u.ops.tail[i]; hd[k] = u.ops.head[i]; cs[k] = u.ops.cost[i]; mu[k] = u.ops.mult[i]; st[k] = (int) u.ops.state[i]; }
            }
        }
        if (it == 0 && hasUpd) {                        // the relabel must be visible before the first gather:
            if (t == 0) while (st_ld_acquire(f.updFlag) != f.seq) { }   // one poller per CTA (2 368 warps on one line would
            st_sync_consumers();                                        // delay the very store they wait for)
        }
#ifdef GNE_STAMPS
        if (it == 0 && t == 0 && f.stamps) atomicMax(reinterpret_cast<unsigned long long *>(f.stamps + 1), (unsigned long long) st_globaltimer());   // last CTA to reach its first gather
#endif
        double pt[ST_ITEMS], ph[ST_ITEMS];
#pragma unroll
        for (int k = 0; k < ST_ITEMS; ++k) { pt[k] = pi[tl[k]]; ph[k] = pi[hd[k]]; }
#pragma unroll
        for (int k = 0; k < ST_ITEMS; ++k) {
            const int o = t + k * ST_THREADS;
            const int64_t g = nb.lo + tile * ST_TILE + o;
            const double c = patched ? cs[k] : sCost[o], m = patched ? mu[k] : sMult[o];
            const int sta = patched ? st[k] : (int) sState[o];
            const double v = (g < nb.end && g >= nb.begin) ? violation(red_cost(c, m, pt[k], ph[k]), sta) : -1.0;
            vv[k] = v;
            if (v > v1) { v2 = v1; v1 = v; p1 = g; }      // tiles come in rising order and positions rise within a tile: the first maximum stays
            else if (v > v2) v2 = v;
        }
        tileOfVv = tile;
        __syncwarp();
        if (lane == 0) st_mbar_arrive(&empty[s]);       // this warp is done with stage s
    }
#ifdef GNE_STAMPS
    const long long st0 = st_globaltimer();
#endif
    // ---------------- once per CTA: maximum, then the candidates within the band of it
    __shared__ int sCnt, sLast;
    __shared__ double sWarpMax[ST_THREADS / 32];
    __shared__ int64_t sPos[ST_CAND];
    __shared__ double sViol[ST_CAND], sSecond[ST_CAND];
    if (t == 0) sCnt = 0;
    double mx = v1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) sWarpMax[t >> 5] = mx;
    st_sync_consumers();
    double Mc = sWarpMax[0];
#pragma unroll
    for (int i = 1; i < ST_THREADS / 32; ++i) Mc = fmax(Mc, sWarpMax[i]);
    if (Mc > 0.0) {
        const double thr = __dsub_rn(Mc, LV_BAND);
        if (f.single) {
#pragma unroll
            for (int k = 0; k < ST_ITEMS; ++k) {
                if (vv[k] >= thr) {
                    const int slot = atomicAdd(&sCnt, 1);
                    if (slot < ST_CAND) { sPos[slot] = nb.lo + tileOfVv * ST_TILE + t + k * ST_THREADS; sViol[slot] = vv[k]; sSecond[slot] = -1.0; }
                }
            }
        } else if (v1 >= thr) {
            const int slot = atomicAdd(&sCnt, 1);
            if (slot < ST_CAND) { sPos[slot] = p1; sViol[slot] = v1; sSecond[slot] = v2; }
        }
    }
    st_sync_consumers();
    const int cnt = sCnt;
    if (cnt > 1) {                                      // rare: several candidates of this CTA within the band
        for (int i = t; i < min(cnt, ST_CAND); i += ST_THREADS) {
            out.candPos[(int64_t) blk * ST_CAND + i] = sPos[i];
            out.candViol[(int64_t) blk * ST_CAND + i] = sViol[i];
            out.candSecond[(int64_t) blk * ST_CAND + i] = sSecond[i];
        }
        __threadfence();
        st_sync_consumers();
    }
    if (t == 0) {
        CtaRec *r = out.rec + blk;
        r->max = Mc > 0.0 ? Mc : -1.0; r->cnt = cnt; r->pad = 0;
        r->pos0 = sPos[0]; r->viol0 = sViol[0]; r->second0 = sSecond[0];       // meaningful when cnt == 1
        __threadfence();
        sLast = (atomicAdd(f.ticket, 1u) == (unsigned int) grd - 1u) ? 1 : 0;
    }
    st_sync_consumers();
    if (!sLast) return;
    // ---------------- the last CTA of the pass: re-arm the counters, reduce, publish (and exchange)
#ifdef GNE_STAMPS
    if (t == 0 && f.stamps) { f.stamps[2] = st0; f.stamps[3] = st_globaltimer(); }     // last CTA: loop done / ticket taken
#endif
    __threadfence();
    if (t == 0) { *f.ticket = 0u; *f.arrive = 0u; *f.tileCtr = 0ull; }
    lv_reduce_records(out, grd, f, stage0);
#ifdef GNE_STAMPS
    if (t == 0 && f.stamps) f.stamps[7] = st_globaltimer();              // last CTA done
#endif
}

template <int STAGES, int MINB>
__global__ void __launch_bounds__(ST_BLOCK, MINB)
k_price_lv_stream(const __grid_constant__ NbView nb, const double *pi, const __grid_constant__ LvCtas out, const int64_t numTiles,
                  const __grid_constant__ LvFused f, const __grid_constant__ SmallUpdate u, const __grid_constant__ NbMut nbMut,
                  const __grid_constant__ PotMut pm)
{
    extern __shared__ __align__(128) unsigned char smem[];
    lv_stream_cta<STAGES>(nb, pi, out, numTiles, f, u, nbMut, pm, (int) blockIdx.x, (int) gridDim.x, smem);
}

// The passes of several INDEPENDENT instances in one launch (configs[3]: a batch of instances on one GPU).  The grid is
// cut into equal shares of `ctasPer` CTAs; share i runs instance i's pass exactly as a launch of its own would (its own
// counters, result block, folded relabel), with `grid` <= ctasPer CTAs taking part.  Everything - up to LVB_MAX complete
// argument sets - travels as kernel parameters: one launch per group of instances instead of one per instance, and no
// copy engine involved.
struct LvPass {
    NbView nb;
    const double *pi;
    LvCtas out;
    int64_t numTiles;
    LvFused f;
    SmallUpdate u;
    NbMut nbMut;
    PotMut pm;
    int grid;
};
constexpr int LVB_MAX = 8;
struct LvBatch {
    int count, ctasPer;
    LvPass p[LVB_MAX];
};
static_assert(sizeof(LvBatch) <= 32764, "kernel parameters are limited to 32 764 bytes");

template <int STAGES, int MINB>
__global__ void __launch_bounds__(ST_BLOCK, MINB)
k_price_lv_batch(const __grid_constant__ LvBatch B)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int inst = (int) blockIdx.x / B.ctasPer;
    const int blk = (int) blockIdx.x - inst * B.ctasPer;
    const LvPass &P = B.p[inst];
    if (blk >= P.grid) return;
    lv_stream_cta<STAGES>(P.nb, P.pi, P.out, P.numTiles, P.f, P.u, P.nbMut, P.pm, blk, P.grid, smem);
}

} // namespace gne
