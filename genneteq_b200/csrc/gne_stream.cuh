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
// Reduction: each CTA keeps a running maximum in shared memory and a short list of the violators
// that reached (running max - band) when they were priced; the exact global maximum and the final
// in-band list are produced by k_price_lv_finish_stream from the per-CTA records.  A violator can
// never be skipped wrongly (the running maximum is a lower bound of the final one); a list that
// overflows is reported and the caller falls back to the ordered compaction.
#pragma once
#include "gne_kernels.cuh"

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

struct LvCtas {
    double *ctaMax;        // [G]  -1 when the CTA saw no violator
    int32_t *ctaCnt;       // [G]  candidates recorded (may exceed ST_CAND: overflow)
    int64_t *candPos;      // [G * ST_CAND] global positions, unordered
    double *candViol;      // [G * ST_CAND] the thread's largest violation
    double *candSecond;    // [G * ST_CAND] the same thread's second largest (in band => the thread may hide more)
};

// Per tile the critical path is: wait for the stage, read the indices from shared memory, gather,
// compare - nothing else.  Each thread keeps its largest violation (first position on ties) and its
// second largest in registers; shared memory and atomics are touched once, after the last tile.
// A thread can therefore contribute one candidate; if its second value is also within the band of
// the global maximum the pass reports overflow (it may hide a third) and the caller falls back to
// the ordered compaction - with ~75 k threads that happens only for long tie lists.
__global__ void __launch_bounds__(ST_THREADS, ST_CTAS_PER_SM)
k_price_lv_stream(NbView nb, const double *__restrict__ pi, LvCtas out, int64_t numTiles)
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
    // once per CTA: maximum, then the threads within the band of it
    __shared__ int sCnt;
    __shared__ int64_t sPos[ST_CAND];
    __shared__ double sViol[ST_CAND], sSecond[ST_CAND];
    if (t == 0) sCnt = 0;
    Best b; b.v = v1; b.p = p1;
    b = block_best(b);                                  // two barriers inside: sCnt = 0 is visible afterwards
    const double M = b.v;
    if (M > 0.0) {
        const double thr = __dsub_rn(M, LV_BAND);
        if (v1 >= thr) {
            const int slot = atomicAdd(&sCnt, 1);
            if (slot < ST_CAND) { sPos[slot] = p1; sViol[slot] = v1; sSecond[slot] = v2; }
        }
    }
    __syncthreads();
    const int cnt = sCnt;
    if (t == 0) { out.ctaMax[blockIdx.x] = M > 0.0 ? M : -1.0; out.ctaCnt[blockIdx.x] = cnt; }
    for (int i = t; i < min(cnt, ST_CAND); i += ST_THREADS) {
        out.candPos[(int64_t) blockIdx.x * ST_CAND + i] = sPos[i];
        out.candViol[(int64_t) blockIdx.x * ST_CAND + i] = sViol[i];
        out.candSecond[(int64_t) blockIdx.x * ST_CAND + i] = sSecond[i];
    }
}

// Reduction over the per-CTA records: exact global maximum, then every recorded candidate within the
// band of it, sorted by position (the CTAs interleave tiles, so their lists are merged by rank).
__global__ void __launch_bounds__(1024)
k_price_lv_finish_stream(LvCtas in, int G, LvResult *res, unsigned long long seq)
{
    __shared__ double sMax[32];
    __shared__ int64_t sPos[LV_LIST_CAP];
    __shared__ double sViol[LV_LIST_CAP];
    __shared__ int sCount, sOvf;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) { sCount = 0; sOvf = 0; }
    pdl_launch_dependents();
    pdl_wait();
    double m = -1.0;
    for (int g = t; g < G; g += 1024) m = fmax(m, in.ctaMax[g]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) sMax[w] = m;
    __syncthreads();
    double M = sMax[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) M = fmax(M, sMax[i]);
    if (M <= 0.0) {
        if (t == 0) {
            res->maxViol = -1.0; res->any = 0; res->overflow = 0; res->count = 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(&res->seq) = seq;
        }
        return;
    }
    const double thr = __dsub_rn(M, LV_BAND);
    // one warp per CTA record whose maximum reaches the band
    for (int g = w; g < G; g += 32) {
        if (in.ctaMax[g] < thr) continue;
        const int c = in.ctaCnt[g];
        if (c > ST_CAND) { if (lane == 0) sOvf = 1; continue; }
        for (int i = lane; i < c; i += 32) {
            const double vv = in.candViol[(int64_t) g * ST_CAND + i];
            if (vv >= thr) {
                if (in.candSecond[(int64_t) g * ST_CAND + i] >= thr) sOvf = 1;      // that thread may hide further members
                const int slot = atomicAdd(&sCount, 1);
                if (slot < LV_LIST_CAP) { sPos[slot] = in.candPos[(int64_t) g * ST_CAND + i]; sViol[slot] = vv; }
            }
        }
    }
    __syncthreads();
    const int total = sCount;
    const bool bad = sOvf || total > 1024;              // longer lists go through the ordered compaction
    __shared__ int sRank[1024];
    if (!bad) {
        for (int i = t; i < total; i += 1024) {         // rank sort by position (positions are distinct)
            const int64_t mine = sPos[i];
            int rank = 0;
            for (int j = 0; j < total; ++j) rank += (sPos[j] < mine);
            sRank[rank] = i;
        }
    }
    __syncthreads();
    // one warp publishes: entries, system-scope fence, header, fence, sequence number (the fences are PCIe
    // round trips, so as few threads as possible execute them)
    if (w == 0) {
        if (!bad) for (int r = lane; r < total; r += 32) { res->pos[r] = sPos[sRank[r]]; res->viol[r] = sViol[sRank[r]]; }
        if (!bad && total > 0) __threadfence_system();
        __syncwarp();
        if (lane == 0) {
            res->maxViol = M; res->any = 1; res->overflow = bad ? 1 : 0; res->count = total;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(&res->seq) = seq;
        }
    }
}

} // namespace gne
