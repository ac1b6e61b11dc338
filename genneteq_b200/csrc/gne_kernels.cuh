// genneteq_b200 — sm_100a kernels of the pricing / potential path.
//
// Everything here is fp64 streaming + gather + reduction: HBM-bound, no dense contraction, so
// no tensor-core path.  Arithmetic uses __dmul_rn/__dadd_rn/__dsub_rn so
// that a multiply is never fused with the following add (the reference's x86-64 -O3 build has
// no FMA: genneteq.h:315-316, gnePotential.cpp:130) — results are bit-identical to the CPU.
//
// Nonbasic arcs live as a structure-of-arrays in setNBarc *position* order (gnePivot.cpp:527-600
// defines that order); a tile is TILE consecutive positions, one CTA per tile, thread t owning
// positions base + t + 256 k (k = 0..ITEMS-1): every load instruction of a warp covers 32
// consecutive elements (one 128-byte line for the 4-byte columns, two for the 8-byte ones), 25
// algorithmic bytes per arc.  The pass is bound by L1TEX wavefronts, not by HBM: the gather of
// pi[head] costs one wavefront per arc (32 distinct lines per warp instruction), against ~0.3 for
// everything else.  Measured on B200 (profiles/): 128-bit vector loads with 48 registers (62%
// occupancy) ran the 20M-arc pass in 121 us, this scalar layout with 32 registers (100%) in 113 us.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gne {

constexpr int TILE_THREADS = 256;
constexpr int ITEMS = 4;                   // positions per thread
constexpr int TILE = ITEMS * TILE_THREADS;  // 1024 positions per CTA
constexpr int CAND_K = 16;                 // in-band candidates kept per tile by the LV pass
constexpr int LV_LIST_CAP = 2048;          // candidates returned by the fused LV pass
constexpr int QTILE = 1024;                // virtual positions per CTA in the quickSelect scan
constexpr double TOL = 1.0e-13;            // main.h:36-43 ROUNDOFF_TOLERANCE
constexpr double LV_BAND = 4.0e-13;        // candidates are kept down to max - 4 tol (see gne_device.cu)

struct NbView {                            // shard of the nonbasic list; index = pos - lo
    const int32_t *tail;
    const int32_t *head;
    const double *cost;
    const double *mult;
    const int8_t *state;
    int64_t lo;                            // first global position of the shard (multiple of TILE)
    int64_t end;                           // min(N, hi): positions >= end are not priced
    int64_t begin;                         // positions < begin are not priced (range-restricted passes; default lo)
    int32_t tile0;                         // tile index of blockIdx.x == 0 (range-restricted passes; default 0)
};

// "fast, non-parity" mode: 0 = the reference's arithmetic, bit for bit (default); 1 = fused multiply-add
// in the reduced cost and the potential finalize - one rounding less per value, so the last bit (and with it a pivot
// tie-break) may differ from the reference; flows / potentials / objective agree to ~1e-9 relative.  Per device,
// set through gne_dev_set_fast_math.
__constant__ int c_gne_fast_math = 0;

// computeArcRedCost (genneteq.h:312-318): (cost - pi[tail]) + (mult * pi[head]), unfused
__device__ __forceinline__ double red_cost(double cost, double mult, double pit, double pih)
{
    if (c_gne_fast_math) return fma(mult, pih, __dsub_rn(cost, pit));
    return __dadd_rn(__dsub_rn(cost, pit), __dmul_rn(mult, pih));
}
// finalizePotentialsTypeI (gnePotential.cpp:130): pi = a0 + a1 * theta, unfused
__device__ __forceinline__ double pot_value(double a0, double a1, double theta)
{
    if (c_gne_fast_math) return fma(a1, theta, a0);
    return __dadd_rn(a0, __dmul_rn(a1, theta));
}

// gnePivotRules.cpp:406-407: violation magnitude |rc| or -1 when the arc is not eligible
__device__ __forceinline__ double violation(double rc, int st)
{
    bool viol = (st == 1 && rc < -TOL) || (st == 2 && rc > TOL);
    return viol ? fabs(rc) : -1.0;
}

__device__ __forceinline__ double ld_pi(const double *pi, int i) { return __ldg(pi + i); }

// Violation magnitudes of this thread's ITEMS positions of the tile starting at local index
// `tileBase`: v[k] belongs to local index tileBase + t + 256 k.
__device__ __forceinline__ void tile_violations(const NbView &nb, const double *__restrict__ pi,
                                                int64_t tileBase, double v[ITEMS], double *rcOut = nullptr)
{
    const int64_t base = tileBase + threadIdx.x;
    int tl[ITEMS], hd[ITEMS], st[ITEMS];
    double cs[ITEMS], mu[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t l = base + k * TILE_THREADS;
        tl[k] = __ldg(nb.tail + l); hd[k] = __ldg(nb.head + l);
        cs[k] = __ldg(nb.cost + l); mu[k] = __ldg(nb.mult + l);
        st[k] = __ldg(nb.state + l);
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t l = base + k * TILE_THREADS;
        const double r = red_cost(cs[k], mu[k], ld_pi(pi, tl[k]), ld_pi(pi, hd[k]));
        v[k] = (nb.lo + l < nb.end && nb.lo + l >= nb.begin) ? violation(r, st[k]) : -1.0;
        if (rcOut) rcOut[k] = r;
    }
}

// local index (within the tile) of slot k of thread t; position order == (k, t) lexicographic
__device__ __forceinline__ int tile_offset(int t, int k) { return t + k * TILE_THREADS; }

// (value, position) argmax with the lower position winning ties
struct Best { double v; int64_t p; };
__device__ __forceinline__ Best best_of(Best a, Best b)
{
    return (b.v > a.v || (b.v == a.v && b.p < a.p)) ? b : a;
}
__device__ __forceinline__ Best warp_best(Best x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
        y.p = __shfl_xor_sync(0xffffffffu, x.p, o);
        x = best_of(x, y);
    }
    return x;
}
// block-wide argmax; result valid in every thread
__device__ __forceinline__ Best block_best(Best x)
{
    __shared__ double sv[TILE_THREADS / 32];
    __shared__ int64_t sp[TILE_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    x = warp_best(x);
    __syncthreads();                       // protect sv/sp reuse across calls
    if (lane == 0) { sv[w] = x.v; sp[w] = x.p; }
    __syncthreads();
    Best r;
    r.v = sv[0]; r.p = sp[0];
#pragma unroll
    for (int i = 1; i < TILE_THREADS / 32; ++i) { Best y; y.v = sv[i]; y.p = sp[i]; r = best_of(r, y); }
    return r;
}
// block-wide exclusive scan of a small non-negative int; *total = block sum
__device__ __forceinline__ int block_excl_scan(int x, int *total)
{
    __shared__ int sw[TILE_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    __syncthreads();
    if (lane == 31) sw[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < TILE_THREADS / 32; ++i) { if (i < w) base += sw[i]; tot += sw[i]; }
    *total = tot;
    return base + inc - x;
}

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): the kernels of one pivot (update -> price -> reduce -> exchange) are
// launched back to back; each lets its successor become resident early and the successor waits here until
// its predecessor has completed and flushed.  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// LV pass, stage 1 (one CTA per tile): tile maximum, number of violators within LV_BAND of it,
// and the first CAND_K of those in position order.
// ------------------------------------------------------------------------------------------
struct LvTiles {
    double *tileMax;       // [G]  -1 when the tile has no violator
    int32_t *tileCnt;      // [G]  in-band candidates (may exceed CAND_K)
    int32_t *candOff;      // [G * CAND_K]  local offsets within the tile, position order
    double *candViol;      // [G * CAND_K]
};

__global__ void __launch_bounds__(TILE_THREADS, 8)
k_price_lv_tiles(NbView nb, const double *__restrict__ pi, LvTiles out)
{
    pdl_launch_dependents();
    pdl_wait();
    const int64_t tileBase = (int64_t) blockIdx.x * TILE;
    double v[ITEMS];
    tile_violations(nb, pi, tileBase, v);
    Best b; b.v = -1.0; b.p = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) if (v[k] > b.v) { b.v = v[k]; b.p = tile_offset(threadIdx.x, k); }
    b = block_best(b);
    __shared__ int sCnt;
    __shared__ int sOff[CAND_K];
    __shared__ double sViol[CAND_K];
    if (threadIdx.x == 0) sCnt = 0;
    __syncthreads();
    if (b.v > 0.0) {
        const double thr = __dsub_rn(b.v, LV_BAND);
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (v[k] > 0.0 && v[k] >= thr) {
                int s = atomicAdd(&sCnt, 1);
                if (s < CAND_K) { sOff[s] = tile_offset(threadIdx.x, k); sViol[s] = v[k]; }
            }
        }
    }
    __syncthreads();
    const int cnt = sCnt;
    if (threadIdx.x == 0) { out.tileMax[blockIdx.x] = b.v; out.tileCnt[blockIdx.x] = cnt; }
    if (cnt <= CAND_K && (int) threadIdx.x < cnt) {            // rank by offset -> position order
        const int mine = sOff[threadIdx.x];
        int rank = 0;
        for (int i = 0; i < cnt; ++i) rank += (sOff[i] < mine);
        out.candOff[(int64_t) blockIdx.x * CAND_K + rank] = mine;
        out.candViol[(int64_t) blockIdx.x * CAND_K + rank] = sViol[threadIdx.x];
    }
}

// Result block of the LV pass (mapped pinned host memory, written by the last CTA of the pass / the finish kernel).
// The host polls the FAST HEADER: two 16-byte chunks, each written by ONE 16-byte store and each carrying the sequence
// number of the pass, so payload and sequence number arrive together and no fence (a PCIe round trip, ~2 us) stands
// between them.  A pass with at most one candidate - nearly all of them - is complete in the header: the candidate is
// the maximum itself (hPos0, hMax).  Longer lists and overflow codes use the full block below, written and fenced
// BEFORE the header.
struct LvResult {
    double hMax;                   // chunk A: exact fp64 maximum of |rc| over violators (-1 if none) ...
    unsigned long long hSeqA;      //          ... and the sequence number of the pass
    long long hPos0;               // chunk B: position of the first candidate ...
    unsigned long long hTagB;      //          ... and seq << 16 | count (12 bits, saturating) << 4 | overflow (2 bits) << 2 | any << 1 | 1
    double maxViol;        // the full block: same maximum
    int32_t any;           // at least one violator
    int32_t overflow;      // 1: candidate list incomplete (too many in a tile / CTA / in total); 2: a peer rank never answered
    int64_t count;         // candidates with |rc| >= maxViol - LV_BAND, position order
    int64_t pos[LV_LIST_CAP];
    double viol[LV_LIST_CAP];
    unsigned long long seq;        // (kept for readers of a device-resident block)
};
__device__ __forceinline__ void lv_store16(void *p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}
// one thread: the fast header of a pass whose full block (if it is needed: count > 1 or overflow) is already visible
__device__ __forceinline__ void lv_fast_header(LvResult *res, double M, long long pos0, unsigned long long seq, long long count, int any, int overflow)
{
    const unsigned long long c = count > 4095 ? 4095ull : (unsigned long long) (count < 0 ? 0 : count);
    const unsigned long long tag = (seq << 16) | (c << 4) | ((unsigned long long) (overflow & 3) << 2) | ((unsigned long long) (any ? 1 : 0) << 1) | 1ull;
    lv_store16(&res->hMax, (unsigned long long) __double_as_longlong(M), seq);
    lv_store16(&res->hPos0, (unsigned long long) pos0, tag);
}

// LV pass, stage 2 (one CTA of 1024 threads): global maximum over the tile maxima, then the
// candidates of the (few) tiles whose maximum reaches the band, gathered in position order.
constexpr int FIN_THREADS = 1024;
constexpr int FIN_QCAP = 1024;             // tiles within the band handled on the device

__global__ void __launch_bounds__(FIN_THREADS)
k_price_lv_finish(LvTiles tiles, int G, int64_t lo, LvResult *res, unsigned long long seq)
{
    __shared__ double sMax[FIN_THREADS / 32];
    __shared__ int sQ[FIN_QCAP];
    __shared__ int sSorted[FIN_QCAP];
    __shared__ int sScan[FIN_THREADS / 32];
    __shared__ int sCount, sOvf;
    __shared__ int64_t sPos0;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) { sCount = 0; sOvf = 0; sPos0 = 0; }
    pdl_launch_dependents();
    pdl_wait();
    // pass 1: maximum of the tile maxima; loads issued in independent batches of 8 (a dependent
    // one-load-per-iteration loop costs a full memory latency per iteration)
    double m = -1.0;
    for (int g0 = t; g0 < G; g0 += 8 * FIN_THREADS) {
        double x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int g = g0 + k * FIN_THREADS; x[k] = (g < G) ? __ldg(tiles.tileMax + g) : -1.0; }
#pragma unroll
        for (int k = 0; k < 8; ++k) m = fmax(m, x[k]);
    }
    const double mine = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) sMax[w] = m;
    __syncthreads();
    double M = sMax[0];
#pragma unroll
    for (int i = 1; i < FIN_THREADS / 32; ++i) M = fmax(M, sMax[i]);
    if (M <= 0.0) {
        if (t == 0) {
            res->maxViol = -1.0; res->any = 0; res->overflow = 0; res->count = 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(&res->seq) = seq;
            lv_fast_header(res, -1.0, 0, seq, 0, 0, 0);
        }
        return;
    }
    const double thr = __dsub_rn(M, LV_BAND);
    // pass 2: only the threads that saw a tile within the band look again
    if (mine >= thr) {
        for (int g = t; g < G; g += FIN_THREADS) {
            if (__ldg(tiles.tileMax + g) >= thr) {
                const int s = atomicAdd(&sCount, 1);
                if (s < FIN_QCAP) sQ[s] = g;
            }
        }
    }
    __syncthreads();
    const int q = sCount;
    if (q > FIN_QCAP) {
        if (t == 0) {
            res->maxViol = M; res->any = 1; res->overflow = 1; res->count = 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(&res->seq) = seq;
            lv_fast_header(res, M, 0, seq, 0, 1, 1);
        }
        return;
    }
    if (t < q) {                                       // rank sort: tile order == position order
        const int mine = sQ[t];
        int rank = 0;
        for (int i = 0; i < q; ++i) rank += (sQ[i] < mine);
        sSorted[rank] = mine;
    }
    __syncthreads();
    int cnt = 0;
    if (t < q) {
        const int g = sSorted[t];
        const int c = tiles.tileCnt[g];
        if (c > CAND_K) atomicExch(&sOvf, 1);
        else for (int i = 0; i < c; ++i) cnt += (tiles.candViol[(int64_t) g * CAND_K + i] >= thr);
    }
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) sScan[w] = inc;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int i = 0; i < FIN_THREADS / 32; ++i) { if (i < w) base += sScan[i]; total += sScan[i]; }
    const bool bad = sOvf || total > LV_LIST_CAP;
    if (!bad && t < q && cnt > 0) {
        const int g = sSorted[t];
        const int c = tiles.tileCnt[g];
        int off = base + inc - cnt;
        for (int i = 0; i < c; ++i) {
            const double vv = tiles.candViol[(int64_t) g * CAND_K + i];
            if (vv >= thr) {
                const int64_t pp = lo + (int64_t) g * TILE + tiles.candOff[(int64_t) g * CAND_K + i];
                res->pos[off] = pp;
                res->viol[off] = vv;
                if (off == 0) sPos0 = pp;
                ++off;
            }
        }
        if (t != 0) __threadfence_system();             // only the threads that wrote entries pay the PCIe round trip (thread 0 fences below)
    }
    if (t == 0) {                                       // header in parallel with the entries: one round trip in all
        res->maxViol = M; res->any = 1; res->overflow = bad ? 1 : 0; res->count = total;
        __threadfence_system();
    }
    __syncthreads();
    if (t == 0) {
        *reinterpret_cast<volatile unsigned long long *>(&res->seq) = seq;
        lv_fast_header(res, M, sPos0, seq, bad ? 0 : total, 1, bad ? 1 : 0);   // (every entry was fenced above)
    }
}

// ------------------------------------------------------------------------------------------
// Ordered compaction of every violator with |rc| >= threshold: count per tile, scan, write.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS)
k_viol_count(NbView nb, const double *__restrict__ pi, double threshold, int32_t *tileCnt, double *tileVmin, double *tileVmax)
{
    double v[ITEMS];
    tile_violations(nb, pi, (int64_t) (nb.tile0 + blockIdx.x) * TILE, v);
    int c = 0;
    double lo = 1.0e308, hi = -1.0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) if (v[k] > 0.0 && v[k] >= threshold) { ++c; lo = fmin(lo, v[k]); hi = fmax(hi, v[k]); }
    int total;
    block_excl_scan(c, &total);
    __shared__ double sLo[TILE_THREADS / 32], sHi[TILE_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { sLo[threadIdx.x >> 5] = lo; sHi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < TILE_THREADS / 32; ++i) { lo = fmin(lo, sLo[i]); hi = fmax(hi, sHi[i]); }
        tileCnt[blockIdx.x] = total; tileVmin[blockIdx.x] = lo; tileVmax[blockIdx.x] = hi;
    }
}

// exclusive scan of G tile counts by one CTA; total -> *totalOut (mapped host) and tileOff[G]
__global__ void __launch_bounds__(1024)
k_scan_tiles(const int32_t *tileCnt, int64_t *tileOff, int G, int64_t *totalOut,
             const double *tileVmin, const double *tileVmax, double *minMaxOut)
{
    __shared__ int64_t sWarp[32];
    __shared__ int64_t sCarry;
    __shared__ double sLo[32], sHi[32];
    if (threadIdx.x == 0) sCarry = 0;
    {
        double lo = 1.0e308, hi = -1.0;
        for (int g = threadIdx.x; g < G; g += 1024) { lo = fmin(lo, tileVmin[g]); hi = fmax(hi, tileVmax[g]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        if ((threadIdx.x & 31) == 0) { sLo[threadIdx.x >> 5] = lo; sHi[threadIdx.x >> 5] = hi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < 32; ++i) { lo = fmin(lo, sLo[i]); hi = fmax(hi, sHi[i]); }
            minMaxOut[0] = lo; minMaxOut[1] = hi;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < G; base += 1024) {
        const int g = base + threadIdx.x;
        const int64_t x = (g < G) ? tileCnt[g] : 0;
        int64_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) sWarp[w] = inc;
        __syncthreads();
        int64_t wb = 0, tot = 0;
        for (int i = 0; i < 32; ++i) { if (i < w) wb += sWarp[i]; tot += sWarp[i]; }
        const int64_t carry = sCarry;
        if (g < G) tileOff[g] = carry + wb + inc - x;
        __syncthreads();
        if (threadIdx.x == 0) sCarry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) { tileOff[G] = sCarry; *totalOut = sCarry; }
}

__global__ void __launch_bounds__(TILE_THREADS)
k_viol_write(NbView nb, const double *__restrict__ pi, double threshold, const int64_t *tileOff,
             int64_t *outPos, double *outViol, int64_t cap)
{
    const int64_t tileBase = (int64_t) (nb.tile0 + blockIdx.x) * TILE;
    double v[ITEMS];
    tile_violations(nb, pi, tileBase, v);
    int64_t base = tileOff[blockIdx.x];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {                   // position order is (k, t) lexicographic
        const int f = (v[k] > 0.0 && v[k] >= threshold) ? 1 : 0;
        int total;
        const int off = block_excl_scan(f, &total);
        if (f) {
            const int64_t o = base + off;
            if (o < cap) {
                outPos[o] = nb.lo + tileBase + tile_offset(threadIdx.x, k);
                if (outViol) outViol[o] = v[k];
            }
        }
        base += total;
    }
}

// Position and violation of the idx-th (and of the last) violator >= threshold in position order,
// from the tile offsets of the compaction — the pick of a long tie list without moving the list.
struct SelectResult { int64_t pos; double viol; int64_t lastPos; double lastViol; };

__global__ void __launch_bounds__(TILE_THREADS)
k_viol_select(NbView nb, const double *__restrict__ pi, double threshold, const int64_t *tileOff, int G,
              int64_t idx, SelectResult *out)
{
    // blockIdx 0 resolves idx, blockIdx 1 the last element
    const int64_t total = tileOff[G];
    const int64_t want = (blockIdx.x == 0) ? idx : total - 1;
    __shared__ int sTile;
    if (threadIdx.x == 0) {
        int lo = 0, hi = G - 1;                          // last tile with tileOff[g] <= want
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tileOff[mid] <= want) lo = mid; else hi = mid - 1; }
        sTile = lo;
    }
    __syncthreads();
    const int g = sTile;
    const int64_t tileBase = (int64_t) g * TILE;
    double v[ITEMS];
    tile_violations(nb, pi, tileBase, v);
    int64_t base = tileOff[g];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int f = (v[k] > 0.0 && v[k] >= threshold) ? 1 : 0;
        int tot;
        const int off = block_excl_scan(f, &tot);
        if (f && base + off == want) {
            const int64_t p = nb.lo + tileBase + tile_offset(threadIdx.x, k);
            if (blockIdx.x == 0) { out->pos = p; out->viol = v[k]; } else { out->lastPos = p; out->lastViol = v[k]; }
        }
        base += tot;
    }
}

// getFirstVarWithViolation: lowest violating position (atomicMin over tiles)
__global__ void __launch_bounds__(TILE_THREADS)
k_price_first(NbView nb, const double *__restrict__ pi, unsigned long long *firstPos)
{
    const int64_t tileBase = (int64_t) blockIdx.x * TILE;
    double v[ITEMS];
    tile_violations(nb, pi, tileBase, v);
    unsigned long long best = ~0ull;
#pragma unroll
    for (int k = ITEMS - 1; k >= 0; --k) if (v[k] > 0.0) best = (unsigned long long) (nb.lo + tileBase + tile_offset(threadIdx.x, k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o); best = y < best ? y : best; }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(firstPos, best);
}

__global__ void __launch_bounds__(TILE_THREADS)
k_reduced_costs(NbView nb, const double *__restrict__ pi, int64_t lo, int64_t hi, double *rc)
{
    const int64_t tileBase = (int64_t) blockIdx.x * TILE;
    double v[ITEMS], r[ITEMS];
    tile_violations(nb, pi, tileBase, v, r);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t g = nb.lo + tileBase + tile_offset(threadIdx.x, k);
        if (g >= lo && g < hi && g < nb.end) rc[g - lo] = r[k];
    }
}

// ------------------------------------------------------------------------------------------
// computeEqfRedCost (genneteq.h:321-330) for every nonbasic equal-flow variable: rc = cost - sum_i nodeVals[i] * pi[i],
// the subtractions in node order.  nodeVals is dense in the reference but zero off the nodes the set's arcs touch, and
// x - 0 * pi == x exactly, so the sparse row (entries sorted by node) subtracted in order gives the same bits.  The sum
// is order-dependent, so one thread walks one row; p (the number of sets) is small and rows are short.
// ------------------------------------------------------------------------------------------
// The reference's loop is DENSE (genneteq.h:324-326): a node outside the set still contributes 0 * pi[i], which is NaN
// when pi[i] is NaN or infinite - and the reference does reach such states (a singular s x s system).  The sparse walk
// therefore counts the non-finite potentials it met; if the potential vector holds more of them (k_count_nonfinite), a
// zero entry would have multiplied one and the dense loop's result is NaN.
__global__ void k_count_nonfinite(int n, const double *__restrict__ pi, unsigned long long *__restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool bad = i < n && !isfinite(pi[i]);
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long) __popc(m));
}

__global__ void k_price_eqf(int p, const int64_t *__restrict__ rowStart, const int32_t *__restrict__ node,
                            const double *__restrict__ val, const double *__restrict__ cost, const double *__restrict__ pi,
                            const unsigned long long *__restrict__ nonFinite, double *__restrict__ rc)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p) return;
    double x = cost[k];
    unsigned long long met = 0;
    for (int64_t q = rowStart[k]; q < rowStart[k + 1]; ++q) {
        const double v = pi[node[q]];
        if (!isfinite(v)) ++met;
        x = __dsub_rn(x, __dmul_rn(val[q], v));
    }
    rc[k] = (*nonFinite > met) ? __longlong_as_double(0x7ff8000000000000ll) : x;
}

// The same against rows that stay on the device (the equal-flow engine prices its nonbasic sets every pivot): which sets, their
// costs and the number of non-finite potentials (the caller computed the potentials, so it knows) travel as kernel arguments,
// the reduced costs land in mapped pinned host memory behind a sequence word - no copy in either direction, no stream
// synchronisation; the host reads them after the arc pass it launched right behind this kernel.
#define EQF_CALL_MAX 128
struct EqfCall {
    int q;
    unsigned long long nonFinite, seq;
    int32_t id[EQF_CALL_MAX];
    double cost[EQF_CALL_MAX];
};

__global__ void __launch_bounds__(EQF_CALL_MAX)
k_price_eqf_resident(const __grid_constant__ EqfCall c, const int64_t *__restrict__ rowStart, const int32_t *__restrict__ node,
                     const double *__restrict__ val, const double *__restrict__ pi, double *out, volatile unsigned long long *seqOut)
{
    const int k = threadIdx.x;
    if (k < c.q) {
        const int r = c.id[k];
        double x = c.cost[k];
        unsigned long long met = 0;
        for (int64_t q = rowStart[r]; q < rowStart[r + 1]; ++q) {
            const double v = pi[node[q]];
            if (!isfinite(v)) ++met;
            x = __dsub_rn(x, __dmul_rn(val[q], v));
        }
        out[k] = (c.nonFinite > met) ? __longlong_as_double(0x7ff8000000000000ll) : x;
        __threadfence_system();
    }
    __syncthreads();
    if (k == 0) *seqOut = c.seq;
}

// ------------------------------------------------------------------------------------------
// quickSelect scan (gnePivotRules.cpp:742-756) over a window [jLo, jHi) of *virtual* indices
// j = (pos - cursor) mod N; one CTA per QTILE virtual indices, thread t owns j = base + t + 256 k.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double qs_violation(const NbView &nb, const double *__restrict__ pi,
                                               int64_t cursor, int64_t N, int64_t j, int64_t jHi)
{
    if (j >= jHi) return -1.0;
    int64_t pos = cursor + j;
    if (pos >= N) pos -= N;
    const int64_t l = pos - nb.lo;
    const int tl = __ldg(nb.tail + l), hd = __ldg(nb.head + l);
    const double rc = red_cost(__ldg(nb.cost + l), __ldg(nb.mult + l), ld_pi(pi, tl), ld_pi(pi, hd));
    return violation(rc, (int) __ldg(nb.state + l));
}

struct QsTiles { int32_t *cnt; double *bestV; int64_t *bestJ; };

__global__ void __launch_bounds__(TILE_THREADS)
k_qs_tiles(NbView nb, const double *__restrict__ pi, int64_t cursor, int64_t N, int64_t jLo, int64_t jHi, QsTiles out)
{
    const int64_t base = jLo + (int64_t) blockIdx.x * QTILE;
    Best b; b.v = -1.0; b.p = 0;
    int c = 0;
#pragma unroll
    for (int k = 0; k < QTILE / TILE_THREADS; ++k) {
        const int64_t j = base + threadIdx.x + k * TILE_THREADS;
        const double v = qs_violation(nb, pi, cursor, N, j, jHi);
        if (v > 0.0) { ++c; if (v > b.v) { b.v = v; b.p = j; } }
    }
    int total;
    block_excl_scan(c, &total);
    b = block_best(b);
    if (threadIdx.x == 0) { out.cnt[blockIdx.x] = total; out.bestV[blockIdx.x] = b.v; out.bestJ[blockIdx.x] = b.p; }
}

struct QsState {            // mapped pinned host memory; carried across windows
    int64_t violations;     // violators seen so far
    double bestV;           // strictly-largest |rc| so far (0 = none, as largestViol starts at 0.0)
    int64_t bestJ;          // its virtual index (-1 = none)
    int64_t stopJ;          // virtual index of the want-th violator (-1 = not reached yet)
    int32_t done;
    int32_t pad;
    unsigned long long seq; // written last, after a system-scope fence (host copy only)
};

__global__ void __launch_bounds__(TILE_THREADS)
k_qs_finish(NbView nb, const double *__restrict__ pi, int64_t cursor, int64_t N, int64_t jLo, int64_t jHi, int64_t jEnd,
            QsTiles tiles, int G, int64_t want, QsState *st, QsState *hostOut, int first, unsigned long long seq)
{
    // `st` is the device-resident carry (first != 0: start from scratch); the result is also written to
    // `hostOut` (mapped pinned) followed by a fence and the sequence number the host spins on
    const int t = threadIdx.x;
    const int64_t carry = first ? 0 : st->violations;
    const int per = (G + TILE_THREADS - 1) / TILE_THREADS;
    const int g0 = min(G, t * per), g1 = min(G, g0 + per);
    int mine = 0;
    for (int g = g0; g < g1; ++g) mine += tiles.cnt[g];
    int total;
    const int off = block_excl_scan(mine, &total);
    __shared__ int sCross;
    __shared__ int64_t sBefore;
    if (t == 0) { sCross = -1; sBefore = 0; }
    __syncthreads();
    {
        int64_t run = carry + off;
        for (int g = g0; g < g1; ++g) {
            const int c = tiles.cnt[g];
            if (run < want && run + c >= want) { sCross = g; sBefore = run; }
            run += c;
        }
    }
    __syncthreads();
    const int cross = sCross;
    const int gEnd = (cross < 0) ? G : cross;          // tiles fully inside the scan
    Best b; b.v = -1.0; b.p = 0;
    for (int g = t; g < gEnd; g += TILE_THREADS) { Best y; y.v = tiles.bestV[g]; y.p = tiles.bestJ[g]; b = best_of(b, y); }
    b = block_best(b);
    int64_t stopJ = -1;
    int64_t seen = carry + total;
    if (cross >= 0) {
        // re-price the crossing tile: find the want-th violator in j order, best up to it
        const int64_t base = jLo + (int64_t) cross * QTILE;
        int64_t need = want - sBefore;                 // >= 1
        __shared__ int64_t sStop;
        if (t == 0) sStop = -1;
        __syncthreads();
        Best bc; bc.v = -1.0; bc.p = 0;
        for (int k = 0; k < QTILE / TILE_THREADS; ++k) {
            const int64_t j = base + t + k * TILE_THREADS;
            const double v = qs_violation(nb, pi, cursor, N, j, jHi);
            const int f = v > 0.0;
            int tot;
            const int ex = block_excl_scan(f, &tot);
            const bool before = sStop < 0;
            __syncthreads();
            if (before) {
                if (f && (int64_t) ex + 1 == need) sStop = j;
                if (f && (int64_t) ex + 1 <= need && v > bc.v) { bc.v = v; bc.p = j; }
                need -= tot;                            // uniform across the block
            }
            __syncthreads();
        }
        bc = block_best(bc);
        b = best_of(b, bc);
        stopJ = sStop;
        seen = want;
    }
    if (t == 0) {
        Best prev; prev.v = first ? 0.0 : st->bestV; prev.p = first ? -1 : st->bestJ;
        // largestViol starts at 0.0 and is replaced on strictly greater: earlier j wins ties
        if (b.v > prev.v) prev = b;
        const int done = (stopJ >= 0 || jHi >= jEnd) ? 1 : 0;
        st->bestV = prev.v; st->bestJ = prev.p; st->violations = seen; st->stopJ = stopJ; st->done = done;
        hostOut->bestV = prev.v; hostOut->bestJ = prev.p; hostOut->violations = seen; hostOut->stopJ = stopJ; hostOut->done = done;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(&hostOut->seq) = seq;
    }
}

// ------------------------------------------------------------------------------------------
// Set mirror and potentials
// ------------------------------------------------------------------------------------------
struct NbMut { int32_t *tail; int32_t *head; double *cost; double *mult; int8_t *state; };

struct NbWriteOps {                        // passed by value: no copy engine involved
    int n;
    int64_t local[8];
    int32_t tail[8], head[8];
    double cost[8], mult[8];
    int8_t state[8];
};

__global__ void k_nb_write(NbMut nb, NbWriteOps ops)
{
    const int i = threadIdx.x;
    if (i < ops.n) {
        const int64_t l = ops.local[i];
        nb.tail[l] = ops.tail[i]; nb.head[l] = ops.head[i];
        nb.cost[l] = ops.cost[i]; nb.mult[l] = ops.mult[i]; nb.state[l] = ops.state[i];
    }
}

__global__ void k_pot_update_nodes(int64_t count, const int32_t *__restrict__ node, const double *__restrict__ a0,
                                   const double *__restrict__ a1, const int32_t *__restrict__ slot,
                                   double *A0, double *A1, int32_t *S)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) { const int v = node[i]; A0[v] = a0[i]; A1[v] = a1[i]; S[v] = slot[i]; }
}

__global__ void k_pot_update_thetas(int64_t count, const int32_t *__restrict__ slot, const double *__restrict__ theta, double *T)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) T[slot[i]] = theta[i];
}

// finalizePotentialsTypeI (gnePotential.cpp:127-135): pi = a0 + a1 * theta, 28 B per node
__global__ void k_pot_finalize(int n, const double *__restrict__ A0, const double *__restrict__ A1,
                               const int32_t *__restrict__ S, const double *__restrict__ T, double *__restrict__ pi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pi[i] = pot_value(A0[i], A1[i], __ldg(T + S[i]));
}

// the same for every node of (at most 8) given trees: one coalesced sweep over the slot column instead of a node
// list - 4 B per node read, 24 B per member; what a changed theta of a big tree costs (gnePotential.cpp:127-135)
struct SlotSet { int n; int32_t slot[8]; };
__global__ void k_pot_finalize_slots(int n, const double *__restrict__ A0, const double *__restrict__ A1,
                                     const int32_t *__restrict__ S, const double *__restrict__ T, double *__restrict__ pi, SlotSet ss)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t s = S[i];
    bool hit = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) hit |= (k < ss.n && ss.slot[k] == s);
    if (hit) pi[i] = pot_value(A0[i], A1[i], __ldg(T + s));
}

__global__ void k_pot_set(int64_t count, const int32_t *__restrict__ node, const double *__restrict__ val, double *pi)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) pi[node[i]] = val[i];
}

// Fused per-pivot update for small dirty sets (one CTA, inputs read straight from mapped pinned
// host memory): queued set-mirror writes, dirty (node, a0, a1, slot) records, dirty thetas, then
// — after a barrier — the finalize of the listed nodes.  One launch instead of four + two copies.
struct FusedUpdate {
    int64_t nNodes;  const int32_t *node;  const double *a0;  const double *a1;  const int32_t *slot;
    int64_t nTheta;  const int32_t *tslot; const double *theta;
    int64_t nFinal;  const int32_t *fnode;              // nodes whose pi must be recomputed
};

__global__ void __launch_bounds__(1024)
k_update_fused(NbMut nb, NbWriteOps ops, FusedUpdate u, double *A0, double *A1, int32_t *S, double *T, double *pi)
{
    const int t = threadIdx.x;
    if (t < ops.n) {
        const int64_t l = ops.local[t];
        nb.tail[l] = ops.tail[t]; nb.head[l] = ops.head[t];
        nb.cost[l] = ops.cost[t]; nb.mult[l] = ops.mult[t]; nb.state[l] = ops.state[t];
    }
    for (int64_t i = t; i < u.nNodes; i += blockDim.x) { const int v = u.node[i]; A0[v] = u.a0[i]; A1[v] = u.a1[i]; S[v] = u.slot[i]; }
    for (int64_t i = t; i < u.nTheta; i += blockDim.x) T[u.tslot[i]] = u.theta[i];
    __syncthreads();
    for (int64_t i = t; i < u.nFinal; i += blockDim.x) {
        const int v = u.fnode[i];
        pi[v] = pot_value(A0[v], A1[v], T[S[v]]);
    }
}

// The same for tiny dirty sets (the common case while trees are small): everything travels as
// kernel arguments, so the kernel touches no host memory at all.
constexpr int SU_NODES = 48, SU_THETAS = 16, SU_FINAL = 160;
struct SmallUpdate {
    NbWriteOps ops;
    int nNodes, nTheta, nFinal;
    int hasPi;                              // fpi[] holds the finalized potentials (computed by the host with the same unfused
                                            // operations): the finalize is then plain stores, no dependent loads
    double a0[SU_NODES], a1[SU_NODES], theta[SU_THETAS], fpi[SU_FINAL];
    int32_t node[SU_NODES], slot[SU_NODES], tslot[SU_THETAS], fnode[SU_FINAL];
};

__global__ void __launch_bounds__(256)
k_update_small(NbMut nb, const __grid_constant__ SmallUpdate u, double *A0, double *A1, int32_t *S, double *T, double *pi)
{
    const int t = threadIdx.x;
    pdl_launch_dependents();
    if (t < u.ops.n) {
        const int64_t l = u.ops.local[t];
        nb.tail[l] = u.ops.tail[t]; nb.head[l] = u.ops.head[t];
        nb.cost[l] = u.ops.cost[t]; nb.mult[l] = u.ops.mult[t]; nb.state[l] = u.ops.state[t];
    }
    if (t < u.nNodes) { const int v = u.node[t]; A0[v] = u.a0[t]; A1[v] = u.a1[t]; S[v] = u.slot[t]; }
    if (t < u.nTheta) T[u.tslot[t]] = u.theta[t];
    __syncthreads();
    if (t < u.nFinal) {
        const int v = u.fnode[t];
        pi[v] = u.hasPi ? u.fpi[t] : pot_value(A0[v], A1[v], T[S[v]]);
    }
}

// finalize of a node list (larger dirty sets, inputs already copied to the device)
__global__ void k_pot_finalize_list(int64_t count, const int32_t *__restrict__ fnode, const double *__restrict__ A0,
                                    const double *__restrict__ A1, const int32_t *__restrict__ S, const double *__restrict__ T,
                                    double *__restrict__ pi)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) { const int v = fnode[i]; pi[v] = pot_value(A0[v], A1[v], __ldg(T + S[v])); }
}

__global__ void k_flush_l2(double *buf, int64_t n)
{
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) buf[i] = (double) i;
}

} // namespace gne
