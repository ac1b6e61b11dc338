// genneteq_b200 — experimental variants of the LV tile kernel, timed side by side by
// gne_bench_variants (tuning harness only: none of these is on the product path; the winner gets
// promoted into gne_kernels.cuh).  All compute the same per-tile maximum of the violation.
#pragma once
#include "gne_kernels.cuh"

namespace gne {

template <bool NOALLOC> __device__ __forceinline__ int4 ld_stream_i4(const int32_t *p)
{
    int4 r;
    if (NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else r = __ldg(reinterpret_cast<const int4 *>(p));
    return r;
}
template <bool NOALLOC> __device__ __forceinline__ double2 ld_stream_d2(const double *p)
{
    double2 r;
    if (NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    else r = __ldg(reinterpret_cast<const double2 *>(p));
    return r;
}
template <bool NOALLOC> __device__ __forceinline__ uint32_t ld_stream_u32(const int8_t *p)
{
    uint32_t r;
    if (NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    else r = __ldg(reinterpret_cast<const uint32_t *>(p));
    return r;
}

struct Chunk { int4 tl, hd; double2 c01, c23, m01, m23; uint32_t st; };

template <bool NOALLOC> __device__ __forceinline__ void load_chunk(const NbView &nb, int64_t l0, Chunk &c)
{
    c.tl = ld_stream_i4<NOALLOC>(nb.tail + l0);
    c.hd = ld_stream_i4<NOALLOC>(nb.head + l0);
    c.c01 = ld_stream_d2<NOALLOC>(nb.cost + l0);
    c.c23 = ld_stream_d2<NOALLOC>(nb.cost + l0 + 2);
    c.m01 = ld_stream_d2<NOALLOC>(nb.mult + l0);
    c.m23 = ld_stream_d2<NOALLOC>(nb.mult + l0 + 2);
    c.st = ld_stream_u32<NOALLOC>(nb.state + l0);
}

__device__ __forceinline__ void price_chunk(const NbView &nb, const double *__restrict__ pi, const Chunk &c, int64_t l0, Best &b)
{
    const double pt0 = ld_pi(pi, c.tl.x);
    const double pt1 = (c.tl.y == c.tl.x) ? pt0 : ld_pi(pi, c.tl.y);
    const double pt2 = (c.tl.z == c.tl.y) ? pt1 : ld_pi(pi, c.tl.z);
    const double pt3 = (c.tl.w == c.tl.z) ? pt2 : ld_pi(pi, c.tl.w);
    const double ph0 = ld_pi(pi, c.hd.x), ph1 = ld_pi(pi, c.hd.y), ph2 = ld_pi(pi, c.hd.z), ph3 = ld_pi(pi, c.hd.w);
    const int64_t g0 = nb.lo + l0;
    double v;
    v = (g0 + 0 < nb.end) ? violation(red_cost(c.c01.x, c.m01.x, pt0, ph0), (int) (c.st & 0xff)) : -1.0;
    if (v > b.v) { b.v = v; b.p = g0; }
    v = (g0 + 1 < nb.end) ? violation(red_cost(c.c01.y, c.m01.y, pt1, ph1), (int) ((c.st >> 8) & 0xff)) : -1.0;
    if (v > b.v) { b.v = v; b.p = g0 + 1; }
    v = (g0 + 2 < nb.end) ? violation(red_cost(c.c23.x, c.m23.x, pt2, ph2), (int) ((c.st >> 16) & 0xff)) : -1.0;
    if (v > b.v) { b.v = v; b.p = g0 + 2; }
    v = (g0 + 3 < nb.end) ? violation(red_cost(c.c23.y, c.m23.y, pt3, ph3), (int) ((c.st >> 24) & 0xff)) : -1.0;
    if (v > b.v) { b.v = v; b.p = g0 + 3; }
}

// one CTA per CHUNKS*1024 positions, all stream loads issued before the first gather
template <int CHUNKS, int MINB, bool NOALLOC>
__global__ void __launch_bounds__(TILE_THREADS, MINB)
k_var_tiles(NbView nb, const double *__restrict__ pi, double *tileMax)
{
    const int64_t base = (int64_t) blockIdx.x * (CHUNKS * 1024);
    Chunk c[CHUNKS];
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) load_chunk<NOALLOC>(nb, base + k * 1024 + 4 * threadIdx.x, c[k]);
    Best b; b.v = -1.0; b.p = 0;
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) price_chunk(nb, pi, c[k], base + k * 1024 + 4 * threadIdx.x, b);
    b = block_best(b);
    if (threadIdx.x == 0) tileMax[blockIdx.x] = b.v;
}

// persistent CTAs looping over 1024-position tiles, next tile's stream loads in flight while the
// current tile's gathers are priced (register double buffering)
template <int MINB, bool NOALLOC>
__global__ void __launch_bounds__(TILE_THREADS, MINB)
k_var_persistent(NbView nb, const double *__restrict__ pi, double *ctaMax, int64_t numTiles)
{
    Best b; b.v = -1.0; b.p = 0;
    int64_t tile = blockIdx.x;
    Chunk cur, nxt;
    if (tile < numTiles) load_chunk<NOALLOC>(nb, tile * 1024 + 4 * threadIdx.x, cur);
    while (tile < numTiles) {
        const int64_t nt = tile + gridDim.x;
        if (nt < numTiles) load_chunk<NOALLOC>(nb, nt * 1024 + 4 * threadIdx.x, nxt);
        price_chunk(nb, pi, cur, tile * 1024 + 4 * threadIdx.x, b);
        cur = nxt;
        tile = nt;
    }
    b = block_best(b);
    if (threadIdx.x == 0) ctaMax[blockIdx.x] = b.v;
}

// scalar strided layout (thread t owns base + t + 256 k): narrower loads, fully coalesced per instruction
template <int ITEMS, int MINB>
__global__ void __launch_bounds__(TILE_THREADS, MINB)
k_var_strided(NbView nb, const double *__restrict__ pi, double *tileMax)
{
    const int64_t base = (int64_t) blockIdx.x * (ITEMS * TILE_THREADS) + threadIdx.x;
    int tl[ITEMS], hd[ITEMS]; double cs[ITEMS], mu[ITEMS]; int st[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t l = base + k * TILE_THREADS;
        tl[k] = __ldg(nb.tail + l); hd[k] = __ldg(nb.head + l); cs[k] = __ldg(nb.cost + l); mu[k] = __ldg(nb.mult + l); st[k] = __ldg(nb.state + l);
    }
    Best b; b.v = -1.0; b.p = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t l = base + k * TILE_THREADS;
        const double v = (nb.lo + l < nb.end) ? violation(red_cost(cs[k], mu[k], ld_pi(pi, tl[k]), ld_pi(pi, hd[k])), st[k]) : -1.0;
        if (v > b.v) { b.v = v; b.p = nb.lo + l; }
    }
    b = block_best(b);
    if (threadIdx.x == 0) tileMax[blockIdx.x] = b.v;
}

} // namespace gne

// ------------------------------------------------------------------------------------------
// TMA-staged persistent variant: one elected thread streams the five columns of the next tiles
// into shared memory with cp.async.bulk (completion on an mbarrier), all threads read their
// indices from shared memory and only the pi gathers go through the LSU / L1TEX path.
// ------------------------------------------------------------------------------------------
namespace gne {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int THREADS, int IT, int STAGES, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_var_tma(NbView nb, const double *__restrict__ pi, double *ctaMax, int64_t numTiles)
{
    constexpr int TT = THREADS * IT;                    // arcs per stage
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);                 // [STAGES]
    unsigned char *stage0 = smem + 128;
    constexpr int STAGE_BYTES = 25 * TT;
    const int t = threadIdx.x;
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int s, int64_t tile) {
        unsigned char *b = stage0 + (size_t) s * STAGE_BYTES;
        const int64_t l0 = tile * TT;
        mbar_expect_tx(&full[s], (uint32_t) STAGE_BYTES);
        bulk_g2s(b, nb.tail + l0, 4 * TT, &full[s]);
        bulk_g2s(b + 4 * TT, nb.head + l0, 4 * TT, &full[s]);
        bulk_g2s(b + 8 * TT, nb.cost + l0, 8 * TT, &full[s]);
        bulk_g2s(b + 16 * TT, nb.mult + l0, 8 * TT, &full[s]);
        bulk_g2s(b + 24 * TT, nb.state + l0, TT, &full[s]);
    };
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) {
            const int64_t tile = blockIdx.x + (int64_t) s * gridDim.x;
            if (tile < numTiles) issue(s, tile);
        }
    }
    Best b; b.v = -1.0; b.p = 0;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        const uint32_t parity = (uint32_t) ((it / STAGES) & 1);
        mbar_wait(&full[s], parity);
        const unsigned char *sb = stage0 + (size_t) s * STAGE_BYTES;
        const int32_t *sTail = reinterpret_cast<const int32_t *>(sb);
        const int32_t *sHead = reinterpret_cast<const int32_t *>(sb + 4 * TT);
        const double *sCost = reinterpret_cast<const double *>(sb + 8 * TT);
        const double *sMult = reinterpret_cast<const double *>(sb + 16 * TT);
        const int8_t *sState = reinterpret_cast<const int8_t *>(sb + 24 * TT);
        double pt[IT], ph[IT];
#pragma unroll
        for (int k = 0; k < IT; ++k) { const int o = t + k * THREADS; pt[k] = ld_pi(pi, sTail[o]); ph[k] = ld_pi(pi, sHead[o]); }
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            const int o = t + k * THREADS;
            const int64_t g = nb.lo + tile * TT + o;
            const double v = (g < nb.end) ? violation(red_cost(sCost[o], sMult[o], pt[k], ph[k]), (int) sState[o]) : -1.0;
            if (v > b.v) { b.v = v; b.p = g; }
        }
        __syncthreads();                                // everyone is done with stage s
        if (t == 0) {
            const int64_t nt = tile + (int64_t) STAGES * gridDim.x;
            if (nt < numTiles) issue(s, nt);
        }
    }
    b = block_best(b);
    if (t == 0) ctaMax[blockIdx.x] = b.v;
}

} // namespace gne
