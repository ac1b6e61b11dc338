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
