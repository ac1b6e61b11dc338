// genneteq_b200 — the per-pivot exchange step of a sharded Dantzig pass.
//
// Every rank prices a contiguous position range of setNBarc; what the ranks exchange per pivot is one
// fixed-size record {shard maximum, in-band candidates in position order}.  The records travel as plain
// stores into the peers' mailboxes (device memory mapped through CUDA IPC, i.e. NVLink / NVSwitch peer
// writes), followed by a sequence flag; the receiving side polls its own mailbox and MERGES the ranks'
// records on the device (global maximum, candidates within the band of it in rank = position order), so
// that the host reads the same small result block as on one GPU instead of nranks records.  No copy
// engine, no collective library on this path.  Included by the pricing kernel, whose last CTA runs the
// exchange itself (gne_stream.cuh), and by the stand-alone exchange kernel (gne_device.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "gne_kernels.cuh"

constexpr int SHARD_LIST = 48;                          // candidates carried per rank in the record
constexpr int MAX_RANKS = 8;
struct ShardRecord {                                    // 24 + 48*16 = 792 bytes
    double maxViol;
    int32_t any, overflow;
    int64_t count;
    int64_t pos[SHARD_LIST];
    double viol[SHARD_LIST];
};
// Receive buffer of one rank: slot r is written by rank r with NVLink stores.  Two generations (seq
// parity) so that a fast rank's next record never overwrites one still being read.
constexpr int BLOB_WORDS = 16;                          // 128-byte blobs of the sharded quickSelect
struct Mailbox {
    ShardRecord rec[2][MAX_RANKS];
    unsigned long long flag[2][MAX_RANKS];
    unsigned long long blob[2][MAX_RANKS][BLOB_WORDS];
    unsigned long long bflag[2][MAX_RANKS];
};

constexpr unsigned long long XCHG_POISON = ~0ull;       // flag value a timed-out rank leaves with every peer: they fail in the same pass
constexpr int XCHG_TIMEOUT = 2;                         // LvResult::overflow value reporting that a peer never answered

// 256 threads synchronised through named barrier 1 (the consumer threads of the pricing kernel, or the
// whole block of the stand-alone exchange kernel).
__device__ __forceinline__ void xchg_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// `rec` (shared memory, complete and visible to the 256 threads) is stored into every rank's mailbox, the
// sequence flag published, the peers' flags awaited (timeoutCycles, default ~4 s: a dead peer fails the pass
// instead of hanging it; the rank that gives up poisons its flag at every peer so that they give up in the
// same pass).  Then the merge every rank computes identically: M = max over the ranks' maxima, the ranks'
// in-band candidates concatenated in rank order (= position order), overflow if an in-band rank overflowed.
// The merged block goes to `out` (mapped pinned host memory) - entries, header, a system-scope fence by every
// writer, then out->seq = outSeq, which the host polls.  mPos / mViol: shared scratch for nranks x SHARD_LIST.
__device__ __forceinline__ void exchange_merge(const ShardRecord *rec, Mailbox *const *peers, int rank, int nranks,
                                               unsigned long long xseq, gne::LvResult *out, unsigned long long outSeq,
                                               int *sTimeout, long long timeoutCycles, int64_t *mPos, double *mViol)
{
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    constexpr int NT = 256;
    const int b = (int) (xseq & 1ull);
    constexpr int WORDS = (int) (sizeof(ShardRecord) / 8);
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(rec);
    static_assert(WORDS == 3 + 2 * SHARD_LIST, "ShardRecord layout");
    // only what the record holds travels: the 24-byte header and `count` (position, violation) pairs - usually one;
    // thread i < 3 + 2 cnt carries one word to every peer
    const int cnt = (rec->count < 0 || rec->overflow) ? 0 : (rec->count > SHARD_LIST ? SHARD_LIST : (int) rec->count);
    if (t < 3 + 2 * cnt) {
        const int wi = t < 3 + cnt ? t : t - cnt + SHARD_LIST;           // words [0, 3 + cnt) and [3 + SHARD_LIST, 3 + SHARD_LIST + cnt)
        const unsigned long long v = src[wi];
        for (int p = 0; p < nranks; ++p) reinterpret_cast<unsigned long long *>(&peers[p]->rec[b][rank])[wi] = v;
    }
    if (t == 0) *sTimeout = 0;
    xchg_sync();
    // the barrier orders the CTA's stores before the flag writers' fence (cumulativity): one system-scope fence per
    // flag-writing thread instead of one per thread
    if (t < nranks) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(&peers[t]->flag[b][rank]) = xseq;
    }
    Mailbox *self = peers[rank];
    if (t < nranks) {
        const long long t0 = clock64();
        while (true) {
            const unsigned long long v = *reinterpret_cast<volatile unsigned long long *>(&self->flag[b][t]);
            if (v == xseq) break;
            if (v == XCHG_POISON || clock64() - t0 > timeoutCycles) { *sTimeout = 1; break; }
        }
    }
    if (t < nranks) __threadfence_system();             // acquire side: the records behind the observed flags
    xchg_sync();
    if (*sTimeout && t < nranks) {                      // fail fast everywhere: both generations of our flag at every peer
        *reinterpret_cast<volatile unsigned long long *>(&peers[t]->flag[0][rank]) = XCHG_POISON;
        *reinterpret_cast<volatile unsigned long long *>(&peers[t]->flag[1][rank]) = XCHG_POISON;
    }
    // ---- merge (identical on every rank)
    __shared__ double sMax[MAX_RANKS];
    __shared__ int sAnyR[MAX_RANKS], sOvfR[MAX_RANKS], sCntR[MAX_RANKS], sWarpTot[NT / 32], sCarry;
    if (t < nranks) {
        const volatile ShardRecord *r = &self->rec[b][t];
        sMax[t] = r->maxViol; sAnyR[t] = r->any; sOvfR[t] = r->overflow; sCntR[t] = (int) r->count;
    }
    if (t == 0) sCarry = 0;
    xchg_sync();
    double M = -1.0;
    bool any = false;
    for (int r = 0; r < nranks; ++r) if (sAnyR[r]) { any = true; M = fmax(M, sMax[r]); }
    const double thr = __dsub_rn(M, gne::LV_BAND);
    bool bad = false;
    for (int r = 0; r < nranks; ++r) if (sAnyR[r] && sMax[r] >= thr && sOvfR[r]) bad = true;
    int total = 0;
    if (any && !bad) {
        for (int base = 0; base < nranks * SHARD_LIST; base += NT) {
            const int idx = base + t, r = idx / SHARD_LIST, i = idx % SHARD_LIST;
            int keep = 0;
            int64_t p = 0;
            double v = 0.0;
            if (r < nranks && sAnyR[r] && sMax[r] >= thr && i < sCntR[r]) {
                v = *reinterpret_cast<const volatile double *>(&self->rec[b][r].viol[i]);
                p = *reinterpret_cast<const volatile int64_t *>(&self->rec[b][r].pos[i]);
                keep = v >= thr;
            }
            int inc = keep;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
            if (lane == 31) sWarpTot[w] = inc;
            xchg_sync();
            int off = sCarry, tot = 0;
#pragma unroll
            for (int q = 0; q < NT / 32; ++q) { if (q < w) off += sWarpTot[q]; tot += sWarpTot[q]; }
            if (keep) { mPos[off + inc - 1] = p; mViol[off + inc - 1] = v; }
            xchg_sync();
            if (t == 0) sCarry += tot;
            xchg_sync();
        }
        total = sCarry;
    }
    // ---- publish (one warp): the fast header alone when it says everything, else the full block first (fenced)
    if (w == 0) {
        const int nOut = (any && !bad) ? total : 0;
        const int ovf = *sTimeout ? XCHG_TIMEOUT : (bad ? 1 : 0);
        if (nOut > 1 || ovf) {
            for (int r = lane; r < nOut; r += 32) { out->pos[r] = mPos[r]; out->viol[r] = mViol[r]; }
            if (lane == 0) { out->maxViol = any ? M : -1.0; out->any = any ? 1 : 0; out->count = nOut; out->overflow = ovf; }
            if (lane == 0 || lane < nOut) __threadfence_system();
            __syncwarp();
        }
        if (lane == 0) gne::lv_fast_header(out, any ? M : -1.0, nOut > 0 ? mPos[0] : 0, outSeq, nOut, any ? 1 : 0, ovf);
    }
}
