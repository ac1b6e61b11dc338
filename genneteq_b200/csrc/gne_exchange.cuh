// genneteq_b200 — the per-pivot exchange step of a sharded Dantzig pass.
//
// Every rank prices a contiguous position range of setNBarc; what the ranks exchange per pivot is one
// fixed-size record {shard maximum, in-band candidates in position order}.  The records travel as plain
// stores into the peers' mailboxes (device memory mapped through CUDA IPC, i.e. NVLink / NVSwitch peer
// writes), followed by a sequence flag; the receiving side polls its own mailbox.  No copy engine, no
// collective library on this path.  Included by the pricing kernels so that the last CTA of a pass can
// run the exchange itself (gne_stream.cuh) and by the stand-alone exchange kernel (gne_device.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

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

// Whole CTA.  `rec` (shared memory, complete and visible to the block) is stored into every rank's
// mailbox, the sequence flag published, the peers' flags awaited (timeoutCycles, default ~4 s: a dead peer
// fails the pass instead of hanging it; the rank that gives up poisons its flag at every peer so that they
// give up in the same pass), and all records copied to mapped pinned host memory; hostSeq[0] is written
// last, after system-scope fences by every writer, so the host can poll it instead of synchronising;
// hostSeq[1] = 1 reports the timeout.
__device__ __forceinline__ void exchange_record(const ShardRecord *rec, Mailbox *const *peers, int rank, int nranks,
                                                unsigned long long seq, ShardRecord *hostOut, unsigned long long *hostSeq,
                                                int *sTimeout, long long timeoutCycles)
{
    const int t = threadIdx.x, nthr = blockDim.x;
    const int b = (int) (seq & 1ull);
    constexpr int WORDS = (int) (sizeof(ShardRecord) / 8);
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(rec);
    for (int p = 0; p < nranks; ++p) {
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(&peers[p]->rec[b][rank]);
        for (int w = t; w < WORDS; w += nthr) dst[w] = src[w];
    }
    if (t == 0) *sTimeout = 0;
    __threadfence_system();
    __syncthreads();
    if (t < nranks) *reinterpret_cast<volatile unsigned long long *>(&peers[t]->flag[b][rank]) = seq;
    Mailbox *self = peers[rank];
    if (t < nranks) {
        const long long t0 = clock64();
        while (true) {
            const unsigned long long v = *reinterpret_cast<volatile unsigned long long *>(&self->flag[b][t]);
            if (v == seq) break;
            if (v == XCHG_POISON || clock64() - t0 > timeoutCycles) { *sTimeout = 1; break; }
        }
    }
    __syncthreads();
    if (*sTimeout && t < nranks) {                      // fail fast everywhere: both generations of our flag at every peer
        *reinterpret_cast<volatile unsigned long long *>(&peers[t]->flag[0][rank]) = XCHG_POISON;
        *reinterpret_cast<volatile unsigned long long *>(&peers[t]->flag[1][rank]) = XCHG_POISON;
    }
    __threadfence_system();
    for (int r = 0; r < nranks; ++r) {
        const volatile unsigned long long *in = reinterpret_cast<const volatile unsigned long long *>(&self->rec[b][r]);
        unsigned long long *out = reinterpret_cast<unsigned long long *>(&hostOut[r]);
        for (int w = t; w < WORDS; w += nthr) out[w] = in[w];
    }
    __syncthreads();
    if (t == 0) hostSeq[1] = *sTimeout ? 1ull : 0ull;
    __threadfence_system();
    __syncthreads();
    if (t == 0) *reinterpret_cast<volatile unsigned long long *>(hostSeq) = seq;
}
