// radix.cuh — stable LSD radix sort passes over (64-bit key, 64-bit payload) pairs, shared by the ordered_set finalisation
// (hashset.cu) and the list aggregator (list.cu).
#pragma once
#include "common.cuh"

namespace b200 {

// ---- LSD radix sort of (64-bit sort key, 64-bit payload) pairs, 8 bits per pass, stable --------------------------------------
// Replaces round 1's bitonic network (210 launches for 2^20 entries) + a 24 MB download + a host loop.  One pass = per-block digit
// histograms (digit-major), one exclusive scan over the 256 x nblocks matrix, a stable scatter.  `from_val`: the digit comes from
// the payload's upper word (the shard) instead of the key.
constexpr int kRadixThreads = 256;

static __global__ void __launch_bounds__(kRadixThreads) k_radix_hist(const unsigned long long *key, const unsigned long long *val, unsigned long long n, int shift,
                                                              int from_val, unsigned *hist, unsigned nblk) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long i = (unsigned long long)blockIdx.x * kRadixThreads + threadIdx.x;
    if (i < n) {
        const unsigned long long x = from_val ? (val[i] >> 32) : key[i];
        atomicAdd(&h[(x >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(unsigned long long)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

static __global__ void __launch_bounds__(kRadixThreads) k_radix_scatter(const unsigned long long *key, const unsigned long long *val, unsigned long long *key_out,
                                                                 unsigned long long *val_out, unsigned long long n, int shift, int from_val,
                                                                 const unsigned *hist, unsigned nblk) {
    __shared__ unsigned wcnt[kRadixThreads / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int w = 0; w < kRadixThreads / 32; w++)
        wcnt[w][threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long i = (unsigned long long)blockIdx.x * kRadixThreads + threadIdx.x;
    const bool live = i < n;
    unsigned long long k = 0, v = 0;
    unsigned d = 0;
    if (live) {
        k = key[i];
        v = val[i];
        d = (unsigned)(((from_val ? (v >> 32) : k) >> shift) & 255u);
    }
    const unsigned act = __ballot_sync(0xffffffffu, live);
    unsigned lrank = 0;
    if (live) {
        const unsigned m = __match_any_sync(act, d);
        lrank = __popc(m & ((1u << lane) - 1u));
        if (lrank == 0)
            wcnt[warp][d] = __popc(m);
    }
    __syncthreads();
    { // exclusive prefix over the warps, per digit (thread t owns digit t)
        unsigned run = 0;
        for (int w = 0; w < kRadixThreads / 32; w++) {
            const unsigned t = wcnt[w][threadIdx.x];
            wcnt[w][threadIdx.x] = run;
            run += t;
        }
    }
    __syncthreads();
    if (live) {
        const unsigned long long pos = (unsigned long long)hist[(unsigned long long)d * nblk + blockIdx.x] + wcnt[warp][d] + lrank;
        key_out[pos] = k;
        val_out[pos] = v;
    }
}


} // namespace b200
