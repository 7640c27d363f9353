// tilecount.cu — count(*) on grids that do not fit in shared memory WITHOUT one L2 atomic per row.
//
// Why: the direct kernel (fast.cu) issues one RED per row, and the L2 retires only ~98 sector requests per clock chip-wide,
// loads included (profiles/r01_ncu_fast_*.txt: 2.5e8 REDs + 6.25e7 load sectors in 3.17e6 cycles = 98.7/clk).  That caps the
// 2-D 1024^2 count at ~1.5e11 rows/s = 18 % of the HBM stream rate.  Shared-memory atomics retire ~6 lanes/clk/SM
// (profiles/r01_microbench.txt), 9x more — but a 1027^2 grid is 4 MB even with 32-bit counters.
//
// Scheme (two kernels per batch of <= 2^28 rows, both on the caller's stream):
//   K1 k_tile_partition  every CTA takes tiles of 4096 rows: 128-bit coalesced loads, the bit-exact fp64 bin index, then a
//                        counting sort of the tile by GRID TILE (flat index >> 15, i.e. 32768 consecutive cells) done with warp
//                        ballots + warp-private counters in shared memory, and one coalesced append of each tile-segment to
//                        that grid tile's bucket in global memory as 16-bit local indices (2 B/row).
//   K2 k_tile_count      CTA (tile, slice) zeroes a private 32768-cell u32 histogram in shared memory (128 KB), streams its
//                        slice of the bucket with 128-bit loads, ATOMS.POPC.INC per entry, then flushes the non-zero cells with
//                        one RED.ADD.64 each into the int64 grid.
// L2 requests per row drop from 1.25 to ~0.4, HBM traffic rises from 8 to 12 B/row; exact integer counts, same grid layout.
// Buckets are provisioned for 4x the uniform share; a segment that would overflow its bucket is applied with direct REDs
// instead (degenerate distributions stay correct, just slower).
#include <algorithm>

#include "binby.cuh"
#include "device_utils.cuh"

namespace b200 {

struct TileParams {
    const void *x[3];
    double vmin[3], scale[3], bins_d[3];
    unsigned bins[3];
    unsigned stride[3];
    long long row0, nrows; // batch
    unsigned cells;
    int nparts, pbits;
    unsigned short *buckets; // nparts * cap entries
    unsigned long long cap;
    unsigned *cursors; // nparts: entries reserved so far (may run past cap)
    unsigned *limits;  // nparts: start of the first segment that straddled cap (0xFFFFFFFF if none) — valid entries end there
    unsigned long long *grid;
};

namespace {

constexpr int kThreads = 256;
constexpr int kRounds = 16;            // rows per thread per tile
constexpr int kTile = kThreads * kRounds;
constexpr int kTileShift = 15;         // 32768 cells per grid tile
constexpr int kTileCells = 1 << kTileShift;
constexpr int kMaxParts = 128;
constexpr int kSlice = 1 << 21;        // bucket entries per K2 CTA

__device__ __forceinline__ unsigned bin_index(double v, double vmin, double scale, double bins_d, unsigned bins) {
    // identical to fast.cu: one saturating round-down conversion + clamp, NaN tested on `scaled` (src/binners.cpp:13-57)
    const double scaled = __dmul_rn(__dsub_rn(v, vmin), scale);
    const int i = __double2int_rd(__dmul_rn(scaled, bins_d));
    const unsigned idx = (unsigned)(min(max(i, -1), (int)bins) + 2);
    return scaled != scaled ? 0u : idx;
}

template <typename T>
__device__ __forceinline__ void load4(const void *p, long long i, double out[4]);
template <>
__device__ __forceinline__ void load4<float>(const void *p, long long i, double out[4]) {
    uint4 a = __ldcs(reinterpret_cast<const uint4 *>(static_cast<const float *>(p) + i));
    out[0] = (double)__uint_as_float(a.x), out[1] = (double)__uint_as_float(a.y), out[2] = (double)__uint_as_float(a.z), out[3] = (double)__uint_as_float(a.w);
}
template <>
__device__ __forceinline__ void load4<double>(const void *p, long long i, double out[4]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(static_cast<const double *>(p) + i);
    uint4 a = __ldcs(q), b = __ldcs(q + 1);
    out[0] = __longlong_as_double(((long long)a.y << 32) | a.x), out[1] = __longlong_as_double(((long long)a.w << 32) | a.z);
    out[2] = __longlong_as_double(((long long)b.y << 32) | b.x), out[3] = __longlong_as_double(((long long)b.w << 32) | b.z);
}

template <typename T, int ND>
__global__ void __launch_bounds__(kThreads) k_tile_partition(const __grid_constant__ TileParams p) {
    __shared__ unsigned short stage[kTile];
    __shared__ unsigned char stage_p[kTile];
    __shared__ unsigned wcnt[kThreads / 32][kMaxParts];
    __shared__ unsigned total[kMaxParts], segstart[kMaxParts + 1], gbase[kMaxParts];
    __shared__ unsigned char ovf[kMaxParts];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < (kThreads / 32) * kMaxParts; i += kThreads)
        (&wcnt[0][0])[i] = 0;
    __syncthreads();

    const long long ntiles = (p.nrows + kTile - 1) / kTile;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long tbase = p.row0 + tile * kTile;
        const long long tend = min(p.row0 + p.nrows, tbase + kTile);
        unsigned packed[kRounds]; // local(15) | part(7) << 15 | slot-in-warp(9) << 22; 0xFFFFFFFF = no row
        // ---- 1. load + index + warp-level multisplit ---------------------------------------------------------------
#pragma unroll
        for (int q = 0; q < kRounds / 4; q++) {
            const long long r0 = tbase + q * (kThreads * 4) + threadIdx.x * 4;
            double c[ND][4];
            if (r0 + 4 <= tend) {
#pragma unroll
                for (int d = 0; d < ND; d++)
                    load4<T>(p.x[d], r0, c[d]);
            } else {
#pragma unroll
                for (int d = 0; d < ND; d++)
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        c[d][j] = r0 + j < tend ? (double)__ldcs(static_cast<const T *>(p.x[d]) + r0 + j) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool valid = r0 + j < tend;
                unsigned idx = 0;
#pragma unroll
                for (int d = 0; d < ND; d++)
                    idx += bin_index(c[d][j], p.vmin[d], p.scale[d], p.bins_d[d], p.bins[d]) * p.stride[d];
                const unsigned part = idx >> kTileShift;
                // lanes with the same grid tile: AND of per-bit ballots
                unsigned peers = __ballot_sync(0xffffffffu, valid);
                for (int b = 0; b < p.pbits; b++) {
                    const unsigned bal = __ballot_sync(0xffffffffu, (part >> b) & 1u);
                    peers &= ((part >> b) & 1u) ? bal : ~bal;
                }
                unsigned old = 0;
                const int leader = __ffs(peers) - 1;
                if (valid && lane == leader) {
                    old = wcnt[warp][part];
                    wcnt[warp][part] = old + __popc(peers);
                }
                old = __shfl_sync(0xffffffffu, old, valid ? leader : 0);
                const unsigned slot = old + __popc(peers & lt_mask);
                packed[q * 4 + j] = valid ? ((idx & (kTileCells - 1)) | (part << kTileShift) | (slot << 22)) : 0xFFFFFFFFu;
            }
        }
        __syncthreads();
        // ---- 2. per-tile offsets: warp bases, segment starts, global reservations -----------------------------------
        if (threadIdx.x < p.nparts) {
            unsigned acc = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; w++) {
                const unsigned t = wcnt[w][threadIdx.x];
                wcnt[w][threadIdx.x] = acc;
                acc += t;
            }
            total[threadIdx.x] = acc;
            unsigned g = 0;
            unsigned char o = 0;
            if (acc) {
                g = atomicAdd(p.cursors + threadIdx.x, acc);
                o = (unsigned long long)g + acc > p.cap;
                if (o && g <= p.cap)
                    p.limits[threadIdx.x] = g; // exactly one segment per bucket straddles cap; everything before it is dense
            }
            gbase[threadIdx.x] = g;
            ovf[threadIdx.x] = o;
        }
        __syncthreads();
        if (warp == 0) { // exclusive scan of total[0..nparts) -> segstart (4 entries per lane)
            unsigned v[4], s = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = lane * 4 + k;
                v[k] = i < p.nparts ? total[i] : 0;
                s += v[k];
            }
            unsigned incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o)
                    incl += n;
            }
            unsigned run = incl - s;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = lane * 4 + k;
                if (i <= p.nparts)
                    segstart[i] = run;
                run += v[k];
            }
        }
        __syncthreads();
        // ---- 3. scatter the tile into shared memory, sorted by grid tile -----------------------------------------------
#pragma unroll
        for (int r = 0; r < kRounds; r++) {
            const unsigned pk = packed[r];
            if (pk != 0xFFFFFFFFu) {
                const unsigned part = (pk >> kTileShift) & 127u;
                const unsigned pos = segstart[part] + wcnt[warp][part] + (pk >> 22);
                stage[pos] = (unsigned short)(pk & (kTileCells - 1));
                stage_p[pos] = (unsigned char)part;
            }
        }
        __syncthreads();
        // ---- 4. coalesced append of every segment to its bucket -----------------------------------------------------------
        const int nvalid = (int)(tend - tbase);
        for (int i = threadIdx.x; i < nvalid; i += kThreads) {
            const unsigned part = stage_p[i];
            const unsigned local = stage[i];
            if (!ovf[part])
                p.buckets[(unsigned long long)part * p.cap + gbase[part] + (i - segstart[part])] = (unsigned short)local;
            else
                atomicAdd(p.grid + ((unsigned long long)part << kTileShift) + local, 1ull);
        }
        for (int i = threadIdx.x; i < (kThreads / 32) * kMaxParts; i += kThreads)
            (&wcnt[0][0])[i] = 0;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_tile_count(const __grid_constant__ TileParams p, int nslices) {
    extern __shared__ __align__(16) unsigned hist[];
    const int part = blockIdx.x / nslices, slice = blockIdx.x % nslices;
    unsigned long long n = p.cursors[part];
    if (n > p.cap) // the excess was applied with direct REDs by k_tile_partition; valid entries end at the straddling segment
        n = min((unsigned long long)p.limits[part], p.cap);
    const unsigned long long begin = (unsigned long long)slice * kSlice;
    if (begin >= n)
        return;
    const unsigned long long end = min(n, begin + (unsigned long long)kSlice);
    for (int i = threadIdx.x; i < kTileCells / 4; i += blockDim.x)
        reinterpret_cast<uint4 *>(hist)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const unsigned short *src = p.buckets + (unsigned long long)part * p.cap;
    const unsigned long long nvec = (end - begin) / 8;
    const uint4 *v = reinterpret_cast<const uint4 *>(src + begin);
    for (unsigned long long i = threadIdx.x; i < nvec; i += blockDim.x) {
        const uint4 a = __ldcs(v + i);
        atomicAdd(hist + (a.x & 0xffffu), 1u);
        atomicAdd(hist + (a.x >> 16), 1u);
        atomicAdd(hist + (a.y & 0xffffu), 1u);
        atomicAdd(hist + (a.y >> 16), 1u);
        atomicAdd(hist + (a.z & 0xffffu), 1u);
        atomicAdd(hist + (a.z >> 16), 1u);
        atomicAdd(hist + (a.w & 0xffffu), 1u);
        atomicAdd(hist + (a.w >> 16), 1u);
    }
    for (unsigned long long i = begin + nvec * 8 + threadIdx.x; i < end; i += blockDim.x)
        atomicAdd(hist + src[i], 1u);
    __syncthreads();
    const unsigned long long cell0 = (unsigned long long)part << kTileShift;
    for (int i = threadIdx.x; i < kTileCells; i += blockDim.x) {
        const unsigned c = hist[i];
        if (c && cell0 + i < p.cells)
            atomicAdd(p.grid + cell0 + i, (unsigned long long)c);
    }
}

template <typename T>
int launch_partition(int nd, int blocks, cudaStream_t st, const TileParams &p) {
    switch (nd) {
    case 1: k_tile_partition<T, 1><<<blocks, kThreads, 0, st>>>(p); break;
    case 2: k_tile_partition<T, 2><<<blocks, kThreads, 0, st>>>(p); break;
    default: k_tile_partition<T, 3><<<blocks, kThreads, 0, st>>>(p); break;
    }
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

} // namespace

// Scratch (buckets + cursors) lives in the slot; grown on demand.
int try_launch_tilecount(b200_ctx *ctx, Slot *slot, const BinParams &bp, bool vec, bool *taken) {
    *taken = false;
    static const bool disabled = getenv("B200_DISABLE_TILECOUNT") && atoi(getenv("B200_DISABLE_TILECOUNT")) != 0;
    if (disabled || !vec || bp.nb < 1 || bp.nb > 3 || bp.na != 1 || bp.nrows < (1ll << 22))
        return B200_OK;
    const DevAgg &a = bp.a[0];
    if (a.op != B200_AGG_COUNT || a.data || a.mask)
        return B200_OK;
    const unsigned long long cells = bp.cells;
    const int nparts = (int)((cells + kTileCells - 1) >> kTileShift);
    if (cells * 4 <= 96 * 1024 || nparts > kMaxParts) // small grids: shared-memory privatisation; huge grids: direct REDs
        return B200_OK;
    const int t = bp.b[0].dtype;
    if (t != B200_F32 && t != B200_F64)
        return B200_OK;
    TileParams p;
    memset(&p, 0, sizeof p);
    for (int i = 0; i < bp.nb; i++) {
        const DevBinner &b = bp.b[i];
        if (b.kind != B200_BINNER_SCALAR || b.dtype != t || b.byteswap || b.mask || b.bins < 1 || b.bins >= (1ull << 30))
            return B200_OK;
        p.x[i] = b.data;
        p.vmin[i] = b.vmin;
        p.scale[i] = b.scale;
        p.bins_d[i] = b.bins_d;
        p.bins[i] = (unsigned)b.bins;
        p.stride[i] = (unsigned)b.stride;
    }
    p.cells = (unsigned)cells;
    p.nparts = nparts;
    p.pbits = 0;
    while ((1 << p.pbits) < nparts)
        p.pbits++;
    p.grid = static_cast<unsigned long long *>(a.grid);

    const long long batch = std::min<long long>(bp.nrows, 1ll << 28);
    const unsigned long long cap = (((unsigned long long)batch * 4 / nparts + 65536) + 7) / 8 * 8;
    const size_t need = (size_t)nparts * cap * 2 + 4096;
    if (slot->scratch_cap < need) {
        if (slot->scratch) {
            B200_CUDA(cudaStreamSynchronize(slot->stream));
            B200_CUDA(cudaFree(slot->scratch));
            slot->scratch = nullptr;
            slot->scratch_cap = 0;
        }
        cudaError_t e = cudaMalloc(&slot->scratch, need);
        if (e != cudaSuccess) { // not enough memory for the buckets: fall back to the direct RED kernel
            cudaGetLastError();
            return B200_OK;
        }
        slot->scratch_cap = need;
    }
    p.cursors = static_cast<unsigned *>(slot->scratch);
    p.limits = p.cursors + 512;
    p.buckets = reinterpret_cast<unsigned short *>(static_cast<char *>(slot->scratch) + 4096);
    p.cap = cap;
    cudaStream_t st = slot->stream;
    static bool attr_set = false;
    if (!attr_set) {
        B200_CUDA(cudaFuncSetAttribute(k_tile_count, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileCells * 4));
        attr_set = true;
    }
    const int nslices = (int)((cap + kSlice - 1) / kSlice);
    for (long long r0 = 0; r0 < bp.nrows; r0 += batch) {
        p.row0 = r0;
        p.nrows = std::min<long long>(batch, bp.nrows - r0);
        B200_CUDA(cudaMemsetAsync(p.cursors, 0, 2048, st));
        B200_CUDA(cudaMemsetAsync(p.limits, 0xff, 2048, st));
        const long long ntiles = (p.nrows + kTile - 1) / kTile;
        const int blocks = (int)std::min<long long>(ntiles, (long long)ctx->sm_count * 6);
        if (t == B200_F32)
            B200_CHECK(launch_partition<float>(bp.nb, blocks, st, p));
        else
            B200_CHECK(launch_partition<double>(bp.nb, blocks, st, p));
        k_tile_count<<<nparts * nslices, 1024, kTileCells * 4, st>>>(p, nslices);
        B200_CUDA(cudaGetLastError());
    }
    *taken = true;
    return B200_OK;
}

} // namespace b200
