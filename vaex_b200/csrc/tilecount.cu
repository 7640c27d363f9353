// tilecount.cu — count(*) on grids that do not fit in shared memory WITHOUT one L2 atomic per row.
//
// Why: the direct kernel (fast.cu) issues one RED per row, and the L2 retires only ~98 sector requests per clock chip-wide,
// loads included (profiles/r01_ncu_fast_*.txt: 2.5e8 REDs + 6.25e7 load sectors in 3.17e6 cycles = 98.7/clk).  That caps the
// 2-D 1024^2 count at ~1.5e11 rows/s = 18 % of the HBM stream rate.  Shared-memory atomics retire ~6 lanes/clk/SM
// (profiles/r01_microbench.txt), 9x more — but a 1027^2 grid is 4 MB even with 32-bit counters.
//
// Scheme (two kernels per batch of <= 2^28 rows, both on the caller's stream); the grid is cut into <= 32 (else 64 / 128)
// GRID TILES of <= 49152 consecutive cells (1027^2 -> 32 tiles of 32,961 cells):
//   K1 k_tile_partition  every WARP takes 512-row tiles: columns staged by TMA (cp.async.bulk + mbarrier, double buffered),
//                        the bit-exact fp64 bin index, a counting sort of the tile by grid tile (one shared-memory atomic per
//                        row for the rank, a warp scan for the segment starts), and a coalesced append of every segment to
//                        that grid tile's bucket in global memory as 16-bit local indices (2 B/row), space reserved in
//                        512-entry chunks.
//   K2 k_tile_count      CTA (tile, slice) zeroes a private u32 histogram of the tile in shared memory (<= 192 KB), streams its
//                        slice of the bucket with 128-bit loads, ATOMS.POPC.INC per entry, then flushes the non-zero cells with
//                        one RED.ADD.64 each into the int64 grid.
// L2 requests per row drop from 1.25 to ~0.4, HBM traffic rises from 8 to ~13 B/row; exact integer counts, same grid layout.
// Buckets are provisioned for 4x the uniform share; a (warp, tile) whose chunk request would overflow the bucket applies its
// rows with direct REDs instead (degenerate distributions stay exact, just slower).
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
    unsigned tile_cells;          // cells per grid tile (<= 49152, multiple of 8): part = idx / tile_cells, local = idx % tile_cells < 2^16
    unsigned magic, magic_shift;  // part = (idx * magic) >> magic_shift with 2^31 <= magic < 2^32: one IMAD.WIDE + one shift, exact for idx < 2^22
    int nparts;
    unsigned short *buckets; // nparts * cap entries
    unsigned long long cap;
    unsigned *cursors; // nparts: entries reserved so far (may run past cap)
    unsigned long long *grid;
};

namespace {

constexpr int kRounds = 16;            // rows per lane per warp tile
constexpr int kWarpTile = 32 * kRounds; // 512 rows sorted per warp at a time
constexpr unsigned kMaxTileCells = 49152; // 192 KB of u32 counters in k_tile_count
constexpr int kMaxParts = 128;
constexpr int kSlice = 1 << 20;        // bucket entries per K2 CTA
constexpr unsigned kChunk = 512;       // bucket entries a warp reserves at a time (>= kWarpTile: any segment fits)
constexpr unsigned kNone = 0xFFFFFFFFu, kOver = 0xFFFFFFFEu;
constexpr unsigned long long kNone64 = ~0ull;
constexpr unsigned short kPad = 0xFFFFu; // padding entry (>= 32768: not a cell)


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

// one warp tile (512 rows): index + rank every row.  FULL tiles carry no per-row validity tests.
template <typename T, int ND, bool FULL>
__device__ __forceinline__ void tile_rank(const TileParams &p, long long tbase, long long tend, int lane, unsigned *seg, unsigned packed[kRounds]) {
#pragma unroll
    for (int q = 0; q < kRounds / 4; q++) {
        const long long r0 = tbase + q * 128 + lane * 4;
        double c[ND][4];
        if (FULL || r0 + 4 <= tend) {
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
            unsigned idx = 0;
#pragma unroll
            for (int d = 0; d < ND; d++)
                idx += bin_index(c[d][j], p.vmin[d], p.scale[d], p.bins_d[d], p.bins[d]) * p.stride[d];
            // rank inside (warp, grid tile): one shared-memory atomic with return per row.  MATCH.ANY + SHFL ranking kept the
            // ADU pipe 70 % busy and 7-bit ballot ranking the ALU pipe 60 % busy (profiles/r01_ncu_tilecount_*.txt)
            if (FULL || r0 + j < tend) {
                const unsigned part = (unsigned)(((unsigned long long)idx * p.magic) >> p.magic_shift);
                const unsigned slot = atomicAdd(seg + part, 1u);
                packed[q * 4 + j] = (idx - part * p.tile_cells) | (part << 16) | (slot << 23); // local(16) | part(7) | slot(9)
            } else {
                packed[q * 4 + j] = 0xFFFFFFFFu;
            }
        }
    }
}

// ---- TMA staging (cp.async.bulk global -> shared, completion on an mbarrier): each warp keeps the NEXT tile's columns in
// flight while it sorts the current one.  No registers, no LSU issue slots, and the copy engine sees 2 KB requests.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile("{\n\t"
                 ".reg .pred P1;\n\t"
                 "WAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                 "@P1 bra DONE;\n\t"
                 "bra WAIT_LOOP;\n\t"
                 "DONE:\n\t"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ void lds4(const T *buf, int i, double out[4]);
template <>
__device__ __forceinline__ void lds4<float>(const float *buf, int i, double out[4]) {
    const float4 a = *reinterpret_cast<const float4 *>(buf + i);
    out[0] = (double)a.x, out[1] = (double)a.y, out[2] = (double)a.z, out[3] = (double)a.w;
}
template <>
__device__ __forceinline__ void lds4<double>(const double *buf, int i, double out[4]) {
    const double2 a = *reinterpret_cast<const double2 *>(buf + i), b = *reinterpret_cast<const double2 *>(buf + i + 2);
    out[0] = a.x, out[1] = a.y, out[2] = b.x, out[3] = b.y;
}

// full tile whose columns already sit in shared memory (TMA staged)
template <typename T, int ND>
__device__ __forceinline__ void tile_rank_staged(const TileParams &p, const T *buf, int lane, unsigned *seg, unsigned packed[kRounds]) {
#pragma unroll
    for (int q = 0; q < kRounds / 4; q++) {
        double c[ND][4];
#pragma unroll
        for (int d = 0; d < ND; d++)
            lds4<T>(buf + d * kWarpTile, q * 128 + lane * 4, c[d]);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            unsigned idx = 0;
#pragma unroll
            for (int d = 0; d < ND; d++)
                idx += bin_index(c[d][j], p.vmin[d], p.scale[d], p.bins_d[d], p.bins[d]) * p.stride[d];
            const unsigned part = (unsigned)(((unsigned long long)idx * p.magic) >> p.magic_shift);
            const unsigned slot = atomicAdd(seg + part, 1u);
            packed[q * 4 + j] = (idx - part * p.tile_cells) | (part << 16) | (slot << 23);
        }
    }
}

// shared memory per warp: [TMA: 2 stages x ND columns x 512 T] [no TMA: 2 KB stage] [seg | base | used: 128 u32 each] [dst: 128 u64]
template <typename T, int ND, bool TMA>
__host__ __device__ constexpr size_t warp_smem_bytes() {
    return (TMA ? 2 * ND * kWarpTile * sizeof(T) : kWarpTile * sizeof(unsigned)) + 3 * kMaxParts * sizeof(unsigned) + kMaxParts * sizeof(unsigned long long);
}

template <typename T, int ND, bool TMA, int WARPS, int PPL>
__global__ void __launch_bounds__(WARPS * 32, TMA ? 2 : 4) k_tile_partition(const __grid_constant__ TileParams p) {
    // Every WARP is autonomous: it sorts its own 512-row tile by grid tile and appends the segments itself, so the kernel has
    // no block-level barrier (only __syncwarp).  Bucket space is handed out in CHUNKS of kChunk entries that a warp owns
    // exclusively: one global atomic per kChunk entries per (warp, grid tile) instead of one per segment — per-segment
    // reservations on 33 shared cursors serialised in the L2 at ~11 ns each (profiles/r01_ncu_tile2_*.txt).
    // A chunk's unused tail is padded with kPad entries, which k_tile_count skips.
    // With TMA staging the sort stage ALIASES the column buffer that was just consumed (it is only refilled by the copy engine
    // one iteration later, after a proxy fence), which is what lets 11 warps x 2 CTAs fit next to the double-buffered columns.
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ __align__(8) unsigned long long bars[WARPS][2];
    constexpr int kThreads = WARPS * 32;
    // PAIR: every segment is padded to an even length so that the copy-out moves TWO 16-bit entries per lane per step with one
    // aligned 32-bit store (needs room for 32*PPL pad entries in the stage: every staged variant except 1-D fp32)
    constexpr bool PAIR = TMA && ND * sizeof(T) >= 8;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr size_t kPerWarp = warp_smem_bytes<T, ND, TMA>();
    unsigned char *mine = dyn_smem + (size_t)warp * kPerWarp;
    T *const mybuf = reinterpret_cast<T *>(mine);                         // TMA: [2][ND][512]
    unsigned char *tables = mine + (TMA ? 2 * ND * kWarpTile * sizeof(T) : kWarpTile * sizeof(unsigned));
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(tables); // bucket entry index of stage[0] per part (kNone64: direct REDs)
    unsigned *seg = reinterpret_cast<unsigned *>(tables + kMaxParts * 8);  // count per part -> (after the scan) segment start in the stage
    unsigned *cbase = seg + kMaxParts;                                     // current chunk of (warp, part): kNone none yet, kOver = bucket full
    unsigned *cused = cbase + kMaxParts;                                   // entries used in the current chunk
    unsigned *stage = reinterpret_cast<unsigned *>(mine);                 // no TMA: dedicated; TMA: re-pointed per iteration
    if (TMA && lane == 0) {
        mbar_init(&bars[warp][0], 1);
        mbar_init(&bars[warp][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned short *const buckets = p.buckets;
    const unsigned long long cap = p.cap;
    const int nparts = p.nparts;
    for (int i = lane; i < kMaxParts; i += 32) {
        seg[i] = 0;
        cbase[i] = kNone;
        cused[i] = 0;
    }
    __syncwarp();

    const long long ntiles = (p.nrows + kWarpTile - 1) / kWarpTile;
    const long long nfull = p.nrows / kWarpTile; // tiles [0, nfull) are complete
    const long long wglobal = (long long)blockIdx.x * (kThreads / 32) + warp, wtotal = (long long)gridDim.x * (kThreads / 32);
    auto issue = [&](long long t, int st) { // lane 0: start the bulk copies of tile t's columns into stage st
        mbar_expect_tx(&bars[warp][st], (unsigned)(ND * kWarpTile * sizeof(T)));
#pragma unroll
        for (int d = 0; d < ND; d++)
            tma_load_1d(mybuf + (st * ND + d) * kWarpTile, static_cast<const T *>(p.x[d]) + p.row0 + t * kWarpTile, (unsigned)(kWarpTile * sizeof(T)), &bars[warp][st]);
    };
    int st = 0;
    unsigned phase0 = 0, phase1 = 0;
    if (TMA && lane == 0 && wglobal < nfull)
        issue(wglobal, 0);
    for (long long tile = wglobal; tile < ntiles; tile += wtotal) {
        const long long tbase = p.row0 + tile * kWarpTile;
        const long long tend = min(p.row0 + p.nrows, tbase + kWarpTile);
        const int nvalid = (int)(tend - tbase);
        unsigned packed[kRounds]; // idx(22) | slot(10) << 22; 0xFFFFFFFF = no row
        // ---- 1. (staged) load + bit-exact index + rank ----------------------------------------------------------------------
        if (TMA) {
            if (lane == 0 && tile + wtotal < nfull)
                issue(tile + wtotal, st ^ 1); // prefetch the next tile while this one is processed
            if (tile < nfull) {
                mbar_wait(&bars[warp][st], st ? phase1 : phase0);
                if (st)
                    phase1 ^= 1;
                else
                    phase0 ^= 1;
                tile_rank_staged<T, ND>(p, mybuf + st * ND * kWarpTile, lane, seg, packed);
                __syncwarp();
            } else {
                tile_rank<T, ND, false>(p, tbase, tend, lane, seg, packed);
            }
            stage = reinterpret_cast<unsigned *>(mybuf + st * ND * kWarpTile); // the consumed column buffer becomes the sort stage
            st ^= 1;
        } else if (nvalid == kWarpTile) {
            tile_rank<T, ND, true>(p, tbase, tend, lane, seg, packed);
        } else {
            tile_rank<T, ND, false>(p, tbase, tend, lane, seg, packed);
        }
        __syncwarp();
        // ---- 2. exclusive scan of the per-part counts (PPL parts per lane: 1 when the grid has <= 32 tiles); place each segment in the warp's current chunk -----
        unsigned nstage = 0; // entries in the stage incl. fillers
        {
            unsigned v[PPL], ve[PPL], s = 0;
#pragma unroll
            for (int k = 0; k < PPL; k++) {
                const int i = lane * PPL + k;
                v[k] = i < nparts ? seg[i] : 0;
                ve[k] = PAIR ? (v[k] + 1u) & ~1u : v[k]; // length the segment occupies in the stage and in its bucket
                s += ve[k];
            }
            unsigned incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o)
                    incl += n;
            }
            unsigned run = incl - s;
            nstage = __shfl_sync(0xffffffffu, incl, 31);
            // which parts need a fresh chunk?  (rare: once per kChunk entries per part)
            bool fresh[PPL];
            unsigned any = 0;
#pragma unroll
            for (int k = 0; k < PPL; k++) {
                const int i = lane * PPL + k;
                fresh[k] = i < nparts && v[k] && cbase[i] != kOver && (cbase[i] == kNone || cused[i] + ve[k] > kChunk);
                any |= fresh[k];
            }
            if (__any_sync(0xffffffffu, any)) {
#pragma unroll
                for (int k = 0; k < PPL; k++) {
                    unsigned need = __ballot_sync(0xffffffffu, fresh[k]);
                    while (need) {
                        const int src = __ffs(need) - 1;
                        need &= need - 1;
                        const int part = src * PPL + k;
                        const unsigned ob = cbase[part], ou = cused[part];
                        if (ob != kNone) // pad the tail of the old chunk
                            for (unsigned e = ou + lane; e < kChunk; e += 32)
                                buckets[(unsigned long long)part * cap + ob + e] = kPad;
                        unsigned nb = 0;
                        if (lane == 0) {
                            nb = atomicAdd(p.cursors + part, (unsigned)kChunk);
                            if ((unsigned long long)nb + kChunk > cap)
                                nb = kOver; // bucket exhausted: this (warp, part) applies its rows with direct REDs from now on
                        }
                        nb = __shfl_sync(0xffffffffu, nb, 0);
                        __syncwarp();
                        if (lane == 0) {
                            cbase[part] = nb;
                            cused[part] = 0;
                        }
                        __syncwarp();
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < PPL; k++) {
                const int i = lane * PPL + k;
                if (i < nparts) {
                    seg[i] = run;
                    if (PAIR && (v[k] & 1u))
                        stage[run + v[k]] = kPad | ((unsigned)i << 16); // the odd segment's filler entry
                    unsigned long long d = kNone64;
                    if (v[k] && cbase[i] != kOver) {
                        d = (unsigned long long)i * cap + cbase[i] + cused[i] - run; // entry index of stage[0] if it belonged to part i
                        cused[i] += ve[k];
                    }
                    dst[i] = d;
                }
                run += ve[k];
            }
        }
        __syncwarp();
        // ---- 3. scatter the tile into the warp's stage, sorted by grid tile ----------------------------------------------
        if (nvalid == kWarpTile) {
#pragma unroll
            for (int r = 0; r < kRounds; r++) {
                const unsigned pk = packed[r];
                stage[seg[(pk >> 16) & 127u] + (pk >> 23)] = pk & 0x7FFFFFu;
            }
        } else {
#pragma unroll
            for (int r = 0; r < kRounds; r++) {
                const unsigned pk = packed[r];
                if (pk != 0xFFFFFFFFu)
                    stage[seg[(pk >> 16) & 127u] + (pk >> 23)] = pk & 0x7FFFFFu;
            }
        }
        __syncwarp();
        // ---- 4. append every segment to its bucket (consecutive lanes -> consecutive 16-bit entries of one segment) -------
        if (PAIR) {
            for (unsigned i2 = lane; i2 < nstage / 2; i2 += 32) {
                const uint2 e = *reinterpret_cast<const uint2 *>(stage + 2 * i2); // never straddles two segments (even starts)
                const unsigned part = e.x >> 16;
                const unsigned long long d = dst[part];
                if (d != kNone64) {
                    *reinterpret_cast<unsigned *>(buckets + d + 2 * i2) = (e.x & 0xffffu) | (e.y << 16);
                } else {
                    atomicAdd(p.grid + (unsigned long long)part * p.tile_cells + (e.x & 0xffffu), 1ull);
                    if ((e.y & 0xffffu) != kPad)
                        atomicAdd(p.grid + (unsigned long long)part * p.tile_cells + (e.y & 0xffffu), 1ull);
                }
            }
        } else {
            for (int i = lane; i < nvalid; i += 32) {
                const unsigned e = stage[i];
                const unsigned part = e >> 16;
                const unsigned long long d = dst[part];
                if (d != kNone64)
                    buckets[d + i] = (unsigned short)e;
                else
                    atomicAdd(p.grid + (unsigned long long)part * p.tile_cells + (e & 0xffffu), 1ull);
            }
        }
        __syncwarp();
        for (int i = lane; i < kMaxParts; i += 32)
            seg[i] = 0;
        __syncwarp();
        if (TMA)
            fence_proxy_async(); // our generic-proxy reads/writes of this buffer are done before the copy engine refills it
    }
    // pad the open chunks so that every reserved chunk is completely written
    for (int part = 0; part < nparts; part++) {
        const unsigned ob = cbase[part], ou = cused[part];
        if (ob != kNone && ob != kOver)
            for (unsigned e = ou + lane; e < kChunk; e += 32)
                buckets[(unsigned long long)part * cap + ob + e] = kPad;
    }
}


__global__ void __launch_bounds__(1024) k_tile_count(const __grid_constant__ TileParams p, int nslices) {
    extern __shared__ __align__(16) unsigned hist[];
    const int part = blockIdx.x / nslices, slice = blockIdx.x % nslices;
    unsigned long long n = p.cursors[part];
    if (n > p.cap) // chunks past cap were refused (those rows were applied with direct REDs); cap is a multiple of kChunk
        n = p.cap;
    const unsigned long long begin = (unsigned long long)slice * kSlice;
    if (begin >= n)
        return;
    const unsigned long long end = min(n, begin + (unsigned long long)kSlice);
    const unsigned tc = p.tile_cells;
    for (int i = threadIdx.x; i < (int)(tc / 4); i += blockDim.x)
        reinterpret_cast<uint4 *>(hist)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const unsigned short *src = p.buckets + (unsigned long long)part * p.cap;
    const unsigned long long nvec = (end - begin) / 8;
    const uint4 *v = reinterpret_cast<const uint4 *>(src + begin);
    auto apply = [&](const uint4 a) {
        const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned lo = w[k] & 0xffffu, hi = w[k] >> 16;
            if (lo < tc) // skip chunk padding (0xFFFF >= tile_cells)
                atomicAdd(hist + lo, 1u);
            if (hi < tc)
                atomicAdd(hist + hi, 1u);
        }
    };
    // four independent 128-bit loads in flight per thread (one load per step left the SM waiting ~1500 cycles per step)
    unsigned long long i = threadIdx.x;
    const unsigned long long stride = blockDim.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        const uint4 a0 = __ldcs(v + i), a1 = __ldcs(v + i + stride), a2 = __ldcs(v + i + 2 * stride), a3 = __ldcs(v + i + 3 * stride);
        apply(a0);
        apply(a1);
        apply(a2);
        apply(a3);
    }
    for (; i < nvec; i += stride)
        apply(__ldcs(v + i));
    for (unsigned long long i = begin + nvec * 8 + threadIdx.x; i < end; i += blockDim.x)
        if (src[i] < tc)
            atomicAdd(hist + src[i], 1u);
    __syncthreads();
    const unsigned long long cell0 = (unsigned long long)part * tc;
    for (int i = threadIdx.x; i < (int)tc; i += blockDim.x) {
        const unsigned c = hist[i];
        if (c && cell0 + i < p.cells)
            atomicAdd(p.grid + cell0 + i, (unsigned long long)c);
    }
}

template <typename T, int ND, int PPL>
int launch_partition_ppl(int sm_count, long long nrows, cudaStream_t st, const TileParams &p) {
    // TMA-staged variant whenever two CTAs of >= 4 warps fit in the 227 KB of an SM; as many warps per CTA as fit (<= 16)
    constexpr size_t per_warp = warp_smem_bytes<T, ND, true>();
    constexpr int fit = (int)((113 * 1024 - 256) / per_warp);
    constexpr int WARPS = fit > 16 ? 16 : fit;
    if constexpr (WARPS >= 4) {
        auto kern = k_tile_partition<T, ND, true, WARPS, PPL>;
        constexpr size_t dyn = per_warp * WARPS;
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        const long long ntiles = (nrows + kWarpTile * WARPS - 1) / (kWarpTile * WARPS);
        const int blocks = (int)std::min<long long>(ntiles, (long long)sm_count * 2);
        kern<<<blocks, WARPS * 32, dyn, st>>>(p);
    } else {
        constexpr int W = 8;
        auto kern = k_tile_partition<T, ND, false, W, PPL>;
        constexpr size_t dyn = warp_smem_bytes<T, ND, false>() * W;
        const long long ntiles = (nrows + kWarpTile * W - 1) / (kWarpTile * W);
        const int blocks = (int)std::min<long long>(ntiles, (long long)sm_count * 4);
        kern<<<blocks, W * 32, dyn, st>>>(p);
    }
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

template <typename T, int ND>
int launch_partition_nd(int sm_count, long long nrows, cudaStream_t st, const TileParams &p) {
    if (p.nparts <= 32)
        return launch_partition_ppl<T, ND, 1>(sm_count, nrows, st, p);
    if (p.nparts <= 64)
        return launch_partition_ppl<T, ND, 2>(sm_count, nrows, st, p);
    return launch_partition_ppl<T, ND, 4>(sm_count, nrows, st, p);
}

template <typename T>
int launch_partition(int nd, int sm_count, long long nrows, cudaStream_t st, const TileParams &p) {
    switch (nd) {
    case 1: return launch_partition_nd<T, 1>(sm_count, nrows, st, p);
    case 2: return launch_partition_nd<T, 2>(sm_count, nrows, st, p);
    default: return launch_partition_nd<T, 3>(sm_count, nrows, st, p);
    }
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
    if (cells * 4 <= 96 * 1024 || cells > (unsigned long long)kMaxTileCells * kMaxParts || cells >= (1ull << 22))
        return B200_OK; // small grids: shared-memory privatisation; huge grids: direct REDs
    // as few grid tiles as possible (32, else 64, else 128): the per-tile bookkeeping of k_tile_partition scales with tiles per lane
    unsigned tile_cells = 0;
    for (int np = 32; np <= kMaxParts; np *= 2) {
        const unsigned long long tcells = ((cells + np - 1) / np + 7) / 8 * 8;
        if (tcells <= kMaxTileCells) {
            tile_cells = (unsigned)tcells;
            break;
        }
    }
    if (!tile_cells)
        return B200_OK;
    const int nparts = (int)((cells + tile_cells - 1) / tile_cells);
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
    p.tile_cells = tile_cells;
    {
        // floor(n / d) == (n * M) >> k for every n < 2^22 when M = ceil(2^k / d) and n * (M*d - 2^k) < 2^k; with
        // k = 31 + ceil(log2 d): M < 2^32 and (M*d - 2^k) < d <= 2^16, so n * d < 2^38 <= 2^k holds for d >= 128
        int lg = 0;
        while ((1u << lg) < tile_cells)
            lg++;
        p.magic_shift = 31 + lg;
        p.magic = (unsigned)((((unsigned long long)1 << p.magic_shift) + tile_cells - 1) / tile_cells);
    }
    p.nparts = nparts;
    p.grid = static_cast<unsigned long long *>(a.grid);

    const long long batch = std::min<long long>(bp.nrows, 1ll << 28);
    const unsigned long long cap = (((unsigned long long)batch * 4 / nparts + 65536) + kChunk - 1) / kChunk * kChunk;
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
    p.buckets = reinterpret_cast<unsigned short *>(static_cast<char *>(slot->scratch) + 4096);
    p.cap = cap;
    cudaStream_t st = slot->stream;
    // per device (function attributes live in the context): set on every call, it costs microseconds
    B200_CUDA(cudaFuncSetAttribute(k_tile_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kMaxTileCells * 4)));
    const int nslices = (int)((cap + kSlice - 1) / kSlice);
    for (long long r0 = 0; r0 < bp.nrows; r0 += batch) {
        p.row0 = r0;
        p.nrows = std::min<long long>(batch, bp.nrows - r0);
        B200_CUDA(cudaMemsetAsync(p.cursors, 0, 2048, st));
        if (t == B200_F32)
            B200_CHECK(launch_partition<float>(bp.nb, ctx->sm_count, p.nrows, st, p));
        else
            B200_CHECK(launch_partition<double>(bp.nb, ctx->sm_count, p.nrows, st, p));
        k_tile_count<<<nparts * nslices, 1024, (size_t)tile_cells * 4, st>>>(p, nslices);
        B200_CUDA(cudaGetLastError());
    }
    *taken = true;
    return B200_OK;
}

} // namespace b200
