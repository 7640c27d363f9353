// ringcount.cu — count(*) on grids that do not fit in shared memory WITHOUT one L2 atomic per row (the headline path).
//
// Why two kernels: a direct kernel issues one RED per row, and the L2 retires only ~98 sector operations per clock chip-wide,
// loads included (profiles/r01_ncu_fast.txt), which caps the 2-D 1024^2 count at ~1.5e11 rows/s = 18 % of the HBM stream rate.
// Shared-memory atomics retire ~6 lanes/clk/SM, 9x more, but a 1027^2 grid is 4 MB even with 32-bit counters.  So the grid is
// cut into <= 32 (else 64) GRID TILES ("parts") of <= 49152 consecutive cells and the rows are partitioned by part first:
//
//   K1 k_ring_partition  every WARP is autonomous.  The keys of the NEXT 256-row group are loaded into registers (128-bit,
//                        evict-first loads) while the current group is placed; a TMA-staged variant (cp.async.bulk + mbarrier,
//                        double buffered, as in round 1) was 8 % slower: 8.5 warp instructions per 32 rows went into issuing
//                        the bulk copies and 4 KB of shared memory per warp into the buffers (profiles/r02_ncu_ring_v4.txt).
//                        Per row: the bit-exact fp64 bin index, part = idx / tile_cells (one IMAD.HI), ONE shared-memory atomic
//                        on the (warp, part) counter whose return value is the row's slot in that part's RING (96 entries), and
//                        one 16-bit store of the local cell index into the ring.  Every 256 rows the rings that hold a full
//                        LINE (64 entries = 128 bytes) append it to the part's current CHUNK in global memory — half a warp per
//                        line: one LDS.64 + one coalesced STG.64 per lane — and slide the <= 32 entries behind it to the front.
//                        There is no scan, no scatter pass and no copy-out pass: round 1's kernel spent 47 of its 77 warp
//                        instructions per 32 rows there (profiles/r01_ncu_final_k1_phases.txt).  (A first version flushed
//                        16-entry granules from the owner lanes: 7.5 of its 16.9 shared-memory/L1 wavefronts per 32 rows went
//                        into those divergent 16-byte accesses, profiles/r02_ncu_ring_v1.txt.)
//                        Chunks (512..2048 entries) come from ONE pool: a warp reserves 32 chunks with one global atomic and
//                        hands them to its parts with a ballot; the chunks of a (warp, part) form a linked list (next[]), its
//                        head and entry count go to head[] / len[].  No per-part bucket provisioning, no overflow fallback, and
//                        2 B/row of scratch whatever the distribution.
//   K2 k_ring_count      one persistent CTA per SM takes a contiguous share of the (part, warp) lists (balanced by entry count:
//                        every warp sees the same row distribution), keeps the part's u32 histogram in <= 192 KB of shared
//                        memory, walks the lists with 128-bit loads + one ATOMS per entry, and flushes the non-zero cells with one
//                        RED.ADD.64 each when the part changes (1-2 flushes per CTA).
//
// HBM traffic: 8 B/row of columns + 2 B/row written + 2 B/row read.  Exact integer counts, same grid layout, any distribution:
// a ring that fills up inside one 256-row group (> 32 rows of a warp in one part on top of a full leftover) sends the excess rows
// to direct REDs.
#include <math.h>

#include <algorithm>
#include <type_traits>

#include "binby.cuh"
#include "device_utils.cuh"

namespace b200 {

struct RingParams {
    const void *x[3];
    double vmin[3], scale[3], bins_d[3];
    unsigned bins[3];
    unsigned stride[3];
    unsigned stride_sum;          // sum(stride): bin_cell_m1 returns cell - 1
    float clamp_lo[3], clamp_hi[3]; // CLAMPED variant (fp32 keys): see bin_cell_m2_clamped
    long long row0, nrows;        // this batch
    unsigned cells;
    unsigned tile_cells;          // cells per part (<= 49151): part = idx / tile_cells, local = idx % tile_cells < 2^16
    unsigned neg_tile_cells;      // 2^32 - tile_cells: local = idx + part * neg_tile_cells in one IMAD
    unsigned magic;               // part == __umulhi(idx, magic) for every idx < cells (verified on the host)
    int nparts;
    int nparts_pad;               // 32 * PPL
    unsigned chunk_shift;         // chunk = 1 << chunk_shift entries (9..11)
    unsigned nchunks_cap;         // chunks in the pool
    unsigned nlists_w;            // warps of K1 (lists per part)
    unsigned short *pool;         // chunk c = pool + (c << chunk_shift)
    unsigned *next;               // per chunk: next chunk of the same (warp, part) list; kNone = last      (memset 0xFF)
    unsigned short *unused;       // per chunk: entries at the tail that were never written                 (memset 0)
    unsigned *head;               // [warp][nparts_pad]: first chunk of the list; kNone = empty               (memset 0xFF)
    unsigned *len;                // [warp][nparts_pad]: entries in the list (pads included)                  (memset 0)
    unsigned *ctl;                // [0] pool cursor in chunks                                                (memset 0)
    unsigned long long *grid;
};

namespace {

constexpr int kGroupRows = 256;   // rows between two ring flushes: 8 per lane
constexpr int kGran = 16;         // entries per lane and step in k_ring_count
#ifndef B200_RING_LINE
#define B200_RING_LINE 64
#endif
constexpr unsigned kLine = B200_RING_LINE; // entries a ring flush moves: 32 = 64 bytes (two sectors), 64 = one 128-byte line
constexpr int kLineLanes = kLine / 8;      // lanes that move one line with 16 bytes each
constexpr unsigned kSuper = 32;   // chunks a warp reserves from the pool at a time
constexpr unsigned kMaxTileCells = 49151; // 192 KB of u32 counters in k_ring_count, one of them the pad cell
constexpr int kMaxParts = 64;
constexpr unsigned kNone = 0xFFFFFFFFu;
constexpr int kCountThreads = 1024;

template <typename T>
__device__ __forceinline__ void ldg4(const void *p, long long i, T out[4]);
template <>
__device__ __forceinline__ void ldg4<float>(const void *p, long long i, float out[4]) {
    const uint4 a = __ldcs(reinterpret_cast<const uint4 *>(static_cast<const float *>(p) + i));
    out[0] = __uint_as_float(a.x), out[1] = __uint_as_float(a.y), out[2] = __uint_as_float(a.z), out[3] = __uint_as_float(a.w);
}
template <>
__device__ __forceinline__ void ldg4<double>(const void *p, long long i, double out[4]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(static_cast<const double *>(p) + i);
    const uint4 a = __ldcs(q), b = __ldcs(q + 1);
    out[0] = __longlong_as_double(((long long)a.y << 32) | a.x), out[1] = __longlong_as_double(((long long)a.w << 32) | a.z);
    out[2] = __longlong_as_double(((long long)b.y << 32) | b.x), out[3] = __longlong_as_double(((long long)b.w << 32) | b.z);
}

// CLAMPED index for fp32 keys: halves the load on the conversion pipe (XU: 16 lanes/clk/SM, the busiest pipe of this kernel at 52 %
// with F2F.F64.F32 + F2I.F64 per key, profiles/r02_ncu_ring_v3.txt).  The cell is a monotone function of the key (every step of
// the reference formula is monotone and correctly rounded).  The host finds by bisection over the fp32 values
//   clamp_lo = the largest fp32 whose cell is 1 (underflow), clamp_hi = the smallest fp32 whose cell is bins+2 (overflow)
// and checks floor(t(clamp_lo)) == -1 and floor(t(clamp_hi)) == bins with the reference formula.  Then for every non-NaN key
// v' = min(max(v, clamp_lo), clamp_hi) has the same cell as v and -1 <= floor(t(v')) <= bins, so no integer clamp is needed and
// the floor can be taken with one round-down addition of 1.5 * 2^52 (low word of the sum) instead of a conversion.
// Returns cell - 2; NaN -> -2.
__device__ __forceinline__ int bin_cell_m2_clamped(float v, float lo, float hi, double vmin, double scale, double bins_d) {
    const float vc = fminf(fmaxf(v, lo), hi); // NaN -> lo (replaced below)
    const double t = __dmul_rn(__dmul_rn(__dsub_rn((double)vc, vmin), scale), bins_d);
    const int i = __double2loint(__dadd_rd(t, 6755399441055744.0));
    return v != v ? -2 : i;
}

// one row: bit-exact index, part, ONE shared-memory atomic for the slot, one 16-bit store
template <typename T, int ND, int RING, int RSTRIDE, bool CLAMPED>
__device__ __forceinline__ void place_row(const RingParams &p, const T (&v)[ND], unsigned *cnt, unsigned short *ring, unsigned &idx, unsigned &slot_out,
                                          unsigned &worst) {
    unsigned id;
    if constexpr (CLAMPED) {
        id = 2 * p.stride_sum;
#pragma unroll
        for (int d = 0; d < ND; d++)
            id += (unsigned)bin_cell_m2_clamped(v[d], p.clamp_lo[d], p.clamp_hi[d], p.vmin[d], p.scale[d], p.bins_d[d]) * p.stride[d];
    } else {
        id = p.stride_sum;
#pragma unroll
        for (int d = 0; d < ND; d++)
            id += (unsigned)bin_cell_m1((double)v[d], p.vmin[d], p.scale[d], p.bins_d[d], p.bins[d]) * p.stride[d];
    }
    idx = id;
    const unsigned part = __umulhi(id, p.magic);
    const unsigned slot = atomicAdd(cnt + part, 1u);
    slot_out = slot;
    worst = max(worst, slot);
    if (slot < (unsigned)RING)
        ring[part * RSTRIDE + slot] = (unsigned short)(id + part * p.neg_tile_cells);
}

// one group of 256 rows (8 per lane) whose keys the caller already holds in registers (loaded one group ahead).
// Returns the largest slot a row of this lane was given (>= RING: that row did not fit, see `overflow`).
template <typename T, int ND, int RING, int RSTRIDE, bool CLAMPED>
__device__ __forceinline__ unsigned group_rows_regs(const RingParams &p, const T (&c)[2][ND][4], unsigned *cnt, unsigned short *ring, unsigned idx[8],
                                                    unsigned slots[8]) {
    unsigned worst = 0;
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            T v[ND];
#pragma unroll
            for (int d = 0; d < ND; d++)
                v[d] = c[q][d][j];
            place_row<T, ND, RING, RSTRIDE, CLAMPED>(p, v, cnt, ring, idx[q * 4 + j], slots[q * 4 + j], worst);
        }
    return worst;
}

// the ragged end of the batch: global loads with bounds; slots[] = kNone marks rows that do not exist
template <typename T, int ND, int RING, int RSTRIDE, bool CLAMPED>
__device__ __forceinline__ unsigned group_rows_tail(const RingParams &p, long long gbase, long long tend, int lane, unsigned *cnt, unsigned short *ring,
                                                    unsigned idx[8], unsigned slots[8]) {
    unsigned worst = 0;
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long row = gbase + q * 128 + lane * 4 + j;
            idx[q * 4 + j] = 0;
            slots[q * 4 + j] = kNone;
            if (row < tend) {
                T v[ND];
#pragma unroll
                for (int d = 0; d < ND; d++)
                    v[d] = __ldcs(static_cast<const T *>(p.x[d]) + row);
                place_row<T, ND, RING, RSTRIDE, CLAMPED>(p, v, cnt, ring, idx[q * 4 + j], slots[q * 4 + j], worst);
            }
        }
    return worst;
}

// shared memory of one warp: [rings: 32*PPL parts x (RING + 8) u16] [counters: 32*PPL u32] [flush map: 32 x {position, owner}]
// A ring holds < 64 entries left over from the last flush + the new ones of one group; a flush moves whole LINES of 64 entries
// (128 bytes).  With <= 32 new entries per part and group the ring never fills (the busiest part of the headline workload takes
// 19 +- 4 of a group's 256 rows); rows that find it full go to direct REDs.
template <typename T, int ND, int PPL, int FG>
struct RingLayout {
    static constexpr int kRing = kLine == 64 ? 96 : 64;
    static constexpr int kRingStride = kRing + 8; // 208 bytes: 16-byte aligned rows that rotate over the banks
    static constexpr size_t kColBytes = 0; // the keys travel through registers
    static constexpr size_t kRingBytes = 32ull * PPL * kRingStride * 2;
    static constexpr size_t kCntBytes = 32ull * PPL * 4 + 256; // + the flush's (write position, owner) map
    static constexpr size_t kPerWarp = kColBytes + kRingBytes + kCntBytes;
};

// state of the parts a lane owns (parts lane, lane+32, ...): where the next granule goes
template <int PPL>
struct Owner {
    unsigned wpos[PPL]; // next entry to write (pool entry index); == wend: the current chunk is full / there is none yet
    unsigned wend[PPL];
    unsigned cur[PPL];  // current chunk id (kNone: none yet)
    unsigned total[PPL]; // entries written to this list so far
};

template <typename T, int ND, int PPL, int FG, int WARPS, bool CLAMPED>
__global__ void __launch_bounds__(WARPS * 32, 2) k_ring_partition(const __grid_constant__ RingParams p) {
    using L = RingLayout<T, ND, PPL, FG>;
    constexpr int RING = L::kRing, RSTRIDE = L::kRingStride;
    static_assert(RING - kLine <= kLine && (RSTRIDE * 2) % 16 == 0, "ring geometry");
    extern __shared__ __align__(128) unsigned char dyn_smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *mine = dyn_smem + (size_t)warp * L::kPerWarp;
    unsigned short *const ring = reinterpret_cast<unsigned short *>(mine + L::kColBytes); // [32*PPL][RING + 8]
    unsigned *const cnt = reinterpret_cast<unsigned *>(mine + L::kColBytes + L::kRingBytes);
    uint2 *const fmap = reinterpret_cast<uint2 *>(cnt + 32 * PPL); // [32] {pool entry index, owner lane} of the lines of one flush
#pragma unroll
    for (int k = 0; k < PPL; k++)
        cnt[lane + 32 * k] = 0;
    __syncwarp();

    const unsigned cshift = p.chunk_shift, csize = 1u << cshift;
    const long long wglobal = (long long)blockIdx.x * WARPS + warp, wtotal = (long long)gridDim.x * WARPS;

    Owner<PPL> own;
#pragma unroll
    for (int k = 0; k < PPL; k++)
        own.wpos[k] = own.wend[k] = 0, own.cur[k] = kNone, own.total[k] = 0;
    // the warp's reservation in the pool: chunks [sc_next, sc_end) are free; lane 0 holds the NEXT reservation (prefetched)
    unsigned sc_next = 0, sc_end = 0, sn_pref = 0;
    if (lane == 0)
        sn_pref = atomicAdd(p.ctl, kSuper);

    // ---- flush: every ring that holds a full LINE (64 entries = 128 bytes) appends it to its chunk list -------------------------
    // The owners (lane l owns parts l, l+32) do the bookkeeping in parallel: chunk switch, write position, counter.  Then the warp
    // moves the lines, FOUR parts per step: a quarter-warp copies one line with one LDS.128 + one coalesced STG.128 per lane and
    // slides the <= 32 entries behind it to the ring's front.  `final`: the incomplete last line too, padded with entries ==
    // tile_cells (k_ring_count's spare cell).
    auto flush = [&](bool final) {
#pragma unroll
        for (int k = 0; k < PPL; k++) {
            const int part = lane + 32 * k;
            unsigned n = min(cnt[part], (unsigned)RING); // rows past RING went to direct REDs
            if (final && n > 0 && n < kLine) {
                for (unsigned e = n; e < kLine; e++)
                    ring[part * RSTRIDE + e] = (unsigned short)p.tile_cells;
                n = kLine;
            }
            const bool has = n >= kLine; // cnt of parts >= nparts stays 0
            const bool need = has && own.wpos[k] == own.wend[k];
            // hand out chunks (rare: once per chunk per part); warp-uniform bookkeeping, ids in lane order
            unsigned newid = kNone;
            const unsigned mneed = __ballot_sync(0xffffffffu, need);
            if (mneed) {
                const unsigned sn = __shfl_sync(0xffffffffu, sn_pref, 0);
                const unsigned i = __popc(mneed & ((1u << lane) - 1u)), tot = __popc(mneed), avail = sc_end - sc_next;
                if (need)
                    newid = i < avail ? sc_next + i : sn + (i - avail);
                if (tot >= avail) { // the current reservation is used up: switch to the prefetched one, prefetch another
                    sc_next = sn + (tot - avail);
                    sc_end = sn + kSuper;
                    if (lane == 0)
                        sn_pref = atomicAdd(p.ctl, kSuper);
                } else {
                    sc_next += tot;
                }
            }
            unsigned wp = kNone; // pool entry index where this part's line goes
            if (has) {
                if (need && newid < p.nchunks_cap) { // open the next chunk of this list
                    if (own.cur[k] == kNone)
                        p.head[(unsigned long long)wglobal * p.nparts_pad + part] = newid;
                    else
                        p.next[own.cur[k]] = newid;
                    own.cur[k] = newid;
                    own.wpos[k] = newid << cshift;
                    own.wend[k] = own.wpos[k] + csize;
                }
                if (own.wpos[k] != own.wend[k]) {
                    wp = own.wpos[k];
                    own.wpos[k] += kLine;
                    own.total[k] += kLine;
                } else { // pool exhausted: cannot happen (sized for the worst case); stay exact anyway
                    for (unsigned e = 0; e < kLine; e++) {
                        const unsigned c = ring[part * RSTRIDE + e];
                        if (c < p.tile_cells)
                            atomicAdd(p.grid + (unsigned long long)part * p.tile_cells + c, 1ull);
                    }
                }
                cnt[part] = n - kLine;
            }
            // the owners publish {write position, lane} in rank order; then every group of kLineLanes lanes moves one line
            const unsigned m = __ballot_sync(0xffffffffu, has);
            if (m) {
                const int nl = __popc(m);
                if (has)
                    fmap[__popc(m & ((1u << lane) - 1u))] = make_uint2(wp, (unsigned)lane);
                __syncwarp();
                const int grp = lane / kLineLanes, gl = lane % kLineLanes;
                for (int j = grp; j < nl; j += 32 / kLineLanes) {
                    const uint2 e = fmap[j];
                    uint4 *r = reinterpret_cast<uint4 *>(ring + (e.y + 32 * k) * RSTRIDE);
                    const uint4 w = r[gl];
                    uint4 up = make_uint4(0, 0, 0, 0);
                    if (gl < (RING - (int)kLine) / 8)
                        up = r[kLineLanes + gl];
                    if (e.x != kNone)
                        reinterpret_cast<uint4 *>(p.pool + e.x)[gl] = w;
                    if (gl < (RING - (int)kLine) / 8)
                        r[gl] = up; // the entries behind the line slide to the front (each lane rewrites the 16 bytes it read)
                }
                __syncwarp();
            }
        }
    };
    // rows that found their ring full (slot >= RING) go straight to the grid; `slots` are the atomics' return values of one group
    auto overflow = [&](const unsigned idx[8], const unsigned slots[8]) {
#pragma unroll
        for (int r = 0; r < 8; r++)
            if (slots[r] >= (unsigned)RING && slots[r] != kNone)
                atomicAdd(p.grid + idx[r], 1ull);
    };

    // ---- complete groups of 256 rows: the keys of the NEXT group are loaded into registers (128-bit, evict-first) while the
    // current one is placed: two register sets, the loop is unrolled by two so that no copy is needed.  (A TMA-staged variant
    // spent 8.5 of its 56 warp instructions per 32 rows on issuing the bulk copies and waiting on the mbarriers, and 4 KB of shared
    // memory per warp on the double buffer: profiles/r02_ncu_ring_v4.txt.)
    const long long ngroups = p.nrows / kGroupRows;
    auto load = [&](long long g, T (&dst)[2][ND][4]) {
        const long long r0 = p.row0 + g * kGroupRows + lane * 4;
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int d = 0; d < ND; d++)
                ldg4<T>(p.x[d], r0 + q * 128, dst[q][d]);
    };
    auto place = [&](const T (&c)[2][ND][4]) {
        unsigned idx[8], slots[8];
        const unsigned worst = group_rows_regs<T, ND, RING, RSTRIDE, CLAMPED>(p, c, cnt, ring, idx, slots);
        __syncwarp();
        flush(false);
        if (__any_sync(0xffffffffu, worst >= (unsigned)RING))
            overflow(idx, slots);
        __syncwarp();
    };
    {
        T ca[2][ND][4], cb[2][ND][4];
        long long g = wglobal;
        if (g < ngroups)
            load(g, ca);
        while (g < ngroups) {
            if (g + wtotal < ngroups)
                load(g + wtotal, cb);
            place(ca);
            g += wtotal;
            if (g >= ngroups)
                break;
            if (g + wtotal < ngroups)
                load(g + wtotal, ca);
            place(cb);
            g += wtotal;
        }
    }
    // ---- the ragged end (< 256 rows), taken by the warp whose turn it is ------------------------------------------------------------
    if (ngroups * kGroupRows < p.nrows && ngroups % wtotal == wglobal) {
        unsigned idx[8], slots[8];
        const unsigned worst = group_rows_tail<T, ND, RING, RSTRIDE, CLAMPED>(p, p.row0 + ngroups * kGroupRows, p.row0 + p.nrows, lane, cnt, ring, idx, slots);
        __syncwarp();
        flush(false);
        if (__any_sync(0xffffffffu, worst >= (unsigned)RING))
            overflow(idx, slots);
        __syncwarp();
    }
    // ---- the incomplete last lines, then the list descriptors ---------------------------------------------------------------
    flush(true);
#pragma unroll
    for (int k = 0; k < PPL; k++) {
        const int part = lane + 32 * k;
        if (part < p.nparts) {
            if (own.cur[k] != kNone)
                p.unused[own.cur[k]] = (unsigned short)(own.wend[k] - own.wpos[k]);
            p.len[(unsigned long long)wglobal * p.nparts_pad + part] = own.total[k];
        }
    }
}

// ---- K2 -------------------------------------------------------------------------------------------------------------------
// The lists are laid out in the order (part, warp); CTA b takes the b-th of gridDim.x equal shares of that sequence measured in
// entries, assuming the entries of a part are spread evenly over its warps (they are: the warps take row tiles round-robin).
__global__ void __launch_bounds__(kCountThreads, 1) k_ring_count(const __grid_constant__ RingParams p) {
    extern __shared__ __align__(16) unsigned hist[];
    __shared__ unsigned long long s_part_entries[kMaxParts + 1]; // exclusive prefix over parts
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nparts = p.nparts, npad = p.nparts_pad;
    const unsigned W = p.nlists_w;
    // entries per part = sum over warps of len[w][part]
    for (int part = warp; part < nparts; part += kCountThreads / 32) {
        unsigned long long s = 0;
        for (unsigned w = lane; w < W; w += 32)
            s += p.len[(unsigned long long)w * npad + part];
#pragma unroll
        for (int o = 16; o; o >>= 1)
            s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0)
            s_part_entries[part + 1] = s;
    }
    __syncthreads();
    if (tid == 0) {
        s_part_entries[0] = 0;
        for (int i = 0; i < nparts; i++)
            s_part_entries[i + 1] += s_part_entries[i];
    }
    __syncthreads();
    const unsigned long long total = s_part_entries[nparts];
    const unsigned long long lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
    if (lo >= hi)
        return;
    const unsigned tc = p.tile_cells;
    const unsigned cshift = p.chunk_shift, csize = 1u << cshift;
    for (int part = 0; part < nparts; part++) {
        const unsigned long long b = s_part_entries[part], e = s_part_entries[part + 1];
        if (e <= lo || b >= hi || e == b)
            continue;
        // the warps [w0, w1) of this part whose lists fall into [lo, hi): boundaries computed identically by both neighbours
        const unsigned long long n = e - b;
        const unsigned w0 = lo <= b ? 0u : (unsigned)(((lo - b) * W) / n);
        const unsigned w1 = hi >= e ? W : (unsigned)(((hi - b) * W) / n);
        if (w0 >= w1)
            continue;
        for (int i = tid; i < (int)(tc / 4 + 1); i += kCountThreads)
            reinterpret_cast<uint4 *>(hist)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        for (unsigned w = w0 + warp; w < w1; w += kCountThreads / 32) {
            unsigned id = p.head[(unsigned long long)w * npad + part];
            while (id != kNone) {
                const unsigned nx = p.next[id];
                const unsigned valid = csize - p.unused[id];
                const uint4 *src = reinterpret_cast<const uint4 *>(p.pool + ((unsigned long long)id << cshift));
                // a chunk = csize/512 steps of 32 lanes x 16 entries; all loads of a chunk are issued before the first atomic
                uint4 v[8];
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const unsigned off = s * 512 + lane * kGran;
                    if (off < valid) {
                        v[2 * s] = __ldcs(src + (off >> 3));
                        v[2 * s + 1] = __ldcs(src + (off >> 3) + 1);
                    }
                }
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const unsigned off = s * 512 + lane * kGran;
                    if (off < valid) {
                        const unsigned wd[8] = {v[2 * s].x, v[2 * s].y, v[2 * s].z, v[2 * s].w, v[2 * s + 1].x, v[2 * s + 1].y, v[2 * s + 1].z, v[2 * s + 1].w};
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            atomicAdd(hist + (wd[k] & 0xffffu), 1u); // pads carry tile_cells: the spare cell, never flushed
                            atomicAdd(hist + (wd[k] >> 16), 1u);
                        }
                    }
                }
                id = nx;
            }
        }
        __syncthreads();
        const unsigned long long cell0 = (unsigned long long)part * tc;
        for (int i = tid; i < (int)tc; i += kCountThreads) {
            const unsigned c = hist[i];
            if (c && cell0 + i < p.cells)
                atomicAdd(p.grid + cell0 + i, (unsigned long long)c);
        }
        __syncthreads();
    }
}

template <typename T, int ND, int PPL, int FG>
int launch_partition_cfg(int sm_count, cudaStream_t st, RingParams &p, int *warps_out, bool dry, bool clamped) {
    using L = RingLayout<T, ND, PPL, FG>;
    constexpr int fit = (int)((113 * 1024 - 512) / L::kPerWarp); // two CTAs per SM
    constexpr int WARPS = fit > 12 ? 12 : fit; // 24 warps per SM leave 85 registers per thread: the two key sets + the placement fit without spills
    if constexpr (WARPS >= 2) {
        const long long ngroups = (p.nrows + kGroupRows - 1) / kGroupRows;
        // at least ~8 groups per warp so that the per-(warp, part) partial chunks stay a small share of the scratch
        long long blocks = std::min<long long>((ngroups / 8 + WARPS - 1) / WARPS, (long long)sm_count * 2);
        if (blocks < 1)
            blocks = 1;
        *warps_out = (int)blocks * WARPS;
        if (dry)
            return B200_OK;
        auto kern = clamped ? k_ring_partition<T, ND, PPL, FG, WARPS, std::is_same<T, float>::value> : k_ring_partition<T, ND, PPL, FG, WARPS, false>;
        constexpr size_t dyn = L::kPerWarp * WARPS;
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<(int)blocks, WARPS * 32, dyn, st>>>(p);
        B200_CUDA(cudaGetLastError());
        return B200_OK;
    } else {
        *warps_out = 0;
        return B200_OK;
    }
}

template <typename T, int ND>
int launch_partition_nd(int sm_count, cudaStream_t st, RingParams &p, int *warps_out, bool dry, bool clamped) {
    if (p.nparts_pad == 32)
        return launch_partition_cfg<T, ND, 1, 1>(sm_count, st, p, warps_out, dry, clamped);
    return launch_partition_cfg<T, ND, 2, 1>(sm_count, st, p, warps_out, dry, clamped);
}

template <typename T>
int launch_partition(int nd, int sm_count, cudaStream_t st, RingParams &p, int *warps_out, bool dry, bool clamped) {
    switch (nd) {
    case 1: return launch_partition_nd<T, 1>(sm_count, st, p, warps_out, dry, clamped);
    case 2: return launch_partition_nd<T, 2>(sm_count, st, p, warps_out, dry, clamped);
    default: return launch_partition_nd<T, 3>(sm_count, st, p, warps_out, dry, clamped);
    }
}

// ---- host side of the CLAMPED index (see bin_cell_m2_clamped) ---------------------------------------------------------------
// the reference formula, evaluated exactly like src/binners.cpp:13-57 does (IEEE double, no contraction: volatile keeps every
// intermediate rounded to double whatever the host compiler flags)
double ref_t(float v, double vmin, double scale, double bins_d) {
    volatile double d = (double)v - vmin;
    volatile double s = d * scale;
    volatile double t = s * bins_d;
    return t;
}
// cell of a non-NaN key
long long ref_cell(float v, double vmin, double scale, double bins_d, unsigned bins) {
    volatile double d = (double)v - vmin;
    volatile double s = d * scale;
    if (s < 0)
        return 1;
    if (s >= 1)
        return (long long)bins + 2;
    volatile double t = s * bins_d;
    return (long long)(int)t + 2;
}
// fp32 values in increasing order <-> integers in increasing order (-inf .. +inf, NaNs excluded)
int float_order(float f) {
    int b;
    memcpy(&b, &f, 4);
    return b >= 0 ? b : (int)(0x80000000u - (unsigned)b);
}
float order_float(int o) {
    const int b = o >= 0 ? o : (int)(0x80000000u - (unsigned)o);
    float f;
    memcpy(&f, &b, 4);
    return f;
}
// the clamp bounds of one dimension; false when the construction does not apply (then the plain index runs)
bool clamp_bounds(double vmin, double scale, double bins_d, unsigned bins, float *lo_out, float *hi_out) {
    const int omin = float_order(-INFINITY), omax = float_order(INFINITY);
    auto first_with_cell_at_least = [&](long long c) { // smallest order o in [omin, omax] with cell >= c; omax + 1 if none
        long long lo = omin, hi = (long long)omax + 1;
        while (lo < hi) {
            const long long mid = lo + (hi - lo) / 2;
            if (ref_cell(order_float((int)mid), vmin, scale, bins_d, bins) >= c)
                hi = mid;
            else
                lo = mid + 1;
        }
        return lo;
    };
    const long long o_in = first_with_cell_at_least(2), o_over = first_with_cell_at_least((long long)bins + 2);
    if (o_in <= omin || o_over > omax || o_in >= o_over)
        return false; // no underflow value, no overflow value, or no value in range
    const float lo = order_float((int)(o_in - 1)), hi = order_float((int)o_over);
    if (!(lo > -INFINITY) || !(hi < INFINITY))
        return false;
    const double tlo = ref_t(lo, vmin, scale, bins_d), thi = ref_t(hi, vmin, scale, bins_d);
    // floor(t(lo)) == -1 and floor(t(hi)) == bins: every clamped key then has -1 <= floor(t) <= bins by monotonicity
    if (!(tlo >= -1.0 && tlo < 0.0) || !(thi >= bins_d && thi < bins_d + 1.0))
        return false;
    *lo_out = lo;
    *hi_out = hi;
    return true;
}

// part == __umulhi(idx, magic) for all idx < cells?  Both sides are monotone step functions of idx, so it suffices to check the
// last index of every part and the first index of the next one.
bool magic_exact(unsigned cells, unsigned tile_cells, unsigned magic) {
    for (unsigned long long k = 0; k * tile_cells < cells; k++) {
        const unsigned long long first = k * tile_cells, last = std::min<unsigned long long>(first + tile_cells, cells) - 1;
        if ((unsigned)((first * magic) >> 32) != k || (unsigned)((last * magic) >> 32) != k)
            return false;
    }
    return true;
}

} // namespace

// Scratch (pool + list tables) lives in the slot; grown on demand.
int try_launch_ringcount(b200_ctx *ctx, Slot *slot, const BinParams &bp, bool vec, bool *taken) {
    *taken = false;
    static const bool disabled = getenv("B200_DISABLE_RINGCOUNT") && atoi(getenv("B200_DISABLE_RINGCOUNT")) != 0;
    static const bool no_clamp = getenv("B200_RING_NO_CLAMP") && atoi(getenv("B200_RING_NO_CLAMP")) != 0; // A/B knob
    if (disabled || !vec || bp.nb < 1 || bp.nb > 3 || bp.na != 1 || bp.nrows < (1ll << 22))
        return B200_OK;
    const DevAgg &a = bp.a[0];
    if (a.op != B200_AGG_COUNT || a.data || a.mask)
        return B200_OK;
    const unsigned long long cells = bp.cells;
    if (cells * 4 <= 96 * 1024 || cells > (unsigned long long)kMaxTileCells * kMaxParts || cells >= (1ull << 22))
        return B200_OK; // small grids: shared-memory privatisation; huge grids: direct REDs / region sort
    const int t = bp.b[0].dtype;
    if (t != B200_F32 && t != B200_F64)
        return B200_OK;
    RingParams p;
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
        p.stride_sum += (unsigned)b.stride;
    }
    bool clamped = t == B200_F32 && !no_clamp;
    for (int i = 0; i < bp.nb && clamped; i++)
        clamped = clamp_bounds(p.vmin[i], p.scale[i], p.bins_d[i], p.bins[i], &p.clamp_lo[i], &p.clamp_hi[i]);
    // as few parts as possible (32, else 64, else 128): the ring memory of k_ring_partition scales with the parts per lane;
    // tile_cells is nudged upwards until the one-instruction division is exact
    unsigned tile_cells = 0, magic = 0;
    for (int np = 32; np <= kMaxParts && !tile_cells; np *= 2) {
        unsigned long long tcells = ((cells + np - 1) / np + 7) / 8 * 8;
        for (int tries = 0; tries < 64 && tcells <= kMaxTileCells; tries++, tcells += 8) {
            const unsigned m = (unsigned)((1ull << 32) / tcells) + 1u;
            if (magic_exact((unsigned)cells, (unsigned)tcells, m)) {
                tile_cells = (unsigned)tcells;
                magic = m;
                break;
            }
        }
    }
    if (!tile_cells)
        return B200_OK;
    p.cells = (unsigned)cells;
    p.tile_cells = tile_cells;
    p.neg_tile_cells = 0u - tile_cells;
    p.magic = magic;
    p.nparts = (int)((cells + tile_cells - 1) / tile_cells);
    p.nparts_pad = p.nparts <= 32 ? 32 : 64;
    p.grid = static_cast<unsigned long long *>(a.grid);

    // batches of <= 2^30 rows (32-bit entry counts), EQUAL in size: a 2^30 + remainder split paid the fixed cost of a pass (memsets,
    // 148 histogram flushes per part) for a small second batch — 1.25e9 rows ran 10 % slower per row than 1e9
    const long long nbatch = (bp.nrows + (1ll << 30) - 1) >> 30;
    const long long batch = std::min<long long>(bp.nrows, (((bp.nrows + nbatch - 1) / nbatch) + 255) & ~255ll);
    p.row0 = 0;
    p.nrows = batch;
    int nwarps = 0;
    if (t == B200_F32)
        B200_CHECK(launch_partition<float>(bp.nb, ctx->sm_count, nullptr, p, &nwarps, true, clamped));
    else
        B200_CHECK(launch_partition<double>(bp.nb, ctx->sm_count, nullptr, p, &nwarps, true, false));
    if (nwarps <= 0)
        return B200_OK;
    // chunk size: about a quarter of a (warp, part) list, within 512..2048 entries
    unsigned cshift = 9;
    while (cshift < 11 && (unsigned long long)batch / ((unsigned long long)nwarps * p.nparts) >= (8ull << cshift))
        cshift++;
    p.chunk_shift = cshift;
    p.nlists_w = (unsigned)nwarps;
    // every list wastes less than one chunk, every warp less than one reservation (+ its prefetched one)
    const unsigned long long nchunks = ((unsigned long long)batch >> cshift) + (unsigned long long)nwarps * (p.nparts + 2 * kSuper) + 2 * kSuper;
    if ((nchunks << cshift) >= (1ull << 32))
        return B200_OK;
    p.nchunks_cap = (unsigned)nchunks;
    const size_t lists = (size_t)nwarps * p.nparts_pad;
    // layout: [ctl 256 B | len | unused] (zeroed)  [head | next] (0xFF)  [pool]
    const size_t off_len = 256, off_unused = align_up(off_len + lists * 4, 256), zero_end = align_up(off_unused + nchunks * 2, 256);
    const size_t off_head = zero_end, off_next = align_up(off_head + lists * 4, 256), ff_end = align_up(off_next + nchunks * 4, 256);
    const size_t off_pool = ff_end, need = off_pool + (nchunks << cshift) * 2;
    if (slot->scratch_cap < need) {
        if (slot->scratch) {
            B200_CUDA(cudaStreamSynchronize(slot->stream));
            B200_CUDA(cudaFree(slot->scratch));
            slot->scratch = nullptr;
            slot->scratch_cap = 0;
        }
        cudaError_t e = cudaMalloc(&slot->scratch, need);
        if (e != cudaSuccess) { // not enough memory for the pool: the direct RED kernel takes the call
            cudaGetLastError();
            return B200_OK;
        }
        slot->scratch_cap = need;
    }
    char *base = static_cast<char *>(slot->scratch);
    p.ctl = reinterpret_cast<unsigned *>(base);
    p.len = reinterpret_cast<unsigned *>(base + off_len);
    p.unused = reinterpret_cast<unsigned short *>(base + off_unused);
    p.head = reinterpret_cast<unsigned *>(base + off_head);
    p.next = reinterpret_cast<unsigned *>(base + off_next);
    p.pool = reinterpret_cast<unsigned short *>(base + off_pool);
    cudaStream_t st = slot->stream;
    B200_CUDA(cudaFuncSetAttribute(k_ring_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((kMaxTileCells + 1) * 4)));
    const size_t hist_bytes = ((size_t)tile_cells / 4 + 1) * 16;
    for (long long r0 = 0; r0 < bp.nrows; r0 += batch) {
        p.row0 = r0;
        p.nrows = std::min<long long>(batch, bp.nrows - r0);
        B200_CUDA(cudaMemsetAsync(base, 0, zero_end, st));
        B200_CUDA(cudaMemsetAsync(base + off_head, 0xFF, ff_end - off_head, st));
        int w = 0;
        if (t == B200_F32)
            B200_CHECK(launch_partition<float>(bp.nb, ctx->sm_count, st, p, &w, false, clamped));
        else
            B200_CHECK(launch_partition<double>(bp.nb, ctx->sm_count, st, p, &w, false, false));
        p.nlists_w = (unsigned)w; // the last batch may launch fewer warps; its lists are the first w rows of head[] / len[]
        k_ring_count<<<ctx->sm_count, kCountThreads, hist_bytes, st>>>(p);
        B200_CUDA(cudaGetLastError());
        slot->ring_len = p.len, slot->ring_ctl = p.ctl, slot->ring_lists = (size_t)w * p.nparts_pad;
        slot->ring_rows = (uint64_t)p.nrows, slot->ring_memset_bytes = zero_end + (ff_end - off_head), slot->ring_chunk_entries = 1u << cshift;
    }
    *taken = true;
    return B200_OK;
}

} // namespace b200
