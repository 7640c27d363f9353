// fast.cu — compile-time specialised binby kernels for the configurations BASELINE.json is quoted on:
//   all binners BinnerScalar over the SAME float type T (fp32 / fp64, native byte order, no masks), ND = 1..3,
//   aggregators drawn from { count(*), count(v), sum(v), sum(v^2) } over at most one value column v (fp32 / fp64).
// That covers df.count / df.sum / df.mean / df.std(binby=[...]) on float columns (configs C1, C2, C3, C5).  Everything
// else takes the descriptor-driven kernel in binby.cu; both are bit-identical on the cases they share (tests).
//
// Why a second kernel: the generic one spends ~143 SASS instructions per row on runtime dtype/op dispatch and 64-bit
// index math and stalls on instruction fetch (profiles/r01_ncu_binby_generic_details.txt); here the whole row costs ~30.
// The bound that remains is the scatter itself: one RED per row per aggregator, and the L2 retires ~98 REDs per clock
// chip-wide (profiles/r01_microbench.txt) whatever the operand width.
//
// Load pattern: each thread owns ROWS = 32/sizeof(T) consecutive... no — two 128-bit loads per column per step, the second
// one 32 lanes further on, so every warp-level LDG.128 is a dense 512-byte segment (fully coalesced, evict-first).
#include <stdlib.h>

#include "binby.cuh"
#include "device_utils.cuh"

namespace b200 {


namespace {

constexpr int kThreads = 256;


template <typename T>
__device__ __forceinline__ double widen(T v);
template <>
__device__ __forceinline__ double widen<float>(float v) { return (double)v; }
template <>
__device__ __forceinline__ double widen<double>(double v) { return v; }

// R consecutive elements starting at element index i (16-byte aligned): R = 4 for float, 2 for double
template <typename T>
struct Vec;
template <>
struct Vec<float> {
    static constexpr int R = 4;
    float v[4];
    __device__ __forceinline__ void load(const void *p, long long i) {
        uint4 a = __ldcs(reinterpret_cast<const uint4 *>(static_cast<const float *>(p) + i));
        v[0] = __uint_as_float(a.x), v[1] = __uint_as_float(a.y), v[2] = __uint_as_float(a.z), v[3] = __uint_as_float(a.w);
    }
};
template <>
struct Vec<double> {
    static constexpr int R = 2;
    double v[2];
    __device__ __forceinline__ void load(const void *p, long long i) {
        uint4 a = __ldcs(reinterpret_cast<const uint4 *>(static_cast<const double *>(p) + i));
        v[0] = __longlong_as_double(((long long)a.y << 32) | a.x), v[1] = __longlong_as_double(((long long)a.w << 32) | a.z);
    }
};

template <typename TV, bool SMEM>
__device__ __forceinline__ void scatter(const FastParams &p, unsigned idx, bool has_v, TV vraw, unsigned *s_count, unsigned *s_vcount, double *s_vsum,
                                        double *s_vm2) {
    if (p.count_star) {
        if (SMEM)
            atomicAdd(s_count + idx, 1u);
        else
            atomicAdd(p.count_star + idx, 1ull);
    }
    if (has_v) {
        const double v = widen<TV>(vraw);
        if (v == v) { // NaN values are skipped by count(v) / sum(v) (src/agg_count.cpp:53-57, src/agg_sum.cpp:118-121)
            if (p.vcount) {
                if (SMEM)
                    atomicAdd(s_vcount + idx, 1u);
                else
                    atomicAdd(p.vcount + idx, 1ull);
            }
            if (p.vsum)
                atomicAdd((SMEM ? s_vsum : p.vsum) + idx, v);
            if (p.vm2)
                atomicAdd((SMEM ? s_vm2 : p.vm2) + idx, v * v); // pow(v, 2)
        }
    }
}

template <typename T, int ND, typename TV, bool HASV, bool SMEM>
__global__ void __launch_bounds__(kThreads) k_binby_fast(const __grid_constant__ FastParams p) {
    extern __shared__ __align__(16) char smem[];
    unsigned *s_count = nullptr, *s_vcount = nullptr;
    double *s_vsum = nullptr, *s_vm2 = nullptr;
    if (SMEM) {
        // layout of one private copy: [vsum f64][vm2 f64][count u32][vcount u32]; `smem_copies` copies, one per warp group
        const size_t c = p.cells;
        const size_t copy_bytes = ((p.vsum ? 8 * c : 0) + (p.vm2 ? 8 * c : 0) + (p.count_star ? 4 * c : 0) + (p.vcount ? 4 * c : 0) + 15) / 16 * 16;
        for (size_t i = threadIdx.x; i < copy_bytes * p.smem_copies / 4; i += kThreads)
            reinterpret_cast<unsigned *>(smem)[i] = 0u;
        __syncthreads();
        char *base = smem + copy_bytes * ((threadIdx.x >> 5) % p.smem_copies);
        if (p.vsum) {
            s_vsum = reinterpret_cast<double *>(base);
            base += 8 * c;
        }
        if (p.vm2) {
            s_vm2 = reinterpret_cast<double *>(base);
            base += 8 * c;
        }
        if (p.count_star) {
            s_count = reinterpret_cast<unsigned *>(base);
            base += 4 * c;
        }
        if (p.vcount)
            s_vcount = reinterpret_cast<unsigned *>(base);
    }

    constexpr int R = Vec<T>::R;           // rows per 128-bit load of a binner column
    constexpr int SPAN = 32 * R;           // rows one warp-level load covers
    constexpr int ROWS_PER_WARP_STEP = 2 * SPAN;
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * kThreads) >> 5;
    const long long full_steps = p.nrows / ROWS_PER_WARP_STEP;

    for (long long s = warp_global; s < full_steps; s += nwarps) {
        const long long base = s * ROWS_PER_WARP_STEP + (long long)lane * R;
        Vec<T> c[ND][2];
#pragma unroll
        for (int d = 0; d < ND; d++) {
            c[d][0].load(p.x[d], base);
            c[d][1].load(p.x[d], base + SPAN);
        }
        TV vv[2][R];
        if (HASV) {
            // the value column may be wider/narrower than T: load element-wise vectors of R rows
            if (sizeof(TV) == sizeof(T)) {
                Vec<TV> a, b;
                a.load(p.v, base);
                b.load(p.v, base + SPAN);
#pragma unroll
                for (int j = 0; j < R; j++) {
                    vv[0][j] = a.v[j < Vec<TV>::R ? j : 0];
                    vv[1][j] = b.v[j < Vec<TV>::R ? j : 0];
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int j = 0; j < R; j++)
                        vv[h][j] = __ldcs(static_cast<const TV *>(p.v) + base + h * SPAN + j);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int j = 0; j < R; j++) {
                unsigned idx = 0;
#pragma unroll
                for (int d = 0; d < ND; d++)
                    idx += bin_index(widen<T>(c[d][h].v[j]), p.vmin[d], p.scale[d], p.bins_d[d], p.bins[d]) * p.stride[d];
                scatter<TV, SMEM>(p, idx, HASV, HASV ? vv[h][j] : TV(0), s_count, s_vcount, s_vsum, s_vm2);
            }
        }
    }
    // ragged tail (< ROWS_PER_WARP_STEP rows): one row per thread of the first CTAs
    const long long tail0 = full_steps * ROWS_PER_WARP_STEP;
    for (long long i = tail0 + (long long)blockIdx.x * kThreads + threadIdx.x; i < p.nrows; i += (long long)gridDim.x * kThreads) {
        unsigned idx = 0;
#pragma unroll
        for (int d = 0; d < ND; d++)
            idx += bin_index(widen<T>(__ldcs(static_cast<const T *>(p.x[d]) + i)), p.vmin[d], p.scale[d], p.bins_d[d], p.bins[d]) * p.stride[d];
        scatter<TV, SMEM>(p, idx, HASV, HASV ? __ldcs(static_cast<const TV *>(p.v) + i) : TV(0), s_count, s_vcount, s_vsum, s_vm2);
    }

    if (SMEM) {
        __syncthreads();
        const size_t c = p.cells;
        const size_t copy_bytes = ((p.vsum ? 8 * c : 0) + (p.vm2 ? 8 * c : 0) + (p.count_star ? 4 * c : 0) + (p.vcount ? 4 * c : 0) + 15) / 16 * 16;
        for (unsigned i = threadIdx.x; i < p.cells; i += kThreads) {
            unsigned long long n0 = 0, n1 = 0;
            double a = 0, b = 0;
            for (int k = 0; k < p.smem_copies; k++) {
                char *base = smem + copy_bytes * k;
                if (p.vsum) {
                    a += reinterpret_cast<double *>(base)[i];
                    base += 8 * c;
                }
                if (p.vm2) {
                    b += reinterpret_cast<double *>(base)[i];
                    base += 8 * c;
                }
                if (p.count_star) {
                    n0 += reinterpret_cast<unsigned *>(base)[i];
                    base += 4 * c;
                }
                if (p.vcount)
                    n1 += reinterpret_cast<unsigned *>(base)[i];
            }
            if (p.count_star && n0)
                atomicAdd(p.count_star + i, n0);
            if (p.vcount && n1)
                atomicAdd(p.vcount + i, n1);
            if (p.vsum && a != 0.0)
                atomicAdd(p.vsum + i, a);
            if (p.vm2 && b != 0.0)
                atomicAdd(p.vm2 + i, b);
        }
    }
}

template <typename T, int ND, typename TV, bool HASV>
int launch3(b200_ctx *ctx, cudaStream_t stream, const FastParams &p, size_t smem) {
    auto go = [&](auto kern) -> int {
        if (smem > 48 * 1024)
            B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
        if (per_sm < 1)
            per_sm = 1;
        const long long rows_per_cta = (long long)kThreads * 2 * Vec<T>::R;
        long long want = (p.nrows + rows_per_cta - 1) / rows_per_cta;
        long long cap = (long long)ctx->sm_count * per_sm;
        int blocks = (int)(want < cap ? want : cap);
        kern<<<blocks < 1 ? 1 : blocks, kThreads, smem, stream>>>(p);
        B200_CUDA(cudaGetLastError());
        return B200_OK;
    };
    if (p.smem_copies > 0)
        return go(k_binby_fast<T, ND, TV, HASV, true>);
    return go(k_binby_fast<T, ND, TV, HASV, false>);
}

template <typename T, int ND>
int launch2(b200_ctx *ctx, cudaStream_t stream, const FastParams &p, int vdtype, size_t smem) {
    if (!p.v)
        return launch3<T, ND, float, false>(ctx, stream, p, smem);
    if (vdtype == B200_F32)
        return launch3<T, ND, float, true>(ctx, stream, p, smem);
    return launch3<T, ND, double, true>(ctx, stream, p, smem);
}

template <typename T>
int launch1(b200_ctx *ctx, cudaStream_t stream, const FastParams &p, int nd, int vdtype, size_t smem) {
    switch (nd) {
    case 1: return launch2<T, 1>(ctx, stream, p, vdtype, smem);
    case 2: return launch2<T, 2>(ctx, stream, p, vdtype, smem);
    default: return launch2<T, 3>(ctx, stream, p, vdtype, smem);
    }
}

} // namespace

// Returns B200_OK and sets *taken when the request matched the fast path and was launched.
int try_launch_fast(b200_ctx *ctx, Slot *slot, const BinParams &bp, bool vec, bool *taken) {
    cudaStream_t stream = slot->stream;
    *taken = false;
    if (!vec || bp.nb < 1 || bp.nb > 3 || bp.na < 1 || bp.cells >= (1ull << 32) || bp.nrows <= 0)
        return B200_OK;
    FastParams p;
    memset(&p, 0, sizeof p);
    const int t = bp.b[0].dtype;
    if (t != B200_F32 && t != B200_F64)
        return B200_OK;
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
    int vdtype = -1;
    for (int k = 0; k < bp.na; k++) {
        const DevAgg &a = bp.a[k];
        if (a.mask || a.byteswap)
            return B200_OK;
        if (a.op == B200_AGG_COUNT && !a.data) {
            if (p.count_star)
                return B200_OK;
            p.count_star = static_cast<unsigned long long *>(a.grid);
            continue;
        }
        if (a.dtype != B200_F32 && a.dtype != B200_F64)
            return B200_OK;
        if (p.v && (p.v != a.data || vdtype != a.dtype))
            return B200_OK; // one value column only
        p.v = a.data;
        vdtype = a.dtype;
        if (a.op == B200_AGG_COUNT && !p.vcount)
            p.vcount = static_cast<unsigned long long *>(a.grid);
        else if (a.op == B200_AGG_SUM && !p.vsum)
            p.vsum = static_cast<double *>(a.grid);
        else if (a.op == B200_AGG_SUM_MOMENT && a.moment == 2 && !p.vm2)
            p.vm2 = static_cast<double *>(a.grid);
        else
            return B200_OK;
    }
    p.nrows = bp.nrows;
    p.cells = (unsigned)bp.cells;
    // shared-memory privatisation for small grids (ATOMS retires ~6 lanes/clk/SM vs ~0.66 for L2 REDs)
    size_t copy = ((p.vsum ? 8ull : 0) + (p.vm2 ? 8ull : 0) + (p.count_star ? 4ull : 0) + (p.vcount ? 4ull : 0)) * bp.cells;
    copy = (copy + 15) / 16 * 16;
    size_t smem = 0;
    if (copy <= 96 * 1024 && bp.nrows >= 8192) {
        int copies = (int)(32 * 1024 / copy);
        p.smem_copies = copies < 1 ? 1 : (copies > 8 ? 8 : copies);
        smem = copy * p.smem_copies;
    }
    const int naggs = (p.count_star != nullptr) + (p.vcount != nullptr) + (p.vsum != nullptr) + (p.vm2 != nullptr);
    if (!smem && ((size_t)naggs * 8 * bp.cells > (100u << 20) || getenv("B200_TILESORT_FORCE"))) {
        // accumulators larger than the L2: sort the rows by grid region first (tilesort.cu), when the input is big enough to pay for it
        bool sorted = false;
        B200_CHECK(try_launch_tilesort(ctx, slot, p, t, bp.nb, vdtype, &sorted));
        if (sorted) {
            *taken = true;
            return B200_OK;
        }
    }
    *taken = true;
    int rc = t == B200_F32 ? launch1<float>(ctx, stream, p, bp.nb, vdtype, smem) : launch1<double>(ctx, stream, p, bp.nb, vdtype, smem);
    return rc;
}

} // namespace b200
