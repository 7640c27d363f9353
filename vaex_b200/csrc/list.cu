// list.cu — AggList_<dtype>: per cell the list of the rows' values.
//
// Reference: AggListPrimitive (src/agg_list.cpp:5-127): `grids` must be 1; aggregate() appends every valid, non-NaN value to its
// cell's std::vector, counts NaN values (unless dropnan) and null rows (data mask == 0, unless dropnull) per cell; get_result()
// returns offsets[cells + 1] + flat values: a cell's values in arrival order, then one NaN per counted NaN, then one (unwritten)
// slot per counted null — handed to vaex.arrow.convert.list_from_arrays.
// Device design: nothing cell-shaped is kept.  Every b200_bin call appends one record per row — key = cell * 4 + category (0 value,
// 1 NaN, 2 null; skipped rows get the all-ones key), payload = the value's bits — to a growing pair of device arrays, at positions
// reserved per call, so arrival order is (call, row) order.  Finishing = one stable LSD radix sort of the records by key (radix.cuh,
// only the bytes that vary) + a per-cell count + a scan: the sorted payloads ARE the flat values.
#include <algorithm>

#include "binby.cuh"
#include "binby_index.cuh"
#include "radix.cuh"
#include "scan.cuh"

namespace b200 {

struct ListParams {
    int nb;
    long long nrows;
    DevBinner b[B200_MAX_BINNERS];
    int dtype, isz, byteswap, dropnan, dropnull;
    const void *data;
    const uint8_t *mask; // aggregator convention: 1 = use the row, 0 = null row
    unsigned long long *keys, *vals;
    unsigned long long base;
    unsigned long long skip_key; // cells * 4: sorts behind every real record, costs no extra radix pass
};

namespace {

template <bool VEC>
__global__ void __launch_bounds__(256) k_list_append(const __grid_constant__ ListParams p) {
    const long long step = (long long)gridDim.x * 256 * 4;
    for (long long base = ((long long)blockIdx.x * 256 + threadIdx.x) * 4; base < p.nrows; base += step) {
        const long long left = p.nrows - base;
        const int nv = left < 4 ? (int)left : 4;
        unsigned long long idx[4];
        binby_indices<VEC>(p.b, p.nb, base, nv, idx);
        uint64_t r[4] = {0, 0, 0, 0};
        unsigned m[4] = {1, 1, 1, 1};
        load4_raw<VEC>(p.data, p.isz, base, nv, r);
        // REFERENCE QUIRK, kept (golden vectors from the compiled reference pin it): AggListPrimitive::aggregate runs per 1024-row
        // block of a bin() call with the block offset applied to the data but NOT to the mask (src/agg_list.cpp:96-97), so row r of
        // a call is judged by mask[r % 1024] — the same slip as AggFirst (src/agg_first.cpp:131).  base is a multiple of 4.
        if (p.mask)
            load4_mask<VEC>(p.mask, base & 1023, nv, m);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j >= nv)
                break;
            const uint64_t raw = p.byteswap ? bswap(r[j], p.isz) : r[j];
            unsigned long long key;
            if (m[j] == 1) {
                if (!raw_isnan(p.dtype, raw))
                    key = idx[j] * 4 + 0;
                else
                    key = p.dropnan ? p.skip_key : idx[j] * 4 + 1;
            } else {
                key = (m[j] == 0 && !p.dropnull) ? idx[j] * 4 + 2 : p.skip_key;
            }
            p.keys[p.base + base + j] = key;
            p.vals[p.base + base + j] = raw;
        }
    }
}

__global__ void k_list_count(const unsigned long long *keys, unsigned long long n, unsigned *counts, unsigned long long cells) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        if ((k >> 2) < cells)
            atomicAdd(counts + (k >> 2), 1u);
    }
}

__global__ void k_list_values(const unsigned long long *keys, const unsigned long long *vals, unsigned long long total, int dtype, int isz, void *out) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned cat = (unsigned)(keys[i] & 3ull);
        unsigned long long v = vals[i];
        if (cat == 1) // std::numeric_limits<T>::quiet_NaN()
            v = dtype == B200_F64 ? 0x7ff8000000000000ull : 0x7fc00000ull;
        else if (cat == 2)
            v = 0; // the reference leaves these slots unwritten
        switch (isz) {
        case 8: static_cast<unsigned long long *>(out)[i] = v; break;
        case 4: static_cast<unsigned *>(out)[i] = (unsigned)v; break;
        case 2: static_cast<unsigned short *>(out)[i] = (unsigned short)v; break;
        default: static_cast<unsigned char *>(out)[i] = (unsigned char)v; break;
        }
    }
}

int nblocks_for(unsigned long long n) {
    const unsigned long long b = (n + 255) / 256;
    return (int)std::max<unsigned long long>(1, std::min<unsigned long long>(b, 148ull * 16));
}

} // namespace

// one b200_bin call: reserve nrows records, append (api.cu calls this for B200_AGG_LIST aggregators)
int bin_list(b200_ctx *ctx, Slot *sl, b200_agg *a, const DevBinner *db, int nbinners, const void *data, const uint8_t *mask, int64_t nrows, bool vec) {
    unsigned long long base;
    {
        std::lock_guard<std::mutex> g(a->nmu);
        if (a->list_n + (uint64_t)nrows > a->list_cap) { // grow: other slots may be appending into the old arrays
            B200_CUDA(cudaDeviceSynchronize());
            const uint64_t cap = std::max<uint64_t>((a->list_n + (uint64_t)nrows) * 2, 1u << 16);
            unsigned long long *nk = nullptr, *nv = nullptr;
            B200_CUDA(cudaMalloc(&nk, cap * 8));
            B200_CUDA(cudaMalloc(&nv, cap * 8));
            if (a->list_n) {
                B200_CUDA(cudaMemcpy(nk, a->list_keys, a->list_n * 8, cudaMemcpyDeviceToDevice));
                B200_CUDA(cudaMemcpy(nv, a->list_vals, a->list_n * 8, cudaMemcpyDeviceToDevice));
            }
            cudaFree(a->list_keys);
            cudaFree(a->list_vals);
            a->list_keys = nk, a->list_vals = nv, a->list_cap = cap;
        }
        base = a->list_n;
        a->list_n += (uint64_t)nrows;
        a->list_sorted = false;
    }
    ListParams p;
    memset(&p, 0, sizeof p);
    p.nb = nbinners;
    p.nrows = nrows;
    memcpy(p.b, db, sizeof(DevBinner) * nbinners);
    p.dtype = a->dtype;
    p.isz = dtype_size(a->dtype);
    p.byteswap = a->byteswap && p.isz > 1;
    p.dropnan = (a->moment & 1) != 0; // AggList_<T>(grid, grids, threads, dropnan, dropnull): carried in `moment` like NUNIQUE's flags
    p.dropnull = (a->moment & 2) != 0;
    p.data = data;
    p.mask = mask;
    p.keys = a->list_keys;
    p.vals = a->list_vals;
    p.base = base;
    p.skip_key = a->cells * 4;
    const bool v = vec && !(reinterpret_cast<uintptr_t>(data) & 15) && !(reinterpret_cast<uintptr_t>(mask) & 15);
    const int blocks = nblocks_for(((unsigned long long)nrows + 3) / 4);
    if (v)
        k_list_append<true><<<blocks, 256, 0, sl->stream>>>(p);
    else
        k_list_append<false><<<blocks, 256, 0, sl->stream>>>(p);
    B200_CUDA(cudaGetLastError());
    (void)ctx;
    return B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" {

/* sorts the records and reports the length of the flat value array (offsets[cells]); every slot is synchronised first */
int b200_agg_list_finish(b200_agg *a, int64_t *total_out) {
    if (!a || a->op != B200_AGG_LIST || !total_out) {
        set_error("b200_agg_list_finish: not a list aggregator");
        return B200_ERR_INVALID;
    }
    B200_CUDA(cudaSetDevice(a->ctx->device));
    B200_CHECK(b200_ctx_sync(a->ctx, -1));
    std::lock_guard<std::mutex> g(a->nmu);
    cudaStream_t st = a->ctx->slots[0]->stream;
    const uint64_t n = a->list_n;
    if (!a->list_sorted && n > 1) {
        if (n >= (1ull << 32)) {
            set_error("AggList: more than 2^32 rows are not supported");
            return B200_ERR_UNSUPPORTED;
        }
        const unsigned nblk = radix_blocks(n), tiles = radix_tiles(n);
        unsigned long long *kb = nullptr, *vb = nullptr;
        unsigned *hist = nullptr;
        B200_CUDA(cudaMalloc(&kb, n * 8));
        B200_CUDA(cudaMalloc(&vb, n * 8));
        B200_CUDA(cudaMalloc(&hist, (size_t)256 * nblk * 4));
        unsigned long long *kin = a->list_keys, *vin = a->list_vals, *kout = kb, *vout = vb;
        // keys are cell * 4 + category (skipped rows: cells * 4): only the bytes that can differ are sorted on
        const unsigned long long maxkey = a->cells * 4 + 3;
        for (int shift = 0; shift < 64 && (maxkey >> shift); shift += 8) {
            k_radix_hist<<<nblk, kRadixThreads, 0, st>>>(kin, vin, n, shift, 0, hist, nblk, tiles);
            k_scan_u32<<<1, 1024, 0, st>>>(hist, 256ull * nblk);
            k_radix_scatter<<<nblk, kRadixThreads, 0, st>>>(kin, vin, kout, vout, n, shift, 0, hist, nblk, tiles);
            B200_CUDA(cudaGetLastError());
            std::swap(kin, kout);
            std::swap(vin, vout);
        }
        B200_CUDA(cudaStreamSynchronize(st));
        if (kin != a->list_keys) { // an odd number of passes: the sorted records sit in the scratch arrays, which have n entries
            B200_CUDA(cudaMemcpy(a->list_keys, kin, n * 8, cudaMemcpyDeviceToDevice));
            B200_CUDA(cudaMemcpy(a->list_vals, vin, n * 8, cudaMemcpyDeviceToDevice));
        }
        cudaFree(kb);
        cudaFree(vb);
        cudaFree(hist);
    }
    a->list_sorted = true;
    // per-cell counts -> offsets (kept on the device until read)
    const size_t cn = (size_t)a->cells + 1;
    if (!a->list_counts)
        B200_CUDA(cudaMalloc((void **)&a->list_counts, cn * 4));
    B200_CUDA(cudaMemsetAsync(a->list_counts, 0, cn * 4, st));
    if (n)
        k_list_count<<<nblocks_for(n), 256, 0, st>>>(a->list_keys, n, a->list_counts, a->cells);
    unsigned long long *d_total = nullptr;
    B200_CUDA(cudaMalloc((void **)&d_total, 8));
    k_scan_u32<<<1, 1024, 0, st>>>(a->list_counts, cn, d_total);
    B200_CUDA(cudaGetLastError());
    unsigned long long total = 0;
    B200_CUDA(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_total);
    a->list_total = total;
    *total_out = (int64_t)total;
    return B200_OK;
}

/* after b200_agg_list_finish: offsets_out = int64[cells + 1], values_out = total elements of the aggregator's dtype */
int b200_agg_list_read(b200_agg *a, int64_t *offsets_out, void *values_out) {
    if (!a || a->op != B200_AGG_LIST || !offsets_out || !a->list_sorted || !a->list_counts) {
        set_error("b200_agg_list_read: call b200_agg_list_finish first");
        return B200_ERR_STATE;
    }
    B200_CUDA(cudaSetDevice(a->ctx->device));
    std::lock_guard<std::mutex> g(a->nmu);
    cudaStream_t st = a->ctx->slots[0]->stream;
    const size_t cn = (size_t)a->cells + 1;
    std::vector<unsigned> off(cn);
    B200_CUDA(cudaMemcpyAsync(off.data(), a->list_counts, cn * 4, cudaMemcpyDeviceToHost, st));
    const int isz = dtype_size(a->dtype);
    void *d_out = nullptr;
    if (a->list_total && values_out) {
        B200_CUDA(cudaMalloc(&d_out, a->list_total * isz));
        k_list_values<<<nblocks_for(a->list_total), 256, 0, st>>>(a->list_keys, a->list_vals, a->list_total, a->dtype, isz, d_out);
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaMemcpyAsync(values_out, d_out, a->list_total * isz, cudaMemcpyDeviceToHost, st));
    }
    B200_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_out);
    for (size_t i = 0; i < cn; i++)
        offsets_out[i] = (int64_t)off[i];
    return B200_OK;
}

} // extern "C"
