// binby.cuh — kernel parameter blocks shared by the binby kernels and the C-ABI layer.
#pragma once
#include "common.cuh"

namespace b200 {

// open-addressing slot of the device ordered_set (hashset.cu)
#define SET_EMPTY 0xFFFFFFFFFFFFFFFFULL
struct SetSlot {
    unsigned long long key;   // canonical 64-bit key pattern
    unsigned long long first; // pass 1: (chunk_seq << 40 | row) of the first occurrence; after finalize: global ordinal
};

struct DevBinner {
    int kind, dtype, isz, byteswap, allow_other, invert;
    double vmin, scale, bins_d;
    unsigned long long bins;
    long long ordinal_count, min_value;
    unsigned long long stride;
    const void *data;
    const uint8_t *mask;
    // B200_BINNER_HASH: finalized set table
    const SetSlot *table;
    unsigned long long table_mask;
    long long nan_ordinal;      // -1 when the set saw no NaN
    long long null_ordinal;     // -1 when the set saw no null (masked rows then fall into the null cell)
    long long sentinel_ordinal; // ordinal of the key whose pattern equals SET_EMPTY, -1 if absent
};

struct DevAgg {
    int op, dtype, isz, byteswap, cell_dtype, smem_off; // smem_off: byte offset of this grid in the CTA's private copy
    unsigned moment;
    int smem_cell; // bytes per cell in the shared-memory copy (COUNT: 4, else the device cell size)
    unsigned long long init_bits; // identity element of the device cell (0, +-inf, numeric limits)
    const void *data;
    const uint8_t *mask;
    void *grid;
};

struct BinParams {
    int nb, na;
    long long nrows;
    unsigned long long cells;
    int smem_copies;     // 0 = global atomics; >= 1: number of per-warp-group private copies in shared memory
    int smem_copy_bytes; // bytes of one private copy (all aggregators)
    DevBinner b[B200_MAX_BINNERS];
    DevAgg a[B200_MAX_AGGS];
};

// FIRST/LAST (first.cu)
struct FirstParams {
    int nb;
    long long nrows, row_offset;
    DevBinner b[B200_MAX_BINNERS];
    int dtype, isz, dtype2, isz2, byteswap, invert;
    const void *data;
    const void *order;
    const uint8_t *mask;
    void *grid;                 // value cells (dtype)
    void *order_grid;           // order cells (dtype2)
    unsigned long long *state;  // {key,row} pairs, 16 B aligned
    uint8_t *cell_masked;
};

// specialised float-binner paths (fast.cu, tilesort.cu)
struct FastParams {
    const void *x[3];
    double vmin[3], scale[3], bins_d[3];
    unsigned bins[3];
    unsigned stride[3];
    long long nrows;
    unsigned cells;
    const void *v;                      // value column or null
    unsigned long long *count_star;     // nullable grids (global)
    unsigned long long *vcount;
    double *vsum;
    double *vm2;
    int smem_copies;                    // > 0: privatise in shared memory (u32 counts, f64 sums)
};

// tilesort.cu: rows sorted by grid region first, so that the scatter works on an L2-resident part of the grids
int try_launch_tilesort(b200_ctx *ctx, Slot *slot, const FastParams &p, int xdtype, int nd, int vdtype, bool *taken);

// NUNIQUE (nunique.cu)
struct NUniqueParams {
    int nb;
    long long row0, nrows;
    DevBinner b[B200_MAX_BINNERS];
    int dtype, isz, byteswap;
    const void *data;
    const uint8_t *valid;      // 1 = value present (nullable)
    const uint8_t *selection;  // 1 = row takes part (nullable)
    unsigned long long *table; // 2 u64 per slot: {cell, canonical value bits}; empty = {~0, ~0}
    unsigned long long tmask;
    unsigned long long *distinct, *nan_rows, *null_rows; // cells each
    unsigned long long *total;
};

int launch_binby(b200_ctx *ctx, Slot *slot, const BinParams &p, bool vec);
int launch_nunique(b200_ctx *ctx, cudaStream_t stream, const NUniqueParams &p, bool vec);
int launch_nunique_rehash(cudaStream_t stream, const unsigned long long *old_table, unsigned long long old_cap, unsigned long long *table, unsigned long long cap);
int launch_first(b200_ctx *ctx, cudaStream_t stream, const FirstParams &p, bool vec);
int launch_fill(cudaStream_t stream, void *ptr, int cell_dtype, uint64_t cells, uint64_t bits);
int launch_fill_elems(cudaStream_t stream, void *ptr, int isz, uint64_t n, uint64_t bits);
int launch_merge(cudaStream_t stream, int op, int cell_dtype, void *dst, const void *src, uint64_t cells);
int launch_merge_first(cudaStream_t stream, b200_agg *dst, const b200_agg *src);

} // namespace b200
