/*
 * binstats_oracle.c — CPU restatement of the vaex binned-statistics / groupby hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (vaex_b200) never does.
 *
 * Every function restates one piece of the reference C++ (paths under
 * /root/reference/packages/vaex-core/src/, cited per function).  Plain C11, scalar, one thread.
 * Parity status: PINNED — checked against the compiled reference (oracle/_ref) and the reference's
 * own known-answer vectors in tests/test_oracle_pinning.py.
 *
 * dtype codes (the order of create_alltypes.hpp):
 *   0 f64, 1 f32, 2 i64, 3 i32, 4 i16, 5 i8, 6 u64, 7 u32, 8 u16, 9 u8, 10 bool
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <limits.h>

enum { ORC_F64 = 0, ORC_F32, ORC_I64, ORC_I32, ORC_I16, ORC_I8, ORC_U64, ORC_U32, ORC_U16, ORC_U8, ORC_BOOL, ORC_NDTYPE };
enum { ORC_BINNER_SCALAR = 0, ORC_BINNER_ORDINAL = 1 };
enum { ORC_COUNT = 0, ORC_SUM, ORC_SUM_MOMENT, ORC_MIN, ORC_MAX, ORC_FIRST };

#define ORC_INDEX_BLOCK_SIZE 1024 /* agg.hpp:28 */

static const int orc_itemsize_[ORC_NDTYPE] = {8, 4, 8, 4, 2, 1, 8, 4, 2, 1, 1};
int orc_itemsize(int dtype) { return (dtype >= 0 && dtype < ORC_NDTYPE) ? orc_itemsize_[dtype] : -1; }

/* agg.hpp:17-26 _to_native: reverse the bytes of a value */
static void flip_bytes(void *p, size_t n) {
    unsigned char *b = (unsigned char *)p;
    for (size_t i = 0; i < n / 2; i++) {
        unsigned char t = b[i];
        b[i] = b[n - 1 - i];
        b[n - 1 - i] = t;
    }
}

/* ------------------------------------------------------------------------------------------------
 * typed element access: FOR_DTYPE(dtype, MACRO) expands MACRO(ctype) for the matching C type.
 * bool is carried as uint8_t (the raw byte the reference's `bool` load sees).
 * ---------------------------------------------------------------------------------------------- */
#define FOR_DTYPE(dtype, M)                                                                                                    \
    switch (dtype) {                                                                                                           \
    case ORC_F64: M(double) break;                                                                                             \
    case ORC_F32: M(float) break;                                                                                              \
    case ORC_I64: M(int64_t) break;                                                                                            \
    case ORC_I32: M(int32_t) break;                                                                                            \
    case ORC_I16: M(int16_t) break;                                                                                            \
    case ORC_I8: M(int8_t) break;                                                                                              \
    case ORC_U64: M(uint64_t) break;                                                                                           \
    case ORC_U32: M(uint32_t) break;                                                                                           \
    case ORC_U16: M(uint16_t) break;                                                                                           \
    case ORC_U8: M(uint8_t) break;                                                                                             \
    case ORC_BOOL: M(uint8_t) break;                                                                                           \
    default: return -1;                                                                                                        \
    }

/* ------------------------------------------------------------------------------------------------
 * BinnerScalar<T>::to_bins — binners.cpp:13-57
 *   scaled = (double(v) - vmin) * (1./(vmax-vmin));  nan/masked -> 0, <0 -> 1, >=1 -> bins+2,
 *   else (int)(scaled*bins)+2;  out[i-offset] += index*stride.   mask: 1 == masked (numpy convention)
 * Compiled with -ffp-contract=off so no FMA is formed (the reference is built without -march=native).
 * ---------------------------------------------------------------------------------------------- */
int orc_scalar_to_bins(int dtype, int flip, const void *data, const uint8_t *mask, double vmin, double vmax, uint64_t bins, uint64_t offset,
                       uint64_t length, uint64_t stride, uint64_t *out) {
    const double scale_v = 1. / (vmax - vmin);
#define BODY(T)                                                                                                                \
    {                                                                                                                          \
        const T *p = (const T *)data;                                                                                          \
        for (uint64_t i = offset; i < offset + length; i++) {                                                                  \
            T value = p[i];                                                                                                    \
            if (flip)                                                                                                          \
                flip_bytes(&value, sizeof(T));                                                                                 \
            double value_double = (double)value;                                                                               \
            double scaled = (value_double - vmin) * scale_v;                                                                   \
            uint64_t index = 0;                                                                                                \
            int masked = mask ? (mask[i] == 1) : 0;                                                                            \
            if (scaled != scaled || masked) {                                                                                  \
            } else if (scaled < 0) {                                                                                           \
                index = 1;                                                                                                     \
            } else if (scaled >= 1) {                                                                                          \
                index = bins - 1 + 3;                                                                                          \
            } else {                                                                                                           \
                index = (uint64_t)(int64_t)((int)(scaled * (double)bins) + 2);                                                 \
            }                                                                                                                  \
            out[i - offset] += index * stride;                                                                                 \
        }                                                                                                                      \
    }
    FOR_DTYPE(dtype, BODY)
#undef BODY
    return 0;
}

/* x86 cvttsd2si / cvttss2si semantics for float->int64 (what the reference binary executes for
 * `int64_t value = data_ptr[i] - min_value` on float/double data): NaN and out-of-range give
 * INT64_MIN ("integer indefinite").  Written out explicitly so the oracle does not lean on UB. */
static int64_t f64_to_i64_x86(double x) {
    if (!(x == x) || x >= 9223372036854775808.0 || x < -9223372036854775808.0)
        return INT64_MIN;
    return (int64_t)x;
}

/* ------------------------------------------------------------------------------------------------
 * BinnerOrdinal<T>::to_bins — binner_ordinal.cpp:20-176
 *   value = int64(data[i] - min_value)   (arithmetic in T's promoted type; see per-type notes)
 *   layout [bin0..binN-1, (other), null, nan];  std::isnan(int64) is never true, so the nan cell is
 *   unreachable; masked -> null (or null cell N+1 when allow_other); out of range -> null / other.
 *   FlipEndian quirk (:28-30): the byte flip is applied to the *int64 difference*, not to the T load.
 * ---------------------------------------------------------------------------------------------- */
int orc_ordinal_to_bins(int dtype, int flip, const void *data, const uint8_t *mask, int64_t ordinal_count, int64_t min_value, int allow_other,
                        int invert, uint64_t offset, uint64_t length, uint64_t stride, uint64_t *out) {
    for (uint64_t i = offset; i < offset + length; i++) {
        int64_t value;
        switch (dtype) {
        case ORC_F64: value = f64_to_i64_x86(((const double *)data)[i] - (double)min_value); break;
        case ORC_F32: value = f64_to_i64_x86((double)(float)(((const float *)data)[i] - (float)min_value)); break;
        case ORC_I64: value = (int64_t)((uint64_t)((const int64_t *)data)[i] - (uint64_t)min_value); break;
        case ORC_I32: value = (int64_t)((const int32_t *)data)[i] - min_value; break;
        case ORC_I16: value = (int64_t)((const int16_t *)data)[i] - min_value; break;
        case ORC_I8: value = (int64_t)((const int8_t *)data)[i] - min_value; break;
        case ORC_U64: value = (int64_t)(((const uint64_t *)data)[i] - (uint64_t)min_value); break;
        case ORC_U32: value = (int64_t)((const uint32_t *)data)[i] - min_value; break;
        case ORC_U16: value = (int64_t)((const uint16_t *)data)[i] - min_value; break;
        case ORC_U8:
        case ORC_BOOL: value = (int64_t)((const uint8_t *)data)[i] - min_value; break;
        default: return -1;
        }
        if (flip)
            flip_bytes(&value, sizeof(value));
        uint64_t index;
        int masked = mask ? (mask[i] == 1) : 0;
        int oob = (value < 0) || (value >= ordinal_count);
        if (allow_other) {
            if (masked)
                index = (uint64_t)(ordinal_count + 1);
            else if (oob)
                index = (uint64_t)ordinal_count;
            else
                index = (uint64_t)(invert ? ordinal_count - 1 - value : value);
        } else {
            if (masked || oob)
                index = (uint64_t)ordinal_count;
            else
                index = (uint64_t)(invert ? ordinal_count - 1 - value : value);
        }
        out[i - offset] += index * stride;
    }
    return 0;
}

/* shape(): binners.cpp:59 (bins+3), binner_ordinal.cpp:178 (N + 2 | 3) */
uint64_t orc_binner_shape(int kind, uint64_t bins_or_count, int allow_other) {
    return kind == ORC_BINNER_SCALAR ? bins_or_count + 3 : bins_or_count + (allow_other ? 3 : 2);
}

/* ------------------------------------------------------------------------------------------------
 * Aggregators.  mask convention here (agg_*.cpp): data_mask[j+offset] == 1 means USE the row.
 * ---------------------------------------------------------------------------------------------- */

/* AggCountPrimitive::aggregate — agg_count.cpp:43-67.  grid: int64 */
int orc_agg_count(int dtype, int flip, const void *data, const uint8_t *mask, int64_t *grid, const uint64_t *idx, size_t length, uint64_t offset) {
    if (!mask && !data) {
        for (size_t j = 0; j < length; j++)
            grid[idx[j]] += 1;
        return 0;
    }
#define BODY(T)                                                                                                                \
    {                                                                                                                          \
        const T *p = (const T *)data;                                                                                          \
        for (size_t j = 0; j < length; j++) {                                                                                  \
            if (mask == NULL || mask[j + offset] == 1) {                                                                       \
                if (p) {                                                                                                       \
                    T value = p[j + offset];                                                                                   \
                    if (flip)                                                                                                  \
                        flip_bytes(&value, sizeof(T));                                                                         \
                    if (value != value)                                                                                        \
                        continue;                                                                                              \
                }                                                                                                              \
                grid[idx[j]] += 1;                                                                                             \
            }                                                                                                                  \
        }                                                                                                                      \
    }
    FOR_DTYPE(dtype, BODY)
#undef BODY
    return 0;
}

/* upcast<T> — agg_sum.cpp:6-62: float/double -> double; signed ints + bool -> int64; unsigned -> uint64 */
int orc_upcast(int dtype) {
    switch (dtype) {
    case ORC_F64:
    case ORC_F32: return ORC_F64;
    case ORC_I64:
    case ORC_I32:
    case ORC_I16:
    case ORC_I8:
    case ORC_BOOL: return ORC_I64;
    default: return ORC_U64;
    }
}

/* AggSumPrimitive / AggSumMomentPrimitive via AggregatorPrimitiveCRTP::aggregate — agg_sum.cpp:98-127,
 * op_mutate :139 (a += b) and :159 (a += pow(b, moment)).  moment == 0 selects plain sum.
 * For integer grids `a += pow(b, moment)` is evaluated in double and truncated back (C++ usual
 * arithmetic conversions): a = (G)((double)a + pow((double)b, (double)moment)). */
int orc_agg_sum(int dtype, int flip, const void *data, const uint8_t *mask, void *grid, const uint64_t *idx, size_t length, uint64_t offset,
                uint32_t moment, int use_moment) {
    if (!data)
        return -2; /* "data not set" */
    int up = orc_upcast(dtype);
#define BODY(T)                                                                                                                \
    {                                                                                                                          \
        const T *p = (const T *)data;                                                                                          \
        for (size_t j = 0; j < length; j++) {                                                                                  \
            if (mask && mask[j + offset] != 1)                                                                                 \
                continue;                                                                                                      \
            T value = p[j + offset];                                                                                           \
            if (flip)                                                                                                          \
                flip_bytes(&value, sizeof(T));                                                                                 \
            if (value != value)                                                                                                \
                continue;                                                                                                      \
            if (up == ORC_F64) {                                                                                               \
                double *g = (double *)grid;                                                                                    \
                double b = (double)value;                                                                                      \
                g[idx[j]] += use_moment ? pow(b, (double)moment) : b;                                                          \
            } else if (up == ORC_I64) {                                                                                        \
                int64_t *g = (int64_t *)grid;                                                                                  \
                int64_t b = (int64_t)value;                                                                                    \
                if (use_moment)                                                                                                \
                    g[idx[j]] = (int64_t)((double)g[idx[j]] + pow((double)b, (double)moment));                                 \
                else                                                                                                           \
                    g[idx[j]] = (int64_t)((uint64_t)g[idx[j]] + (uint64_t)b);                                                  \
            } else {                                                                                                           \
                uint64_t *g = (uint64_t *)grid;                                                                                \
                uint64_t b = (uint64_t)value;                                                                                  \
                if (use_moment)                                                                                                \
                    g[idx[j]] = (uint64_t)((double)g[idx[j]] + pow((double)b, (double)moment));                                \
                else                                                                                                           \
                    g[idx[j]] += b;                                                                                            \
            }                                                                                                                  \
        }                                                                                                                      \
    }
    FOR_DTYPE(dtype, BODY)
#undef BODY
    return 0;
}

/* AggMinPrimitive / AggMaxPrimitive::aggregate — agg_minmax.cpp:45-74 (max), :120-145 (min).
 * grid dtype == data dtype;  std::max(value, cell) = (value < cell) ? cell : value. */
int orc_agg_minmax(int dtype, int flip, const void *data, const uint8_t *mask, void *grid, const uint64_t *idx, size_t length, uint64_t offset,
                   int is_max) {
    if (!data)
        return -2;
#define BODY(T)                                                                                                                \
    {                                                                                                                          \
        const T *p = (const T *)data;                                                                                          \
        T *g = (T *)grid;                                                                                                      \
        for (size_t j = 0; j < length; j++) {                                                                                  \
            if (mask && mask[j + offset] != 1)                                                                                 \
                continue;                                                                                                      \
            T value = p[j + offset];                                                                                           \
            if (flip)                                                                                                          \
                flip_bytes(&value, sizeof(T));                                                                                 \
            if (value != value)                                                                                                \
                continue;                                                                                                      \
            T cell = g[idx[j]];                                                                                                \
            if (is_max)                                                                                                        \
                g[idx[j]] = (value < cell) ? cell : value;                                                                     \
            else                                                                                                               \
                g[idx[j]] = (cell < value) ? cell : value;                                                                     \
        }                                                                                                                      \
    }
    FOR_DTYPE(dtype, BODY)
#undef BODY
    return 0;
}

/* initial_fill — agg_count.cpp:13 (0), agg_sum.cpp:137 (0), agg_minmax.cpp:13-18 / :83-87
 * (max: -inf or numeric_limits::min(); min: +inf or numeric_limits::max()).  bool: min()=false,max()=true */
int orc_fill_minmax(int dtype, void *grid, uint64_t cells, int is_max) {
    for (uint64_t i = 0; i < cells; i++) {
        switch (dtype) {
        case ORC_F64: ((double *)grid)[i] = is_max ? -INFINITY : INFINITY; break;
        case ORC_F32: ((float *)grid)[i] = is_max ? -INFINITY : INFINITY; break;
        case ORC_I64: ((int64_t *)grid)[i] = is_max ? INT64_MIN : INT64_MAX; break;
        case ORC_I32: ((int32_t *)grid)[i] = is_max ? INT32_MIN : INT32_MAX; break;
        case ORC_I16: ((int16_t *)grid)[i] = is_max ? INT16_MIN : INT16_MAX; break;
        case ORC_I8: ((int8_t *)grid)[i] = is_max ? INT8_MIN : INT8_MAX; break;
        case ORC_U64: ((uint64_t *)grid)[i] = is_max ? 0 : UINT64_MAX; break;
        case ORC_U32: ((uint32_t *)grid)[i] = is_max ? 0 : UINT32_MAX; break;
        case ORC_U16: ((uint16_t *)grid)[i] = is_max ? 0 : UINT16_MAX; break;
        case ORC_U8: ((uint8_t *)grid)[i] = is_max ? 0 : UINT8_MAX; break;
        case ORC_BOOL: ((uint8_t *)grid)[i] = is_max ? 0 : 1; break;
        default: return -1;
        }
    }
    return 0;
}

/* order-column helpers for AggFirst: load element i of a typed column as (is_float, double, int64/uint64) */
typedef struct {
    int cls; /* 0 float, 1 signed, 2 unsigned */
    double f;
    int64_t s;
    uint64_t u;
} orc_num;

static int load_num(int dtype, const void *data, uint64_t i, int flip, orc_num *o) {
#define BODY(T)                                                                                                                \
    {                                                                                                                          \
        T v = ((const T *)data)[i];                                                                                            \
        if (flip)                                                                                                              \
            flip_bytes(&v, sizeof(T));                                                                                         \
        if (dtype == ORC_F64 || dtype == ORC_F32) {                                                                            \
            o->cls = 0;                                                                                                        \
            o->f = (double)v;                                                                                                  \
        } else if (dtype == ORC_I64 || dtype == ORC_I32 || dtype == ORC_I16 || dtype == ORC_I8) {                              \
            o->cls = 1;                                                                                                        \
            o->s = (int64_t)v;                                                                                                 \
        } else {                                                                                                               \
            o->cls = 2;                                                                                                        \
            o->u = (uint64_t)v;                                                                                                \
        }                                                                                                                      \
    }
    FOR_DTYPE(dtype, BODY)
#undef BODY
    return 0;
}

static int num_less(const orc_num *a, const orc_num *b) { return a->cls == 0 ? a->f < b->f : (a->cls == 1 ? a->s < b->s : a->u < b->u); }
static int num_isnan(const orc_num *a) { return a->cls == 0 && a->f != a->f; }

static void store_typed(int dtype, void *arr, uint64_t i, const orc_num *v) {
    switch (dtype) {
    case ORC_F64: ((double *)arr)[i] = v->f; break;
    case ORC_F32: ((float *)arr)[i] = (float)v->f; break;
    case ORC_I64: ((int64_t *)arr)[i] = v->s; break;
    case ORC_I32: ((int32_t *)arr)[i] = (int32_t)v->s; break;
    case ORC_I16: ((int16_t *)arr)[i] = (int16_t)v->s; break;
    case ORC_I8: ((int8_t *)arr)[i] = (int8_t)v->s; break;
    case ORC_U64: ((uint64_t *)arr)[i] = v->u; break;
    case ORC_U32: ((uint32_t *)arr)[i] = (uint32_t)v->u; break;
    case ORC_U16: ((uint16_t *)arr)[i] = (uint16_t)v->u; break;
    default: ((uint8_t *)arr)[i] = (uint8_t)v->u; break;
    }
}

/* AggFirstPrimitive::initial_fill — agg_first.cpp:19-26: value 99, order = invert ? limits::min() : limits::max()
 * (NB numeric_limits<double>::min() is DBL_MIN, the smallest positive normal — kept as is), cell_masked 1. */
int orc_fill_first(int dtype, int dtype2, void *grid, void *grid_order, uint8_t *cell_masked, uint64_t cells, int invert) {
    for (uint64_t i = 0; i < cells; i++) {
        orc_num v;
        memset(&v, 0, sizeof v);
        v.f = 99;
        v.s = 99;
        v.u = (dtype == ORC_BOOL) ? 1 : 99;
        store_typed(dtype, grid, i, &v);
        switch (dtype2) {
        case ORC_F64: ((double *)grid_order)[i] = invert ? DBL_MIN : DBL_MAX; break;
        case ORC_F32: ((float *)grid_order)[i] = invert ? FLT_MIN : FLT_MAX; break;
        case ORC_I64: ((int64_t *)grid_order)[i] = invert ? INT64_MIN : INT64_MAX; break;
        case ORC_I32: ((int32_t *)grid_order)[i] = invert ? INT32_MIN : INT32_MAX; break;
        case ORC_I16: ((int16_t *)grid_order)[i] = invert ? INT16_MIN : INT16_MAX; break;
        case ORC_I8: ((int8_t *)grid_order)[i] = invert ? INT8_MIN : INT8_MAX; break;
        case ORC_U64: ((uint64_t *)grid_order)[i] = invert ? 0 : UINT64_MAX; break;
        case ORC_U32: ((uint32_t *)grid_order)[i] = invert ? 0 : UINT32_MAX; break;
        case ORC_U16: ((uint16_t *)grid_order)[i] = invert ? 0 : UINT16_MAX; break;
        case ORC_U8: ((uint8_t *)grid_order)[i] = invert ? 0 : UINT8_MAX; break;
        case ORC_BOOL: ((uint8_t *)grid_order)[i] = invert ? 0 : 1; break;
        default: return -1;
        }
        cell_masked[i] = 1;
    }
    return 0;
}

/* AggFirstPrimitive::aggregate — agg_first.cpp:115-165.
 *   mask test uses data_mask_ptr[j] WITHOUT the block offset (:131) — reference behaviour, kept.
 *   order value = data2 ? data2[offset+j] : (DataType2)(offset+j)  (:134; chunk-local row index)
 *   strict < (first) / > (last, invert) against the stored order; masked cells take the value directly. */
int orc_agg_first(int dtype, int dtype2, int flip, const void *data, const void *data2, const uint8_t *mask, void *grid, void *grid_order,
                  uint8_t *cell_masked, int invert, const uint64_t *idx, size_t length, uint64_t offset) {
    if (!data)
        return -2;
    for (size_t j = 0; j < length; j++) {
        if (mask && mask[j] != 1)
            continue;
        orc_num value, order, cur;
        memset(&value, 0, sizeof value);
        memset(&order, 0, sizeof order);
        memset(&cur, 0, sizeof cur);
        load_num(dtype, data, offset + j, flip, &value);
        if (data2) {
            load_num(dtype2, data2, offset + j, flip, &order);
        } else {
            /* DataType2 value_order = offset + j, then _to_native() if FlipEndian: materialise in dtype2 and reload */
            unsigned char tmp[8];
            orc_num r;
            memset(&r, 0, sizeof r);
            r.f = (double)(offset + j);
            r.s = (int64_t)(offset + j);
            r.u = (uint64_t)(offset + j);
            store_typed(dtype2, tmp, 0, &r);
            load_num(dtype2, tmp, 0, flip, &order);
        }
        if (num_isnan(&value) || num_isnan(&order))
            continue;
        uint64_t i = idx[j];
        if (cell_masked[i] == 1) {
            store_typed(dtype, grid, i, &value);
            cell_masked[i] = 0;
            store_typed(dtype2, grid_order, i, &order);
        } else {
            load_num(dtype2, grid_order, i, 0, &cur);
            int better = invert ? num_less(&cur, &order) : num_less(&order, &cur);
            if (better) {
                store_typed(dtype, grid, i, &value);
                cell_masked[i] = 0;
                store_typed(dtype2, grid_order, i, &order);
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Grid: strides (first binner fastest) + blocked driver loop — agg.hpp:63-73, :106-137
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t kind, dtype, flip, allow_other, invert, pad_;
    const void *data;
    const uint8_t *mask; /* 1 = masked */
    double vmin, vmax;
    uint64_t bins;
    int64_t ordinal_count, min_value;
} orc_binner;

typedef struct {
    int32_t op, dtype, dtype2, flip, invert, use_moment;
    uint32_t moment, pad_;
    const void *data;
    const void *data2;
    const uint8_t *mask; /* 1 = use */
    void *grid;
    void *grid_order;
    uint8_t *cell_masked;
} orc_agg;

uint64_t orc_grid_layout(const orc_binner *b, int nb, uint64_t *shapes, uint64_t *strides) {
    uint64_t length1d = 1;
    for (int i = 0; i < nb; i++) {
        shapes[i] = orc_binner_shape(b[i].kind, b[i].kind == ORC_BINNER_SCALAR ? b[i].bins : (uint64_t)b[i].ordinal_count, b[i].allow_other);
        length1d *= shapes[i];
    }
    if (nb > 0) {
        strides[0] = 1;
        for (int i = 1; i < nb; i++)
            strides[i] = strides[i - 1] * shapes[i - 1];
    }
    return length1d;
}

int orc_bin(const orc_binner *b, int nb, const orc_agg *a, int na, uint64_t length) {
    uint64_t shapes[16], strides[16];
    uint64_t idx[ORC_INDEX_BLOCK_SIZE];
    if (nb > 16)
        return -1;
    orc_grid_layout(b, nb, shapes, strides);
    uint64_t offset = 0;
    while (offset < length) {
        uint64_t n = length - offset < ORC_INDEX_BLOCK_SIZE ? length - offset : ORC_INDEX_BLOCK_SIZE;
        memset(idx, 0, n * sizeof(uint64_t));
        for (int i = 0; i < nb; i++) {
            int rc;
            if (b[i].kind == ORC_BINNER_SCALAR)
                rc = orc_scalar_to_bins(b[i].dtype, b[i].flip, b[i].data, b[i].mask, b[i].vmin, b[i].vmax, b[i].bins, offset, n, strides[i], idx);
            else
                rc = orc_ordinal_to_bins(b[i].dtype, b[i].flip, b[i].data, b[i].mask, b[i].ordinal_count, b[i].min_value, b[i].allow_other,
                                         b[i].invert, offset, n, strides[i], idx);
            if (rc)
                return rc;
        }
        for (int k = 0; k < na; k++) {
            int rc;
            switch (a[k].op) {
            case ORC_COUNT: rc = orc_agg_count(a[k].dtype, a[k].flip, a[k].data, a[k].mask, (int64_t *)a[k].grid, idx, n, offset); break;
            case ORC_SUM: rc = orc_agg_sum(a[k].dtype, a[k].flip, a[k].data, a[k].mask, a[k].grid, idx, n, offset, 0, 0); break;
            case ORC_SUM_MOMENT: rc = orc_agg_sum(a[k].dtype, a[k].flip, a[k].data, a[k].mask, a[k].grid, idx, n, offset, a[k].moment, 1); break;
            case ORC_MIN: rc = orc_agg_minmax(a[k].dtype, a[k].flip, a[k].data, a[k].mask, a[k].grid, idx, n, offset, 0); break;
            case ORC_MAX: rc = orc_agg_minmax(a[k].dtype, a[k].flip, a[k].data, a[k].mask, a[k].grid, idx, n, offset, 1); break;
            case ORC_FIRST:
                rc = orc_agg_first(a[k].dtype, a[k].dtype2, a[k].flip, a[k].data, a[k].data2, a[k].mask, a[k].grid, a[k].grid_order,
                                   a[k].cell_masked, a[k].invert, idx, n, offset);
                break;
            default: rc = -1;
            }
            if (rc)
                return rc;
        }
        offset += n;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Hash functors — hash.hpp:40-45 (_hash64, splitmix64 finaliser), :50-152 (per-type key widening).
 *   int64/uint64/double: the 64 raw bits;  int32: sign-extended;  uint32: zero-extended;
 *   float: 32 raw bits zero-extended;  int8/int16/uint8/uint16/bool: std::hash identity (NO mixing).
 * ---------------------------------------------------------------------------------------------- */
uint64_t orc_hash64(uint64_t x) {
    x = (x ^ (x >> 30)) * UINT64_C(0xbf58476d1ce4e5b9);
    x = (x ^ (x >> 27)) * UINT64_C(0x94d049bb133111eb);
    x = x ^ (x >> 31);
    return x;
}

/* key -> canonical 64-bit pattern ("bits") used for storage/equality, and its hash */
static uint64_t key_bits(int dtype, const void *data, uint64_t i) {
    switch (dtype) {
    case ORC_F64:
    case ORC_I64:
    case ORC_U64: return ((const uint64_t *)data)[i];
    case ORC_F32:
    case ORC_U32: return (uint64_t)((const uint32_t *)data)[i];
    case ORC_I32: return (uint64_t)(int64_t)((const int32_t *)data)[i];
    case ORC_I16: return (uint64_t)(int64_t)((const int16_t *)data)[i];
    case ORC_I8: return (uint64_t)(int64_t)((const int8_t *)data)[i];
    case ORC_U16: return (uint64_t)((const uint16_t *)data)[i];
    default: return (uint64_t)((const uint8_t *)data)[i];
    }
}

uint64_t orc_hash_bits(int dtype, uint64_t bits) {
    switch (dtype) {
    case ORC_F64:
    case ORC_I64:
    case ORC_U64:
    case ORC_F32:
    case ORC_U32:
    case ORC_I32: return orc_hash64(bits);
    default: return bits; /* std::hash<int8/16/uint8/16/bool> is the identity in libstdc++ */
    }
}

static int key_isnan(int dtype, uint64_t bits) {
    if (dtype == ORC_F64) {
        double d;
        memcpy(&d, &bits, 8);
        return d != d;
    }
    if (dtype == ORC_F32) {
        float f;
        uint32_t b = (uint32_t)bits;
        memcpy(&f, &b, 4);
        return f != f;
    }
    return 0;
}

static void key_store(int dtype, void *out, uint64_t i, uint64_t bits) {
    switch (orc_itemsize_[dtype]) {
    case 8: ((uint64_t *)out)[i] = bits; break;
    case 4: ((uint32_t *)out)[i] = (uint32_t)bits; break;
    case 2: ((uint16_t *)out)[i] = (uint16_t)bits; break;
    default: ((uint8_t *)out)[i] = (uint8_t)bits; break;
    }
}

/* one shard: insertion-ordered open-addressing map bits -> ordinal (the reference uses tsl::hopscotch_map
 * purely as a container, hash.hpp:154-164; observable results do not depend on its internals) */
typedef struct {
    uint64_t *keys; /* insertion order */
    int64_t *vals;
    int64_t size, cap;
    int64_t *table; /* index into keys, -1 empty */
    int64_t tcap;   /* power of two */
} orc_shard;

static void shard_init(orc_shard *s) {
    memset(s, 0, sizeof *s);
    s->cap = 16;
    s->keys = (uint64_t *)malloc(sizeof(uint64_t) * s->cap);
    s->vals = (int64_t *)malloc(sizeof(int64_t) * s->cap);
    s->tcap = 32;
    s->table = (int64_t *)malloc(sizeof(int64_t) * s->tcap);
    for (int64_t i = 0; i < s->tcap; i++)
        s->table[i] = -1;
}
static void shard_free(orc_shard *s) {
    free(s->keys);
    free(s->vals);
    free(s->table);
}
static int64_t shard_find(const orc_shard *s, uint64_t bits) {
    uint64_t h = orc_hash64(bits) & (uint64_t)(s->tcap - 1);
    for (;;) {
        int64_t e = s->table[h];
        if (e < 0)
            return -1;
        if (s->keys[e] == bits)
            return e;
        h = (h + 1) & (uint64_t)(s->tcap - 1);
    }
}
static void shard_rehash(orc_shard *s) {
    int64_t ncap = s->tcap * 2;
    free(s->table);
    s->table = (int64_t *)malloc(sizeof(int64_t) * ncap);
    for (int64_t i = 0; i < ncap; i++)
        s->table[i] = -1;
    s->tcap = ncap;
    for (int64_t e = 0; e < s->size; e++) {
        uint64_t h = orc_hash64(s->keys[e]) & (uint64_t)(ncap - 1);
        while (s->table[h] >= 0)
            h = (h + 1) & (uint64_t)(ncap - 1);
        s->table[h] = e;
    }
}
static void shard_emplace(orc_shard *s, uint64_t bits, int64_t val) {
    if (s->size == s->cap) {
        s->cap *= 2;
        s->keys = (uint64_t *)realloc(s->keys, sizeof(uint64_t) * s->cap);
        s->vals = (int64_t *)realloc(s->vals, sizeof(int64_t) * s->cap);
    }
    s->keys[s->size] = bits;
    s->vals[s->size] = val;
    if ((s->size + 1) * 2 > s->tcap)
        shard_rehash(s);
    uint64_t h = orc_hash64(bits) & (uint64_t)(s->tcap - 1);
    while (s->table[h] >= 0)
        h = (h + 1) & (uint64_t)(s->tcap - 1);
    s->table[h] = s->size;
    s->size++;
}

/* ordered_set<T> — hash_primitives.hpp:437-725 on hash_base (:41-330) / hash_common (hash.hpp:234-372) */
typedef struct {
    int dtype;
    int nmaps;
    int64_t limit;
    orc_shard *maps;
    int64_t nan_count, null_count;
    int64_t nan_value, null_value;          /* init 0x7fffffff (:447) */
    int64_t ordinal_code_offset_null_nan;   /* :447 */
} orc_set;

orc_set *orc_set_create(int dtype, int nmaps, int64_t limit) {
    orc_set *s = (orc_set *)calloc(1, sizeof *s);
    s->dtype = dtype;
    s->nmaps = nmaps;
    s->limit = limit;
    s->maps = (orc_shard *)malloc(sizeof(orc_shard) * nmaps);
    for (int i = 0; i < nmaps; i++)
        shard_init(&s->maps[i]);
    s->nan_value = 0x7fffffff;
    s->null_value = 0x7fffffff;
    return s;
}
void orc_set_destroy(orc_set *s) {
    if (!s)
        return;
    for (int i = 0; i < s->nmaps; i++)
        shard_free(&s->maps[i]);
    free(s->maps);
    free(s);
}

/* hash_common::count — hash.hpp:321-335 */
int64_t orc_set_count(const orc_set *s) {
    int64_t c = 0;
    for (int i = 0; i < s->nmaps; i++) {
        c += s->maps[i].size;
        if (i == 0) {
            if (s->null_count)
                c++;
            if (s->nan_count)
                c++;
        }
    }
    return c;
}
/* hash_common::offsets — hash.hpp:337-353 */
void orc_set_offsets(const orc_set *s, int64_t *out) {
    int64_t offset = 0;
    for (int i = 0; i < s->nmaps; i++) {
        out[i] = offset;
        offset += s->maps[i].size;
        if (i == 0) {
            if (s->null_count)
                offset++;
            if (s->nan_count)
                offset++;
        }
    }
}
int64_t orc_set_nan_count(const orc_set *s) { return s->nan_count; }
int64_t orc_set_null_count(const orc_set *s) { return s->null_count; }
int64_t orc_set_nan_index(const orc_set *s) { return s->nan_value; }
int64_t orc_set_null_index(const orc_set *s) { return s->null_value; }

/* ordered_set::add_new :471-479, add_existing :481-484 */
static int64_t set_add(orc_set *s, int map_index, uint64_t bits) {
    orc_shard *m = &s->maps[map_index];
    int64_t e = shard_find(m, bits);
    if (e >= 0)
        return m->vals[e];
    int64_t code = m->size;
    if (map_index == 0)
        code += s->ordinal_code_offset_null_nan;
    shard_emplace(m, bits, code);
    return code;
}
/* update1_nan (:297-300) + add_nan (:454-461); update1_null (hash.hpp:262-265) + add_null (:462-468) */
static int64_t set_add_nan(orc_set *s) {
    s->nan_count++;
    if (s->nan_count == 1) {
        s->nan_value = s->maps[0].size + s->ordinal_code_offset_null_nan;
        s->ordinal_code_offset_null_nan++;
    }
    return s->nan_value;
}
static int64_t set_add_null(orc_set *s) {
    s->null_count++;
    if (s->null_count == 1) {
        s->null_value = s->maps[0].size + s->ordinal_code_offset_null_nan;
        s->ordinal_code_offset_null_nan++;
    }
    return s->null_value;
}

/* hash_base::_update — hash_primitives.hpp:98-295.
 *   keys are bucketed by hash % nmaps in row order, each bucket flushed in map order, THEN nulls/NaNs
 *   are applied to map 0 (null before nan when use_offsets, nan before null otherwise, :264-287).
 *   masks: 1 = null.  out_values/out_map_index (nullable) = return_values.
 *   `limit`: if count() >= limit before the flush the whole flush is skipped (:237-249). */
int orc_set_update(orc_set *s, const void *keys, const uint8_t *masks, int64_t n, int64_t start_index, int return_values, int64_t *out_values,
                   int16_t *out_map_index) {
    if (s->limit >= 0 && return_values)
        return -3;
    const int use_offsets = return_values || (start_index != -1);
    int is_integer = !(s->dtype == ORC_F64 || s->dtype == ORC_F32);
    int full = (s->limit >= 0 && orc_set_count(s) >= s->limit);
    /* flush map by map, rows in order within a map (== bucket order) */
    if (!full) {
        for (int m = 0; m < s->nmaps; m++) {
            for (int64_t i = 0; i < n; i++) {
                if (masks && masks[i])
                    continue;
                uint64_t bits = key_bits(s->dtype, keys, (uint64_t)i);
                if (!is_integer && key_isnan(s->dtype, bits))
                    continue;
                if ((int)(orc_hash_bits(s->dtype, bits) % (uint64_t)s->nmaps) != m)
                    continue;
                int64_t v = set_add(s, m, bits);
                if (return_values) {
                    out_values[i] = v;
                    out_map_index[i] = (int16_t)m;
                }
            }
        }
    }
    /* specials */
    for (int pass = 0; pass < 2; pass++) {
        int do_null = use_offsets ? (pass == 0) : (pass == 1);
        for (int64_t i = 0; i < n; i++) {
            int isnull = masks && masks[i];
            if (do_null) {
                if (!isnull)
                    continue;
                int64_t v = set_add_null(s);
                if (return_values) {
                    out_values[i] = v;
                    out_map_index[i] = 0;
                }
            } else {
                if (isnull || is_integer)
                    continue;
                uint64_t bits = key_bits(s->dtype, keys, (uint64_t)i);
                if (!key_isnan(s->dtype, bits))
                    continue;
                int64_t v = set_add_nan(s);
                if (return_values) {
                    out_values[i] = v;
                    out_map_index[i] = 0;
                }
            }
        }
    }
    return 0;
}

/* hash_base::key_array — hash_primitives.hpp:302-328 (index = ordinal + offsets[map]; NaN slot gets NaN,
 * null slot gets -1 cast to the key type) */
int orc_set_key_array(const orc_set *s, void *out) {
    int64_t *offsets = (int64_t *)malloc(sizeof(int64_t) * s->nmaps);
    orc_set_offsets(s, offsets);
    for (int m = 0; m < s->nmaps; m++)
        for (int64_t e = 0; e < s->maps[m].size; e++)
            key_store(s->dtype, out, (uint64_t)(s->maps[m].vals[e] + offsets[m]), s->maps[m].keys[e]);
    if (s->nan_count) {
        if (s->dtype == ORC_F64) {
            ((double *)out)[s->nan_value] = NAN;
        } else if (s->dtype == ORC_F32) {
            ((float *)out)[s->nan_value] = NAN;
        }
    }
    if (s->null_count) {
        if (s->dtype == ORC_F64)
            ((double *)out)[s->null_value] = -1.0;
        else if (s->dtype == ORC_F32)
            ((float *)out)[s->null_value] = -1.0f;
        else if (s->dtype == ORC_BOOL)
            ((uint8_t *)out)[s->null_value] = 1; /* bool(-1) */
        else
            key_store(s->dtype, out, (uint64_t)s->null_value, UINT64_MAX);
    }
    free(offsets);
    return 0;
}

/* ordered_set::_map_ordinal — hash_primitives.hpp:624-691: NaN -> nan_value (or -1 when no NaN seen),
 * missing -> -1, else ordinal + offsets[map].  Output widened to int64 here; the caller narrows by
 * length() (<2^7 int8, <2^15 int16, <2^31 int32, :611-622). */
int orc_set_map_ordinal(const orc_set *s, const void *keys, int64_t n, int64_t *out) {
    int64_t *offsets = (int64_t *)malloc(sizeof(int64_t) * s->nmaps);
    orc_set_offsets(s, offsets);
    for (int64_t i = 0; i < n; i++) {
        uint64_t bits = key_bits(s->dtype, keys, (uint64_t)i);
        if (key_isnan(s->dtype, bits)) {
            out[i] = s->nan_count > 0 ? s->nan_value : -1;
            continue;
        }
        int m = (int)(orc_hash_bits(s->dtype, bits) % (uint64_t)s->nmaps);
        int64_t e = shard_find(&s->maps[m], bits);
        out[i] = e < 0 ? -1 : s->maps[m].vals[e] + (s->nmaps == 1 ? 0 : offsets[m]);
    }
    free(offsets);
    return 0;
}

/* ordered_set::isin — hash_primitives.hpp:539-565 */
int orc_set_isin(const orc_set *s, const void *keys, int64_t n, uint8_t *out) {
    for (int64_t i = 0; i < n; i++) {
        uint64_t bits = key_bits(s->dtype, keys, (uint64_t)i);
        if (key_isnan(s->dtype, bits)) {
            out[i] = s->nan_count > 0;
            continue;
        }
        int m = (int)(orc_hash_bits(s->dtype, bits) % (uint64_t)s->nmaps);
        out[i] = shard_find(&s->maps[m], bits) >= 0;
    }
    return 0;
}

/* ordered_set::merge — hash_primitives.hpp:693-720 (new keys appended with ordinal = map.size(); other cleared) */
int orc_set_merge(orc_set *s, orc_set *other) {
    if (s->nmaps != other->nmaps)
        return -4;
    for (int m = 0; m < s->nmaps; m++) {
        for (int64_t e = 0; e < other->maps[m].size; e++) {
            uint64_t bits = other->maps[m].keys[e];
            if (shard_find(&s->maps[m], bits) < 0)
                shard_emplace(&s->maps[m], bits, s->maps[m].size);
        }
        shard_free(&other->maps[m]);
        shard_init(&other->maps[m]);
    }
    s->nan_count += other->nan_count;
    s->null_count += other->null_count;
    return 0;
}

/* ordered_set::create (flatten / unpickle) — hash_primitives.hpp:486-537: one map; key i gets ordinal i;
 * i == null_value is the null slot, NaN keys take the nan slot. Returns NULL on the reference's error cases. */
orc_set *orc_set_from_keys(int dtype, const void *keys, int64_t n, int64_t null_value, int64_t nan_count, int64_t null_count) {
    orc_set *s = orc_set_create(dtype, 1, -1);
    for (int64_t i = 0; i < n; i++) {
        uint64_t bits = key_bits(dtype, keys, (uint64_t)i);
        if (i == null_value)
            set_add_null(s);
        else if (key_isnan(dtype, bits))
            set_add_nan(s);
        else
            set_add(s, 0, bits);
    }
    int bad = 0;
    if ((nan_count == 0) != (s->nan_count == 0))
        bad = 1;
    if ((null_count == 0) != (s->null_count == 0))
        bad = 1;
    if (null_count != 0 && s->null_value != null_value)
        bad = 1;
    if (orc_set_count(s) != n)
        bad = 1;
    if (bad) {
        orc_set_destroy(s);
        return NULL;
    }
    s->null_count = null_count;
    s->nan_count = nan_count;
    return s;
}

/* ---- limits pre-pass: df.minmax -> TaskStatistic(OP_MIN_MAX) -> vaexfast.statisticNd ---------------
 * vaex/cpu.py:487-626 (TaskPartStatistic.process): rows masked in the column are dropped (:533-537); the column is cast with
 * as_flat_array to float64 when its dtype is float64 or int64 and to FLOAT32 for every other dtype (:519-531 — int32 / uint64 / ...
 * values are rounded to fp32 on the way, a quirk that is part of the observable result); byte-swapped input is read through
 * functor_double_to_native (src/vaexfast.cpp:1339-1358).  The kernel op_min_max (src/vaexfast.cpp:1089-1101) starts from
 * (+inf, -inf) (vaex/tasks.py StatOpMinMax.init) and keeps `value < min` / `value > max`: NaN never wins.
 * out[0] = min, out[1] = max as doubles; DataFrame.minmax casts them back to the column dtype (vaex/dataframe.py:1524-1528),
 * which the Python wrapper does. */
int orc_minmax(int dtype, int flip, const void *data, const uint8_t *mask, uint64_t length, double out[2]) {
    double lo = INFINITY, hi = -INFINITY;
#define BODY(T)                                                                                                                \
    for (uint64_t i = 0; i < length; i++) {                                                                                    \
        if (mask && mask[i])                                                                                                   \
            continue;                                                                                                          \
        T v = ((const T *)data)[i];                                                                                            \
        if (flip)                                                                                                              \
            flip_bytes(&v, sizeof(T));                                                                                         \
        double value;                                                                                                          \
        if (dtype == ORC_F64 || dtype == ORC_I64)                                                                              \
            value = (double)v; /* numpy astype(float64): round to nearest even */                                              \
        else                                                                                                                   \
            value = (double)(float)v; /* numpy astype(float32), then the kernel widens to double */                            \
        if (value < lo)                                                                                                        \
            lo = value;                                                                                                        \
        if (value > hi)                                                                                                        \
            hi = value;                                                                                                        \
    }
    FOR_DTYPE(dtype, BODY)
#undef BODY
    out[0] = lo;
    out[1] = hi;
    return 0;
}
