// microbench.cu — measures the B200 primitives the binby design rests on (run under gpurun; prints one line per test).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench.bin tools/microbench.cu
// Results are summarised in profiles/ and DESIGN.md ("what bounds the kernel").
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                          \
    do {                                                                                               \
        cudaError_t e = (x);                                                                           \
        if (e != cudaSuccess) {                                                                        \
            printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__);                          \
            exit(1);                                                                                   \
        }                                                                                              \
    } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

// K scattered REDs per thread into `cells` cells
template <typename T, int MODE> // MODE 0: RED (no return), 1: ATOM (return used)
__global__ void k_red(T *grid, uint64_t cells, int iters, uint64_t *sink) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = mix(tid + 1);
    T acc = 0;
    for (int i = 0; i < iters; i++) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        uint64_t idx = (s >> 20) % cells;
        if (MODE == 0)
            atomicAdd(grid + idx, (T)1);
        else
            acc += atomicAdd(grid + idx, (T)1);
    }
    if (MODE == 1 && acc == (T)123456789)
        sink[0] = 1;
}

// cheaper index (power-of-two cells) to make sure the generator is not the limiter
template <typename T>
__global__ void k_red_pow2(T *grid, uint64_t mask, int iters) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = (uint32_t)mix(tid + 1) | 1u;
    for (int i = 0; i < iters; i++) {
        s ^= s << 13;
        s ^= s >> 17;
        s ^= s << 5;
        atomicAdd(grid + (s & mask), (T)1);
    }
}

template <typename T>
__global__ void k_atoms(int cells, int iters, T *out) {
    extern __shared__ unsigned char sm[];
    T *h = reinterpret_cast<T *>(sm);
    for (int i = threadIdx.x; i < cells; i += blockDim.x)
        h[i] = 0;
    __syncthreads();
    uint32_t s = (uint32_t)mix((uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1) | 1u;
    for (int i = 0; i < iters; i++) {
        s ^= s << 13;
        s ^= s >> 17;
        s ^= s << 5;
        atomicAdd(h + (s % (uint32_t)cells), (T)1);
    }
    __syncthreads();
    if (threadIdx.x == 0)
        out[blockIdx.x] = h[0];
}

// streaming read (the HBM side of the binby kernel): 128-bit evict-first loads
__global__ void k_stream(const uint4 *p, uint64_t n16, uint64_t *sink) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t i = tid; i < n16; i += step) {
        uint4 v = __ldcs(p + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u)
        sink[0] = acc;
}

// random 16-byte gathers (the probe side of the hash path)
__global__ void k_gather(const ulonglong2 *table, uint64_t mask, int iters, uint64_t *sink) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = (uint32_t)mix(tid + 1) | 1u;
    uint64_t acc = 0;
    for (int i = 0; i < iters; i++) {
        s ^= s << 13;
        s ^= s >> 17;
        s ^= s << 5;
        ulonglong2 v = __ldg(table + (((uint64_t)s * 2654435761ULL) & mask));
        acc += v.x ^ v.y;
    }
    if (acc == 0x12345678u)
        sink[0] = acc;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best)
            best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock %d kHz l2 %d MB\n", prop.name, sms, prop.clockRate, prop.l2CacheSize >> 20);
    uint64_t *sink;
    CK(cudaMalloc(&sink, 64));
    const int threads = 256, blocks = sms * 8, iters = 2048;
    const double lanes = (double)threads * blocks * iters;

    uint64_t sizes[] = {131, 1054729, 17373979, 69495916};
    for (uint64_t cells : sizes) {
        unsigned long long *g64;
        CK(cudaMalloc(&g64, cells * 8));
        CK(cudaMemset(g64, 0, cells * 8));
        float ms = time_ms([&] { k_red<unsigned long long, 0><<<blocks, threads>>>(g64, cells, iters, sink); });
        printf("RED.ADD.64   cells %10llu (%7.1f MB): %8.3f ms  %.3e lanes/s  %.2f lanes/clk/SM@1.9GHz\n", (unsigned long long)cells, cells * 8 / 1e6, ms,
               lanes / ms * 1e3, lanes / ms * 1e3 / sms / 1.9e9);
        ms = time_ms([&] { k_red<unsigned long long, 1><<<blocks, threads>>>(g64, cells, iters, sink); });
        printf("ATOM.ADD.64  cells %10llu (%7.1f MB): %8.3f ms  %.3e lanes/s\n", (unsigned long long)cells, cells * 8 / 1e6, ms, lanes / ms * 1e3);
        ms = time_ms([&] { k_red<double, 0><<<blocks, threads>>>((double *)g64, cells, iters, sink); });
        printf("RED.ADD.F64  cells %10llu (%7.1f MB): %8.3f ms  %.3e lanes/s\n", (unsigned long long)cells, cells * 8 / 1e6, ms, lanes / ms * 1e3);
        ms = time_ms([&] { k_red<unsigned, 0><<<blocks, threads>>>((unsigned *)g64, cells, iters, sink); });
        printf("RED.ADD.32   cells %10llu (%7.1f MB): %8.3f ms  %.3e lanes/s\n", (unsigned long long)cells, cells * 4 / 1e6, ms, lanes / ms * 1e3);
        CK(cudaFree(g64));
    }
    {   // is the RED rate an SM-side or an L2-side limit?  shrink the number of SMs issuing
        uint64_t cells = 1 << 20;
        unsigned long long *g64;
        CK(cudaMalloc(&g64, cells * 8));
        for (int nsm : {37, 74, 111, 148}) {
            int b = nsm * 8; // 8 CTAs x 256 thr = one full SM each; the block scheduler spreads CTAs one per SM first
            double l = (double)threads * b * iters;
            float ms = time_ms([&] { k_red_pow2<unsigned long long><<<b, threads>>>(g64, cells - 1, iters); });
            printf("RED.ADD.64 pow2 1M cells, %4d CTAs (~%d SMs x8): %8.3f ms  %.3e lanes/s\n", b, nsm, ms, l / ms * 1e3);
        }
        for (int nsm : {37, 74, 148}) {
            int b = nsm; // one CTA per SM -> all SMs busy but 1/8 of the warps
            double l = (double)threads * b * iters;
            float ms = time_ms([&] { k_red_pow2<unsigned long long><<<b, threads>>>(g64, cells - 1, iters); });
            printf("RED.ADD.64 pow2 1M cells, %4d CTAs (1 CTA/SM):   %8.3f ms  %.3e lanes/s\n", b, ms, l / ms * 1e3);
        }
        for (int occ : {2, 4, 8}) {
            int b = sms * occ;
            double l = (double)threads * b * iters;
            float ms = time_ms([&] { k_red_pow2<unsigned long long><<<b, threads>>>(g64, cells - 1, iters); });
            printf("RED.ADD.64 pow2 1M cells, %d CTAs/SM: %8.3f ms  %.3e lanes/s\n", occ, ms, l / ms * 1e3);
            ms = time_ms([&] { k_red_pow2<unsigned><<<b, threads>>>((unsigned *)g64, cells - 1, iters); });
            printf("RED.ADD.32 pow2 1M cells, %d CTAs/SM: %8.3f ms  %.3e lanes/s\n", occ, ms, l / ms * 1e3);
        }
        CK(cudaFree(g64));
    }
    {
        unsigned *out;
        CK(cudaMalloc(&out, blocks * 8));
        CK(cudaFuncSetAttribute(k_atoms<unsigned>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int cells : {131, 4096, 16384, 49152}) {
            int b2 = cells * 4 > 100 * 1024 ? sms : (cells * 4 > 48 * 1024 ? sms * 2 : blocks);
            double lanes = (double)threads * b2 * iters;
            float ms = time_ms([&] { k_atoms<unsigned><<<b2, threads, cells * 4>>>(cells, iters, out); });
            printf("ATOMS.ADD.32 cells %6d: %8.3f ms  %.3e lanes/s  %.2f lanes/clk/SM@1.9GHz\n", cells, ms, lanes / ms * 1e3, lanes / ms * 1e3 / sms / 1.9e9);
        }
        CK(cudaFuncSetAttribute(k_atoms<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        for (int cells : {131, 4096}) {
            float ms = time_ms([&] { k_atoms<unsigned long long><<<blocks, threads, cells * 8>>>(cells, iters, (unsigned long long *)out); });
            printf("ATOMS.ADD.64 cells %6d: %8.3f ms  %.3e lanes/s\n", cells, ms, lanes / ms * 1e3);
        }
        CK(cudaFree(out));
    }
    {
        uint64_t bytes = 8ull << 30;
        uint4 *p;
        CK(cudaMalloc(&p, bytes));
        CK(cudaMemset(p, 1, bytes));
        for (int occ : {4, 8}) {
            float ms = time_ms([&] { k_stream<<<sms * occ, 256>>>(p, bytes / 16, sink); });
            printf("stream read 8 GiB ldcs.128, %d CTAs/SM: %8.3f ms  %.1f GB/s\n", occ, ms, bytes / ms / 1e6);
        }
        CK(cudaFree(p));
    }
    {
        for (uint64_t slots : {1ull << 21, 1ull << 25}) {
            ulonglong2 *t;
            CK(cudaMalloc(&t, slots * 16));
            CK(cudaMemset(t, 0, slots * 16));
            float ms = time_ms([&] { k_gather<<<blocks, threads>>>(t, slots - 1, iters, sink); });
            printf("gather 16 B random, table %6.1f MB: %8.3f ms  %.3e lanes/s\n", slots * 16 / 1e6, ms, lanes / ms * 1e3);
            CK(cudaFree(t));
        }
    }
    return 0;
}
