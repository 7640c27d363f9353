// microbench_dsmem.cu — how fast can a thread-block cluster scatter 32-bit increments into its DISTRIBUTED shared memory?
// (question behind it: a 1027^2 u32 count grid is 4.2 MB; a cluster of 16 CTAs owns 3.6 MB of shared memory, so a histogram
//  that lives in the cluster would replace one L2 RED per row (capped at ~98/clk chip-wide) by one remote shared-memory RED.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_dsmem tools/microbench_dsmem.cu && ./microbench_dsmem
//
// Every thread draws pseudo-random cell numbers in [0, cluster_size * cells_per_cta), maps the owner CTA's shared window with
// `mapa` and issues `red.shared::cluster.add.u32`.  Variants: cluster size 1..16, all-local targets, a spatially coherent
// stream (runs of equal owners), and the loop without the RED (index cost only).  The result is verified (sum of all
// counters == number of REDs issued).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e = (x);                                                                    \
        if (e != cudaSuccess) {                                                                 \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);      \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

constexpr int kCells = 48 * 1024; // u32 counters per CTA (192 KB)
constexpr int kThreads = 1024;

__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// mode 0: uniform random over the whole cluster; 1: always the own CTA (local ATOMS through the cluster window);
// 2: no RED at all (loop + index cost); 3: random, but only every 8th draw changes the owner CTA
__global__ void __launch_bounds__(kThreads, 1) k_dsmem(unsigned long long *total, int iters, int mode) {
    extern __shared__ __align__(16) unsigned hist[];
    for (int i = threadIdx.x; i < kCells; i += kThreads)
        hist[i] = 0;
    cluster_barrier();
    const unsigned csize = cluster_size(), me = cluster_rank();
    const unsigned base = (unsigned)__cvta_generic_to_shared(hist);
    unsigned s = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 12345u;
    unsigned acc = 0, owner = me;
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
        s = s * 1664525u + 1013904223u;
        const unsigned cell = (s >> 8) % kCells;
        if (mode == 0)
            owner = (s >> 3) % csize;
        else if (mode == 3) {
            if ((it & 7) == 0)
                owner = (s >> 3) % csize;
        }
        if (mode == 2) {
            acc += cell + owner;
            continue;
        }
        unsigned raddr;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(base + cell * 4u), "r"(owner));
        asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(raddr), "r"(1u) : "memory");
    }
    cluster_barrier();
    unsigned long long sum = acc & 1u ? 0ull : 0ull;
    for (int i = threadIdx.x; i < kCells; i += kThreads)
        sum += hist[i];
    atomicAdd(total, sum);
    if (mode == 2 && acc == 0xdeadbeefu)
        atomicAdd(total, 1ull);
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("# %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    unsigned long long *total;
    CK(cudaMalloc(&total, 8));
    const size_t smem = (size_t)kCells * 4;
    CK(cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const int iters = 8192;
    printf("%-8s %-6s %-9s %-8s %-12s %-14s %-10s\n", "cluster", "mode", "clusters", "CTAs", "ms", "REDs/s", "REDs/clk/SM");
    for (int csize : {1, 2, 4, 8, 16}) {
        for (int mode : {0, 1, 3, 2}) {
            if (csize == 1 && (mode == 0 || mode == 3))
                continue;
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = csize;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            cfg.gridDim = dim3(csize);
            int nclusters = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_dsmem, &cfg);
            if (e != cudaSuccess || nclusters < 1) {
                printf("%-8d %-6d not launchable (%s)\n", csize, mode, cudaGetErrorString(e));
                cudaGetLastError();
                continue;
            }
            cfg.gridDim = dim3(nclusters * csize);
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            float best = 1e30f;
            unsigned long long got = 0;
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaMemset(total, 0, 8));
                CK(cudaEventRecord(a));
                CK(cudaLaunchKernelEx(&cfg, k_dsmem, total, iters, mode));
                CK(cudaEventRecord(b));
                CK(cudaEventSynchronize(b));
                float ms;
                CK(cudaEventElapsedTime(&ms, a, b));
                best = ms < best ? ms : best;
                CK(cudaMemcpy(&got, total, 8, cudaMemcpyDeviceToHost));
            }
            const double n = (double)nclusters * csize * kThreads * iters;
            const bool ok = mode == 2 ? true : got == (unsigned long long)n;
            printf("%-8d %-6d %-9d %-8d %-12.3f %-14.3e %-10.2f %s\n", csize, mode, nclusters, nclusters * csize, best, n / (best * 1e-3),
                   n / (best * 1e-3) / (prop.clockRate * 1e3) / (nclusters * csize), ok ? "" : "SUM MISMATCH");
        }
    }
    return 0;
}
