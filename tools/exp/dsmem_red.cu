// dsmem_red.cu -- developer micro-benchmark: throughput of red.shared::cluster.add.u32 to random words of the histograms
// of the CTAs of a cluster (the k = 9..10 alternative to multi-pass partitioning: one pass, k-mers forwarded to the SM
// that owns their partition).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/exp/dsmem_red.cu -o tools/exp/dsmem_red
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int CL>
__global__ void __launch_bounds__(1024, 1) k_red(int iters, int remote_frac16, unsigned long long *out, long long *clk) {
    extern __shared__ uint32_t hist[];   // 32768 words
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) hist[i] = 0;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    uint32_t base = (uint32_t)__cvta_generic_to_shared(hist);
    uint32_t rb[CL];
#pragma unroll
    for (int r = 0; r < CL; r++) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb[r]) : "r"(base), "r"(r));
    asm volatile("barrier.cluster.arrive.aligned; barrier.cluster.wait.aligned;" ::: "memory");
    uint32_t x = (blockIdx.x * 1024u + threadIdx.x) * 2654435761u + 12345u;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x = x * 1664525u + 1013904223u;
            const uint32_t w = (x >> 8) & 0x7FFFu;
            // owner: with probability remote_frac16/16 a uniformly chosen CTA of the cluster, else this CTA
            const uint32_t sel = (x >> 24) & 15u;
            uint32_t owner = rank;
            if (sel < (uint32_t)remote_frac16) owner = (x >> 28) % CL;
            uint32_t a = rb[0];
#pragma unroll
            for (int r = 1; r < CL; r++) a = owner == (uint32_t)r ? rb[r] : a;
            asm volatile("red.shared::cluster.add.u32 [%0], %1;" ::"r"(a + w * 4u), "r"(1u + ((x & 1u) << 16)) : "memory");
        }
    }
    const long long t1 = clock64();
    asm volatile("barrier.cluster.arrive.aligned; barrier.cluster.wait.aligned;" ::: "memory");
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) s += (hist[i] & 0xFFFFu) + (hist[i] >> 16);
    atomicAdd(out, s);
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int CL>
void run(int remote_frac16) {
    const int iters = 256, grid = 148 / CL * CL;
    unsigned long long *out; long long *clk;
    CK(cudaMalloc(&out, 8)); CK(cudaMemset(out, 0, 8)); CK(cudaMalloc(&clk, 8 * 148));
    auto kern = k_red<CL>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    if (CL > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 131072;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaMemset(out, 0, 8));
        CK(cudaLaunchKernelEx(&cfg, kern, iters, remote_frac16, out, clk));
        CK(cudaDeviceSynchronize());
    }
    long long h[148]; unsigned long long s;
    CK(cudaMemcpy(h, clk, 8 * grid, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&s, out, 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
    const double reds = 1024.0 * iters * 16;
    printf("cluster %2d  remote share %5.1f%%  %.2f REDs/clk/SM  (sum %llu of %.0f)\n", CL, 100.0 * remote_frac16 / 16 * (CL - 1) / CL, reds / mx, s, reds * grid);
}

int main() {
    run<1>(0);
    run<2>(16); run<4>(16); run<8>(16); run<16>(16);
    run<4>(0);
    return 0;
}
