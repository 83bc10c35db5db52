// gather_probe.cu — micro-benchmark behind the scan kernel's design: how fast can one B200 answer
// N random membership probes into a table of TABLE_BYTES?  Variants: LDG width / cache hints,
// LDGSTS (cp.async) into shared memory, probes in flight per thread, occupancy, DSMEM pull across a
// 16-CTA cluster holding the table in distributed shared memory.  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu && ./gather_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int PER, int MODE>
__global__ void __launch_bounds__(256) k_ldg(const uint2 *__restrict__ t8, const uint32_t *__restrict__ t4, uint32_t mask8, uint32_t mask4,
                                             int64_t n, unsigned long long *out)
{
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * PER;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PER; base < n; base += stride) {
        uint32_t v[PER];
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const uint32_t idx = mix((uint32_t)(base + i) * 2654435761u + 12345u);
            if (MODE == 0) { const uint2 w = __ldg(&t8[idx & mask8]); v[i] = w.x ^ w.y; }
            else if (MODE == 1) v[i] = __ldg(&t4[idx & mask4]);
            else if (MODE == 2) { uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(t4 + (idx & mask4))); v[i] = r; }
            else if (MODE == 3) { uint32_t r; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(t4 + (idx & mask4))); v[i] = r; }
            else { uint32_t r; asm volatile("ld.global.nc.L2::256B.u32 %0, [%1];" : "=r"(r) : "l"(t4 + (idx & mask4))); v[i] = r; }
        }
#pragma unroll
        for (int i = 0; i < PER; i++) acc += v[i] & 1u;
    }
    if (acc == 0xFFFFFFFFFFFFull) *out = acc;
}

// LDGSTS: every thread keeps PER 4-byte cp.async copies in flight into its own smem slots
template <int PER>
__global__ void __launch_bounds__(256) k_cpasync(const uint32_t *__restrict__ t4, uint32_t mask4, int64_t n, unsigned long long *out)
{
    __shared__ uint32_t buf[256 * PER];
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * PER;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PER; base < n; base += stride) {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const uint32_t idx = mix((uint32_t)(base + i) * 2654435761u + 12345u);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&buf[i * 256 + threadIdx.x]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(t4 + (idx & mask4)) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int i = 0; i < PER; i++) acc += buf[i * 256 + threadIdx.x] & 1u;
    }
    if (acc == 0xFFFFFFFFFFFFull) *out = acc;
}

// DSMEM: a cluster of CS CTAs holds the whole table (TABLE_WORDS / CS words each); probes read remote smem
template <int PER, int CS>
__global__ void __launch_bounds__(1024) k_dsmem(const uint32_t *__restrict__ t4, uint32_t table_words, int64_t n, unsigned long long *out)
{
    extern __shared__ uint32_t slice[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    const uint32_t per = table_words / CS;
    for (uint32_t i = threadIdx.x; i < per; i += blockDim.x) slice[i] = t4[rank * per + i];
    cluster.sync();
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * PER;
    const uint32_t slice_addr = (uint32_t)__cvta_generic_to_shared(slice);
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PER; base < n; base += stride) {
        uint32_t v[PER];
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const uint32_t idx = mix((uint32_t)(base + i) * 2654435761u + 12345u) & (table_words - 1);
            const uint32_t owner = idx / per, off = idx % per;
            uint32_t ra, r;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(slice_addr + off * 4), "r"(owner));
            asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(r) : "r"(ra));
            v[i] = r;
        }
#pragma unroll
        for (int i = 0; i < PER; i++) acc += v[i] & 1u;
    }
    if (acc == 0xFFFFFFFFFFFFull) *out = acc;
    cluster.sync();
}

// texture path: the same probes through tex1Dfetch on a linear texture object (TEX pipe instead of the LSU pipe)
template <int PER, int WIDE>
__global__ void __launch_bounds__(256) k_tex(cudaTextureObject_t tex, uint32_t mask, int64_t n, unsigned long long *out)
{
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * PER;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PER; base < n; base += stride) {
        uint32_t v[PER];
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const uint32_t idx = mix((uint32_t)(base + i) * 2654435761u + 12345u);
            if (WIDE) { const uint2 w = tex1Dfetch<uint2>(tex, (int)(idx & mask)); v[i] = w.x ^ w.y; }
            else v[i] = tex1Dfetch<uint32_t>(tex, (int)(idx & mask));
        }
#pragma unroll
        for (int i = 0; i < PER; i++) acc += v[i] & 1u;
    }
    if (acc == 0xFFFFFFFFFFFFull) *out = acc;
}

template <typename F>
static float timeit(F f, int iters = 5)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int i = 0; i < iters; i++) {
        cudaEventRecord(a); f(); cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    return best;
}

int main()
{
    const int64_t N = 14705888;                 // scan positions of config C2
    const uint32_t W8 = 1u << 19, W4 = 1u << 19; // 4 MB of uint2, 2 MB of uint32
    uint2 *t8; uint32_t *t4; unsigned long long *out;
    CK(cudaMalloc(&t8, (size_t)W8 * 8)); CK(cudaMalloc(&t4, (size_t)W4 * 4)); CK(cudaMalloc(&out, 8));
    CK(cudaMemset(t8, 0x5a, (size_t)W8 * 8)); CK(cudaMemset(t4, 0xa5, (size_t)W4 * 4)); CK(cudaMemset(out, 0, 8));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const double clk = p.clockRate * 1e3;
    auto report = [&](const char *name, float ms) {
        printf("%-44s %8.1f us  %6.2f Gprobe/s  %.3f probes/clk/SM\n", name, ms * 1e3, N / ms / 1e6, N / (ms * 1e-3) / clk / p.multiProcessorCount);
    };
    for (int bps : {2, 4, 6, 8}) {
        const int grid = 148 * bps;
        char nm[96];
        snprintf(nm, sizeof nm, "LDG.64 4MB per=8 blocks/SM=%d", bps);
        report(nm, timeit([&] { k_ldg<8, 0><<<grid, 256>>>(t8, t4, W8 - 1, W4 - 1, N, out); }));
        snprintf(nm, sizeof nm, "LDG.32 2MB per=8 blocks/SM=%d", bps);
        report(nm, timeit([&] { k_ldg<8, 1><<<grid, 256>>>(t8, t4, W8 - 1, W4 - 1, N, out); }));
        snprintf(nm, sizeof nm, "LDG.32 2MB per=16 blocks/SM=%d", bps);
        report(nm, timeit([&] { k_ldg<16, 1><<<grid, 256>>>(t8, t4, W8 - 1, W4 - 1, N, out); }));
        snprintf(nm, sizeof nm, "LDG.32 no_allocate per=8 blocks/SM=%d", bps);
        report(nm, timeit([&] { k_ldg<8, 2><<<grid, 256>>>(t8, t4, W8 - 1, W4 - 1, N, out); }));
        snprintf(nm, sizeof nm, "LDG.32 .cg per=8 blocks/SM=%d", bps);
        report(nm, timeit([&] { k_ldg<8, 3><<<grid, 256>>>(t8, t4, W8 - 1, W4 - 1, N, out); }));
        snprintf(nm, sizeof nm, "LDG.32 L2::256B per=8 blocks/SM=%d", bps);
        report(nm, timeit([&] { k_ldg<8, 4><<<grid, 256>>>(t8, t4, W8 - 1, W4 - 1, N, out); }));
        snprintf(nm, sizeof nm, "LDGSTS 4B per=8 blocks/SM=%d", bps);
        report(nm, timeit([&] { k_cpasync<8><<<grid, 256>>>(t4, W4 - 1, N, out); }));
        snprintf(nm, sizeof nm, "LDGSTS 4B per=16 blocks/SM=%d", bps);
        report(nm, timeit([&] { k_cpasync<16><<<grid, 256>>>(t4, W4 - 1, N, out); }));
    }
    {   // texture objects over the same tables
        cudaTextureObject_t tex4 = 0, tex8 = 0;
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = t4; rd.res.linear.sizeInBytes = (size_t)W4 * 4;
        rd.res.linear.desc = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindUnsigned);
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&tex4, &rd, &td, nullptr));
        rd.res.linear.devPtr = t8; rd.res.linear.sizeInBytes = (size_t)W8 * 8;
        rd.res.linear.desc = cudaCreateChannelDesc(32, 32, 0, 0, cudaChannelFormatKindUnsigned);
        CK(cudaCreateTextureObject(&tex8, &rd, &td, nullptr));
        for (int bps : {2, 4, 8}) {
            const int grid = 148 * bps;
            char nm[96];
            snprintf(nm, sizeof nm, "TEX.32 2MB per=8 blocks/SM=%d", bps);
            report(nm, timeit([&] { k_tex<8, 0><<<grid, 256>>>(tex4, W4 - 1, N, out); }));
            snprintf(nm, sizeof nm, "TEX.32 2MB per=16 blocks/SM=%d", bps);
            report(nm, timeit([&] { k_tex<16, 0><<<grid, 256>>>(tex4, W4 - 1, N, out); }));
            snprintf(nm, sizeof nm, "TEX.64 4MB per=8 blocks/SM=%d", bps);
            report(nm, timeit([&] { k_tex<8, 1><<<grid, 256>>>(tex8, W8 - 1, N, out); }));
        }
    }
    {   // DSMEM, cluster of 16 x 128 KB
        constexpr int CS = 16;
        const size_t smem = (size_t)W4 * 4 / CS;
        auto kern = k_dsmem<8, CS>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
            int nclusters = 0;
            cfg.gridDim = dim3(CS * 8);
            e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
            printf("DSMEM: max active clusters of %d = %d (%s)\n", CS, nclusters, cudaGetErrorString(e));
            if (e == cudaSuccess && nclusters > 0) {
                cfg.gridDim = dim3(CS * nclusters);
                const uint32_t words = W4;
                float ms = timeit([&] { cudaLaunchKernelEx(&cfg, kern, (const uint32_t *)t4, words, N, out); });
                CK(cudaGetLastError());
                report("DSMEM pull cluster16 per=8 1024thr", ms);
            }
        } else printf("DSMEM setup failed: %s\n", cudaGetErrorString(e));
        cudaGetLastError();
    }
    {   // DSMEM, cluster of 8 x 227 KB would need a 1.8 MB table: use 8 x 128 KB = half the table (1 MB) to see the rate
        constexpr int CS = 8;
        const uint32_t words = W4 / 2;
        const size_t smem = (size_t)words * 4 / CS;
        auto kern = k_dsmem<8, CS>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
            int nclusters = 0;
            cfg.gridDim = dim3(CS * 8);
            e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
            printf("DSMEM: max active clusters of %d = %d (%s)\n", CS, nclusters, cudaGetErrorString(e));
            if (e == cudaSuccess && nclusters > 0) {
                cfg.gridDim = dim3(CS * nclusters);
                float ms = timeit([&] { cudaLaunchKernelEx(&cfg, kern, (const uint32_t *)t4, words, N, out); });
                CK(cudaGetLastError());
                report("DSMEM pull cluster8 (1MB) per=8 1024thr", ms);
            }
        }
        cudaGetLastError();
    }
    return 0;
}
