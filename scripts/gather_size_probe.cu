// gather_size_probe.cu — how does the random-probe rate of one B200 depend on the TABLE SIZE (L2 residency)?
// 14.7 M random 2-byte probes (the scan positions of config C2) into tables of 2 MB .. 256 MB, L2 warm (best of 5
// back-to-back launches) and L2 cold (a 512 MB memset between launches).  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_size_probe gather_size_probe.cu && ./gather_size_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int PER>
__global__ void __launch_bounds__(256) k_u16(const uint16_t *__restrict__ t, uint32_t mask, int64_t n, unsigned long long *out)
{
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * PER;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PER; base < n; base += stride) {
        uint32_t v[PER];
#pragma unroll
        for (int i = 0; i < PER; i++) v[i] = __ldg(&t[mix((uint32_t)(base + i) * 2654435761u + 12345u) & mask]);
#pragma unroll
        for (int i = 0; i < PER; i++) acc += v[i] & 1u;
    }
    if (acc == 0xFFFFFFFFFFFFull) *out = acc;
}

// Does a SERIAL trip through the load/store pipe in front of every burst of gathers cost throughput?  TRIPS = 0: the
// plain gather loop; 1: a shared-memory read (whose result the gather addresses depend on) in front of each burst;
// 2: a small global read (L1 hit) in front of that read; PIPE = 1: the same reads issued one round AHEAD, together with
// the previous burst (software pipelined), so that a round still makes one trip.
template <int TRIPS, int PIPE>
__global__ void __launch_bounds__(256) k_trips(const uint16_t *__restrict__ t, uint32_t mask, const uint32_t *__restrict__ small, int64_t n,
                                               unsigned long long *out)
{
    __shared__ uint32_t sm[256 * 11];
    for (int i = threadIdx.x; i < 256 * 11; i += 256) sm[i] = i * 2654435761u;
    __syncthreads();
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
    uint32_t salt_next = 0;
    int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (PIPE && TRIPS) {
        uint32_t d = TRIPS >= 2 ? __ldg(&small[(base >> 3) & 1023]) : 0u;
        salt_next = sm[((threadIdx.x * 11 + (d & 1u)) % (256 * 11))] & 0xFFu;
    }
    for (; base < n; base += stride) {
        uint32_t salt = 0;
        if (TRIPS && !PIPE) {
            uint32_t d = TRIPS >= 2 ? __ldg(&small[(base >> 3) & 1023]) : 0u;
            salt = sm[((threadIdx.x * 11 + (d & 1u)) % (256 * 11))] & 0xFFu;
        } else if (TRIPS) salt = salt_next;
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __ldg(&t[(mix((uint32_t)(base + i) * 2654435761u + 12345u) ^ salt) & mask]);
        if (PIPE && TRIPS) {       // next round's reads travel with this round's gathers
            uint32_t d = TRIPS >= 2 ? __ldg(&small[((base + stride) >> 3) & 1023]) : 0u;
            salt_next = sm[((threadIdx.x * 11 + (d & 1u)) % (256 * 11))] & 0xFFu;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) acc += v[i] & 1u;
    }
    if (acc == 0xFFFFFFFFFFFFull) *out = acc;
}

// Does the instruction stream AROUND the gathers cost throughput?  PRE dependent integer operations in front of a round's
// eight gathers (their addresses depend on the result, like the lookup words formed from the staged slice), POST
// dependent operations behind them (like the votes / queue / bookkeeping), per thread and round.
template <int PRE, int POST>
__global__ void __launch_bounds__(256) k_pad(const uint16_t *__restrict__ t, uint32_t mask, int64_t n, unsigned long long *out)
{
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; base < n; base += stride) {
        uint32_t x = (uint32_t)base;
#pragma unroll 16
        for (int i = 0; i < PRE; i++) x = x * 1664525u + 1013904223u;
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __ldg(&t[(mix((uint32_t)(base + i) * 2654435761u + 12345u) ^ (PRE ? (x & 0xFFu) : 0u)) & mask]);
        uint32_t y = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) y += v[i];
#pragma unroll 16
        for (int i = 0; i < POST; i++) y = y * 1664525u + 1013904223u;
        acc += y & 1u;
    }
    if (acc == 0xFFFFFFFFFFFFull) *out = acc;
}

// How much do DEPENDENT shared-memory round trips per round cost (each read's address comes from the previous read)?
// The load/store pipe they go through is the one the gathers' 32 wavefronts per instruction occupy.
template <int K>
__global__ void __launch_bounds__(256) k_lds(const uint16_t *__restrict__ t, uint32_t mask, int64_t n, unsigned long long *out)
{
    __shared__ uint32_t sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = (i * 2654435761u) >> 21;       // values in 0 .. 2047
    __syncthreads();
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; base < n; base += stride) {
        uint32_t x = (uint32_t)(base >> 3) & 2047u;
#pragma unroll
        for (int i = 0; i < K; i++) x = sm[x];
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __ldg(&t[(mix((uint32_t)(base + i) * 2654435761u + 12345u) ^ (x & 0xFFu)) & mask]);
#pragma unroll
        for (int i = 0; i < 8; i++) acc += v[i] & 1u;
    }
    if (acc == 0xFFFFFFFFFFFFull) *out = acc;
}

int main()
{
    const int64_t N = 14705888;
    unsigned long long *out; CK(cudaMalloc(&out, 8));
    uint8_t *flush; const size_t FL = 512u << 20; CK(cudaMalloc(&flush, FL));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const double clk = p.clockRate * 1e3;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mb : {2, 4, 8, 16, 32, 64, 128, 256}) {
        const size_t bytes = (size_t)mb << 20;
        uint16_t *t; CK(cudaMalloc(&t, bytes)); CK(cudaMemset(t, 0x5a, bytes));
        const uint32_t mask = (uint32_t)(bytes / 2 - 1);
        for (int bps : {4, 6}) {
            float warm = 1e9f, cold = 1e9f;
            for (int it = 0; it < 6; it++) {
                cudaEventRecord(a); k_u16<8><<<148 * bps, 256>>>(t, mask, N, out); cudaEventRecord(b);
                CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b);
                if (it) warm = ms < warm ? ms : warm;
            }
            for (int it = 0; it < 4; it++) {
                CK(cudaMemsetAsync(flush, it, FL));
                cudaEventRecord(a); k_u16<8><<<148 * bps, 256>>>(t, mask, N, out); cudaEventRecord(b);
                CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b);
                cold = ms < cold ? ms : cold;
            }
            printf("U16 table %3d MB blocks/SM=%d  warm %7.1f us (%.3f probes/clk/SM)   cold %7.1f us (%.3f)\n", mb, bps,
                   warm * 1e3, N / (warm * 1e-3) / clk / p.multiProcessorCount, cold * 1e3, N / (cold * 1e-3) / clk / p.multiProcessorCount);
        }
        CK(cudaFree(t));
    }
    {
        const size_t bytes = (size_t)32 << 20;
        uint16_t *t; CK(cudaMalloc(&t, bytes)); CK(cudaMemset(t, 0x5a, bytes));
        uint32_t *small; CK(cudaMalloc(&small, 4096)); CK(cudaMemset(small, 0, 4096));
        const uint32_t mask = (uint32_t)(bytes / 2 - 1);
        auto run = [&](const char *name, auto kern) {
            float best = 1e9f;
            for (int it = 0; it < 6; it++) {
                cudaEventRecord(a); kern(); cudaEventRecord(b);
                CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b);
                if (it) best = ms < best ? ms : best;
            }
            printf("%-60s %7.1f us (%.3f probes/clk/SM)\n", name, best * 1e3, N / (best * 1e-3) / clk / p.multiProcessorCount);
        };
        {   // the same table from the stream-ordered pool (cudaMallocAsync), as the engine allocates its query tables
            uint16_t *ta = nullptr;
            CK(cudaMallocAsync((void **)&ta, bytes, 0)); CK(cudaMemsetAsync(ta, 0x5a, bytes, 0)); CK(cudaDeviceSynchronize());
            run("plain gather loop, table from cudaMallocAsync, 4 blocks/SM", [&] { k_pad<0, 0><<<148 * 4, 256>>>(ta, mask, N, out); });
            // and with 200 MB of other pool allocations made first (fragmented pool, like a lane that holds several tables)
            void *other[8];
            for (int i = 0; i < 8; i++) CK(cudaMallocAsync(&other[i], (size_t)25 << 20, 0));
            uint16_t *tb = nullptr;
            CK(cudaMallocAsync((void **)&tb, bytes, 0)); CK(cudaMemsetAsync(tb, 0x5a, bytes, 0)); CK(cudaDeviceSynchronize());
            run("plain gather loop, pool table allocated behind 200 MB of others", [&] { k_pad<0, 0><<<148 * 4, 256>>>(tb, mask, N, out); });
        }
        run("dependent shared-memory reads per round:  0, 6 blocks/SM", [&] { k_lds<0><<<148 * 6, 256>>>(t, mask, N, out); });
        run("dependent shared-memory reads per round:  2", [&] { k_lds<2><<<148 * 6, 256>>>(t, mask, N, out); });
        run("dependent shared-memory reads per round:  4", [&] { k_lds<4><<<148 * 6, 256>>>(t, mask, N, out); });
        run("dependent shared-memory reads per round:  8", [&] { k_lds<8><<<148 * 6, 256>>>(t, mask, N, out); });
        run("dependent shared-memory reads per round: 16", [&] { k_lds<16><<<148 * 6, 256>>>(t, mask, N, out); });
        run("dependent shared-memory reads per round: 32", [&] { k_lds<32><<<148 * 6, 256>>>(t, mask, N, out); });
        run("pad   0 /   0 dependent ops around a round, 4 blocks/SM", [&] { k_pad<0, 0><<<148 * 4, 256>>>(t, mask, N, out); });
        run("pad 150 /   0", [&] { k_pad<150, 0><<<148 * 4, 256>>>(t, mask, N, out); });
        run("pad 150 / 400", [&] { k_pad<150, 400><<<148 * 4, 256>>>(t, mask, N, out); });
        run("pad 150 / 400, 6 blocks/SM", [&] { k_pad<150, 400><<<148 * 6, 256>>>(t, mask, N, out); });
        run("pad 300 / 800, 6 blocks/SM", [&] { k_pad<300, 800><<<148 * 6, 256>>>(t, mask, N, out); });
        run("pad 150 / 400, 2 blocks/SM", [&] { k_pad<150, 400><<<148 * 2, 256>>>(t, mask, N, out); });
        run("serial trips 0 (plain gather loop), 4 blocks/SM", [&] { k_trips<0, 0><<<148 * 4, 256>>>(t, mask, small, N, out); });
        run("serial trips 1 (LDS in front of each burst)", [&] { k_trips<1, 0><<<148 * 4, 256>>>(t, mask, small, N, out); });
        run("serial trips 2 (LDG L1-hit -> LDS -> burst)", [&] { k_trips<2, 0><<<148 * 4, 256>>>(t, mask, small, N, out); });
        run("trips 1, issued a round ahead (pipelined)", [&] { k_trips<1, 1><<<148 * 4, 256>>>(t, mask, small, N, out); });
        run("trips 2, issued a round ahead (pipelined)", [&] { k_trips<2, 1><<<148 * 4, 256>>>(t, mask, small, N, out); });
    }
    return 0;
}
