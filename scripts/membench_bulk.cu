// Microbenchmark: per-warp ring of shared-memory stages filled by cp.async.bulk (one elected lane issues NR bulk copies
// of CH bytes from pseudo-random places of a multi-GB buffer, completion on an mbarrier), consumed with LDS and written
// back with plain vector stores — the staging scheme of the LDPC kernel.  Build: nvcc -O3 -arch=sm_100a.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t
hash32(uint32_t x)
{
    x ^= x >> 16, x *= 0x7feb352du, x ^= x >> 15, x *= 0x846ca68bu, x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t
smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void
mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void
mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void
mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    do
    {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void
bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template<int NR, int S, int WARPS, typename T> // T = per-lane vector (8 B -> 256 B chunks, 16 B -> 512 B)
__global__ void __launch_bounds__(WARPS * 32) bulk_kernel(T* buf, uint32_t n_chunks, int iters, int write)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int CH = 32 * sizeof(T);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t warp = blockIdx.x * WARPS + w;
    unsigned char* ring = smem + (size_t)w * S * NR * CH;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * S * NR * CH) + w * S;
    if (lane == 0)
        for (int s = 0; s < S; s++)
            mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    auto issue = [&](int it) {
        const int s = it % S;
        if (lane == 0)
        {
            mbar_expect_tx(bars + s, NR * CH);
#pragma unroll
            for (int r = 0; r < NR; r++)
            {
                const uint32_t idx = hash32(warp * 7919u + (uint32_t)(it * NR + r) * 104729u) % n_chunks;
                bulk_g2s(ring + ((size_t)s * NR + r) * CH, buf + (size_t)idx * 32, CH, bars + s);
            }
        }
    };
    for (int it = 0; it < S - 1 && it < iters; it++)
        issue(it);
    for (int it = 0; it < iters; it++)
    {
        const int s = it % S;
        mbar_wait(bars + s, (it / S) & 1);
        T v[NR];
#pragma unroll
        for (int r = 0; r < NR; r++)
            v[r] = reinterpret_cast<const T*>(ring + ((size_t)s * NR + r) * CH)[lane];
        // all lanes have their values in registers before the slot is refilled
        unsigned acc = 0;
#pragma unroll
        for (int r = 0; r < NR; r++)
            acc ^= *reinterpret_cast<unsigned*>(&v[r]);
        if (__any_sync(0xffffffffu, acc == 0x12345678u)) buf[0] = v[0];
        if (it + S - 1 < iters) issue(it + S - 1);
        if (write)
        {
#pragma unroll
            for (int r = 0; r < NR; r++)
            {
                const uint32_t idx = hash32(warp * 7919u + (uint32_t)(it * NR + r) * 104729u) % n_chunks;
                __stcg(buf + (size_t)idx * 32 + lane, v[r]);
            }
        }
    }
}

template<int NR, int S, int WARPS, typename T>
void
run(void* buf, size_t bytes, int ctas_per_sm, int write)
{
    const uint32_t n_chunks = (uint32_t)(bytes / (32 * sizeof(T)));
    const int iters = 300;
    const size_t smem = (size_t)WARPS * S * NR * 32 * sizeof(T) + WARPS * S * 8;
    auto k = bulk_kernel<NR, S, WARPS, T>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, WARPS * 32, smem);
    const int blocks = 148 * (occ < ctas_per_sm ? occ : ctas_per_sm);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k<<<blocks, WARPS * 32, smem>>>((T*)buf, n_chunks, 50, write);
    cudaEventRecord(e0);
    k<<<blocks, WARPS * 32, smem>>>((T*)buf, n_chunks, iters, write);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double moved = (double)blocks * WARPS * iters * NR * 32 * sizeof(T) * (write ? 2 : 1);
    printf("bulk chunk %4zu B rows %2d stages %d warps/CTA %d CTAs/SM %d (occ %d) smem %6zu : %7.1f GB/s (%s) %.2f ms err=%d\n",
           32 * sizeof(T), NR, S, WARPS, blocks / 148, occ, smem, moved / ms / 1e6, write ? "read+write" : "read only", ms,
           (int)cudaGetLastError());
}

int
main()
{
    const size_t bytes = (size_t)6 << 30;
    void* buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    for (int wr = 0; wr < 2; wr++)
    {
        run<26, 2, 4, float2>(buf, bytes, 4, wr);
        run<26, 3, 4, float2>(buf, bytes, 2, wr);
        run<26, 3, 3, float2>(buf, bytes, 3, wr);
        run<26, 4, 4, float2>(buf, bytes, 2, wr);
        run<26, 4, 8, float2>(buf, bytes, 1, wr);
        run<26, 2, 4, float>(buf, bytes, 8, wr);
        run<26, 3, 4, float>(buf, bytes, 5, wr);
        run<26, 2, 4, float4>(buf, bytes, 2, wr);
        run<26, 3, 4, float4>(buf, bytes, 1, wr);
        run<10, 3, 4, float2>(buf, bytes, 7, wr);
        run<10, 4, 8, float2>(buf, bytes, 2, wr);
    }
    return 0;
}
