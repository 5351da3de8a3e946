// Random-chunk read-modify-write microbenchmark: what HBM3e delivers for the LDPC kernel's access pattern
// (warp-wide contiguous chunks of 128/256/512 B at pseudo-random places of a multi-GB buffer, each read then
// written back), as a function of chunk size and loads in flight per warp.  Build: nvcc -O3 -arch=sm_100a.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t
hash32(uint32_t x)
{
    x ^= x >> 16, x *= 0x7feb352du, x ^= x >> 15, x *= 0x846ca68bu, x ^= x >> 16;
    return x;
}

template<typename T, int BATCH, bool WRITE>
__global__ void
rmw_kernel(T* buf, uint32_t n_chunks, int iters)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    T acc[BATCH];
    for (int it = 0; it < iters; it++)
    {
        uint32_t idx[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; b++)
        {
            idx[b] = hash32(warp * 7919u + (uint32_t)(it * BATCH + b) * 104729u) % n_chunks;
            acc[b] = __ldcg(buf + (size_t)idx[b] * 32 + lane);
        }
#pragma unroll
        for (int b = 0; b < BATCH; b++)
        {
            T v = acc[b];
            if (WRITE)
                __stcg(buf + (size_t)idx[b] * 32 + lane, v);
            else if (*(const int*)&v == 0x7fffffff)
                buf[0] = v;
        }
    }
}

template<typename T, int BATCH, bool WRITE>
void
run(const char* name, void* buf, size_t bytes, int warps_per_sm)
{
    const uint32_t n_chunks = (uint32_t)(bytes / (32 * sizeof(T)));
    const int iters = 4000 / BATCH;
    const int blocks = 148 * warps_per_sm / 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    rmw_kernel<T, BATCH, WRITE><<<blocks, 128>>>((T*)buf, n_chunks, iters / 4);
    cudaEventRecord(e0);
    rmw_kernel<T, BATCH, WRITE><<<blocks, 128>>>((T*)buf, n_chunks, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double moved = (double)blocks * 4 * iters * BATCH * 32 * sizeof(T) * (WRITE ? 2 : 1);
    printf("%-28s chunk %4zu B batch %2d warps/SM %2d : %7.1f GB/s (%s)  %.2f ms  err=%d\n", name, 32 * sizeof(T), BATCH, warps_per_sm,
           moved / ms / 1e6, WRITE ? "read+write" : "read only", ms, (int)cudaGetLastError());
}

int
main()
{
    const size_t bytes = (size_t)6 << 30;
    void* buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    for (int w : { 16, 32, 48 })
    {
        run<float, 8, true>("f32  rmw", buf, bytes, w);
        run<float, 26, true>("f32  rmw", buf, bytes, w);
        run<float2, 8, true>("f32x2 rmw", buf, bytes, w);
        run<float2, 26, true>("f32x2 rmw", buf, bytes, w);
        run<float4, 8, true>("f32x4 rmw", buf, bytes, w);
        run<float4, 26, true>("f32x4 rmw", buf, bytes, w);
        run<float, 26, false>("f32  read", buf, bytes, w);
        run<float2, 26, false>("f32x2 read", buf, bytes, w);
        run<float4, 26, false>("f32x4 read", buf, bytes, w);
        run<short, 26, true>("i16  rmw", buf, bytes, w);
        run<int, 26, true>("i16x2 rmw", buf, bytes, w);
    }
    return 0;
}
