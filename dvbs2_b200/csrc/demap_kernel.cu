// Soft demapper (log-MAP / max*) for QPSK, 8PSK, 16APSK + column-row deinterleaver + pow2 quantiser, sm_100a.
//
// Semantics: lib/aff3ct/include/Module/Modem/Generic/Modem_generic.hxx:100-110 (inv_sigma2 = 1/(2 sigma^2) per frame),
// Modem_generic_fast.hxx:72-138,193-225 (L_b = max*_{j: bit_b(j)=0} d_j - max*_{j: bit_b(j)=1} d_j with
// d_j = -|y - c_j|^2 inv_sigma2, points visited j = 0..2^bps-1), lib/aff3ct/include/Tools/Math/max.hxx:29-36
// (max*(a,b) = max(a,b) + log1p(exp(-|a-b|))); deinterleaver: Interleaver.hxx:228-258 with the LUT of
// Interleaver_core_column_row.cpp:59-87 in closed form (code_tables.h deinterleaved_pos); quantiser:
// lib/aff3ct/src/Module/Quantizer/Pow2/Quantizer_pow2.cpp:184-190.
// The reference's AVX body uses polynomial exp/log, so parity with it is to float rounding (tests: 1e-4 relative).
//
// Two output layouts: frame-major (the Modem::demodulate socket) and the LDPC kernel's lane-interleaved
// var[group][v][32] layout, written directly so LLRs never make a second trip through HBM in the fused chain.
// Both are pure streaming kernels (HBM-bound: 8 B/symbol in, 4 B/LLR out).
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.h"

namespace dvbs2
{
namespace
{
__device__ __forceinline__ float
max_star(float a, float b)
{
    return fmaxf(a, b) + log1pf(expf(-fabsf(a - b)));
}

// LLRs of one received symbol: out[b], b = 0..BPS-1
template<int BPS>
__device__ __forceinline__ void
demap_symbol(float yr, float yi, float inv, const float* __restrict__ cs, float* out)
{
    float d[1 << BPS];
#pragma unroll
    for (int j = 0; j < (1 << BPS); j++)
    {
        const float dr = yr - cs[2 * j], di = yi - cs[2 * j + 1];
        d[j] = -(dr * dr + di * di) * inv;
    }
#pragma unroll
    for (int b = 0; b < BPS; b++)
    {
        float L0 = -INFINITY, L1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < (1 << BPS); j++)
        {
            if (((j >> b) & 1) == 0)
                L0 = max_star(L0, d[j]);
            else
                L1 = max_star(L1, d[j]);
        }
        out[b] = L0 - L1;
    }
}

__device__ __forceinline__ int
natural_pos(int k, int b, int bps, int n_sym, int itl_cols, int itl_rev)
{
    if (itl_cols == 0) return k * bps + b;
    const int col = itl_rev ? (itl_cols - 1 - b) : b;
    return col * n_sym + k;
}

template<typename R>
__device__ __forceinline__ R
quantize(float y, float factor, float vmax)
{
    float v = roundf(factor * y); // std::round: half away from zero
    v = fminf(fmaxf(v, -vmax), vmax);
    return (R)v;
}
template<>
__device__ __forceinline__ float
quantize<float>(float y, float, float)
{
    return y;
}

template<int BPS>
__global__ void
demap_frames_kernel(const DemapLaunch a, float* __restrict__ llr, int deinterleave)
{
    __shared__ float cs[2 << BPS];
    if (threadIdx.x < (2 << BPS)) cs[threadIdx.x] = a.cstl[threadIdx.x];
    __syncthreads();
    const int f = blockIdx.x; // frames on grid.x: batches exceed the 65535 limit of grid.y
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k >= a.n_sym) return;
    const float sg = a.sigma[f];
    const float inv = 1.0f / (2 * sg * sg);
    const float2 y = reinterpret_cast<const float2*>(a.sym + (size_t)f * 2 * a.n_sym)[k];
    float out[BPS];
    demap_symbol<BPS>(y.x, y.y, inv, cs, out);
    float* o = llr + (size_t)f * N_LDPC;
#pragma unroll
    for (int b = 0; b < BPS; b++)
        o[deinterleave ? natural_pos(k, b, BPS, a.n_sym, a.itl_cols, a.itl_rev) : k * BPS + b] = out[b];
}

// warp = (subgroup of 32 frames, 16 consecutive symbols); lane = frame.  Symbols are staged through shared memory so the
// global reads are 128-byte rows of one frame and the global writes are 32-lane rows of one variable.
constexpr int DEMAP_WARPS = 4;
template<int BPS, typename R>
__global__ void __launch_bounds__(DEMAP_WARPS * 32) demap_to_var_kernel(const DemapLaunch a, R* __restrict__ var, int nf)
{
    __shared__ float cs[2 << BPS];
    __shared__ float tile[DEMAP_WARPS][32][33];
    if (threadIdx.x < (2 << BPS)) cs[threadIdx.x] = a.cstl[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y;
    const int k0 = (blockIdx.x * DEMAP_WARPS + warp) * 16;
    if (k0 >= a.n_sym) return;
    const int nfl = 2 * a.n_sym;
    for (int j = 0; j < 32; j++)
    {
        const int f = g * 32 + j, e = 2 * k0 + lane;
        tile[warp][j][lane] = (f < a.n_frames && e < nfl) ? a.sym[(size_t)f * nfl + e] : 0.f;
    }
    __syncwarp();
    const int frame = g * 32 + lane;
    const float sg = frame < a.n_frames ? a.sigma[frame] : 1.f;
    const float inv = 1.0f / (2 * sg * sg);
    const float factor = (float)(1 << a.quant_dec);
    const float vmax = (float)((1 << (a.quant_bits - 1)) - 1);
    // subgroup g = group g / nf, slot g % nf of the decoder's var[group][v][lane][nf] layout
    R* vg = var + ((size_t)(g / nf) * N_LDPC * 32 + lane) * nf + g % nf;
#pragma unroll 4
    for (int s = 0; s < 16; s++)
    {
        const int k = k0 + s;
        if (k >= a.n_sym) break;
        float out[BPS];
        demap_symbol<BPS>(tile[warp][lane][2 * s], tile[warp][lane][2 * s + 1], inv, cs, out);
#pragma unroll
        for (int b = 0; b < BPS; b++)
        {
            const int v = natural_pos(k, b, BPS, a.n_sym, a.itl_cols, a.itl_rev);
            vg[(size_t)v * 32 * nf] = frame < a.n_frames ? quantize<R>(out[b], factor, vmax) : (R)0;
        }
    }
}

__global__ void
deinterleave_kernel(const float* __restrict__ itl, float* __restrict__ nat, int bps, int n_sym, int itl_cols, int itl_rev)
{
    const int f = blockIdx.x;
    const int n = blockIdx.y * blockDim.x + threadIdx.x;
    if (n >= N_LDPC) return;
    const int k = n / bps, b = n % bps;
    nat[(size_t)f * N_LDPC + natural_pos(k, b, bps, n_sym, itl_cols, itl_rev)] = itl[(size_t)f * N_LDPC + n];
}

template<typename R>
__global__ void
quantize_kernel(const float* __restrict__ in, R* __restrict__ out, size_t n, float factor, float vmax)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = quantize<R>(in[i], factor, vmax);
}
} // namespace

cudaError_t
launch_demap_frames(const DemapLaunch& a, float* llr, int deinterleave, cudaStream_t s)
{
    const dim3 grid(a.n_frames, (a.n_sym + 255) / 256);
    switch (a.bps)
    {
        case 2: demap_frames_kernel<2><<<grid, 256, 0, s>>>(a, llr, deinterleave); break;
        case 3: demap_frames_kernel<3><<<grid, 256, 0, s>>>(a, llr, deinterleave); break;
        case 4: demap_frames_kernel<4><<<grid, 256, 0, s>>>(a, llr, deinterleave); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template<int BPS>
static cudaError_t
demap_to_var_bps(const DemapLaunch& a, void* var, int precision, int nf, cudaStream_t s)
{
    const int groups = (a.n_frames + 32 * nf - 1) / (32 * nf) * nf; // whole decoder groups: pad subgroups get zeros
    const dim3 grid((a.n_sym + 16 * DEMAP_WARPS - 1) / (16 * DEMAP_WARPS), groups);
    switch (precision)
    {
        case 0: demap_to_var_kernel<BPS, float><<<grid, DEMAP_WARPS * 32, 0, s>>>(a, (float*)var, nf); break;
        case 1: demap_to_var_kernel<BPS, int16_t><<<grid, DEMAP_WARPS * 32, 0, s>>>(a, (int16_t*)var, nf); break;
        case 2: demap_to_var_kernel<BPS, int8_t><<<grid, DEMAP_WARPS * 32, 0, s>>>(a, (int8_t*)var, nf); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t
launch_demap_to_var(const DemapLaunch& a, void* var, int precision, int nf, cudaStream_t s)
{
    switch (a.bps)
    {
        case 2: return demap_to_var_bps<2>(a, var, precision, nf, s);
        case 3: return demap_to_var_bps<3>(a, var, precision, nf, s);
        case 4: return demap_to_var_bps<4>(a, var, precision, nf, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t
launch_deinterleave(const float* itl, float* nat, int n_frames, int n_sym, int itl_cols, int itl_rev, cudaStream_t s)
{
    const int bps = N_LDPC / n_sym;
    const dim3 grid(n_frames, (N_LDPC + 255) / 256);
    deinterleave_kernel<<<grid, 256, 0, s>>>(itl, nat, bps, n_sym, itl_cols, itl_rev);
    return cudaGetLastError();
}

cudaError_t
launch_quantize(const float* in, void* out, int precision, size_t n, int dec, int bits, cudaStream_t s)
{
    const float factor = (float)(1 << dec), vmax = (float)((1 << (bits - 1)) - 1);
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    if (precision == 1)
        quantize_kernel<int16_t><<<blocks, 256, 0, s>>>(in, (int16_t*)out, n, factor, vmax);
    else if (precision == 2)
        quantize_kernel<int8_t><<<blocks, 256, 0, s>>>(in, (int8_t*)out, n, factor, vmax);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}
} // namespace dvbs2
