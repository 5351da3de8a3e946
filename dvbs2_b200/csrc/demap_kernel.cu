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
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "kernels.h"

namespace dvbs2
{
namespace
{
// max*(a,b) = max(a,b) + log(1 + exp(-|a-b|)) (max.hxx:29-36).  The correction term lies in (0, ln 2]; it is evaluated
// with the SFU approximations (ex2.approx / lg2.approx, absolute error below 1e-6 here), as the reference's own SIMD body
// evaluates it with MIPP's polynomial exp/log rather than libm (Modem_generic_fast.hxx:72-138).  The precise
// expf/log1pf pair made this kernel ALU-bound at 4x its streaming time.
__device__ __forceinline__ float
max_star(float a, float b)
{
    return fmaxf(a, b) + __logf(1.0f + __expf(-fabsf(a - b)));
}

// Component i of the symbol buffer in its wire format (compact input sockets).  The integer formats are
// converted exactly as a host-side adaptor would before handing floats to Modem::demodulate — (float)x * scale, one rounded
// multiply that is never fused into the distance computation — so transporting sc16 / sc8 / half instead of float changes
// no bit of the result for the same sample values.
// FMT is a compile-time parameter: a run-time switch inside the staging loops kept the compiler from issuing the loads of a
// tile back to back (measured: the float demapper went from 9 to 29 ms per 341k-frame batch).
template<int FMT>
__device__ __forceinline__ float
load_sym(const DemapLaunch& a, size_t i)
{
    if constexpr (FMT == 1)
        return __half2float(__ldg(reinterpret_cast<const __half*>(a.sym) + i));
    else if constexpr (FMT == 2)
        return __fmul_rn((float)__ldg(reinterpret_cast<const int16_t*>(a.sym) + i), a.iq_scale);
    else if constexpr (FMT == 3)
        return __fmul_rn((float)__ldg(reinterpret_cast<const int8_t*>(a.sym) + i), a.iq_scale);
    else if constexpr (FMT == 4)
    { // component i of symbol i / 2: three bytes per symbol, I in bits 0..11 and Q in bits 12..23 of the little-endian word
        const uint8_t* p = reinterpret_cast<const uint8_t*>(a.sym) + (i >> 1) * 3;
        const uint32_t b1 = __ldg(p + 1);
        const uint32_t v = (i & 1) ? ((b1 >> 4) | ((uint32_t)__ldg(p + 2) << 4)) : ((uint32_t)__ldg(p) | ((b1 & 0xfu) << 8));
        return __fmul_rn((float)(((int)(v << 20)) >> 20), a.iq_scale);
    }
    else
        return __ldg(a.sym + i);
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
    if constexpr (BPS >= 3)
    {
        // 8PSK / 16APSK: the max* chain over a class IS log(sum_j exp(d_j)) (max.hxx:29-36 is the exact Jacobian logarithm), so
        // all classes can share one exponential per constellation point: e_j = exp(d_j - max_j d_j), L_b = log(sum of the
        // class-0 e_j) - log(sum of the class-1 e_j) — 2^BPS + 2 BPS SFU operations per symbol instead of 4 (2^BPS - 2) BPS / 2
        // (14 instead of 36 for 8PSK, 24 instead of 112 for 16APSK; the chain made this kernel SFU-bound at 4x its streaming
        // time).  Same value up to float rounding.  Only when a whole class lies more than ~69 nats below the best point
        // would its sum leave the normal float range: those symbols (noise outliers at high SNR) take the chain below.
        float m = d[0];
#pragma unroll
        for (int j = 1; j < (1 << BPS); j++)
            m = fmaxf(m, d[j]);
        float e[1 << BPS];
#pragma unroll
        for (int j = 0; j < (1 << BPS); j++)
            e[j] = __expf(d[j] - m);
        bool ok = true;
        float fast[BPS];
#pragma unroll
        for (int b = 0; b < BPS; b++)
        {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < (1 << BPS); j++)
                if (((j >> b) & 1) == 0)
                    s0 += e[j];
                else
                    s1 += e[j];
            ok = ok && s0 > 1e-30f && s1 > 1e-30f;
            fast[b] = __logf(s0) - __logf(s1);
        }
        if (ok)
        {
#pragma unroll
            for (int b = 0; b < BPS; b++)
                out[b] = fast[b];
            return;
        }
    }
#pragma unroll
    for (int b = 0; b < BPS; b++)
    {
        // max*(-inf, x) = x: the first point of each class starts the accumulation (resolved at compile time)
        float L0 = 0.f, L1 = 0.f;
        bool h0 = false, h1 = false;
#pragma unroll
        for (int j = 0; j < (1 << BPS); j++)
        {
            if (((j >> b) & 1) == 0)
            {
                L0 = h0 ? max_star(L0, d[j]) : d[j];
                h0 = true;
            }
            else
            {
                L1 = h1 ? max_star(L1, d[j]) : d[j];
                h1 = true;
            }
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

template<int BPS, int FMT>
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
    const size_t at = a.pl_sym > 0 ? (size_t)f * 2 * a.pl_sym + 2 * pl_symbol_pos(k, a.n_sym) : (size_t)f * 2 * a.n_sym + 2 * k;
    const float2 y = make_float2(load_sym<FMT>(a, at), load_sym<FMT>(a, at + 1));
    float out[BPS];
    demap_symbol<BPS>(y.x, y.y, inv, cs, out);
    float* o = llr + (size_t)f * N_LDPC;
#pragma unroll
    for (int b = 0; b < BPS; b++)
        o[deinterleave ? natural_pos(k, b, BPS, a.n_sym, a.itl_cols, a.itl_rev) : k * BPS + b] = out[b];
}

// NF values of R as one 4- or 8-byte store
template<typename R, int NF>
struct LaneVec;
template<>
struct LaneVec<float, 1>
{
    using T = float;
    __device__ static T make(const float* x) { return x[0]; }
};
template<>
struct LaneVec<float, 2>
{
    using T = float2;
    __device__ static T make(const float* x) { return make_float2(x[0], x[1]); }
};
template<>
struct LaneVec<int16_t, 1> // one frame per lane: the cooperative small-batch LDPC kernel
{
    using T = int16_t;
    __device__ static T make(const int16_t* x) { return x[0]; }
};
template<>
struct LaneVec<int16_t, 4>
{
    using T = uint2;
    __device__ static T make(const int16_t* x)
    {
        return make_uint2((uint32_t)(uint16_t)x[0] | ((uint32_t)(uint16_t)x[1] << 16),
                          (uint32_t)(uint16_t)x[2] | ((uint32_t)(uint16_t)x[3] << 16));
    }
};
template<>
struct LaneVec<int8_t, 4>
{
    using T = uint32_t;
    __device__ static T make(const int8_t* x)
    {
        return (uint32_t)(uint8_t)x[0] | ((uint32_t)(uint8_t)x[1] << 8) | ((uint32_t)(uint8_t)x[2] << 16) | ((uint32_t)(uint8_t)x[3] << 24);
    }
};

// warp = (decoder group of 32*NF frames, SYMS consecutive symbols); lane l demaps frames l, 32+l, ... of the group.
// Symbols are staged through shared memory so the global reads are 128-byte (SYMS = 16) or 64-byte (SYMS = 8, two frames per
// warp load) rows of one frame and every global write is one full row var[group][v][0..31][0..NF-1] of the decoder's layout
// (256 B for float and int16).  Four-frame lanes use the shorter tile so that as many warps fit on an SM (shared memory).
template<int NF>
struct DemapCfg
{
    static constexpr int SYMS = NF == 4 ? 8 : 16;        // symbols per warp tile
    static constexpr int WARPS = NF == 1 ? 8 : 4;        // <= 34.8 KB of shared memory per block
};
template<int BPS, typename R, int NF, int FMT>
__global__ void __launch_bounds__(DemapCfg<NF>::WARPS * 32) demap_to_var_kernel(const DemapLaunch a, R* __restrict__ var)
{
    constexpr int WARPS = DemapCfg<NF>::WARPS, SYMS = DemapCfg<NF>::SYMS;
    constexpr int RPL = 16 / SYMS; // frame rows per warp load: 1 (32 floats per row) or 2 (16 floats per row)
    using Vec = typename LaneVec<R, NF>::T;
    __shared__ float cs[2 << BPS];
    __shared__ float tile[WARPS][NF][32][2 * SYMS + 1];
    if (threadIdx.x < (2 << BPS)) cs[threadIdx.x] = a.cstl[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y; // decoder group
    const int k0 = (blockIdx.x * WARPS + warp) * SYMS;
    if (k0 >= a.n_sym) return;
    // row length and position of symbol k0 in it; a tile never straddles a pilot block (1440 = 90 * 16)
    const int nfl = a.pl_sym > 0 ? 2 * a.pl_sym : 2 * a.n_sym;
    const int col = lane % (2 * SYMS), sub = lane / (2 * SYMS);
    const int e = (a.pl_sym > 0 ? 2 * pl_symbol_pos(k0, a.n_sym) : 2 * k0) + col;
    const bool in_frame = 2 * k0 + col < 2 * a.n_sym;
#pragma unroll
    for (int j = 0; j < NF; j++)
#pragma unroll 16
        for (int r0 = 0; r0 < 32; r0 += RPL)
        {
            const int r = r0 + sub;
            const int f = (g * NF + j) * 32 + r;
            tile[warp][j][r][col] = (f < a.n_frames && in_frame) ? load_sym<FMT>(a, (size_t)f * nfl + e) : 0.f;
        }
    __syncwarp();
    float inv[NF];
    bool live[NF];
#pragma unroll
    for (int j = 0; j < NF; j++)
    {
        const int frame = (g * NF + j) * 32 + lane;
        live[j] = frame < a.n_frames;
        const float sg = live[j] ? a.sigma[frame] : 1.f;
        inv[j] = 1.0f / (2 * sg * sg);
    }
    const float factor = (float)(1 << a.quant_dec);
    const float vmax = (float)((1 << (a.quant_bits - 1)) - 1);
    Vec* vg = reinterpret_cast<Vec*>(var) + (size_t)g * N_LDPC * 32 + lane;
#pragma unroll 2
    for (int s = 0; s < SYMS; s++)
    {
        const int k = k0 + s;
        if (k >= a.n_sym) break;
        R q[BPS][NF];
#pragma unroll
        for (int j = 0; j < NF; j++)
        {
            float out[BPS];
            demap_symbol<BPS>(tile[warp][j][lane][2 * s], tile[warp][j][lane][2 * s + 1], inv[j], cs, out);
#pragma unroll
            for (int b = 0; b < BPS; b++)
                q[b][j] = live[j] ? quantize<R>(out[b], factor, vmax) : (R)0;
        }
#pragma unroll
        for (int b = 0; b < BPS; b++)
        {
            const int v = natural_pos(k, b, BPS, a.n_sym, a.itl_cols, a.itl_rev);
            vg[(size_t)v * 32] = LaneVec<R, NF>::make(q[b]);
        }
    }
}

// Framer::remove_plh (src/common/Module/Framer/Framer.hxx:330-343) as a plain copy
__global__ void
remove_plh_kernel(const float2* __restrict__ pl, float2* __restrict__ out, int n_sym, int pl_sym)
{
    const int f = blockIdx.x;
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k >= n_sym) return;
    out[(size_t)f * n_sym + k] = pl[(size_t)f * pl_sym + pl_symbol_pos(k, n_sym)];
}

// Estimator_DVBS2::_estimate (src/common/Module/Estimator/Estimator_DVBS2.hxx:31-58): M2/M4 moments of the frame's payload
// symbols, summed in float in symbol order exactly like the reference loop (no FMA contraction), then
// Se = sqrt|2 M2^2 - M4|, Ne = |M2 - Se|, Es/N0 = 10 log10(Se / Ne) (100 dB if infinite), sigma = esn0_to_sigma(Es/N0),
// Eb/N0 = esn0_to_ebn0(Es/N0, K_bch / 16200, bps) (lib/aff3ct/src/Tools/general_utils.cpp:78-113).
// warp = 32 frames (lane = frame, so every frame's sum runs in the reference's order); the symbols pass through shared
// memory in tiles of 16 so that global reads are 128-byte rows of one frame.  Two tiles are in flight per warp.
template<int FMT>
__global__ void __launch_bounds__(128) estimate_kernel(const DemapLaunch a, int K_bch, float* __restrict__ sigma,
                                                       float* __restrict__ ebn0, float* __restrict__ esn0)
{
    __shared__ float tile[4][2][32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sg = blockIdx.x * 4 + warp;
    if (sg * 32 >= a.n_frames) return;
    const int nfl = a.pl_sym > 0 ? 2 * a.pl_sym : 2 * a.n_sym;
    const int n_tiles = (a.n_sym + 15) / 16;
    auto stage = [&](int tl, int buf) {
        const int k0 = tl * 16;
        const int e = (a.pl_sym > 0 ? 2 * pl_symbol_pos(k0, a.n_sym) : 2 * k0) + lane;
        const bool in_frame = 2 * k0 + lane < 2 * a.n_sym;
#pragma unroll 16
        for (int r = 0; r < 32; r++)
        {
            const int f = sg * 32 + r;
            tile[warp][buf][r][lane] = (f < a.n_frames && in_frame) ? load_sym<FMT>(a, (size_t)f * nfl + e) : 0.f;
        }
    };
    float m2 = 0.f, m4 = 0.f;
    stage(0, 0);
    for (int tl = 0; tl < n_tiles; tl++)
    {
        __syncwarp();
        if (tl + 1 < n_tiles) stage(tl + 1, (tl + 1) & 1);
        const int ns = min(16, a.n_sym - tl * 16);
        for (int s = 0; s < ns; s++)
        {
            const float re = tile[warp][tl & 1][lane][2 * s], im = tile[warp][tl & 1][lane][2 * s + 1];
            const float tmp = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
            m2 = __fadd_rn(m2, tmp);
            m4 = __fadd_rn(m4, __fmul_rn(tmp, tmp));
        }
        __syncwarp();
    }
    const int f = sg * 32 + lane;
    if (f >= a.n_frames) return;
    m2 = __fdiv_rn(m2, (float)a.n_sym);
    m4 = __fdiv_rn(m4, (float)a.n_sym);
    const float Se = sqrtf(fabsf(__fsub_rn(__fmul_rn(__fmul_rn(2.f, m2), m2), m4)));
    const float Ne = fabsf(__fsub_rn(m2, Se));
    float e = 10.f * log10f(__fdiv_rn(Se, Ne));
    if (isinf(e)) e = 100.f;
    sigma[f] = sqrtf(__fdiv_rn(1.f, 2.f * powf(10.f, __fdiv_rn(e, 10.f))));
    if (ebn0) ebn0[f] = e - 10.f * log10f(((float)K_bch / 16200.f) * (float)a.bps);
    if (esn0) esn0[f] = e;
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
#define DEMAP_FRAMES(BPS)                                                                                              \
    switch (a.sym_fmt)                                                                                                 \
    {                                                                                                                  \
        case 1: demap_frames_kernel<BPS, 1><<<grid, 256, 0, s>>>(a, llr, deinterleave); break;                         \
        case 2: demap_frames_kernel<BPS, 2><<<grid, 256, 0, s>>>(a, llr, deinterleave); break;                         \
        case 3: demap_frames_kernel<BPS, 3><<<grid, 256, 0, s>>>(a, llr, deinterleave); break;                         \
        case 4: demap_frames_kernel<BPS, 4><<<grid, 256, 0, s>>>(a, llr, deinterleave); break;                         \
        default: demap_frames_kernel<BPS, 0><<<grid, 256, 0, s>>>(a, llr, deinterleave); break;                        \
    }
    switch (a.bps)
    {
        case 2: DEMAP_FRAMES(2) break;
        case 3: DEMAP_FRAMES(3) break;
        case 4: DEMAP_FRAMES(4) break;
        default: return cudaErrorInvalidValue;
    }
#undef DEMAP_FRAMES
    count_launch();
    return cudaGetLastError();
}

template<int BPS, typename R, int NF>
static void
demap_to_var_launch(const DemapLaunch& a, void* var, cudaStream_t s)
{
    constexpr int WARPS = DemapCfg<NF>::WARPS;
    const int groups = (a.n_frames + 32 * NF - 1) / (32 * NF); // whole decoder groups: pad frames get zeros
    const dim3 grid((a.n_sym + DemapCfg<NF>::SYMS * WARPS - 1) / (DemapCfg<NF>::SYMS * WARPS), groups);
    switch (a.sym_fmt)
    {
        case 1: demap_to_var_kernel<BPS, R, NF, 1><<<grid, WARPS * 32, 0, s>>>(a, (R*)var); break;
        case 2: demap_to_var_kernel<BPS, R, NF, 2><<<grid, WARPS * 32, 0, s>>>(a, (R*)var); break;
        case 3: demap_to_var_kernel<BPS, R, NF, 3><<<grid, WARPS * 32, 0, s>>>(a, (R*)var); break;
        case 4: demap_to_var_kernel<BPS, R, NF, 4><<<grid, WARPS * 32, 0, s>>>(a, (R*)var); break;
        default: demap_to_var_kernel<BPS, R, NF, 0><<<grid, WARPS * 32, 0, s>>>(a, (R*)var); break;
    }
    count_launch();
}

template<int BPS>
static cudaError_t
demap_to_var_bps(const DemapLaunch& a, void* var, int precision, int nf, cudaStream_t s)
{
    // the group width follows the precision: 1 or 2 (chosen per call) for float, 4 (or 1: small batches) for int16, 4 for int8
    if (precision == 0 ? (nf != 1 && nf != 2) : precision == 1 ? (nf != 1 && nf != 4) : nf != 4) return cudaErrorInvalidValue;
    switch (precision)
    {
        case 0:
            if (nf == 1)
                demap_to_var_launch<BPS, float, 1>(a, var, s);
            else
                demap_to_var_launch<BPS, float, 2>(a, var, s);
            break;
        case 1:
            if (nf == 1)
                demap_to_var_launch<BPS, int16_t, 1>(a, var, s);
            else
                demap_to_var_launch<BPS, int16_t, 4>(a, var, s);
            break;
        case 2: demap_to_var_launch<BPS, int8_t, 4>(a, var, s); break;
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
launch_remove_plh(const float* pl, float* out, int n_sym, int n_frames, cudaStream_t s)
{
    const dim3 grid(n_frames, (n_sym + 255) / 256);
    remove_plh_kernel<<<grid, 256, 0, s>>>((const float2*)pl, (float2*)out, n_sym, pl_frame_symbols(n_sym)); count_launch();
    return cudaGetLastError();
}

cudaError_t
launch_estimate(const DemapLaunch& a, int K_bch, float* sigma, float* ebn0, float* esn0, cudaStream_t s)
{
    const int subgroups = (a.n_frames + 31) / 32;
    const int blocks = (subgroups + 3) / 4;
    switch (a.sym_fmt)
    {
        case 1: estimate_kernel<1><<<blocks, 128, 0, s>>>(a, K_bch, sigma, ebn0, esn0); break;
        case 2: estimate_kernel<2><<<blocks, 128, 0, s>>>(a, K_bch, sigma, ebn0, esn0); break;
        case 3: estimate_kernel<3><<<blocks, 128, 0, s>>>(a, K_bch, sigma, ebn0, esn0); break;
        case 4: estimate_kernel<4><<<blocks, 128, 0, s>>>(a, K_bch, sigma, ebn0, esn0); break;
        default: estimate_kernel<0><<<blocks, 128, 0, s>>>(a, K_bch, sigma, ebn0, esn0); break;
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t
launch_deinterleave(const float* itl, float* nat, int n_frames, int n_sym, int itl_cols, int itl_rev, cudaStream_t s)
{
    const int bps = N_LDPC / n_sym;
    const dim3 grid(n_frames, (N_LDPC + 255) / 256);
    deinterleave_kernel<<<grid, 256, 0, s>>>(itl, nat, bps, n_sym, itl_cols, itl_rev); count_launch();
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
    count_launch();
    return cudaGetLastError();
}
} // namespace dvbs2
