// On-device transmitter, AWGN channel and BER/FER monitor for Monte-Carlo runs and benchmarks, sm_100a.
//
// What the reference chains on the CPU in src/mains/TX_RX_BB/main.cpp:75-84,92-94:
//   Source_random -> Scrambler_BB::scramble (src/common/Module/Scrambler/Scrambler_BB/Scrambler_BB.hxx:51-72)
//   -> Encoder_BCH_DVBS2 (src/common/Module/Encoder_BCH_DVBS2/Encoder_BCH_DVBS2.cpp:29-44 around
//      lib/aff3ct/src/Module/Encoder/BCH/Encoder_BCH.cpp:46-68)
//   -> Encoder_LDPC_DVBS2 (lib/aff3ct/src/Module/Encoder/LDPC/DVBS2/Encoder_LDPC_DVBS2.cpp:29-63)
//   -> Interleaver::interleave -> Modem_generic::modulate (lib/aff3ct/include/Module/Modem/Generic/Modem_generic.hxx:150-174)
//   -> Channel_AWGN_LLR::add_noise, and behind the decoders Scrambler_BB::descramble + Monitor_BFER::check_errors
//      (lib/aff3ct/src/Module/Monitor/BFER/Monitor_BFER.cpp:177-188).
// None of this is on the decode hot path; it exists so that frames are produced and scored where they are decoded
// (no PCIe) and FER curves can be pushed to low error rates on the GPUs.
//
// B200 mapping: everything up to the mapper works on the decoder's bit-sliced layout (one u32 per code bit holds
// that bit of the 32 frames of a group), so scrambling is one XOR per word, the BCH encoder is a 168-stage bit-sliced
// LFSR spread over the lanes of a warp, and the IRA encoder is an XOR-gather per check followed by a prefix XOR.
#include <cstdint>
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "kernels.h"

namespace dvbs2
{
namespace
{
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t
mix32(uint64_t x)
{ // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (uint32_t)(x >> 16);
}

// payload words: src[g][i], bit f = payload bit i of frame 32 g + f
__global__ void
tx_random_payload_kernel(uint32_t* __restrict__ src, int K, int n_groups, uint64_t seed)
{
    const size_t n = (size_t)K * n_groups;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        src[i] = mix32(seed * 0x100000001B3ull + i);
}

// BB scrambling + systematic BCH encoding.  One warp per group; lane L keeps LFSR stages 6L..6L+5 (168 = 28 x 6).
// cw[g][0..K) = scrambled payload, cw[g][K..N) = parity (first transmitted parity bit = highest LFSR stage).
__global__ void __launch_bounds__(128) tx_bch_encode_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ cw,
                                                            const uint32_t* __restrict__ scr, const uint32_t* __restrict__ gmask,
                                                            int K, int N, int n_groups)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const uint32_t* s = src + (size_t)g * K;
    uint32_t* c = cw + (size_t)g * N_LDPC;
    uint32_t gm[6], p[6];
#pragma unroll
    for (int r = 0; r < 6; r++)
        gm[r] = gmask[6 * lane + r], p[r] = 0;
    for (int i0 = 0; i0 < K; i0 += 32)
    {
        uint32_t w = 0;
        if (i0 + lane < K)
        {
            w = s[i0 + lane] ^ scr[i0 + lane];
            c[i0 + lane] = w;
        }
        const int n = min(32, K - i0);
        for (int j = 0; j < n; j++)
        {
            const uint32_t in = __shfl_sync(FULL, w, j);
            const uint32_t fb = in ^ __shfl_sync(FULL, p[5], 27); // stage 167
            uint32_t up = __shfl_up_sync(FULL, p[5], 1);
            if (lane == 0) up = 0;
            p[5] = p[4] ^ (gm[5] & fb);
            p[4] = p[3] ^ (gm[4] & fb);
            p[3] = p[2] ^ (gm[3] & fb);
            p[2] = p[1] ^ (gm[2] & fb);
            p[1] = p[0] ^ (gm[1] & fb);
            p[0] = up ^ (gm[0] & fb);
        }
    }
    const int R = N - K; // 168
#pragma unroll
    for (int r = 0; r < 6; r++)
    {
        const int stage = 6 * lane + r;
        if (stage < R) c[K + (R - 1 - stage)] = p[r];
    }
}

// IRA encoding: parity[c] = XOR of the information words of check c, then running XOR over c.  One block per group.
__global__ void __launch_bounds__(256) tx_ldpc_encode_kernel(uint32_t* __restrict__ cw, const uint16_t* __restrict__ idx, int K,
                                                             int M, int d_info, int d_pad)
{
    __shared__ uint32_t tot[256];
    const int g = blockIdx.x, t = threadIdx.x;
    uint32_t* c = cw + (size_t)g * N_LDPC;
    const int per = (M + 255) / 256;
    const int c0 = t * per, c1 = min(M, c0 + per);
    uint32_t run = 0;
    for (int k = c0; k < c1; k++)
    {
        uint32_t a = 0;
        const uint16_t* ic = idx + (size_t)k * d_pad;
        for (int i = 0; i < d_info; i++)
            a ^= c[ic[i]];
        run ^= a;
        c[K + k] = run; // local prefix, fixed up below
    }
    tot[t] = run;
    __syncthreads();
    uint32_t carry = 0;
    for (int j = 0; j < t; j++)
        carry ^= tot[j];
    if (carry)
        for (int k = c0; k < c1; k++)
            c[K + k] ^= carry;
}

// interleave + map + AWGN: thread = (frame, symbol).  Bit b of the symbol label is the code bit at natural position
// deinterleaved_pos(k, b) (the interleaver's inverse is the same closed form the demapper uses).
__global__ void __launch_bounds__(256) tx_modulate_awgn_kernel(const uint32_t* __restrict__ cw, const float* __restrict__ cstl,
                                                              float* __restrict__ sym, float* __restrict__ cp, int bps, int n_sym,
                                                              int itl_cols, int itl_rev, int n_frames, float sigma,
                                                              uint64_t seed)
{
    const int f = blockIdx.x; // frames on grid.x: batches exceed the 65535 limit of grid.y
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k == 0) cp[f] = sigma;
    if (k >= n_sym) return;
    const uint32_t* c = cw + (size_t)(f >> 5) * N_LDPC;
    unsigned label = 0;
    for (int b = 0; b < bps; b++)
    {
        int pos;
        if (itl_cols == 0)
            pos = k * bps + b;
        else
            pos = (itl_rev ? (itl_cols - 1 - b) : b) * n_sym + k;
        label |= ((c[pos] >> (f & 31)) & 1u) << b;
    }
    float re = cstl[2 * label], im = cstl[2 * label + 1];
    if (sigma > 0.f)
    {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)f * n_sym + k, 0, &st);
        const float2 n = curand_normal2(&st);
        re += sigma * n.x, im += sigma * n.y;
    }
    reinterpret_cast<float2*>(sym + (size_t)f * 2 * n_sym)[k] = make_float2(re, im);
}

// descramble + compare with the payload: one warp per group, lane L scans words L, L+32, ...; counters in registers
__global__ void __launch_bounds__(128) monitor_kernel(const uint32_t* __restrict__ dec, int dec_stride, const uint32_t* __restrict__ src,
                                                      const uint32_t* __restrict__ scr, int K, int n_frames, int n_groups,
                                                      unsigned long long* __restrict__ mon)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const uint32_t* d = dec + (size_t)g * dec_stride;
    const uint32_t* s = src + (size_t)g * K;
    unsigned valid = FULL;
    if ((g + 1) * 32 > n_frames) valid = (1u << (n_frames - g * 32)) - 1u;
    // bit-sliced vertical counters: cnt[j] holds bit j of the per-frame error count (K < 2^14 ... 2^15)
    uint32_t cnt[15];
#pragma unroll
    for (int j = 0; j < 15; j++)
        cnt[j] = 0;
    for (int i = lane; i < K; i += 32)
    {
        uint32_t carry = (d[i] ^ scr[i] ^ s[i]) & valid;
#pragma unroll
        for (int j = 0; j < 15; j++)
        {
            const uint32_t t = cnt[j] & carry;
            cnt[j] ^= carry;
            carry = t;
        }
    }
    // lane f extracts frame f's count from every lane's counters
    unsigned long long be = 0;
    unsigned bad = 0;
    for (int l = 0; l < 32; l++)
    {
        unsigned mine = 0;
#pragma unroll
        for (int j = 0; j < 15; j++)
            mine |= ((__shfl_sync(FULL, cnt[j], l) >> lane) & 1u) << j;
        be += mine;
    }
    bad = __ballot_sync(FULL, be != 0);
    for (int o = 16; o; o >>= 1)
        be += __shfl_xor_sync(FULL, be, o);
    if (lane == 0)
    {
        atomicAdd(mon + 0, (unsigned long long)__popc(valid));
        atomicAdd(mon + 1, be);
        atomicAdd(mon + 2, (unsigned long long)__popc(bad));
    }
}
} // namespace

cudaError_t
launch_tx_random_payload(uint32_t* src, int K, int n_groups, uint64_t seed, cudaStream_t s)
{
    tx_random_payload_kernel<<<148 * 8, 256, 0, s>>>(src, K, n_groups, seed); count_launch();
    return cudaGetLastError();
}

cudaError_t
launch_tx_encode(const TxLaunch& a, cudaStream_t s)
{
    tx_bch_encode_kernel<<<(a.n_groups + 3) / 4, 128, 0, s>>>(a.src, a.cw, a.scr, a.gmask, a.K_bch, a.N_bch, a.n_groups); count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    tx_ldpc_encode_kernel<<<a.n_groups, 256, 0, s>>>(a.cw, a.idx, a.N_bch, N_LDPC - a.N_bch, a.d_info, a.d_pad); count_launch();
    return cudaGetLastError();
}

cudaError_t
launch_tx_modulate_awgn(const TxLaunch& a, const float* cstl, float* sym, float* cp, int bps, int n_sym, int itl_cols,
                        int itl_rev, float sigma, uint64_t seed, cudaStream_t s)
{
    const dim3 grid(a.n_frames, (n_sym + 255) / 256);
    tx_modulate_awgn_kernel<<<grid, 256, 0, s>>>(a.cw, cstl, sym, cp, bps, n_sym, itl_cols, itl_rev, a.n_frames, sigma, seed); count_launch();
    return cudaGetLastError();
}

cudaError_t
launch_monitor(const uint32_t* dec, int dec_stride, const TxLaunch& a, unsigned long long* mon, cudaStream_t s)
{
    monitor_kernel<<<(a.n_groups + 3) / 4, 128, 0, s>>>(dec, dec_stride, a.src, a.scr, a.K_bch, a.n_frames, a.n_groups, mon); count_launch();
    return cudaGetLastError();
}
} // namespace dvbs2
