// BCH (N_bch, K_bch, t = 12) hard decoder over GF(2^14), shortened from 16383, sm_100a.
//
// Semantics (bit-exact): src/common/Module/Decoder_BCH_DVBS2/Decoder_BCH_DVBS2.cpp:28-40 (DVB-S2 bit order: the first
// transmitted bit is the highest-degree coefficient) around lib/aff3ct/src/Module/Decoder/BCH/Standard/
// Decoder_BCH_std.cpp:52-221: syndromes S_1..S_24, Berlekamp's iterative algorithm aborted once the locator degree
// exceeds t, Chien search over ALL 16383 field elements, success iff #roots == locator degree, flips only inside the
// shortened length, payload = first K bits, CWD = success.
//
// B200 mapping:
//   * Syndromes are where the reference spends its time (24 x N table look-ups per frame).  Here the received word
//     stays bit-sliced across the 32 frames of a group (one u32 per code bit, exactly the hd words the LDPC kernel
//     leaves in HBM), and each thread divides it by ONE of the 12 degree-14 minimal polynomials with a bit-sliced
//     LFSR: 14 state words, ~7 XORs per code bit for 32 frames at once.  S_j is then a fixed GF(2)-linear map of the
//     remainder modulo the minimal polynomial of alpha^j (conjugates share it), so all 24 syndromes cost 12 divisions.
//   * Only frames with a non-zero remainder (rare after a converged LDPC word) go on: one warp per such frame runs
//     Berlekamp-Massey on lane 0 (the minimal LFSR is unique whenever its length <= t, so any BM yields the
//     reference's polynomial; longer ones are failures in both) and a 32-lane Chien search; corrections are applied
//     to the bit-sliced words with atomicXor.
#include <algorithm>
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.h"

namespace dvbs2
{
namespace
{
// minimal polynomials of alpha^(2p+1), p = 0..11, for p(x) = 1 + x + x^3 + x^5 + x^14 (bit k = coefficient of x^k);
// dvbs2_cuda_create() checks them against the ones it derives from the field (code_tables.cpp).
__host__ __device__ constexpr unsigned
minpoly_const(int p)
{
    constexpr unsigned t[BCH_NPOLY] = { 0x402b, 0x4941, 0x4647, 0x5591, 0x6b55, 0x6389,
                                        0x6ce5, 0x4f21, 0x460f, 0x5a49, 0x5811, 0x65ef };
    return t[p];
}

template<int P>
__device__ __forceinline__ void
lfsr_step(uint32_t (&s)[GF_M], uint32_t in)
{
    constexpr unsigned MP = minpoly_const(P);
    const uint32_t fb = s[GF_M - 1];
#pragma unroll
    for (int k = GF_M - 1; k >= 1; k--)
        s[k] = ((MP >> k) & 1u) ? (s[k - 1] ^ fb) : s[k - 1];
    s[0] = in ^ fb; // every minimal polynomial has constant term 1
}

template<int P>
__device__ void
remainder_serial(const BchLaunch& a, int g)
{
    uint32_t s[GF_M];
#pragma unroll
    for (int k = 0; k < GF_M; k++)
        s[k] = 0;
    const uint4* w = reinterpret_cast<const uint4*>(a.words + (size_t)g * a.stride);
    const int n4 = a.N / 4;
#pragma unroll 2
    for (int i = 0; i < n4; i++)
    {
        const uint4 q = __ldg(w + i);
        lfsr_step<P>(s, q.x);
        lfsr_step<P>(s, q.y);
        lfsr_step<P>(s, q.z);
        lfsr_step<P>(s, q.w);
    }
    for (int i = n4 * 4; i < a.N; i++)
        lfsr_step<P>(s, a.words[(size_t)g * a.stride + i]);
    uint32_t any = 0;
    uint32_t* r = a.rem + ((size_t)g * BCH_NPOLY + P) * GF_M;
#pragma unroll
    for (int k = 0; k < GF_M; k++)
    {
        r[k] = s[k];
        any |= s[k];
    }
    if (any) atomicOr(a.nz + g, any);
}

// Large batches: thread = (group, polynomial), the whole word through one LFSR — sequential 16-byte loads of the group's own
// words (the segmented kernel's 32 lanes read 32 distant lines per load, which costs more than it saves once there are
// thousands of groups to hide the serial division behind).  grid: (ceil(groups / 128), 12 polynomials)
__global__ void __launch_bounds__(128) bch_remainder_serial_kernel(const BchLaunch a)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n_groups) return;
    switch (blockIdx.y)
    {
        case 0: remainder_serial<0>(a, g); break;
        case 1: remainder_serial<1>(a, g); break;
        case 2: remainder_serial<2>(a, g); break;
        case 3: remainder_serial<3>(a, g); break;
        case 4: remainder_serial<4>(a, g); break;
        case 5: remainder_serial<5>(a, g); break;
        case 6: remainder_serial<6>(a, g); break;
        case 7: remainder_serial<7>(a, g); break;
        case 8: remainder_serial<8>(a, g); break;
        case 9: remainder_serial<9>(a, g); break;
        case 10: remainder_serial<10>(a, g); break;
        case 11: remainder_serial<11>(a, g); break;
    }
}

// One warp per (group, polynomial): lane j runs the LFSR over its own segment of the word from a zero state, scales the
// remainder by x^(bits after the segment) mod m(x) (a constant 14-bit polynomial per lane: 14 shift-and-add steps on the
// bit-sliced state), and the 32 partial remainders are XORed across the warp.  The division of one word is then 450 dependent
// steps instead of 14 400 — a serial LFSR per thread cost ~1 ms however small the batch (17 % of a 4096-frame call).
template<int P>
__device__ void
remainder_one(const BchLaunch& a, int g, int lane)
{
    uint32_t s[GF_M], acc[GF_M];
#pragma unroll
    for (int k = 0; k < GF_M; k++)
        s[k] = acc[k] = 0;
    const uint32_t* w = a.words + (size_t)g * a.stride;
    const int b0 = min(lane * a.seg_len, a.N), b1 = min(b0 + a.seg_len, a.N);
#pragma unroll 4
    for (int i = b0; i < b1; i++)
        lfsr_step<P>(s, __ldg(w + i));
    const unsigned c = a.seg_mul[P][lane]; // x^(N - b1) mod m_P
#pragma unroll
    for (int t = 0; t < GF_M; t++)
    {
        if ((c >> t) & 1u)
        {
#pragma unroll
            for (int k = 0; k < GF_M; k++)
                acc[k] ^= s[k];
        }
        lfsr_step<P>(s, 0u); // s <- s * x
    }
    uint32_t any = 0;
    uint32_t* r = a.rem + ((size_t)g * BCH_NPOLY + P) * GF_M;
#pragma unroll
    for (int k = 0; k < GF_M; k++)
    {
        const uint32_t v = __reduce_xor_sync(0xffffffffu, acc[k]);
        if (lane == 0) r[k] = v;
        any |= v;
    }
    if (lane == 0 && any) atomicOr(a.nz + g, any);
}

// grid: (ceil(groups / 4), 12 polynomials); warp = (group, polynomial)
__global__ void __launch_bounds__(128) bch_remainder_kernel(const BchLaunch a)
{
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= a.n_groups) return;
    switch (blockIdx.y)
    {
        case 0: remainder_one<0>(a, g, lane); break;
        case 1: remainder_one<1>(a, g, lane); break;
        case 2: remainder_one<2>(a, g, lane); break;
        case 3: remainder_one<3>(a, g, lane); break;
        case 4: remainder_one<4>(a, g, lane); break;
        case 5: remainder_one<5>(a, g, lane); break;
        case 6: remainder_one<6>(a, g, lane); break;
        case 7: remainder_one<7>(a, g, lane); break;
        case 8: remainder_one<8>(a, g, lane); break;
        case 9: remainder_one<9>(a, g, lane); break;
        case 10: remainder_one<10>(a, g, lane); break;
        case 11: remainder_one<11>(a, g, lane); break;
    }
}

struct GF
{
    const uint16_t* __restrict__ exp;
    const int16_t* __restrict__ log;
    __device__ unsigned mul(unsigned x, unsigned y) const
    {
        if (!x || !y) return 0;
        int e = __ldg(log + x) + __ldg(log + y);
        if (e >= GF_N) e -= GF_N;
        return __ldg(exp + e);
    }
    __device__ unsigned div(unsigned x, unsigned y) const // y != 0
    {
        if (!x) return 0;
        int e = __ldg(log + x) - __ldg(log + y);
        if (e < 0) e += GF_N;
        return __ldg(exp + e);
    }
};

constexpr int BCH_FRAMES_PER_BLOCK = 4;
// one warp per frame; frames with a zero syndrome leave immediately
__global__ void __launch_bounds__(BCH_FRAMES_PER_BLOCK * 32) bch_correct_kernel(const BchLaunch a, const BchDevTables t)
{
    __shared__ unsigned sh_S[BCH_FRAMES_PER_BLOCK][2 * BCH_T + 2];
    __shared__ int sh_lam_log[BCH_FRAMES_PER_BLOCK][BCH_T + 2]; // log of locator coefficients, -1 = zero
    __shared__ int sh_L[BCH_FRAMES_PER_BLOCK];
    __shared__ int sh_loc[BCH_FRAMES_PER_BLOCK][BCH_T + 4];
    __shared__ int sh_cnt[BCH_FRAMES_PER_BLOCK];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int frame = blockIdx.x * BCH_FRAMES_PER_BLOCK + warp;
    if (frame >= a.n_frames) return;
    const int g = frame >> 5, f = frame & 31;
    if (!((a.nz[g] >> f) & 1u))
    {
        if (lane == 0 && a.cwd) a.cwd[frame] = 1;
        return;
    }
    const GF gf{ t.alpha_to, t.index_of };

    // syndromes in polynomial form: S_j = sum_k rem_{poly_of[j]}[k] alpha^(jk)
    if (lane >= 1 && lane <= 2 * BCH_T)
    {
        const uint32_t* r = a.rem + ((size_t)g * BCH_NPOLY + t.poly_of[lane]) * GF_M;
        unsigned acc = 0;
#pragma unroll
        for (int k = 0; k < GF_M; k++)
            if ((r[k] >> f) & 1u) acc ^= t.synd_mat[lane][k];
        sh_S[warp][lane] = acc;
    }
    if (lane == 0) sh_cnt[warp] = 0;
    __syncwarp();

    if (lane == 0)
    {
        // Berlekamp-Massey; C = current connection polynomial, B = last one before a length change
        unsigned Cp[BCH_T + 2], Bp[BCH_T + 2], Tp[BCH_T + 2];
#pragma unroll
        for (int i = 0; i < BCH_T + 2; i++)
            Cp[i] = Bp[i] = 0;
        Cp[0] = Bp[0] = 1;
        int L = 0, m = 1;
        unsigned b = 1;
        bool fail = false;
        const unsigned* S = sh_S[warp];
        for (int n = 0; n < 2 * BCH_T && !fail; n++)
        {
            unsigned d = S[n + 1];
            for (int i = 1; i <= L; i++)
                d ^= gf.mul(Cp[i], S[n + 1 - i]);
            if (d == 0)
            {
                m++;
                continue;
            }
            const unsigned coef = gf.div(d, b);
            if (2 * L <= n)
            {
                const int newL = n + 1 - L;
                if (newL > BCH_T)
                {
                    fail = true;
                    L = newL;
                    break;
                }
                for (int i = 0; i < BCH_T + 2; i++)
                    Tp[i] = Cp[i];
                for (int i = 0; i + m < BCH_T + 2; i++)
                    Cp[i + m] ^= gf.mul(coef, Bp[i]);
                for (int i = 0; i < BCH_T + 2; i++)
                    Bp[i] = Tp[i];
                L = newL, b = d, m = 1;
            }
            else
            {
                for (int i = 0; i + m < BCH_T + 2; i++)
                    Cp[i + m] ^= gf.mul(coef, Bp[i]);
                m++;
            }
        }
        sh_L[warp] = fail ? -1 : L;
        for (int j = 0; j <= BCH_T; j++)
            sh_lam_log[warp][j] = Cp[j] ? (int)__ldg(t.index_of + Cp[j]) : -1;
    }
    __syncwarp();

    const int L = sh_L[warp];
    bool ok = false;
    if (L >= 0)
    {
        // Chien search: Lambda(alpha^i) for i = 1..16383, lanes interleaved; root at i <-> error location 16383 - i
        int e[BCH_T + 1];
#pragma unroll
        for (int j = 1; j <= BCH_T; j++)
        {
            const int lg = sh_lam_log[warp][j];
            e[j] = (j <= L && lg >= 0) ? (lg + j * (lane + 1)) % GF_N : -1;
        }
        for (int i = lane + 1; i <= GF_N; i += 32)
        {
            unsigned q = 1;
#pragma unroll
            for (int j = 1; j <= BCH_T; j++)
                if (e[j] >= 0)
                {
                    q ^= __ldg(t.alpha_to + e[j]);
                    e[j] += 32 * j;
                    if (e[j] >= GF_N) e[j] -= GF_N; // 32*12 < 16383: one subtraction suffices
                }
            if (q == 0)
            {
                const int slot = atomicAdd(&sh_cnt[warp], 1);
                if (slot < BCH_T + 4) sh_loc[warp][slot] = GF_N - i;
            }
        }
        __syncwarp();
        const int cnt = sh_cnt[warp];
        if (cnt == L)
        {
            ok = true;
            if (lane < cnt)
            {
                const int loc = sh_loc[warp][lane];
                if (loc < a.N) atomicXor(a.words + (size_t)g * a.stride + (a.N - 1 - loc), 1u << f);
            }
        }
    }
    if (lane == 0)
    {
        if (a.cwd) a.cwd[frame] = (int8_t)ok;
        if (a.stats)
        {
            atomicAdd(a.stats + STAT_BCH_NONZERO, 1ull);
            if (!ok) atomicAdd(a.stats + STAT_BCH_FAIL, 1ull);
        }
    }
}
} // namespace

bool
bch_minpoly_matches(const BchTables& t)
{
    for (int p = 0; p < BCH_NPOLY; p++)
        if ((minpoly_const(p) & 0x3FFFu) != t.minpoly[p] || !(minpoly_const(p) >> GF_M)) return false;
    return true;
}

// x^e mod (x^14 + low14(x)) by square-and-multiply; polynomials as bit masks
static unsigned
xpow_mod(unsigned e, unsigned low14)
{
    auto mulx = [&](unsigned c) { c <<= 1; return (c & (1u << GF_M)) ? (c ^ (1u << GF_M) ^ low14) : c; };
    unsigned c = 1;
    for (int bit = 31; bit >= 0; bit--)
    {
        unsigned sq = 0; // c^2: Horner over the coefficients of c, multiplying by x^2 per step
        for (int k = GF_M - 1; k >= 0; k--)
            sq = mulx(mulx(sq)) ^ ((c >> k) & 1u);
        c = sq;
        if ((e >> bit) & 1u) c = mulx(c);
    }
    return c;
}

void
bch_segment_table(BchLaunch& a, const uint16_t* minpoly_low14)
{
    a.seg_len = (a.N + 31) / 32;
    for (int p = 0; p < BCH_NPOLY; p++)
        for (int j = 0; j < 32; j++)
        {
            const int b0 = std::min(j * a.seg_len, a.N), b1 = std::min(b0 + a.seg_len, a.N);
            a.seg_mul[p][j] = (uint16_t)xpow_mod((unsigned)(a.N - b1), minpoly_low14[p]);
        }
}

cudaError_t
launch_bch_decode(const BchLaunch& a, const BchDevTables& t, cudaStream_t s)
{
    if (a.seg_len <= 0 || a.seg_len * 32 < a.N) return cudaErrorInvalidValue; // bch_segment_table() was not called
    cudaError_t e = cudaMemsetAsync(a.nz, 0, sizeof(uint32_t) * a.n_groups, s);
    if (e != cudaSuccess) return e;
    if (a.n_groups <= 2048)
    { // up to 65536 frames: the segmented division (short dependency chain)
        const dim3 grid((a.n_groups + 3) / 4, BCH_NPOLY);
        bch_remainder_kernel<<<grid, 128, 0, s>>>(a); count_launch();
    }
    else
    {
        const dim3 grid((a.n_groups + 127) / 128, BCH_NPOLY);
        bch_remainder_serial_kernel<<<grid, 128, 0, s>>>(a); count_launch();
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int blocks = (a.n_frames + BCH_FRAMES_PER_BLOCK - 1) / BCH_FRAMES_PER_BLOCK;
    bch_correct_kernel<<<blocks, BCH_FRAMES_PER_BLOCK * 32, 0, s>>>(a, t); count_launch();
    return cudaGetLastError();
}
} // namespace dvbs2
