// Horizontal-layered (normalised) min-sum LDPC decoder for the DVB-S2 short-frame IRA codes, sm_100a.
//
// Semantics (bit-exact): lib/aff3ct/include/Module/Decoder/LDPC/BP/Horizontal_layered/
// Decoder_LDPC_BP_horizontal_layered.hxx:89-123,156-207 with Update_rule_MS.hxx:48-92 / Update_rule_NMS.hxx:14-56,
// 100-106 and LDPC_syndrome.hxx:43-69 — checks are visited in natural order c = 0..M-1 (the IRA dual diagonal makes
// that a strict chain), per-frame syndrome early termination.
//
// B200 mapping ("lane = frame"): a warp owns a group of 32 frames and walks the checks in order; lane f holds frame f.
//   * Variable-node values live in HBM as var[group][v][32 lanes] so every access of the warp is one full line
//     (128 B for float, 64 B for int16, 32 B for int8): all traffic is sector-exact although the per-frame access
//     pattern is a gather.
//   * Check-to-variable messages are NOT stored per edge.  Per check and frame we keep the two output magnitudes,
//     one sign bit and one "|t| == min1" bit per edge (16 B) — the reference's output rule is value-equality based,
//     so this reconstructs every message exactly.  Iteration 0 reads no check state at all (messages are zero).
//   * The parity-chain value var[K+c] produced by check c is carried to check c+1 in a register.
//   * Hard decisions are kept bit-sliced across the 32 frames: hd[group][v] is a 32-bit word written with one
//     __ballot_sync per variable update, so the syndrome test after an iteration is E word-XORs per warp instead of
//     E loads per frame, and the BCH stage consumes the same words directly.
//   * A lane whose frame has converged stops writing; its bits were snapshotted into out_hd at that moment.
// The kernel is HBM-bound: algorithmic bytes per frame and iteration = 8 B * information edges + 8 B * M (parity
// r/w) + 32 B * M (check state r/w)   (float).
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.h"

namespace dvbs2
{
namespace
{
constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------- arithmetic traits
template<typename R>
struct Arith;

template<>
struct Arith<float>
{
    using C = float;
    __device__ static C max_value() { return FLT_MAX; }
    __device__ static C sub(C a, C b) { return __fsub_rn(a, b); }
    __device__ static C add(C a, C b) { return __fadd_rn(a, b); }
    __device__ static C absv(C t) { return fabsf(t); }
    __device__ static unsigned neg(C t) { return __float_as_uint(t) >> 31; } // std::signbit (Update_rule_MS.hxx:62)
    __device__ static bool lt0(C v) { return v < 0.f; }                    // LDPC_syndrome.hxx:56
    __device__ static C vmin(C a, C b) { return fminf(a, b); }
    __device__ static C vmax(C a, C b) { return fmaxf(a, b); }
    __device__ static C normalize(C v, float f, int) { return __fmul_rn(v, f); } // Update_rule_NMS.hxx:41-46
    __device__ static C with_sign(C m, unsigned s) { return __uint_as_float(__float_as_uint(m) | (s << 31)); }
    __device__ static unsigned pack(C a, C b, unsigned& hi)
    {
        hi = __float_as_uint(b);
        return __float_as_uint(a);
    }
    __device__ static void unpack(unsigned lo, unsigned hi, C& a, C& b) { a = __uint_as_float(lo), b = __uint_as_float(hi); }
};

// Integer R: C++ integral promotion to int, truncation back to R on every store (wraps), exactly like the
// reference's sequential decoder instantiated with R = int16_t / int8_t.
template<typename R>
struct ArithInt
{
    using C = int;
    __device__ static C wrap(int x) { return (int)(R)x; }
    __device__ static C max_value() { return sizeof(R) == 2 ? 32767 : 127; }
    __device__ static C sub(C a, C b) { return wrap(a - b); }
    __device__ static C add(C a, C b) { return wrap(a + b); }
    __device__ static C absv(C t) { return wrap(t < 0 ? -t : t); }
    __device__ static unsigned neg(C t) { return t < 0; }
    __device__ static bool lt0(C v) { return v < 0; }
    __device__ static C vmin(C a, C b) { return a < b ? a : b; }
    __device__ static C vmax(C a, C b) { return a > b ? a : b; }
    __device__ static C normalize(C v, float, int k8)
    { // Update_rule_NMS.hxx:14-39 with div2/4/8 = arithmetic shifts (lib/streampu/include/Tools/Math/utils.hxx:18-78)
        if (k8 == 8) return v;
        int r = 0;
        if (k8 & 4) r += wrap(v >> 1);
        if (k8 & 2) r += wrap(v >> 2);
        if (k8 & 1) r += wrap(v >> 3);
        return wrap(r);
    }
    __device__ static C with_sign(C m, unsigned s) { return wrap(s ? -m : m); }
    __device__ static unsigned pack(C a, C b, unsigned& hi)
    {
        hi = 0;
        return ((unsigned)a & 0xffffu) | ((unsigned)b << 16);
    }
    __device__ static void unpack(unsigned lo, unsigned, C& a, C& b) { a = (int)(short)(lo & 0xffffu), b = (int)(short)(lo >> 16); }
};
template<>
struct Arith<int16_t> : ArithInt<int16_t>
{
};
template<>
struct Arith<int8_t> : ArithInt<int8_t>
{
};

template<typename R>
__device__ __forceinline__ typename Arith<R>::C
ld_var(const R* p)
{
    return (typename Arith<R>::C)__ldcg(p);
}
template<typename R>
__device__ __forceinline__ void
st_var(R* p, typename Arith<R>::C v)
{
    __stcg(p, (R)v);
}

// ---------------------------------------------------------------------------------------------- the decoder
template<typename R, int DI>
__global__ void __launch_bounds__(LDPC_WARPS_PER_BLOCK * 32) ldpc_decode_kernel(const LdpcLaunch a)
{
    using A = Arith<R>;
    using C = typename A::C;
    constexpr int DP = (DI + 7) / 8 * 8; // row stride of the index table
    constexpr int D = DI + 2;            // edges per check incl. the two parity edges (bit DI = in, DI+1 = out)

    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * LDPC_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (g >= a.n_groups) return;

    R* __restrict__ var = (R*)a.var + (size_t)g * N_LDPC * 32 + lane;
    uint4* __restrict__ cn = a.cn + (size_t)g * a.M * 32 + lane;
    uint32_t* __restrict__ hd = a.hd + (size_t)g * N_LDPC;
    uint32_t* __restrict__ out_hd = a.out_hd + (size_t)g * N_LDPC;
    const uint16_t* __restrict__ idx = a.idx;
    const int K = a.K, M = a.M;

    const int frame = g * 32 + lane;
    unsigned active = __ballot_sync(FULL, frame < a.n_frames);
    unsigned done = 0;   // lanes whose result is already in out_hd
    unsigned failed = 0; // per-lane: syndrome non-zero at its last iteration
    int my_iters = 0;

    // hard decisions of the channel values (needed only if n_ite == 0 is ever allowed; also primes hd for reads)
    for (int ite = 0; ite < a.n_ite && active; ite++)
    {
        const bool mine = (active >> lane) & 1;
        C chain = A::max_value(); // var[K+c-1] after check c-1; for c = 0 a neutral dummy edge

        for (int c = 0; c < M; c++)
        {
            // ---- this check's information-variable indices (warp-uniform 16-byte loads)
            uint32_t iw[DP / 2];
            const uint4* ip = reinterpret_cast<const uint4*>(idx + (size_t)c * DP);
#pragma unroll
            for (int k = 0; k < DP / 8; k++)
            {
                const uint4 q = __ldg(ip + k);
                iw[4 * k + 0] = q.x, iw[4 * k + 1] = q.y, iw[4 * k + 2] = q.z, iw[4 * k + 3] = q.w;
            }
            // ---- previous messages of this check, compressed (none yet in iteration 0)
            uint4 st = make_uint4(0u, 0u, 0u, 0u);
            if (ite > 0) st = __ldcg(cn + (size_t)c * 32);
            C o1, o2;
            A::unpack(st.x, st.y, o1, o2);
            const unsigned osg = st.z, oeq = st.w;

            // ---- pass 1: t = var - msg, running sign / min1 / min2 (Update_rule_MS.hxx:57-67)
            C t[D];
            unsigned sacc = 0;
            C min1 = A::max_value(), min2 = A::max_value();
#pragma unroll
            for (int i = 0; i < D; i++)
            {
                C v;
                if (i < DI)
                {
                    const unsigned vi = (iw[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                    v = ld_var<R>(var + (size_t)vi * 32);
                }
                else if (i == DI)
                    v = chain;
                else
                    v = ld_var<R>(var + (size_t)(K + c) * 32);
                const C m = A::with_sign(((oeq >> i) & 1u) ? o1 : o2, (osg >> i) & 1u);
                C ti = A::sub(v, m);
                if (i == DI && c == 0) ti = A::max_value(); // check 0 has no incoming chain edge
                t[i] = ti;
                const C ab = A::absv(ti);
                sacc ^= A::neg(ti);
                min2 = A::vmin(min2, A::vmax(ab, min1));
                min1 = A::vmin(min1, ab);
            }
            // ---- normalise, clamp (Update_rule_NMS.hxx:100-106)
            const C c1 = A::vmax((C)0, A::normalize(min2, a.norm, a.k8));
            const C c2 = A::vmax((C)0, A::normalize(min1, a.norm, a.k8));

            // ---- pass 2: new messages and variable values (Update_rule_MS.hxx:82-92, ...layered.hxx:198-203)
            unsigned nsg = 0, neq = 0;
#pragma unroll
            for (int i = 0; i < D; i++)
            {
                const C ti = t[i];
                const unsigned e = (A::absv(ti) == min1);
                const unsigned s = sacc ^ A::neg(ti);
                const C nv = A::add(ti, A::with_sign(e ? c1 : c2, s));
                neq |= e << i;
                nsg |= s << i;
                int vi;
                if (i < DI)
                    vi = (iw[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                else if (i == DI)
                    vi = K + c - 1;
                else
                    vi = K + c;
                if (i == DI + 1)
                {
                    chain = nv; // consumed (and stored) by check c+1; only the last check must store it itself
                    if (c != M - 1) continue;
                }
                if (i == DI && c == 0) continue;
                if (mine) st_var<R>(var + (size_t)vi * 32, nv);
                const unsigned w = __ballot_sync(FULL, A::lt0(nv));
                if (lane == 0) __stcg(hd + vi, w);
            }
            if (mine)
            {
                unsigned hi;
                const unsigned lo = A::pack(c1, c2, hi);
                __stcg(cn + (size_t)c * 32, make_uint4(lo, hi, nsg, neq));
            }
        }
        my_iters += mine ? 1 : 0;

        // ---- syndrome on the bit-sliced hard decisions: lane j tests checks j, j+32, ...
        __syncwarp();
        unsigned unsat = 0;
        for (int c = lane; c < M; c += 32)
        {
            unsigned p = __ldcg(hd + K + c);
            if (c > 0) p ^= __ldcg(hd + K + c - 1);
            const uint16_t* ic = idx + (size_t)c * DP;
#pragma unroll
            for (int i = 0; i < DI; i++)
                p ^= __ldcg(hd + ic[i]);
            unsat |= p;
        }
        unsat = __reduce_or_sync(FULL, unsat);

        const bool last = (ite == a.n_ite - 1);
        unsigned fin = active & ~unsat;
        if (!a.early_stop) fin = 0;
        if (last) fin = active;
        if (fin)
        {
            // snapshot the finishing lanes' hard decisions
            for (int v = lane; v < N_LDPC; v += 32)
            {
                const unsigned w = __ldcg(hd + v);
                const unsigned o = done ? __ldcg(out_hd + v) : 0u;
                __stcg(out_hd + v, (o & done) | (w & fin));
            }
            __syncwarp();
            failed |= fin & unsat;
            done |= fin;
            active &= ~fin;
        }
    }

    if (frame < a.n_frames)
    {
        const bool ok = !((failed >> lane) & 1);
        if (a.cwd) a.cwd[frame] = (int8_t)ok;
        if (a.n_iter) a.n_iter[frame] = my_iters;
        // stats: iterations and failures (one atomic per warp)
    }
    const unsigned valid = __ballot_sync(FULL, frame < a.n_frames);
    int it_sum = __reduce_add_sync(FULL, frame < a.n_frames ? my_iters : 0);
    if (lane == 0 && a.stats)
    {
        atomicAdd(a.stats + STAT_FRAMES, (unsigned long long)__popc(valid));
        atomicAdd(a.stats + STAT_LDPC_ITERS, (unsigned long long)it_sum);
        atomicAdd(a.stats + STAT_LDPC_FAIL, (unsigned long long)__popc(failed & valid));
    }
}

// ---------------------------------------------------------------------------------------------- layout kernels
// frame-major LLRs [F][16200] -> var[group][v][lane]; frames past n_frames are zero-filled.
template<typename R>
__global__ void
pack_llr_kernel(const R* __restrict__ in, R* __restrict__ var, int n_frames)
{
    __shared__ R tile[32][33];
    const int g = blockIdx.y, v0 = blockIdx.x * 32;
    const int x = threadIdx.x, y = threadIdx.y;
    for (int j = y; j < 32; j += blockDim.y)
    {
        const int f = g * 32 + j, v = v0 + x;
        tile[j][x] = (f < n_frames && v < N_LDPC) ? in[(size_t)f * N_LDPC + v] : (R)0;
    }
    __syncthreads();
    for (int j = y; j < 32; j += blockDim.y)
    {
        const int v = v0 + j;
        if (v < N_LDPC) var[((size_t)g * N_LDPC + v) * 32 + x] = tile[x][j];
    }
}

// bit-sliced words [group][stride] -> int32 bits [F][K] (first K variables)
__global__ void
unpack_bits_kernel(const uint32_t* __restrict__ words, int32_t* __restrict__ out, int stride, int K, int n_frames)
{
    const int g = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= K) return;
    const uint32_t w = words[(size_t)g * stride + v];
#pragma unroll 8
    for (int f = 0; f < 32; f++)
    {
        const int frame = g * 32 + f;
        if (frame < n_frames) out[(size_t)frame * K + v] = (int32_t)((w >> f) & 1u);
    }
}

// int32 bits [F][N] -> bit-sliced words [group][stride]
__global__ void
pack_bits_kernel(const int32_t* __restrict__ in, uint32_t* __restrict__ words, int stride, int N, int n_frames)
{
    __shared__ int tile[32][33];
    const int g = blockIdx.y, v0 = blockIdx.x * 32;
    const int x = threadIdx.x, y = threadIdx.y;
    for (int j = y; j < 32; j += blockDim.y)
    {
        const int f = g * 32 + j, v = v0 + x;
        tile[j][x] = (f < n_frames && v < N) ? in[(size_t)f * N + v] : 0;
    }
    __syncthreads();
    for (int j = y; j < 32; j += blockDim.y)
    {
        const unsigned w = __ballot_sync(FULL, tile[x][j] != 0);
        if (x == 0 && v0 + j < N) words[(size_t)g * stride + v0 + j] = w;
    }
}
} // namespace

// ---------------------------------------------------------------------------------------------- launchers
template<typename R>
static cudaError_t
launch_ldpc_t(const LdpcLaunch& a, cudaStream_t s)
{
    const int blocks = (a.n_groups + LDPC_WARPS_PER_BLOCK - 1) / LDPC_WARPS_PER_BLOCK;
    if (a.d_info == 25)
        ldpc_decode_kernel<R, 25><<<blocks, LDPC_WARPS_PER_BLOCK * 32, 0, s>>>(a);
    else if (a.d_info == 9)
        ldpc_decode_kernel<R, 9><<<blocks, LDPC_WARPS_PER_BLOCK * 32, 0, s>>>(a);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t
launch_ldpc_decode(const LdpcLaunch& a, int precision, cudaStream_t s)
{
    switch (precision)
    {
        case 0: return launch_ldpc_t<float>(a, s);
        case 1: return launch_ldpc_t<int16_t>(a, s);
        case 2: return launch_ldpc_t<int8_t>(a, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t
launch_pack_llr(const void* in, void* var, int precision, int n_frames, cudaStream_t s)
{
    const int groups = (n_frames + 31) / 32;
    const dim3 grid((N_LDPC + 31) / 32, groups), block(32, 8);
    switch (precision)
    {
        case 0: pack_llr_kernel<float><<<grid, block, 0, s>>>((const float*)in, (float*)var, n_frames); break;
        case 1: pack_llr_kernel<int16_t><<<grid, block, 0, s>>>((const int16_t*)in, (int16_t*)var, n_frames); break;
        case 2: pack_llr_kernel<int8_t><<<grid, block, 0, s>>>((const int8_t*)in, (int8_t*)var, n_frames); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t
launch_unpack_bits(const uint32_t* words, int32_t* out, int stride, int K, int n_frames, cudaStream_t s)
{
    const int groups = (n_frames + 31) / 32;
    const dim3 grid((K + 127) / 128, groups);
    unpack_bits_kernel<<<grid, 128, 0, s>>>(words, out, stride, K, n_frames);
    return cudaGetLastError();
}

cudaError_t
launch_pack_bits(const int32_t* in, uint32_t* words, int stride, int N, int n_frames, cudaStream_t s)
{
    const int groups = (n_frames + 31) / 32;
    const dim3 grid((N + 31) / 32, groups), block(32, 8);
    pack_bits_kernel<<<grid, block, 0, s>>>(in, words, stride, N, n_frames);
    return cudaGetLastError();
}
} // namespace dvbs2
