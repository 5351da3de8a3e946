// "Design B" of the LDPC decoder: variable nodes ON CHIP, float, bit-exact.
//
// Same semantics as ldpc_kernel.cu (Decoder_LDPC_BP_horizontal_layered.hxx:89-123,156-207 with Update_rule_NMS, natural check
// order, per-frame syndrome stop); a different mapping:
//   * a CTA decodes F frames at a time (F = 3 for rate 8/9, 5 for rate 3/5) whose information variables live in shared memory
//     (57.6 / 38.9 KB per frame); parity variables are touched by two consecutive checks only and stream through L2, as does the
//     compressed check state (16 B per check, the same exact encoding as in ldpc_kernel.cu);
//   * lanes = CHECKS: a "frame warp" owns one frame and processes a window of W = 26 consecutive checks per step, one check per
//     lane.  Checks closer than min_reuse_distance (52 / 55 for these codes) share no information variable, so the lanes of a
//     window — and two neighbouring windows — never touch the same shared-memory word;
//   * the only dependency between the checks of a window is the parity chain var[K+c-1] -> var[K+c].  For check c it is
//         x_c = t_out + sign * min(A_c, norm(|x_{c-1} - m_in|))          (three dependent float operations)
//     with t_out, m_in, A_c = max(0, norm(min over the information edges)) and the sign parity known before the chain arrives
//     (the value-equality rule of Update_rule_MS.hxx:82-92 gives every edge the minimum over the OTHER edges).  One "chain warp"
//     per CTA walks these recurrences for all F frames at once (lane = frame) from what the frame warps posted in shared
//     memory, and hands back t_in = x_{c-1} - m_in per check; the frame warps then finish the window (new messages for every
//     edge, check state).  Frame warps post window w+1 before they finish window w, so the chain never waits for them;
//   * no frame compaction, no rounds, no riding frames: a frame leaves its slot the moment its syndrome is zero and the slot
//     loads the next frame of the queue.
// HBM traffic is the compulsory 64.8 KB in + 1.8 KB out per frame; the kernel is bound by instruction issue / the chain.
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "kernels.h"

namespace dvbs2
{
namespace
{
constexpr unsigned FULL = 0xffffffffu;
constexpr int W = ONCHIP_WINDOW; // checks per window

__device__ __forceinline__ void
mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void
mbar_arrive_release(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void
mbar_wait_acquire(uint64_t* bar, unsigned parity)
{
    unsigned ok;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
// sign(a) ^ sign(b) applied to min(|a|, |b|)
__device__ __forceinline__ float
min_xorsign_abs(float a, float b)
{
    float d;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float
with_sign_of(float mag, unsigned w) // mag >= 0, sign taken from bit 31 of w
{
    return __uint_as_float(__float_as_uint(mag) | (w & 0x80000000u));
}
__device__ __forceinline__ void
fold(float t, float& m1, float& m2, unsigned& sa)
{
    const float ab = fabsf(t);
    sa ^= __float_as_uint(t);
    m2 = fminf(m2, fmaxf(ab, m1));
    m1 = fminf(m1, ab);
}

// what pass 1 of a window needs from L2: requested two windows ahead, so that its latency lies under the work of the windows
// in between (a warp issues in order: a load consumed right after its request would stall everything behind it)
template<int DI>
struct LoadRegs
{
    uint32_t ix[(DI + 1) / 2];
    uint4 st;
    float pvo;
};

template<int DI>
struct WinRegs
{
    float t[DI];                 // t = var - msg of the information edges
    uint32_t ix[(DI + 1) / 2];   // byte offsets of their variables in the frame's shared-memory block, two per word
    float m1, m2, t_out;
    unsigned sa;
};

template<int DI, int F, bool NORM1>
__global__ void __launch_bounds__((F + 1) * 32, 1) ldpc_onchip_kernel(const OnchipLaunch a)
{
    constexpr int IWP = (DI + 1) / 2;
    const int K = a.K, M = a.M;
    const int n_win = (M + W - 1) / W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_var = reinterpret_cast<float*>(smem_raw);                  // [F][K]
    uint4* s_post = reinterpret_cast<uint4*>(s_var + (size_t)F * K);     // [2][W][F] {t_out, m_in, A with the sign parity, -}
    float* s_ans = reinterpret_cast<float*>(s_post + 2 * W * F);         // [2][W][F] t_in per check
    uint32_t* s_pw = reinterpret_cast<uint32_t*>(s_ans + 2 * W * F);     // [F][n_win + 1] hard decisions of the parity chain
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_pw + (((size_t)F * (n_win + 1) + 1) & ~(size_t)1)); // posted[2], answered[2]
    int* s_frame = reinterpret_cast<int*>(s_bar + 4);                    // [F] frame in each slot, -1 = empty

    const int lane = threadIdx.x & 31, w_id = threadIdx.x >> 5;
    const bool chain_warp = w_id == F;
    if (threadIdx.x < 2) mbar_init(s_bar + threadIdx.x, F);      // posted: one arrival per frame warp
    if (threadIdx.x >= 2 && threadIdx.x < 4) mbar_init(s_bar + threadIdx.x, 1); // answered: the chain warp
    uint64_t* const posted = s_bar;
    uint64_t* const answered = s_bar + 2;
    // Windows are numbered g = 0, 1, 2, ... across iterations; window g uses buffer g & 1 of the post / answer slots and of the
    // two barrier pairs, and is the (g >> 1)-th use of that buffer: the phase parity every waiter needs.
    unsigned g_chain = 0, g_p1 = 0, g_p2 = 0;
    __syncthreads();

    // ---- per frame warp
    const int f = chain_warp ? 0 : w_id;
    const int slot = blockIdx.x * F + f;
    float* const sv = s_var + (size_t)f * K;
    uint4* const cn = a.cn + (size_t)slot * M;
    float* const pv = a.pv + (size_t)slot * M;
    uint32_t* const pw = s_pw + f * (n_win + 1);
    int frame = -1, iters = 0;
    bool have_state = false;

    for (;;)
    {
        if (!chain_warp)
        {
            if (frame < 0)
            { // take the next frame of the queue into this slot
                int fr = 0;
                if (lane == 0) fr = atomicAdd(a.queue, 1);
                fr = __shfl_sync(FULL, fr, 0);
                if (fr < a.n_frames)
                {
                    frame = fr, iters = 0, have_state = false;
                    const float* src = a.llr + (size_t)fr * N_LDPC;
                    const float4* s4 = reinterpret_cast<const float4*>(src);
                    float4* d4 = reinterpret_cast<float4*>(sv);
                    for (int i = lane; i < K / 4; i += 32)
                        d4[i] = __ldcs(s4 + i);
                    for (int i = lane; i < M; i += 32)
                        pv[i] = __ldcs(src + K + i);
                }
            }
            if (lane == 0) s_frame[f] = frame;
        }
        __syncthreads();
        bool any = false;
#pragma unroll
        for (int k = 0; k < F; k++)
            any |= s_frame[k] >= 0;
        if (!any) break;

        if (chain_warp)
        {
            // ---- the parity chain of all F frames, lane = frame.  Check 0 has no incoming chain edge: the frame warps post
            //      m_in = -FLT_MAX for it and the chain starts from x = 0, so that t_in = x - m_in = FLT_MAX exactly, the
            //      neutral edge ldpc_kernel.cu uses, without a special case inside the dependent sequence.
            const int fl = lane < F ? lane : 0;
            float x = 0.f;
#pragma unroll 1
            for (int w = 0; w < n_win; w++, g_chain++)
            {
                const int b = g_chain & 1u;
                mbar_wait_acquire(posted + b, (g_chain >> 1) & 1u);
                const uint4* post = s_post + (size_t)b * W * F + fl;
                float* ans = s_ans + (size_t)b * W * F + fl;
                const int nk = min(W, M - w * W);
                uint4 p = post[0];
#pragma unroll 2
                for (int k = 0; k < nk; k++)
                {
                    const uint4 q = p;
                    if (k + 1 < nk) p = post[(k + 1) * F];
                    const float t_out = __uint_as_float(q.x), m_in = __uint_as_float(q.y), AnS = __uint_as_float(q.z);
                    const float ti = __fsub_rn(x, m_in);
                    const float u = NORM1 ? ti : __fmul_rn(ti, a.norm);
                    x = __fadd_rn(t_out, min_xorsign_abs(u, AnS));
                    if (lane < F) ans[k * F] = ti;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_release(answered + b);
            }
        }
        else
        {
            const bool live = frame >= 0;
            WinRegs<DI> RA, RB;
            LoadRegs<DI> LA, LB;

            // ---- requests of window w (index words, check state, the outgoing parity variable's old value)
            auto LD = [&](int w, LoadRegs<DI>& L) {
                const int c = w * W + lane;
                const uint32_t* it = a.idxT + ((size_t)w * IWP) * 32 + lane;
#pragma unroll
                for (int j = 0; j < IWP; j++)
                    L.ix[j] = __ldg(it + j * 32);
                L.st = make_uint4(0u, 0u, 0u, 0u);
                L.pvo = 0.f;
                if (lane < W && c < M && live)
                {
                    if (have_state) L.st = __ldcg(cn + c);
                    L.pvo = __ldcg(pv + c);
                }
            };

            // ---- pass 1 of window w: everything but the incoming chain edge (Update_rule_MS.hxx:57-67 on D - 1 edges)
            auto P1 = [&](int w, const LoadRegs<DI>& L, WinRegs<DI>& R) {
                const int b = g_p1 & 1u;
                g_p1++;
                const int c = w * W + lane;
                const float o1 = __uint_as_float(L.st.x), o2 = __uint_as_float(L.st.y);
                const unsigned osg = L.st.z, oeq = L.st.w;
                float m1 = FLT_MAX, m2 = FLT_MAX;
                unsigned sa = 0;
#pragma unroll
                for (int j = 0; j < IWP; j++)
                    R.ix[j] = L.ix[j];
#pragma unroll
                for (int i = 0; i < DI; i++)
                {
                    const unsigned off = (L.ix[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                    const float v = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(sv) + off);
                    const float mag = (oeq & (1u << i)) ? o1 : o2;
                    const float t = __fsub_rn(v, with_sign_of(mag, osg << (31 - i)));
                    R.t[i] = t;
                    fold(t, m1, m2, sa);
                }
                const float An = fmaxf(0.f, NORM1 ? m1 : __fmul_rn(m1, a.norm)); // min over the information edges, normalised
                const float AnS = with_sign_of(An, sa);
                const float t_out = __fsub_rn(L.pvo, with_sign_of((oeq & (1u << (DI + 1))) ? o1 : o2, osg << (31 - (DI + 1))));
                float m_in = with_sign_of((oeq & (1u << DI)) ? o1 : o2, osg << (31 - DI));
                if (c == 0) m_in = -FLT_MAX; // see the chain warp
                fold(t_out, m1, m2, sa);
                R.m1 = m1, R.m2 = m2, R.sa = sa, R.t_out = t_out;
                if (lane < W)
                    s_post[((size_t)b * W + lane) * F + f] = make_uint4(__float_as_uint(t_out), __float_as_uint(m_in), __float_as_uint(AnS), 0u);
                __syncwarp();
                if (lane == 0) mbar_arrive_release(posted + b);
            };

            // ---- pass 2 of window w: fold the chain edge in, new messages and variables (Update_rule_MS.hxx:82-92)
            auto P2 = [&](int w, WinRegs<DI>& R) {
                const int b = g_p2 & 1u;
                const int c = w * W + lane;
                const bool valid = lane < W && c < M;
                mbar_wait_acquire(answered + b, (g_p2 >> 1) & 1u);
                g_p2++;
                float ti = FLT_MAX;
                if (lane < W && c > 0) ti = s_ans[((size_t)b * W + lane) * F + f];
                float m1 = R.m1, m2 = R.m2;
                unsigned sa = R.sa;
                fold(ti, m1, m2, sa);
                const float c1 = fmaxf(0.f, NORM1 ? m2 : __fmul_rn(m2, a.norm)); // Update_rule_NMS.hxx:100-106
                const float c2 = fmaxf(0.f, NORM1 ? m1 : __fmul_rn(m1, a.norm));
                unsigned nsg = 0, neq = 0;
                const bool wr = valid && live;
#pragma unroll
                for (int i = 0; i < DI; i++)
                {
                    const float tt = R.t[i];
                    const bool e = fabsf(tt) == m1;
                    const unsigned xs = sa ^ __float_as_uint(tt);
                    const float nv = __fadd_rn(tt, with_sign_of(e ? c1 : c2, xs));
                    const unsigned off = (R.ix[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                    if (wr) *reinterpret_cast<float*>(reinterpret_cast<char*>(sv) + off) = nv;
                    neq |= (e ? 1u : 0u) << i;
                    nsg |= (xs >> 31) << i;
                }
                // incoming chain edge: check c is the last one to touch var[K+c-1]
                const bool e_in = fabsf(ti) == m1;
                const unsigned xs_in = sa ^ __float_as_uint(ti);
                const float nv_in = __fadd_rn(ti, with_sign_of(e_in ? c1 : c2, xs_in));
                neq |= (e_in ? 1u : 0u) << DI;
                nsg |= (xs_in >> 31) << DI;
                // outgoing chain edge: its new value is what the chain warp carried on; only the last check stores it itself
                const bool e_out = fabsf(R.t_out) == m1;
                const unsigned xs_out = sa ^ __float_as_uint(R.t_out);
                const float nv_out = __fadd_rn(R.t_out, with_sign_of(e_out ? c1 : c2, xs_out));
                neq |= (e_out ? 1u : 0u) << (DI + 1);
                nsg |= (xs_out >> 31) << (DI + 1);
                if (wr)
                {
                    if (c > 0) __stcg(pv + c - 1, nv_in);
                    if (c == M - 1) __stcg(pv + c, nv_out);
                    __stcg(cn + c, make_uint4(__float_as_uint(c1), __float_as_uint(c2), nsg, neq));
                }
                // hard decisions of the parity variables, for the syndrome: bit l of pw[w] = (var[K + w*W + l - 1] < 0)
                const unsigned word = __ballot_sync(FULL, valid && c > 0 && nv_in < 0.f);
                if (lane == 0) pw[w] = word;
                if (valid && c == M - 1) pw[n_win] = nv_out < 0.f ? 1u : 0u;
            };

            // three windows in flight per frame warp: requests of w + 2, pass 1 of w + 1, pass 2 of w
            LD(0, LA);
            if (1 < n_win) LD(1, LB);
            P1(0, LA, RA);
#pragma unroll 1
            for (int w = 0; w < n_win; w += 2)
            {
                if (w + 2 < n_win) LD(w + 2, LA);
                if (w + 1 < n_win) P1(w + 1, LB, RB);
                P2(w, RA);
                if (w + 1 < n_win)
                {
                    if (w + 3 < n_win) LD(w + 3, LB);
                    if (w + 2 < n_win) P1(w + 2, LA, RA);
                    P2(w + 1, RB);
                }
            }

            if (live)
            {
                // ---- syndrome on the signs of the updated variables (LDPC_syndrome.hxx:43-69), lanes = checks
                __syncwarp();
                bool unsat = false;
                uint32_t ia[IWP], ib[IWP]; // index words of the current / next window (requested one window ahead)
                auto LI = [&](int w, uint32_t* ix) {
                    const uint32_t* it = a.idxT + ((size_t)w * IWP) * 32 + lane;
#pragma unroll
                    for (int j = 0; j < IWP; j++)
                        ix[j] = __ldg(it + j * 32);
                };
                auto SY = [&](int w, const uint32_t* ix) {
                    const int c = w * W + lane;
                    const bool valid = lane < W && c < M;
                    unsigned par = 0;
#pragma unroll
                    for (int i = 0; i < DI; i++)
                    {
                        const unsigned off = (ix[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                        const float v = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(sv) + off);
                        par ^= v < 0.f ? 1u : 0u;
                    }
                    const unsigned cur = pw[w], nxt = pw[w + 1];
                    const unsigned hd_in = (cur >> lane) & 1u; // var[K+c-1]; 0 for check 0
                    const unsigned hd_out = (lane + 1 < W && c + 1 < M) ? (cur >> (lane + 1)) & 1u : nxt & 1u; // var[K+c]
                    par ^= hd_in ^ hd_out;
                    unsat |= valid && par != 0u;
                };
                LI(0, ia);
#pragma unroll 1
                for (int w = 0; w < n_win; w += 2)
                {
                    if (w + 1 < n_win) LI(w + 1, ib);
                    SY(w, ia);
                    if (w + 1 < n_win)
                    {
                        if (w + 2 < n_win) LI(w + 2, ia);
                        SY(w + 1, ib);
                    }
                }
                const bool bad = __any_sync(FULL, unsat);
                iters++;
                have_state = true;
                if ((a.early_stop && !bad) || iters >= a.n_ite)
                {
                    // ---- the frame is done: hard decisions of the information bits, packed 32 per word
                    const int KW = (K + 31) / 32;
                    uint32_t* ob = a.out_bits + (size_t)frame * KW;
                    for (int j0 = 0; j0 < KW; j0 += 32)
                    {
                        uint32_t keep = 0;
                        const int nj = min(32, KW - j0);
                        for (int jj = 0; jj < nj; jj++)
                        {
                            const int v = (j0 + jj) * 32 + lane;
                            const unsigned word = __ballot_sync(FULL, v < K && sv[v] < 0.f);
                            if (lane == jj) keep = word;
                        }
                        if (lane < nj) ob[j0 + lane] = keep;
                    }
                    if (lane == 0)
                    {
                        // without the syndrome test the reference reports every frame as valid (Decoder_LDPC_BP.hxx:8-20)
                        if (a.cwd) a.cwd[frame] = (int8_t)(!a.early_stop || !bad);
                        if (a.n_iter) a.n_iter[frame] = iters;
                        if (a.stats)
                        {
                            atomicAdd(a.stats + STAT_FRAMES, 1ull);
                            atomicAdd(a.stats + STAT_LDPC_ITERS, (unsigned long long)iters);
                            if (bad) atomicAdd(a.stats + STAT_LDPC_FAIL, 1ull);
                            atomicAdd(a.stats + STAT_ITER_HIST + min(iters, LDPC_HIST_BINS - 1), 1ull);
                        }
                    }
                    frame = -1;
                }
            }
        }
        __syncthreads();
    }
}

// per-frame packed bits [frames][KW] -> bit-sliced words out_hd[subgroup][stride] (what the BCH stage and the unpackers read)
__global__ void __launch_bounds__(128) bits_to_sliced_kernel(const uint32_t* __restrict__ bits, uint32_t* __restrict__ out, int KW,
                                                             int K, int stride, int n_frames)
{
    const int lane = threadIdx.x & 31;
    const int sg = blockIdx.y;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= KW) return;
    const int frame = sg * 32 + lane;
    const uint32_t w = frame < n_frames ? bits[(size_t)frame * KW + j] : 0u;
    uint32_t keep = 0;
#pragma unroll
    for (int b = 0; b < 32; b++)
    {
        const unsigned o = __ballot_sync(FULL, (w >> b) & 1u);
        if (lane == b) keep = o;
    }
    const int v = j * 32 + lane;
    if (v < K) out[(size_t)sg * stride + v] = keep;
}
} // namespace

int
ldpc_onchip_frames_per_cta(int d_info)
{
    return d_info == 25 ? 3 : 5;
}

// idxT[window][word][lane]: byte offsets (4 * variable) of the information variables of check window*W + lane, two per word
std::vector<uint32_t>
build_onchip_index(const LdpcTables& t)
{
    const int n_win = (t.M + W - 1) / W, iwp = (t.d_info + 1) / 2;
    std::vector<uint32_t> out((size_t)n_win * iwp * 32, 0u);
    for (int c = 0; c < t.M; c++)
        for (int i = 0; i < t.d_info; i++)
        {
            const uint32_t off = 4u * t.idx[(size_t)c * t.d_pad + i];
            out[((size_t)(c / W) * iwp + i / 2) * 32 + c % W] |= off << ((i & 1) * 16);
        }
    return out;
}

size_t
ldpc_onchip_smem_bytes(int K, int M, int F)
{
    const int n_win = (M + W - 1) / W;
    return (size_t)F * K * 4 + (size_t)2 * W * F * 16 + (size_t)2 * W * F * 4 + (((size_t)F * (n_win + 1) + 1) & ~(size_t)1) * 4 + 4 * 8 +
           (size_t)F * 4 + 16;
}

template<int DI, int F>
static cudaError_t
launch_onchip_t(const OnchipLaunch& a, cudaStream_t s)
{
    const size_t smem = ldpc_onchip_smem_bytes(a.K, a.M, F);
    const int blocks = std::max(1, std::min(a.n_sms, (a.n_frames + F - 1) / F));
    cudaError_t e = cudaMemsetAsync(a.queue, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    if (a.norm == 1.0f)
    {
        auto k = ldpc_onchip_kernel<DI, F, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<blocks, (F + 1) * 32, smem, s>>>(a);
    }
    else
    {
        auto k = ldpc_onchip_kernel<DI, F, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<blocks, (F + 1) * 32, smem, s>>>(a);
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t
launch_ldpc_onchip(const OnchipLaunch& a, cudaStream_t s)
{
    if (a.d_info == 25) return launch_onchip_t<25, 3>(a, s);
    if (a.d_info == 9) return launch_onchip_t<9, 5>(a, s);
    return cudaErrorInvalidValue;
}

cudaError_t
launch_bits_to_sliced(const uint32_t* bits, uint32_t* out_hd, int K, int stride, int n_frames, cudaStream_t s)
{
    const int KW = (K + 31) / 32;
    const dim3 grid((KW + 3) / 4, (n_frames + 31) / 32);
    bits_to_sliced_kernel<<<grid, 128, 0, s>>>(bits, out_hd, KW, K, stride, n_frames);
    count_launch();
    return cudaGetLastError();
}
} // namespace dvbs2
