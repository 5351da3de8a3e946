// Horizontal-layered (normalised) min-sum LDPC decoder for the DVB-S2 short-frame IRA codes, sm_100a.
//
// Semantics (bit-exact): lib/aff3ct/include/Module/Decoder/LDPC/BP/Horizontal_layered/
// Decoder_LDPC_BP_horizontal_layered.hxx:89-123,156-207 with Update_rule_MS.hxx:48-92 / Update_rule_NMS.hxx:14-56,
// 100-106 and LDPC_syndrome.hxx:43-69 — checks are visited in natural order c = 0..M-1 (the IRA dual diagonal makes
// that a strict chain), per-frame syndrome early termination.
//
// B200 mapping ("lane = frames"): a warp owns a group of 32*NF frames and walks the checks in order; lane l holds
// frames l, 32+l, ... of the group (NF of them, so every access is one NF-wide vector per lane).
//   * Variable-node values live in HBM as var[group][v][lane][NF]: every access of the warp is one contiguous chunk of
//     32*NF*sizeof(R) bytes (256 B for float).  The access pattern per frame is a gather, so HBM sees random chunks;
//     measured (profiles/r1_membench_random_rmw.txt) random read-modify-write of 128 B chunks tops out near 4.5 TB/s
//     on this part and 256 B chunks near 5.7 TB/s, which is why NF = 2 for float.
//   * Two checks sharing an information variable are never closer than 52 checks apart in these codes
//     (LdpcTables::min_reuse_distance), so the values of check c+1 can be requested before check c is computed: the
//     check loop is software-pipelined over two register sets, loads of the next check in flight during the arithmetic
//     of the current one.  Only the parity-chain value var[K+c] links consecutive checks, and that one is carried in a
//     register.
//   * The per-check index rows (and "last update" masks) sit in shared memory, loaded once per persistent CTA.
//   * Check-to-variable messages are NOT stored per edge.  Per check and frame we keep the two output magnitudes,
//     one sign bit and one "|t| == min1" bit per edge (16 B) — the reference's output rule is value-equality based,
//     so this reconstructs every message exactly.  Iteration 0 reads no check state at all (messages are zero).
//   * Hard decisions are kept bit-sliced across 32 frames: hd[subgroup][v] is a 32-bit word written with one
//     __ballot_sync when a variable receives its last update of the iteration, so the syndrome test after an
//     iteration is E word-XORs per warp instead of E loads per frame, and the BCH stage consumes the same words.
//   * A lane whose frame has converged stops writing; its bits were snapshotted into out_hd at that moment.
//   * Groups are handed to warps through an atomic queue (persistent CTAs, one per SM).
// The kernel is HBM-bound on message passing: bytes per frame and iteration = 2*sizeof(R)*(E_info + M) + 32*M
// (DESIGN.md §5).
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstring>
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
    __device__ static unsigned raw(C t) { return __float_as_uint(t); } // bit 31 = neg(t)
    __device__ static C from_bits(unsigned w) { return __uint_as_float(w); } // inverse of raw()
    // magnitude m >= 0 with the sign held in bit 31 of w
    __device__ static C apply_sign(C m, unsigned w) { return __uint_as_float(__float_as_uint(m) | (w & 0x80000000u)); }
    __device__ static unsigned pack(C a, C b, unsigned& hi)
    {
        hi = __float_as_uint(b);
        return __float_as_uint(a);
    }
    __device__ static void unpack(unsigned lo, unsigned hi, C& a, C& b) { a = __uint_as_float(lo), b = __uint_as_float(hi); }
    // element j of a raw vector of NF values / its inverse
    __device__ static C from_raw(const uint32_t* w, int j) { return __uint_as_float(w[j]); }
    __device__ static void to_raw(uint32_t* w, int j, C v) { w[j] = __float_as_uint(v); }
};

// Integer R: C++ integral promotion to int, truncation back to R on every store (wraps), exactly like the
// reference's sequential decoder instantiated with R = int16_t / int8_t.
template<typename R>
struct ArithInt
{
    using C = int;
    static constexpr int BITS = 8 * (int)sizeof(R);
    static constexpr int PER = 32 / BITS; // values per raw word
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
    __device__ static unsigned raw(C t) { return (unsigned)t; } // bit 31 = neg(t): values are sign-extended ints
    __device__ static C from_bits(unsigned w) { return (int)w; }  // inverse of raw()
    __device__ static C apply_sign(C m, unsigned w) { return wrap((int)w < 0 ? -m : m); }
    __device__ static unsigned pack(C a, C b, unsigned& hi)
    {
        hi = 0;
        return ((unsigned)a & 0xffffu) | ((unsigned)b << 16);
    }
    __device__ static void unpack(unsigned lo, unsigned, C& a, C& b) { a = (int)(short)(lo & 0xffffu), b = (int)(short)(lo >> 16); }
    __device__ static C from_raw(const uint32_t* w, int j) { return (int)(R)(w[j / PER] >> ((j % PER) * BITS)); }
    __device__ static void to_raw(uint32_t* w, int j, C v)
    {
        const uint32_t m = (BITS == 32) ? 0xffffffffu : ((1u << (BITS % 32)) - 1u);
        if (j % PER == 0) w[j / PER] = 0;
        w[j / PER] |= ((uint32_t)v & m) << ((j % PER) * BITS);
    }
};
template<>
struct Arith<int16_t> : ArithInt<int16_t>
{
};
template<>
struct Arith<int8_t> : ArithInt<int8_t>
{
};

// NF consecutive values of R as one 4/8/16-byte global access
template<int BYTES>
struct RawVec;
template<>
struct RawVec<4>
{
    using T = uint32_t;
};
template<>
struct RawVec<8>
{
    using T = uint2;
};
template<>
struct RawVec<16>
{
    using T = uint4;
};


// ---------------------------------------------------------------------------------------------- packed int16 pairs
// Two frames per 32-bit word (low / high half), sm_100a packed-halfword instructions (VIADD.16x2, VIMNMX.S16x2).  Every
// operation is the ring operation modulo 2^16 per half, i.e. exactly the reference's int16_t arithmetic with its wraps.
namespace pk
{
__device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return __vadd2(a, b); }
__device__ __forceinline__ uint32_t neg(uint32_t a) { return __vadd2(~a, 0x00010001u); }
__device__ __forceinline__ uint32_t vmin(uint32_t a, uint32_t b) { return __vmins2(a, b); }
__device__ __forceinline__ uint32_t vmax(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
// |t| with the reference's wrap (|-32768| = -32768): max(t, -t) as signed halves
__device__ __forceinline__ uint32_t absv(uint32_t t) { return __vmaxs2(t, neg(t)); }
// 0xffff in every half whose bit 15 is set
// (prmt's sign-replicate selector, bit 3 of a nibble; __byte_perm() masks that bit off, hence the PTX)
__device__ __forceinline__ uint32_t signmask(uint32_t x)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xbb99;" : "=r"(r) : "r"(x));
    return r;
}
// 0xffff in the low / high half if bit b / bit 16 + b of w is set (b = 0..15, compile-time)
__device__ __forceinline__ uint32_t bitmask(uint32_t w, int b) { return signmask(w << (15 - b)); }
// per half: m ? a : b
__device__ __forceinline__ uint32_t mux(uint32_t a, uint32_t b, uint32_t m) { return (a & m) | (b & ~m); }
__device__ __forceinline__ int lo(uint32_t x) { return (int)(short)(x & 0xffffu); }
__device__ __forceinline__ int hi(uint32_t x) { return (int)(short)(x >> 16); }
__device__ __forceinline__ uint32_t pair(int l, int h) { return ((uint32_t)l & 0xffffu) | ((uint32_t)h << 16); }
} // namespace pk

// ---------------------------------------------------------------------------------------------- check-state storage
// Compressed check state of one (check, frame) — or (check, pair of frames) in the packed int16 kernel — is four 32-bit words
// for the degree-27 checks of the rate-8/9 codes and THREE for the degree-11 checks of the rate-3/5 codes, whose sign and
// "== min1" bits (or sign bits and argmin index) fit one word: 12 instead of 16 bytes, -25 % of the state traffic, -7 % of
// all bytes these codes move per iteration.  Element e of a group = check * (frames or pairs per lane) + index in the lane.
//   SW = 4: [element][lane] uint4 (one 16-byte access per lane, 512 contiguous bytes per warp)
//   SW = 3: [element][word][lane] uint32 (three 128-byte rows per warp)
// ldpc_state_words(d_info) in kernels.h tells the host side which one a code uses.
template<int SW>
struct StateIO
{
    static __device__ __forceinline__ size_t group_words(int M, int per_lane) { return (size_t)M * per_lane * SW * 32; }
    // `at` = first word of the group + the lane's offset (lane_offset()); w[3] of the result is 0 for SW = 3
    static __device__ __forceinline__ int lane_offset(int lane) { return SW == 4 ? 4 * lane : lane; }
    static __device__ __forceinline__ uint4 load(const uint32_t* __restrict__ at, size_t e)
    {
        if constexpr (SW == 4)
            return __ldcg(reinterpret_cast<const uint4*>(at) + e * 32);
        else
        {
            const uint32_t* p = at + e * (3 * 32);
            return make_uint4(__ldcg(p), __ldcg(p + 32), __ldcg(p + 64), 0u);
        }
    }
    static __device__ __forceinline__ void store(uint32_t* __restrict__ at, size_t e, uint4 v)
    {
        if constexpr (SW == 4)
            __stcg(reinterpret_cast<uint4*>(at) + e * 32, v);
        else
        {
            uint32_t* p = at + e * (3 * 32);
            __stcg(p, v.x), __stcg(p + 32, v.y), __stcg(p + 64, v.z);
        }
    }
};
// degree D = d_info + 2 edges: the sign bits and the "== min1" bits (16 each), or the sign bits and a 5-bit index per half in the
// packed kernel (11 + 5), share a word when D <= 11
template<int DI>
constexpr int STATE_WORDS = (DI + 2 <= 11) ? 3 : 4;

// ---------------------------------------------------------------------------------------------- the decoder
template<typename R, int DI, int NF>
struct CheckRegs
{
    typename RawVec<sizeof(R) * NF>::T v[DI + 1]; // information variables of the check, then var[K+c]
    uint4 st[NF];                                 // compressed check state of the previous iteration
};

// PK (int16_t, NF = 4 only): the check loop works on packed pairs of frames — lane l holds frames (l, 32 + l) in word 0
// and (64 + l, 96 + l) in word 1 of its 8-byte vector, so a variable is again one 256-byte chunk per warp but for 128 frames,
// and every arithmetic instruction serves two frames.  Check state per pair and check (one uint4): the two output
// magnitudes as pairs, the sign bits of edges 0..15 (bit e: first frame, bit 16 + e: second frame), the sign bits of edges
// 16.. in the same arrangement with the index of one edge attaining min1 in bits 11..15 / 27..31 (with ties min2 = min1,
// so any such edge reproduces the reference's value-equality rule).
template<typename R, int DI, int NF, bool PK = false>
__global__ void __launch_bounds__(LDPC_WARPS * 32, 1) ldpc_decode_kernel(const LdpcLaunch a)
{
    static_assert(!PK || (sizeof(R) == 2 && (NF == 4 || NF == 2)), "packed pairs: int16_t, two or four frames per lane");
    using A = Arith<R>;
    using C = typename A::C;
    using Raw = typename RawVec<sizeof(R) * NF>::T;
    constexpr int NW = sizeof(Raw) / 4;  // 32-bit words per raw vector
    constexpr int D = DI + 2;            // edges per check incl. the two parity edges (bit DI = in, DI+1 = out)
    constexpr int IW = (DI + 1) / 2;     // u32 words per index row
    constexpr unsigned VB = 32 * NF * sizeof(R); // bytes between consecutive variables of a group
    using HdVec = typename RawVec<4 * NF>::T;    // the NF hard-decision words of one variable

    extern __shared__ uint32_t smem[];
    const int K = a.K, M = a.M;
    uint32_t* s_idx = smem;                    // [M][IW]
    uint32_t* s_last = smem + (size_t)M * IW;  // [M]
    for (int i = threadIdx.x; i < M * IW; i += blockDim.x)
        s_idx[i] = a.idx32[i];
    for (int i = threadIdx.x; i < M; i += blockDim.x)
        s_last[i] = a.last[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;

    // slots of this launch: round 1 decodes frames in place (slot = frame); round 2 decodes the survivors that the
    // compaction step packed into fresh slots, their count known only on the device
    const int n_slots = a.n_slots_dev ? min(*a.n_slots_dev, a.n_frames) : a.n_frames;
    const int n_groups = (n_slots + 32 * NF - 1) / (32 * NF);
    const bool resumed = a.slot_iters != nullptr; // resumed frames carry check state from their first round

    for (;;)
    {
        int g = 0;
        if (lane == 0) g = atomicAdd(a.queue, 1);
        g = __shfl_sync(FULL, g, 0);
        if (g >= n_groups) break;

        // byte offsets inside a group fit 32 bits: (warp-uniform base) + variable * VB + lane offset
        char* __restrict__ vbytes = (char*)a.var + (size_t)g * N_LDPC * VB;
        const unsigned loff = lane * NF * (unsigned)sizeof(R);
        constexpr int SW = STATE_WORDS<DI>;
        using SIO = StateIO<SW>;
        uint32_t* __restrict__ cn = reinterpret_cast<uint32_t*>(a.cn) + (size_t)g * SIO::group_words(M, PK ? NF / 2 : NF) + SIO::lane_offset(lane);
        uint32_t* __restrict__ hd = a.hd + (size_t)g * NF * N_LDPC;       // running hard decisions, [v][NF] words
        uint32_t* __restrict__ out_hd = a.out_hd + (size_t)g * NF * N_LDPC; // final ones, [NF][v] (subgroup-major)

        unsigned active[NF], done[NF], failed[NF];
        int my_iters[NF];
        bool any_active = false;
#pragma unroll
        for (int j = 0; j < NF; j++)
        {
            const int slot = (g * NF + j) * 32 + lane;
            active[j] = __ballot_sync(FULL, slot < n_slots);
            done[j] = failed[j] = 0;
            my_iters[j] = (resumed && slot < n_slots) ? a.slot_iters[slot] : 0;
            any_active |= active[j] != 0;
        }
        bool exported = false;

        for (int ite = 0; any_active; ite++)
        {
            bool mine[NF];
#pragma unroll
            for (int j = 0; j < NF; j++)
                mine[j] = (active[j] >> lane) & 1;
            bool mine_any = false;
#pragma unroll
            for (int j = 0; j < NF; j++)
                mine_any |= mine[j];

            if constexpr (PK)
            {
                constexpr int NP = NF / 2;    // pairs per lane (2 in round 1; 1 in the narrower groups of the later rounds)
                constexpr uint32_t ONE = 0x00010001u, MAXV = 0x7fff7fffu;
                const bool have_state = resumed || ite > 0;
                bool pmine[NP];
#pragma unroll
                for (int p = 0; p < NP; p++)
                    pmine[p] = mine[2 * p] || mine[2 * p + 1];
                uint32_t chain[NP];
#pragma unroll
                for (int p = 0; p < NP; p++)
                    chain[p] = MAXV;
                // x[i]: what the warp holds for edge i of the current check — first the loaded variable (both pairs), then, in
                // place, t = var - msg.  Pass 2 consumes x[i] and immediately requests edge i of the NEXT check into the same
                // registers (hazard-free: checks c and c+1 share no information variable, see header), so the loads of check
                // c+1 trickle out during pass 2 of check c and only ~D vector registers are live instead of 2 D.
                uint2 x[D];
                uint4 rst[NP];
                auto load_var = [&](int c, int i) { // i < DI: i-th information variable of check c;  i == DI + 1: var[K + c]
                    unsigned vi = (unsigned)(K + c);
                    if (i < DI) vi = (s_idx[c * IW + (i >> 1)] >> ((i & 1) * 16)) & 0xffffu;
                    if constexpr (NP == 2)
                        x[i] = __ldcg(reinterpret_cast<const uint2*>(vbytes + (vi * VB + loff)));
                    else
                        x[i].x = __ldcg(reinterpret_cast<const uint32_t*>(vbytes + (vi * VB + loff)));
                };
                auto load_state = [&](int c) {
#pragma unroll
                    for (int p = 0; p < NP; p++)
                    {
                        rst[p] = have_state ? SIO::load(cn, (size_t)(c * NP + p)) : make_uint4(0u, 0u, 0u, 0u);
                        if constexpr (SW == 3) rst[p].w = rst[p].z; // one word: signs in bits 0..10, argmin index in bits 11..15 of each half
                    }
                };
#pragma unroll
                for (int i = 0; i < D; i++)
                    if (i != DI) load_var(0, i);
                load_state(0);
#pragma unroll 1
                for (int c = 0; c < M; c++)
                {
                    const uint32_t* row = s_idx + c * IW;
                    const uint32_t lastm = s_last[c];
                    const bool more = c + 1 < M;
                    uint32_t c1[NP], c2[NP], sacc[NP], F0[NP], F1[NP], nidx[NP];
#pragma unroll
                    for (int p = 0; p < NP; p++)
                    {
                        const uint32_t o1 = rst[p].x, o2 = rst[p].y, S0 = rst[p].z, S1 = rst[p].w;
                        uint32_t ia = 0, ib = 0; // an edge attaining min1, per half (the last one; any will do, see header)
                        const uint32_t no1 = pk::neg(o1), no2 = pk::neg(o2);
                        const uint32_t ohA = 1u << ((S1 >> 11) & 31u), ohB = 1u << (S1 >> 27);
                        const uint32_t E0 = __byte_perm(ohA, ohB, 0x5410), E1 = __byte_perm(ohA, ohB, 0x7632);
                        uint32_t sa = 0, m1 = MAXV, m2 = MAXV;
#pragma unroll
                        for (int i = 0; i < D; i++)
                        {
                            const uint32_t v = (i == DI) ? chain[p] : (p == 0 ? x[i].x : x[i].y);
                            const int b = i & 15;
                            const uint32_t sm = pk::bitmask(i < 16 ? S0 : S1, b), em = pk::bitmask(i < 16 ? E0 : E1, b);
                            // t = v - msg (Update_rule_MS.hxx:57-67): add -mag for a positive message, +mag for a negative one
                            uint32_t ti = pk::add(v, pk::mux(pk::mux(o1, o2, em), pk::mux(no1, no2, em), sm));
                            if (i == DI && c == 0) ti = MAXV; // check 0 has no incoming chain edge
                            if (p == 0)
                                x[i].x = ti;
                            else
                                x[i].y = ti;
                            const uint32_t ab = pk::absv(ti);
                            sa ^= ti;
                            m2 = pk::vmin(m2, pk::vmax(ab, m1));
                            bool ph, pl;
                            m1 = __vibmin_s16x2(ab, m1, &ph, &pl); // min + "ab <= m1" per half in one VIMNMX
                            ia = pl ? (uint32_t)i : ia;
                            ib = ph ? (uint32_t)i : ib;
                        }
                        sacc[p] = sa;
                        nidx[p] = ia | (ib << 16);
                        { // one-hot masks of those edges in the sign-word arrangement (edges 0..15 / 16.., first | second frame)
                            const uint32_t ha = 1u << ia, hb = 1u << ib;
                            F0[p] = __byte_perm(ha, hb, 0x5410), F1[p] = __byte_perm(ha, hb, 0x7632);
                        }
                        uint32_t n2 = m2, n1 = m1;
                        if (a.k8 != 8)
                        { // Update_rule_NMS.hxx:14-39, per half
                            n2 = pk::pair(Arith<int16_t>::normalize(pk::lo(m2), a.norm, a.k8), Arith<int16_t>::normalize(pk::hi(m2), a.norm, a.k8));
                            n1 = pk::pair(Arith<int16_t>::normalize(pk::lo(m1), a.norm, a.k8), Arith<int16_t>::normalize(pk::hi(m1), a.norm, a.k8));
                        }
                        c1[p] = pk::vmax(0u, n2), c2[p] = pk::vmax(0u, n1); // Update_rule_NMS.hxx:100-106
                    }
                    if (more) load_state(c + 1);

                    uint32_t ns0[NP], ns1[NP], nc1[NP], nc2[NP];
#pragma unroll
                    for (int p = 0; p < NP; p++)
                    {
                        ns0[p] = ns1[p] = 0;
                        nc1[p] = pk::neg(c1[p]), nc2[p] = pk::neg(c2[p]);
                    }
#pragma unroll
                    for (int i = 0; i < D; i++)
                    {
                        uint32_t nv[NP];
                        const int b = i & 15;
#pragma unroll
                        for (int p = 0; p < NP; p++)
                        {
                            const uint32_t ti = p == 0 ? x[i].x : x[i].y;
                            // the edge that attained min1 gets the magnitude built from min2 (with ties min2 = min1, so every
                            // edge with |t| == min1 gets the same value, as Update_rule_MS.hxx:82-92 demands)
                            const uint32_t em = pk::bitmask(i < 16 ? F0[p] : F1[p], b);
                            const uint32_t xs = sacc[p] ^ ti; // bits 15 / 31: sign of the new message
                            const uint32_t sm = pk::signmask(xs);
                            // t + sign * mag: the signed magnitude is picked from {c1, c2, -c1, -c2} with three bitwise selects
                            nv[p] = pk::add(ti, pk::mux(pk::mux(nc1[p], c1[p], sm), pk::mux(nc2[p], c2[p], sm), em));
                            const uint32_t sb = sm & (ONE << b); // one LOP3 with the accumulation below
                            if (i < 16)
                                ns0[p] |= sb;
                            else
                                ns1[p] |= sb;
                        }
                        if (i != DI && more) load_var(c + 1, i); // x[i] is free from here on
                        unsigned vi;
                        bool ballot;
                        if (i < DI)
                        {
                            vi = (row[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                            ballot = (lastm >> i) & 1u;
                        }
                        else if (i == DI)
                        {
                            if (c == 0) continue;
                            vi = K + c - 1;
                            ballot = true;
                        }
                        else
                        {
#pragma unroll
                            for (int p = 0; p < NP; p++)
                                chain[p] = nv[p];
                            if (c != M - 1) continue;
                            vi = K + c;
                            ballot = true;
                        }
                        if (mine_any)
                        {
                            if constexpr (NP == 2)
                                __stcg(reinterpret_cast<uint2*>(vbytes + (vi * VB + loff)), make_uint2(nv[0], nv[NP - 1]));
                            else
                                __stcg(reinterpret_cast<uint32_t*>(vbytes + (vi * VB + loff)), nv[0]);
                        }
                        if (ballot) // warp-uniform: only the last update of a variable in the iteration fixes its hard decision
                        {
                            uint32_t hww[NF];
#pragma unroll
                            for (int p = 0; p < NP; p++)
                            {
                                hww[2 * p] = __ballot_sync(FULL, (nv[p] & 0x8000u) != 0);
                                hww[2 * p + 1] = __ballot_sync(FULL, (int)nv[p] < 0);
                            }
                            HdVec hw;
                            memcpy(&hw, hww, sizeof(HdVec));
                            if (lane == 0) __stcg(reinterpret_cast<HdVec*>(hd + (size_t)vi * NF), hw);
                        }
                    }
#pragma unroll
                    for (int p = 0; p < NP; p++)
                        if (pmine[p])
                        {
                            const uint32_t w3 = ns1[p] | ((nidx[p] & 0x001f001fu) << 11);
                            SIO::store(cn, (size_t)(c * NP + p), SW == 3 ? make_uint4(c1[p], c2[p], ns0[p] | w3, 0u) : make_uint4(c1[p], c2[p], ns0[p], w3));
                        }
                }
            }
            else
            {
            C chain[NF]; // var[K+c-1] after check c-1; for c = 0 a neutral dummy edge
#pragma unroll
            for (int j = 0; j < NF; j++)
                chain[j] = A::max_value();
            const bool have_state = resumed || ite > 0;

            // ---- requests everything check c reads from HBM (hazard-free ahead of time, see header)
            auto issue = [&](CheckRegs<R, DI, NF>& r, int c) {
                const uint32_t* row = s_idx + c * IW;
#pragma unroll
                for (int i = 0; i < DI; i++)
                {
                    const unsigned vi = (row[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                    r.v[i] = __ldcg(reinterpret_cast<const Raw*>(vbytes + (vi * VB + loff)));
                }
                r.v[DI] = __ldcg(reinterpret_cast<const Raw*>(vbytes + ((unsigned)(K + c) * VB + loff)));
#pragma unroll
                for (int j = 0; j < NF; j++)
                {
                    r.st[j] = have_state ? SIO::load(cn, (size_t)(c * NF + j)) : make_uint4(0u, 0u, 0u, 0u);
                    if constexpr (SW == 3) r.st[j].w = r.st[j].z >> 16, r.st[j].z &= 0xffffu; // sign bits | "== min1" bits << 16
                }
            };

            // One register set: pass 1 turns the loaded values into t = var - msg (Update_rule_MS.hxx:57-67), which frees
            // them for the loads of check c+1; those are in flight during pass 2 (new messages, stores; :82-92 and
            // ...layered.hxx:198-203).  Sign and "== min1" bits of edge i live at bit D-1-i of the state words.
            CheckRegs<R, DI, NF> r;
            issue(r, 0);
#pragma unroll 1
            for (int c = 0; c < M; c++)
            {
                const uint32_t* row = s_idx + c * IW;
                const uint32_t lastm = s_last[c];
                C t[D][NF];
                C min1[NF], c1[NF], c2[NF];
                unsigned sacc[NF];
#pragma unroll
                for (int j = 0; j < NF; j++)
                {
                    C o1, o2;
                    A::unpack(r.st[j].x, r.st[j].y, o1, o2);
                    const unsigned osg = r.st[j].z, oeq = r.st[j].w;
                    unsigned sa = 0;
                    C m1 = A::max_value(), m2 = A::max_value();
#pragma unroll
                    for (int i = 0; i < D; i++)
                    {
                        C v;
                        if (i == DI)
                            v = chain[j];
                        else
                        {
                            uint32_t w[NW];
                            memcpy(w, &r.v[i < DI ? i : DI], sizeof(Raw));
                            v = A::from_raw(w, j);
                        }
                        constexpr int dummy = 0;
                        (void)dummy;
                        const int b = D - 1 - i;
                        const C mag = (oeq & (1u << b)) ? o1 : o2;
                        C ti = A::sub(v, A::apply_sign(mag, osg << (31 - b)));
                        if (i == DI && c == 0) ti = A::max_value(); // check 0 has no incoming chain edge
                        t[i][j] = ti;
                        const C ab = A::absv(ti);
                        sa ^= A::raw(ti);
                        m2 = A::vmin(m2, A::vmax(ab, m1));
                        m1 = A::vmin(m1, ab);
                    }
                    min1[j] = m1, sacc[j] = sa;
                    c1[j] = A::vmax((C)0, A::normalize(m2, a.norm, a.k8)); // Update_rule_NMS.hxx:100-106
                    c2[j] = A::vmax((C)0, A::normalize(m1, a.norm, a.k8));
                }

                if (c + 1 < M) issue(r, c + 1);

                unsigned nsg[NF], neq[NF];
#pragma unroll
                for (int j = 0; j < NF; j++)
                    nsg[j] = neq[j] = 0;
#pragma unroll
                for (int i = 0; i < D; i++)
                {
                    C nv[NF];
#pragma unroll
                    for (int j = 0; j < NF; j++)
                    {
                        const C ti = t[i][j];
                        const bool e = A::absv(ti) == min1[j];
                        const unsigned x = sacc[j] ^ A::raw(ti); // bit 31: sign of the new message
                        nv[j] = A::add(ti, A::apply_sign(e ? c1[j] : c2[j], x));
                        if (e) neq[j] |= 1u << (D - 1 - i);
                        nsg[j] = __funnelshift_l(x, nsg[j], 1);
                    }
                    unsigned vi;
                    bool ballot;
                    if (i < DI)
                    {
                        vi = (row[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                        ballot = (lastm >> i) & 1u;
                    }
                    else if (i == DI)
                    {
                        if (c == 0) continue;
                        vi = K + c - 1;
                        ballot = true; // check c is the last one to touch parity variable K+c-1
                    }
                    else
                    {
#pragma unroll
                        for (int j = 0; j < NF; j++)
                            chain[j] = nv[j]; // consumed (and stored) by check c+1; only the last check stores it itself
                        if (c != M - 1) continue;
                        vi = K + c;
                        ballot = true;
                    }
                    if (mine_any)
                    { // a finished frame sharing the lane gets a value it never reads again
                        uint32_t w[NW];
#pragma unroll
                        for (int j = 0; j < NF; j++)
                            A::to_raw(w, j, nv[j]);
                        Raw out;
                        memcpy(&out, w, sizeof(Raw));
                        __stcg(reinterpret_cast<Raw*>(vbytes + (vi * VB + loff)), out);
                    }
                    // the votes are unconditional (no divergence scaffolding around them); one predicated NF-word store
                    HdVec hw;
                    uint32_t hww[NF];
#pragma unroll
                    for (int j = 0; j < NF; j++)
                        hww[j] = __ballot_sync(FULL, A::lt0(nv[j]));
                    memcpy(&hw, hww, sizeof(HdVec));
                    if (ballot && lane == 0) __stcg(reinterpret_cast<HdVec*>(hd + (size_t)vi * NF), hw);
                }
#pragma unroll
                for (int j = 0; j < NF; j++)
                    if (mine[j])
                    {
                        unsigned hi;
                        const unsigned lo = A::pack(c1[j], c2[j], hi);
                        SIO::store(cn, (size_t)(c * NF + j), SW == 3 ? make_uint4(lo, hi, nsg[j] | (neq[j] << 16), 0u) : make_uint4(lo, hi, nsg[j], neq[j]));
                    }
            }
            }
#pragma unroll
            for (int j = 0; j < NF; j++)
                my_iters[j] += mine[j] ? 1 : 0;

            // ---- syndrome on the bit-sliced hard decisions: lane l tests checks l, l+32, ...
            __syncwarp();
            unsigned unsat[NF];
#pragma unroll
            for (int j = 0; j < NF; j++)
                unsat[j] = 0;
            for (int c = lane; c < M; c += 32)
            {
                const uint32_t* row = s_idx + c * IW;
                uint32_t p[NF];
                auto fold = [&](unsigned v) {
                    const HdVec h = __ldcg(reinterpret_cast<const HdVec*>(hd + (size_t)v * NF));
                    uint32_t w[NF];
                    memcpy(w, &h, sizeof(HdVec));
#pragma unroll
                    for (int j = 0; j < NF; j++)
                        p[j] ^= w[j];
                };
#pragma unroll
                for (int j = 0; j < NF; j++)
                    p[j] = 0;
                fold(K + c);
                if (c > 0) fold(K + c - 1);
#pragma unroll
                for (int i = 0; i < DI; i++)
                    fold((row[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
#pragma unroll
                for (int j = 0; j < NF; j++)
                    unsat[j] |= p[j];
            }
            any_active = false;
            int n_active = 0;
#pragma unroll
            for (int j = 0; j < NF; j++)
            {
                unsat[j] = __reduce_or_sync(FULL, unsat[j]);
                const unsigned spent = __ballot_sync(FULL, mine[j] && my_iters[j] >= a.n_ite); // iteration budget used up
                unsigned fin = a.early_stop ? (active[j] & ~unsat[j]) : 0u;
                fin |= spent;
                if (fin)
                {
                    // snapshot the finishing lanes' hard decisions
                    uint32_t* oj = out_hd + (size_t)j * N_LDPC;
                    for (int v = lane; v < N_LDPC; v += 32)
                    {
                        const unsigned w = __ldcg(hd + (size_t)v * NF + j);
                        const unsigned o = done[j] ? __ldcg(oj + v) : 0u;
                        __stcg(oj + v, (o & done[j]) | (w & fin));
                    }
                    failed[j] |= fin & unsat[j];
                    done[j] |= fin;
                    active[j] &= ~fin;
                }
                any_active |= active[j] != 0;
                n_active += __popc(active[j]);
            }
            __syncwarp();
            // ---- hand the stragglers over to the compaction step once most of the group has converged: they continue in
            //      a dense group instead of keeping this warp (and its 256-byte chunks) busy for a few live lanes
            if (a.exit_below > 0 && n_active > 0 && n_active < a.exit_below)
            {
                int left = 0; // iterations the group's live frames may still run (they differ after a compaction)
#pragma unroll
                for (int j = 0; j < NF; j++)
                    if ((active[j] >> lane) & 1) left = max(left, a.n_ite - my_iters[j]);
                if (__reduce_max_sync(FULL, left) >= 2)
                {
                    exported = true;
                    break;
                }
            }
        }

        // survivors of this group (none unless exported): mask per subgroup + iterations already spent per slot
        if (a.surv_mask)
        {
#pragma unroll
            for (int j = 0; j < NF; j++)
            {
                const unsigned m = exported ? active[j] : 0u;
                if (lane == 0) a.surv_mask[g * NF + j] = m;
                if ((m >> lane) & 1) a.surv_iters[(g * NF + j) * 32 + lane] = my_iters[j];
            }
        }

        int it_sum = 0, n_valid = 0, n_failed = 0;
#pragma unroll
        for (int j = 0; j < NF; j++)
        {
            const int slot = (g * NF + j) * 32 + lane;
            const bool fin = (done[j] >> lane) & 1;
            if (fin)
            {
                const int frame = a.slot_frame ? a.slot_frame[slot] : slot;
                // without the syndrome test the reference reports every frame as valid (Decoder_LDPC_BP.hxx:8-20: status 0, CWD 1)
                const bool ok = !a.early_stop || !((failed[j] >> lane) & 1);
                if (a.cwd) a.cwd[frame] = (int8_t)ok;
                if (a.n_iter) a.n_iter[frame] = my_iters[j];
                if (a.stats) atomicAdd(a.stats + STAT_ITER_HIST + min(my_iters[j], LDPC_HIST_BINS - 1), 1ull);
            }
            it_sum += __reduce_add_sync(FULL, fin ? my_iters[j] : 0);
            n_valid += __popc(done[j]);
            n_failed += __popc(failed[j] & done[j]);
        }
        if (lane == 0 && a.stats)
        {
            atomicAdd(a.stats + STAT_FRAMES, (unsigned long long)n_valid);
            atomicAdd(a.stats + STAT_LDPC_ITERS, (unsigned long long)it_sum);
            atomicAdd(a.stats + STAT_LDPC_FAIL, (unsigned long long)n_failed);
        }
    }
}

// ---------------------------------------------------------------------------------------------- small batches
// Cooperative variant for small batches (float, one frame per lane): a CTA decodes ONE 32-frame group.  Everything a check does
// except its incoming parity-chain edge is independent of the chain (the information variables of checks closer than
// min_reuse_distance are disjoint, and var[K+c] still holds last iteration's value).  So the work is split by role:
//   * COOP_WORKERS worker warps: worker u takes checks u, u + 15, ...; it loads its check 15 checks ahead, turns the loads into
//     t = var - msg for every edge but the chain edge, reduces them to (min1, min2, sign parity) and posts those, the outgoing
//     parity edge's t and the chain edge's old message in its shared-memory slot; when the chain warp has answered, it computes
//     and stores the new messages of all its edges (pass 2) and the compressed check state.
//   * ONE chain warp walks the checks in natural order with var[K+c-1] in a register: per check it reads the five posted words,
//     folds the chain edge in, forms the two output magnitudes and var[K+c], and posts (t_chain, min1, c1, c2, sign parity)
//     back — about ten dependent instructions per check instead of a whole check (~2400 cycles: HBM latency + 900 instructions),
//     which is what bounds a call when there are too few groups to fill the machine.
// Slots are guarded by mbarriers (one "posted", one "answered" per worker), so waiting warps are suspended by the hardware
// rather than spinning.  Same arithmetic, same order of the operations that do not commute (min / max / XOR are exact and
// order-free), so results are bit-identical to the kernel above.  One round only (no frame compaction).
constexpr int COOP_WORKERS = 15;
constexpr int COOP_WARPS = COOP_WORKERS + 1;
// Hazard bound of the worker pipeline: a worker requests the variables of check c + COOP_WORKERS right after posting check c,
// while the answers (and hence the stores) of checks c - 2 * COOP_WORKERS + 1 .. c + COOP_WORKERS - 1 may still be outstanding
// (its own previous check has stored; every other worker can be one full turn behind).  So no two checks closer than
// 3 * COOP_WORKERS may share an information variable.  The DVB-S2 short-frame tables give 55 (rate 8/9) and 52 (rate 3/5);
// dvbs2_cuda_create turns the cooperative kernel off for a table that does not satisfy this (ldpc_coop_min_reuse()).
constexpr int COOP_MIN_REUSE = 3 * COOP_WORKERS;
// Ordering of the slot words (what compute-sanitizer's racecheck cannot see, profiles/r2_racecheck.log): per worker u there is
// one single-count mbarrier per direction.
//   post:    every lane of worker u writes sl[0..4]; __syncwarp(); lane 0: mbarrier.arrive.release.cta(posted[u])
//            chain warp: mbarrier.try_wait.parity.acquire.cta(posted[u], n-th phase) -> reads sl[0..4]               (RAW ordered)
//   answer:  every lane of the chain warp writes sl[5..9]; __syncwarp(); lane 0: arrive.release(answered[u])
//            worker u: try_wait.acquire(answered[u], n-th phase) -> reads sl[5..9]                                   (RAW ordered)
//   reuse:   worker u writes the post of its NEXT check only after it has read the answer to the current one, and the chain
//            warp read that check's post before it answered; the chain warp writes the next answer for u only after fetching
//            u's next post, which u made after reading the previous answer                                            (WAR ordered)
// __syncwarp() orders the lanes' stores before lane 0's release (bar.warp.sync is a morally-strong synchronisation, release
// is cumulative), so one arrival per warp suffices.  Phase parity: the n-th completion of a single-count barrier is awaited with
// parity n & 1, and a waiter never runs more than one phase ahead of the arriver (it consumed phase n - 1 before).

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

// Compressed check state of the cooperative kernel (one frame per lane, its own layout: the kernel runs one round, nothing else
// reads it).  float: the words of ldpc_decode_kernel (STATE_WORDS<DI>).  int16_t: three words per (check, frame) — both output
// magnitudes in one, then the sign bits and the "== min1" bits (sharing a word when the check has at most 16 edges).
template<typename R, int DI>
struct CoopState
{
    using A = Arith<R>;
    using C = typename A::C;
    static constexpr bool FLT = sizeof(R) == 4;
    static constexpr int D = DI + 2;
    static constexpr int SW = FLT ? STATE_WORDS<DI> : 3;
    static __device__ __forceinline__ void decode(const uint4 st, C& o1, C& o2, unsigned& osg, unsigned& oeq)
    {
        if constexpr (FLT)
        {
            A::unpack(st.x, st.y, o1, o2);
            if constexpr (SW == 4)
                osg = st.z, oeq = st.w;
            else
                osg = st.z & 0xffffu, oeq = st.z >> 16;
        }
        else
        {
            A::unpack(st.x, 0u, o1, o2);
            if constexpr (D > 16)
                osg = st.y, oeq = st.z;
            else
                osg = st.y & 0xffffu, oeq = st.y >> 16;
        }
    }
    static __device__ __forceinline__ uint4 encode(const C c1, const C c2, const unsigned nsg, const unsigned neq)
    {
        unsigned hi;
        const unsigned lo = A::pack(c1, c2, hi);
        if constexpr (FLT)
            return SW == 4 ? make_uint4(lo, hi, nsg, neq) : make_uint4(lo, hi, nsg | (neq << 16), 0u);
        else
            return D > 16 ? make_uint4(lo, nsg, neq, 0u) : make_uint4(lo, nsg | (neq << 16), 0u, 0u);
    }
};

template<typename R, int DI>
__global__ void __launch_bounds__(COOP_WARPS * 32, 1) ldpc_decode_coop_kernel(const LdpcLaunch a)
{
    static_assert(sizeof(R) == 4 || sizeof(R) == 2, "float (the dvbs2 binaries) and int16_t (AFF3CT's functional fixed-point path)");
    using A = Arith<R>;
    using C = typename A::C;
    using ST = CoopState<R, DI>;
    constexpr int NWK = COOP_WORKERS;
    constexpr int D = DI + 2;        // edges per check incl. the two parity edges (bit DI = in, DI+1 = out)
    constexpr int IW = (DI + 1) / 2; // u32 words per index row
    constexpr unsigned VB = 32 * sizeof(R);
    constexpr int SLOT = 10 * 32;    // words per worker slot: 5 posted + 5 answered, 32 lanes each

    extern __shared__ uint32_t smem[];
    const int K = a.K, M = a.M;
    uint32_t* s_idx = smem;                   // [M][IW]
    uint32_t* s_last = smem + (size_t)M * IW; // [M]
    uint64_t* s_posted = reinterpret_cast<uint64_t*>(smem + (((size_t)M * IW + M + 1) & ~(size_t)1)); // [NWK]
    uint64_t* s_answered = s_posted + NWK;                                                              // [NWK]
    uint32_t* s_slot = reinterpret_cast<uint32_t*>(s_answered + NWK);                                   // [NWK][SLOT]
    unsigned* s_unsat = s_slot + NWK * SLOT; // [1] OR of the unsatisfied-check masks of all warps
    int* s_group = reinterpret_cast<int*>(s_unsat + 1);
    for (int i = threadIdx.x; i < M * IW; i += blockDim.x)
        s_idx[i] = a.idx32[i];
    for (int i = threadIdx.x; i < M; i += blockDim.x)
        s_last[i] = a.last[i];
    if (threadIdx.x < 2 * NWK) mbar_init(s_posted + threadIdx.x, 1);
    __syncthreads();

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool chain_warp = w == NWK; // the last warp walks the chain, warps 0 .. NWK-1 are workers
    const int n_groups = (a.n_frames + 31) / 32;
    unsigned n_turns = 0;   // worker: checks posted so far (phase parity of its two barriers)
    unsigned visits = 0;    // chain warp, lane u: checks of worker u answered so far

    for (;;)
    {
        if (threadIdx.x == 0) *s_group = atomicAdd(a.queue, 1);
        __syncthreads();
        const int g = *s_group;
        __syncthreads();
        if (g >= n_groups) break;

        char* __restrict__ vbytes = (char*)a.var + (size_t)g * N_LDPC * VB;
        const unsigned loff = lane * (unsigned)sizeof(R);
        using SIO = StateIO<ST::SW>;
        uint32_t* __restrict__ cn = reinterpret_cast<uint32_t*>(a.cn) + (size_t)g * SIO::group_words(M, 1) + SIO::lane_offset(lane);
        uint32_t* __restrict__ hd = a.hd + (size_t)g * N_LDPC;
        uint32_t* __restrict__ out_hd = a.out_hd + (size_t)g * N_LDPC;

        const int slot = g * 32 + lane;
        unsigned active = __ballot_sync(FULL, slot < a.n_frames), done = 0, failed = 0; // identical in every warp
        int my_iters = 0;

        for (int ite = 0; active; ite++)
        {
            const bool mine = (active >> lane) & 1;
            const bool have_state = ite > 0;
            if (chain_warp)
            {
                // ---- the chain: x = var[K+c-1] after check c-1, per lane (= frame).  The words posted for check c+1 are
                //      fetched before check c is computed: they do not depend on the chain, and their shared-memory latency
                //      (and the barrier test) then lie beside the dependent instructions instead of between them.
                C x = A::max_value();
                C n_m1, n_m2, n_tout, n_msg;
                unsigned n_sa;
                auto fetch = [&](int c) {
                    const int u = c % NWK;
                    const uint32_t* sl = s_slot + u * SLOT + lane;
                    const unsigned par = __shfl_sync(FULL, visits, u) & 1u;
                    if (lane == u) visits++;
                    mbar_wait_acquire(s_posted + u, par);
                    n_m1 = A::from_bits(sl[0 * 32]), n_m2 = A::from_bits(sl[1 * 32]);
                    n_sa = sl[2 * 32];
                    n_tout = A::from_bits(sl[3 * 32]);
                    n_msg = A::from_bits(sl[4 * 32]);
                };
                fetch(0);
#pragma unroll 1
                for (int c = 0; c < M; c++)
                {
                    uint32_t* sl = s_slot + (c % NWK) * SLOT + lane;
                    C m1 = n_m1, m2 = n_m2;
                    unsigned sa = n_sa;
                    const C t_out = n_tout, msg_in = n_msg;
                    if (c + 1 < M) fetch(c + 1);
                    const C ti = c == 0 ? A::max_value() : A::sub(x, msg_in); // check 0 has no incoming chain edge
                    const C ab = A::absv(ti);
                    sa ^= A::raw(ti);
                    m2 = A::vmin(m2, A::vmax(ab, m1));
                    m1 = A::vmin(m1, ab);
                    const C c1 = A::vmax((C)0, A::normalize(m2, a.norm, a.k8)); // Update_rule_NMS.hxx:100-106
                    const C c2 = A::vmax((C)0, A::normalize(m1, a.norm, a.k8));
                    const bool e = A::absv(t_out) == m1;
                    x = A::add(t_out, A::apply_sign(e ? c1 : c2, sa ^ A::raw(t_out))); // var[K+c]
                    sl[5 * 32] = A::raw(ti);
                    sl[6 * 32] = A::raw(m1);
                    sl[7 * 32] = A::raw(c1);
                    sl[8 * 32] = A::raw(c2);
                    sl[9 * 32] = sa;
                    __syncwarp();
                    if (lane == 0) mbar_arrive_release(s_answered + (c % NWK));
                }
            }
            else
            {
                C v[DI + 1]; // information variables of the check, then var[K+c]
                uint4 st;
                auto issue = [&](int c) {
                    const uint32_t* row = s_idx + c * IW;
#pragma unroll
                    for (int i = 0; i < DI; i++)
                    {
                        const unsigned vi = (row[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                        v[i] = (C)__ldcg(reinterpret_cast<const R*>(vbytes + (vi * VB + loff)));
                    }
                    v[DI] = (C)__ldcg(reinterpret_cast<const R*>(vbytes + ((unsigned)(K + c) * VB + loff)));
                    st = have_state ? SIO::load(cn, (size_t)c) : make_uint4(0u, 0u, 0u, 0u);
                };
                uint32_t* sl = s_slot + w * SLOT + lane;
                if (w < M) issue(w);
#pragma unroll 1
                for (int c = w; c < M; c += NWK)
                {
                    const uint32_t* row = s_idx + c * IW;
                    const uint32_t lastm = s_last[c];
                    // ---- every edge but the incoming chain edge (pass 1 of Update_rule_MS.hxx:57-67 on D - 1 edges)
                    C o1, o2;
                    unsigned osg, oeq;
                    ST::decode(st, o1, o2, osg, oeq);
                    C t[D];
                    unsigned sa = 0;
                    C m1 = A::max_value(), m2 = A::max_value();
#pragma unroll
                    for (int i = 0; i < D; i++)
                    {
                        if (i == DI) continue;
                        const int b = D - 1 - i;
                        const C mag = (oeq & (1u << b)) ? o1 : o2;
                        const C ti = A::sub(v[i < DI ? i : DI], A::apply_sign(mag, osg << (31 - b)));
                        t[i] = ti;
                        const C ab = A::absv(ti);
                        sa ^= A::raw(ti);
                        m2 = A::vmin(m2, A::vmax(ab, m1));
                        m1 = A::vmin(m1, ab);
                    }
                    // the chain edge is edge DI: bit D - 1 - DI = 1 of the state words
                    sl[0 * 32] = A::raw(m1);
                    sl[1 * 32] = A::raw(m2);
                    sl[2 * 32] = sa;
                    sl[3 * 32] = A::raw(t[DI + 1]);
                    sl[4 * 32] = A::raw(A::apply_sign((oeq & 2u) ? o1 : o2, osg << 30));
                    __syncwarp();
                    if (lane == 0) mbar_arrive_release(s_posted + w);
                    if (c + NWK < M) issue(c + NWK);

                    mbar_wait_acquire(s_answered + w, n_turns & 1u);
                    n_turns++;
                    t[DI] = A::from_bits(sl[5 * 32]);
                    m1 = A::from_bits(sl[6 * 32]);
                    const C c1 = A::from_bits(sl[7 * 32]), c2 = A::from_bits(sl[8 * 32]);
                    sa = sl[9 * 32];

                    // ---- new messages (Update_rule_MS.hxx:82-92, ...layered.hxx:198-203)
                    unsigned nsg = 0, neq = 0;
#pragma unroll
                    for (int i = 0; i < D; i++)
                    {
                        const C tt = t[i];
                        const bool e = A::absv(tt) == m1;
                        const unsigned xs = sa ^ A::raw(tt); // bit 31: sign of the new message
                        const C nv = A::add(tt, A::apply_sign(e ? c1 : c2, xs));
                        if (e) neq |= 1u << (D - 1 - i);
                        nsg |= (xs >> 31) << (D - 1 - i);
                        unsigned vi;
                        bool ballot;
                        if (i < DI)
                        {
                            vi = (row[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                            ballot = (lastm >> i) & 1u;
                        }
                        else if (i == DI)
                        {
                            if (c == 0) continue;
                            vi = K + c - 1;
                            ballot = true; // check c is the last one to touch parity variable K+c-1
                        }
                        else
                        {
                            if (c != M - 1) continue; // stored by check c+1 (as its incoming edge); only the last check stores it
                            vi = K + c;
                            ballot = true;
                        }
                        if (mine) __stcg(reinterpret_cast<R*>(vbytes + (vi * VB + loff)), (R)nv);
                        if (ballot)
                        {
                            const unsigned hw = __ballot_sync(FULL, A::lt0(nv));
                            if (lane == 0) __stcg(hd + vi, hw);
                        }
                    }
                    if (mine) SIO::store(cn, (size_t)c, ST::encode(c1, c2, nsg, neq));
                }
            }
            my_iters += mine ? 1 : 0;
            if (threadIdx.x == 0) *s_unsat = 0;
            __syncthreads(); // every hard decision of the iteration is written

            // ---- syndrome on the bit-sliced hard decisions: thread t tests checks t, t + 512, ...
            unsigned unsat = 0;
            for (int c = threadIdx.x; c < M; c += blockDim.x)
            {
                const uint32_t* row = s_idx + c * IW;
                unsigned p = __ldcg(hd + K + c);
                if (c > 0) p ^= __ldcg(hd + K + c - 1);
#pragma unroll
                for (int i = 0; i < DI; i++)
                    p ^= __ldcg(hd + ((row[i >> 1] >> ((i & 1) * 16)) & 0xffffu));
                unsat |= p;
            }
            unsat = __reduce_or_sync(FULL, unsat);
            if (lane == 0 && unsat) atomicOr(s_unsat, unsat);
            __syncthreads();
            unsat = *s_unsat;
            const unsigned spent = __ballot_sync(FULL, mine && my_iters >= a.n_ite);
            unsigned fin = a.early_stop ? (active & ~unsat) : 0u;
            fin |= spent;
            if (fin)
            { // snapshot the finishing lanes' hard decisions (all threads share the variables)
                for (int vv = threadIdx.x; vv < N_LDPC; vv += blockDim.x)
                {
                    const unsigned hwv = __ldcg(hd + vv);
                    const unsigned o = done ? __ldcg(out_hd + vv) : 0u;
                    __stcg(out_hd + vv, (o & done) | (hwv & fin));
                }
                failed |= fin & unsat;
                done |= fin;
                active &= ~fin;
            }
            __syncthreads(); // hd and s_unsat are reused by the next iteration
        }

        if (w == 0)
        {
            const bool fin = (done >> lane) & 1;
            if (fin)
            {
                const bool ok = !a.early_stop || !((failed >> lane) & 1);
                if (a.cwd) a.cwd[slot] = (int8_t)ok;
                if (a.n_iter) a.n_iter[slot] = my_iters;
                if (a.stats) atomicAdd(a.stats + STAT_ITER_HIST + min(my_iters, LDPC_HIST_BINS - 1), 1ull);
            }
            const int it_sum = __reduce_add_sync(FULL, fin ? my_iters : 0);
            if (lane == 0 && a.stats)
            {
                atomicAdd(a.stats + STAT_FRAMES, (unsigned long long)__popc(done));
                atomicAdd(a.stats + STAT_LDPC_ITERS, (unsigned long long)it_sum);
                atomicAdd(a.stats + STAT_LDPC_FAIL, (unsigned long long)__popc(failed & done));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- frame compaction
__device__ __forceinline__ int
occupied_groups(const CompactLaunch& c)
{
    if (!c.n_src_slots) return c.n_groups;
    const int per = 32 * c.nf;
    return min(c.n_groups, (*c.n_src_slots + per - 1) / per);
}

// survivors per group -> exclusive prefix (one block; groups are few thousand)
__global__ void __launch_bounds__(1024) compact_scan_kernel(const CompactLaunch c)
{
    __shared__ int part[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int n_groups = occupied_groups(c);
    for (int g0 = 0; g0 < n_groups; g0 += 1024)
    {
        const int g = g0 + threadIdx.x;
        int n = 0;
        if (g < n_groups)
            for (int j = 0; j < c.nf; j++)
                n += __popc(c.surv_mask[g * c.nf + j]);
        part[threadIdx.x] = n;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) // Hillis-Steele inclusive scan
        {
            const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
            __syncthreads();
            part[threadIdx.x] += v;
            __syncthreads();
        }
        if (g < n_groups) c.dst_base[g] = carry + part[threadIdx.x] - n;
        __syncthreads();
        if (threadIdx.x == 1023) carry += part[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *c.n_surv = min(carry, c.cap);
}

// one thread per source subgroup: destination slots of its survivors, in slot order
__global__ void
compact_expand_kernel(const CompactLaunch c)
{
    const int sg = blockIdx.x * blockDim.x + threadIdx.x;
    if (sg >= occupied_groups(c) * c.nf) return;
    const int g = sg / c.nf, j = sg % c.nf;
    int d = c.dst_base[g];
    for (int k = 0; k < j; k++)
        d += __popc(c.surv_mask[g * c.nf + k]);
    unsigned m = c.surv_mask[sg];
    while (m)
    {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        if (d < c.cap)
        {
            const int s = sg * 32 + l;
            const int frame = c.src_frame ? c.src_frame[s] : s;
            c.src_of_dst[d] = s;
            c.dst_frame[d] = frame;
            c.dst_iters[d] = c.surv_iters[s];
            c.final_loc[frame] = c.dst_slot0 + d;
        }
        d++;
    }
}

// warp = (destination group, part): gathers the state of its 32*ND survivors.  Survivors of one source group are
// neighbours in destination order, so the scattered element loads of a warp fall into a handful of source chunks.
constexpr int COMPACT_PARTS = 8;
template<typename R, int NF, int ND, int SW = 4> // NF: frames per lane of the source groups, ND: of the destination groups; SW: state words
__global__ void __launch_bounds__(128) compact_copy_kernel(const CompactLaunch c)
{
    using SIO = StateIO<SW>;
    const int lane = threadIdx.x & 31;
    const int d = blockIdx.x;
    const int part = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n_surv = *c.n_surv;
    if (d * 32 * ND >= n_surv || part >= COMPACT_PARTS) return;
    size_t vbase[ND], cbase[ND];
    bool valid[ND];
#pragma unroll
    for (int j = 0; j < ND; j++)
    {
        const int ds = (d * ND + j) * 32 + lane;
        valid[j] = ds < n_surv;
        const int s = valid[j] ? c.src_of_dst[ds] : 0;
        const int gs = s / (32 * NF), js = (s / 32) % NF, ls = s % 32;
        vbase[j] = ((size_t)gs * N_LDPC * 32 + ls) * NF + js;
        cbase[j] = (size_t)gs; // source group; element and lane follow from js, ls below
        (void)ls;
    }
    const R* vs = (const R*)c.var_src;
    R* vd = (R*)c.var_dst + ((size_t)d * N_LDPC * 32 + lane) * ND;
    const int v0 = (int)((long)N_LDPC * part / COMPACT_PARTS), v1 = (int)((long)N_LDPC * (part + 1) / COMPACT_PARTS);
#pragma unroll 4
    for (int v = v0; v < v1; v++)
    {
        R x[ND];
#pragma unroll
        for (int j = 0; j < ND; j++)
            x[j] = valid[j] ? __ldcg(vs + vbase[j] + (size_t)v * 32 * NF) : (R)0;
#pragma unroll
        for (int j = 0; j < ND; j++)
            vd[(size_t)v * 32 * ND + j] = x[j];
    }
    constexpr int PLS = sizeof(R) == 2 ? NF / 2 : NF, PLD = sizeof(R) == 2 ? ND / 2 : ND; // state elements per lane: pairs for int16
    const uint32_t* cs = reinterpret_cast<const uint32_t*>(c.cn_src);
    uint32_t* cd = reinterpret_cast<uint32_t*>(c.cn_dst) + (size_t)d * SIO::group_words(c.M, PLD) + SIO::lane_offset(lane);
    int src_js[ND], src_ls[ND];
#pragma unroll
    for (int j = 0; j < ND; j++)
    {
        const int ds = (d * ND + j) * 32 + lane;
        const int sl = valid[j] ? c.src_of_dst[ds] : 0;
        src_js[j] = (sl / 32) % NF, src_ls[j] = sl % 32;
    }
    const int c0 = (int)((long)c.M * part / COMPACT_PARTS), c1 = (int)((long)c.M * (part + 1) / COMPACT_PARTS);
    if constexpr (sizeof(R) == 2)
    { // packed pairs (ldpc_decode_kernel<int16_t, ., 4, true>): state [check][pair][lane], one half of every word per frame
        static_assert((NF == 4 || NF == 2) && (ND == 4 || ND == 2), "int16_t groups hold two or four frames per lane");
        int sh[ND];
#pragma unroll
        for (int j = 0; j < ND; j++)
            sh[j] = (src_js[j] & 1) * 16;
#pragma unroll 2
        for (int k = c0; k < c1; k++)
#pragma unroll
            for (int p = 0; p < ND / 2; p++)
            {
                uint32_t w[4] = { 0u, 0u, 0u, 0u };
#pragma unroll
                for (int hf = 0; hf < 2; hf++)
                {
                    const int j = 2 * p + hf;
                    if (!valid[j]) continue;
                    const uint4 x = SIO::load(cs + cbase[j] * SIO::group_words(c.M, PLS) + SIO::lane_offset(src_ls[j]), (size_t)k * PLS + (src_js[j] >> 1));
                    w[0] |= ((x.x >> sh[j]) & 0xffffu) << (16 * hf);
                    w[1] |= ((x.y >> sh[j]) & 0xffffu) << (16 * hf);
                    w[2] |= ((x.z >> sh[j]) & 0xffffu) << (16 * hf);
                    w[3] |= ((x.w >> sh[j]) & 0xffffu) << (16 * hf);
                }
                SIO::store(cd, (size_t)k * PLD + p, make_uint4(w[0], w[1], w[2], w[3]));
            }
    }
    else
    {
#pragma unroll 2
        for (int k = c0; k < c1; k++)
#pragma unroll
            for (int j = 0; j < ND; j++)
                SIO::store(cd, (size_t)k * PLD + j,
                           valid[j] ? SIO::load(cs + cbase[j] * SIO::group_words(c.M, PLS) + SIO::lane_offset(src_ls[j]), (size_t)k * PLS + src_js[j])
                                    : make_uint4(0u, 0u, 0u, 0u));
    }
}

// after the last round: the final hard decisions of every frame handed over by round 1 go back to the bit lane of its
// round-1 subgroup.  warp = (round-1 group, part), lanes = consecutive variables.
__global__ void __launch_bounds__(128) merge_back_kernel(const MergeLaunch c)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x;
    const int part = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (part >= COMPACT_PARTS) return;
    const int v0 = (int)((long)N_LDPC * part / COMPACT_PARTS), v1 = (int)((long)N_LDPC * (part + 1) / COMPACT_PARTS);
    for (int j = 0; j < c.nf; j++)
    {
        const unsigned m = c.surv_mask[g * c.nf + j];
        if (!m) continue; // warp-uniform
        const int my_loc = ((m >> lane) & 1) ? c.final_loc[(g * c.nf + j) * 32 + lane] : 0;
        uint32_t* dst = c.out_hd + (size_t)(g * c.nf + j) * N_LDPC;
        for (int v0w = v0; v0w < v1; v0w += 32)
        {
            const int v = v0w + lane;
            const bool in = v < v1;
            uint32_t w = in ? dst[v] : 0u;
            unsigned mm = m;
            while (mm)
            {
                const int l = __ffs(mm) - 1;
                mm &= mm - 1;
                const int loc = __shfl_sync(FULL, my_loc, l);
                if (in)
                {
                    const uint32_t bit = (c.out_hd[(size_t)(loc >> 5) * N_LDPC + v] >> (loc & 31)) & 1u;
                    w = (w & ~(1u << l)) | (bit << l);
                }
            }
            if (in) dst[v] = w;
        }
    }
}

// ---------------------------------------------------------------------------------------------- layout kernels
// frame-major LLRs [F][16200] -> var[group][v][lane][NF] (frame = (group*NF + j)*32 + lane); frames past n_frames = 0.
// One block per (32 variables, 32-frame subgroup).
template<typename R, typename RI = R> // RI: element type on the wire (int8 LLRs feeding the int16 decoder)
__global__ void
pack_llr_kernel(const RI* __restrict__ in, R* __restrict__ var, int n_frames, int nf)
{
    __shared__ R tile[32][33];
    const int sg = blockIdx.y, v0 = blockIdx.x * 32; // sg = 32-frame subgroup index = group*nf + j
    const int g = sg / nf, j = sg % nf;
    const int x = threadIdx.x, y = threadIdx.y;
    for (int k = y; k < 32; k += blockDim.y)
    {
        const int f = sg * 32 + k, v = v0 + x;
        tile[k][x] = (f < n_frames && v < N_LDPC) ? (R)in[(size_t)f * N_LDPC + v] : (R)0;
    }
    __syncthreads();
    for (int k = y; k < 32; k += blockDim.y)
    {
        const int v = v0 + k;
        if (v < N_LDPC) var[(((size_t)g * N_LDPC + v) * 32 + x) * nf + j] = tile[x][k];
    }
}

// bit-sliced words [subgroup][stride] -> int32 bits [F][K] (first K variables)
__global__ void
unpack_bits_kernel(const uint32_t* __restrict__ words, const uint8_t* __restrict__ scr, int32_t* __restrict__ out, int stride,
                   int K, int n_frames)
{
    const int g = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= K) return;
    uint32_t w = words[(size_t)g * stride + v];
    if (scr && ((scr[v >> 3] >> (v & 7)) & 1u)) w = ~w; // Scrambler_BB::descramble: XOR with the PRBS bit of position v
#pragma unroll 8
    for (int f = 0; f < 32; f++)
    {
        const int frame = g * 32 + f;
        if (frame < n_frames) out[(size_t)frame * K + v] = (int32_t)((w >> f) & 1u);
    }
}

// bit-sliced words [subgroup][stride] -> packed bytes [F][K/8], bit i of a frame at byte i/8, bit i%8 (the layout of
// spu::tools::Bit_packer::pack with msb_to_lsb = false, lib/streampu/include/Tools/Algo/Bit_packer/Bit_packer.hxx, which
// Sink_user_binary writes to disk); scr (or null) = the BB scrambler sequence packed the same way, XORed in
// (Scrambler_BB::descramble, src/common/Module/Scrambler/Scrambler_BB/Scrambler_BB.hxx:51-72).  thread = (subgroup, byte).
__global__ void
unpack_bytes_kernel(const uint32_t* __restrict__ words, const uint8_t* __restrict__ scr, uint8_t* __restrict__ out, int stride,
                    int KB, int n_frames)
{
    const int g = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= KB) return;
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
        w[k] = words[(size_t)g * stride + 8 * b + k];
    const uint32_t x = scr ? scr[b] : 0u;
    const int nf = min(32, n_frames - g * 32);
    for (int f = 0; f < nf; f++)
    {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            v |= ((w[k] >> f) & 1u) << k;
        out[(size_t)(g * 32 + f) * KB + b] = (uint8_t)(v ^ x);
    }
}

// int32 bits [F][N] -> bit-sliced words [subgroup][stride]
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
int
ldpc_frames_per_lane(int precision)
{
    return precision == 0 ? 2 : 4; // float: 2 x 4 B, int16: 4 x 2 B (packed pairs), int8: 4 x 1 B per lane and variable
}

template<typename R, int DI, int NF, bool PK = false>
static cudaError_t
launch_ldpc_k(const LdpcLaunch& a, cudaStream_t s)
{
    auto k = ldpc_decode_kernel<R, DI, NF, PK>;
    const size_t smem = ((size_t)a.M * ((DI + 1) / 2) + a.M) * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); // per device
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(a.queue, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    const int groups = (a.n_frames + 32 * NF - 1) / (32 * NF); // capacity when the real count is on the device
    // spread the groups over all SMs (a warp's walk through the checks is latency-bound, so 4 warps on each of 148 SMs
    // finish sooner than 12 warps on each of 50); the queue hands groups to whichever warp is free
    const int blocks = groups < a.n_sms ? groups : a.n_sms;
    // ... and a CTA carries only the warps it can feed, so that launches of other streams (pipeline chunks, other module
    // clones) fit beside it on the SM instead of waiting for a mostly idle 12-warp CTA to drain
    const int warps = std::min(LDPC_WARPS, std::max(1, (groups + blocks - 1) / std::max(1, blocks)));
    k<<<blocks, warps * 32, smem, s>>>(a); count_launch();
    return cudaGetLastError();
}

template<typename R, int NF, bool PK = false>
static cudaError_t
launch_ldpc_t(const LdpcLaunch& a, cudaStream_t s)
{
    if (a.d_info == 25) return launch_ldpc_k<R, 25, NF, PK>(a, s);
    if (a.d_info == 9) return launch_ldpc_k<R, 9, NF, PK>(a, s);
    return cudaErrorInvalidValue;
}

template<typename R, int DI>
static cudaError_t
launch_ldpc_coop_k(const LdpcLaunch& a, cudaStream_t s)
{
    auto k = ldpc_decode_coop_kernel<R, DI>;
    const size_t smem = ((size_t)a.M * ((DI + 1) / 2) + a.M + 1 + 4 * COOP_WORKERS + COOP_WORKERS * 320 + 2) * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(a.queue, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    const int groups = (a.n_frames + 31) / 32;
    k<<<groups < a.n_sms ? groups : a.n_sms, COOP_WARPS * 32, smem, s>>>(a); count_launch();
    return cudaGetLastError();
}

int
ldpc_coop_min_reuse()
{
    return COOP_MIN_REUSE;
}

// float or int16, one frame per lane, one round: the cooperative small-batch kernel
cudaError_t
launch_ldpc_decode_coop(const LdpcLaunch& a, int precision, cudaStream_t s)
{
    if (precision == 0)
    {
        if (a.d_info == 25) return launch_ldpc_coop_k<float, 25>(a, s);
        if (a.d_info == 9) return launch_ldpc_coop_k<float, 9>(a, s);
    }
    else if (precision == 1)
    {
        if (a.d_info == 25) return launch_ldpc_coop_k<int16_t, 25>(a, s);
        if (a.d_info == 9) return launch_ldpc_coop_k<int16_t, 9>(a, s);
    }
    return cudaErrorInvalidValue;
}

// bytes of the cooperative kernel's check state per (check, frame)
int
ldpc_coop_state_bytes(int precision, int d_info)
{
    return precision == 0 ? 4 * ldpc_state_words(d_info) : 12;
}

cudaError_t
launch_ldpc_decode(const LdpcLaunch& a, int precision, cudaStream_t s)
{
    switch (precision)
    {
        case 0: return a.nf == 1 ? launch_ldpc_t<float, 1>(a, s) : launch_ldpc_t<float, 2>(a, s);
        case 1: return a.nf == 2 ? launch_ldpc_t<int16_t, 2, true>(a, s) : launch_ldpc_t<int16_t, 4, true>(a, s);
        case 2: return launch_ldpc_t<int8_t, 4>(a, s);
    }
    return cudaErrorInvalidValue;
}

template<typename R, int NF, int ND>
static cudaError_t
launch_compact_t(const CompactLaunch& c, cudaStream_t s)
{
    compact_scan_kernel<<<1, 1024, 0, s>>>(c); count_launch();
    compact_expand_kernel<<<(c.n_groups * c.nf + 127) / 128, 128, 0, s>>>(c); count_launch();
    const dim3 grid(c.cap / (32 * ND), (COMPACT_PARTS + 3) / 4);
    if (grid.x > 0)
    {
        if (c.state_words == 3)
            compact_copy_kernel<R, NF, ND, 3><<<grid, 128, 0, s>>>(c);
        else
            compact_copy_kernel<R, NF, ND, 4><<<grid, 128, 0, s>>>(c);
        count_launch();
    }
    return cudaGetLastError();
}

int
ldpc_state_words(int d_info)
{
    return d_info == 9 ? STATE_WORDS<9> : STATE_WORDS<25>;
}

int
ldpc_later_frames_per_lane(int precision, int preferred)
{
    // float and int16: narrower groups in the later rounds (their few warps are latency-bound, so halving the work per
    // check halves the time a warp needs per check) unless asked otherwise; int8 keeps its 4-byte lanes
    if (precision == 0) return preferred == 2 ? 2 : 1;
    if (precision == 1) return preferred == 4 ? 4 : 2;
    return ldpc_frames_per_lane(precision);
}

cudaError_t
launch_compact(const CompactLaunch& c, int precision, cudaStream_t s)
{
    if (c.cap % (32 * c.nf_dst)) return cudaErrorInvalidValue;
    switch (precision)
    {
        case 0:
            if (c.nf == 2 && c.nf_dst == 1) return launch_compact_t<float, 2, 1>(c, s);
            if (c.nf == 2 && c.nf_dst == 2) return launch_compact_t<float, 2, 2>(c, s);
            if (c.nf == 1 && c.nf_dst == 1) return launch_compact_t<float, 1, 1>(c, s);
            break;
        case 1:
            if (c.nf == 4 && c.nf_dst == 4) return launch_compact_t<int16_t, 4, 4>(c, s);
            if (c.nf == 4 && c.nf_dst == 2) return launch_compact_t<int16_t, 4, 2>(c, s);
            if (c.nf == 2 && c.nf_dst == 2) return launch_compact_t<int16_t, 2, 2>(c, s);
            break;
        case 2:
            if (c.nf == 4 && c.nf_dst == 4) return launch_compact_t<int8_t, 4, 4>(c, s);
            break;
    }
    return cudaErrorInvalidValue;
}

cudaError_t
launch_merge_back(const MergeLaunch& c, cudaStream_t s)
{
    const dim3 grid(c.n_groups, (COMPACT_PARTS + 3) / 4);
    merge_back_kernel<<<grid, 128, 0, s>>>(c); count_launch();
    return cudaGetLastError();
}

cudaError_t
launch_pack_llr(const void* in, void* var, int precision, int n_frames, int nf, cudaStream_t s)
{
    const int groups = (n_frames + 32 * nf - 1) / (32 * nf);
    const dim3 grid((N_LDPC + 31) / 32, groups * nf), block(32, 8);
    switch (precision)
    {
        case 0: pack_llr_kernel<float><<<grid, block, 0, s>>>((const float*)in, (float*)var, n_frames, nf); count_launch(); break;
        case 1: pack_llr_kernel<int16_t><<<grid, block, 0, s>>>((const int16_t*)in, (int16_t*)var, n_frames, nf); count_launch(); break;
        case 2: pack_llr_kernel<int8_t><<<grid, block, 0, s>>>((const int8_t*)in, (int8_t*)var, n_frames, nf); count_launch(); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t
launch_pack_llr8_as_i16(const int8_t* in, void* var, int n_frames, int nf, cudaStream_t s)
{
    const int groups = (n_frames + 32 * nf - 1) / (32 * nf);
    const dim3 grid((N_LDPC + 31) / 32, groups * nf), block(32, 8);
    pack_llr_kernel<int16_t, int8_t><<<grid, block, 0, s>>>(in, (int16_t*)var, n_frames, nf); count_launch();
    return cudaGetLastError();
}

cudaError_t
launch_unpack_bits(const uint32_t* words, int32_t* out, int stride, int K, int n_frames, cudaStream_t s, const uint8_t* scr)
{
    const int groups = (n_frames + 31) / 32;
    const dim3 grid((K + 127) / 128, groups);
    unpack_bits_kernel<<<grid, 128, 0, s>>>(words, scr, out, stride, K, n_frames); count_launch();
    return cudaGetLastError();
}

cudaError_t
launch_unpack_bytes(const uint32_t* words, const uint8_t* scr, uint8_t* out, int stride, int K, int n_frames, cudaStream_t s)
{
    if (K % 8) return cudaErrorInvalidValue;
    const int groups = (n_frames + 31) / 32;
    const dim3 grid((K / 8 + 127) / 128, groups);
    unpack_bytes_kernel<<<grid, 128, 0, s>>>(words, scr, out, stride, K / 8, n_frames); count_launch();
    return cudaGetLastError();
}

cudaError_t
launch_pack_bits(const int32_t* in, uint32_t* words, int stride, int N, int n_frames, cudaStream_t s)
{
    const int groups = (n_frames + 31) / 32;
    const dim3 grid((N + 31) / 32, groups), block(32, 8);
    pack_bits_kernel<<<grid, block, 0, s>>>(in, words, stride, N, n_frames); count_launch();
    return cudaGetLastError();
}
} // namespace dvbs2
