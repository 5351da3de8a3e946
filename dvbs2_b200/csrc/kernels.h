// Internal launch interfaces between capi.cu and the kernel translation units.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "code_tables.h"

namespace dvbs2
{
constexpr int LDPC_WARPS = 12; // warps of the persistent LDPC CTA (one CTA per SM)

// Every kernel launch of the library is counted at its <<<>>> site (dvbs2_cuda_stats::kernel_launches).  The counter is
// per host thread — a handle is driven by one thread at a time — and capi.cu credits the difference to the handle in use.
uint64_t& launch_counter();
inline void
count_launch()
{
    launch_counter()++;
}

enum StatSlot
{
    STAT_FRAMES = 0,
    STAT_LDPC_ITERS,
    STAT_LDPC_FAIL,
    STAT_BCH_NONZERO,
    STAT_BCH_FAIL,
    STAT_ITER_HIST,                       // [LDPC_HIST_BINS] frames that stopped after 0, 1, 2, ... iterations (last bin: >= 32)
    STAT_COUNT = STAT_ITER_HIST + 33
};
constexpr int LDPC_HIST_BINS = 33;

// Device-resident layout shared by the kernels.  "subgroup" = 32 consecutive frames (bit f of a bit-sliced word =
// frame 32*subgroup + f); "group" = NF consecutive subgroups decoded by one warp, NF = ldpc_frames_per_lane(precision):
//   var    [groups][16200][32][NF]   R      variable-node values (LLR + messages); element (lane l, j) = frame
//                                            (group*NF + j)*32 + l
//   cn     [groups][M][NF][32]       uint4  compressed check state {out1, out2, sign bits, eq bits}
//   hd     [groups][16200][NF]       u32    running hard decisions (the NF subgroup words of a variable are adjacent)
//   out_hd [subgroups][16200]        u32    hard decisions snapshotted when each frame stopped iterating
struct LdpcLaunch
{
    void* var;
    uint4* cn;
    uint32_t* hd;
    uint32_t* out_hd;
    const uint32_t* idx32; // [M][(d_info+1)/2] information-variable indices, two per word
    const uint32_t* last;  // [M] bit i: check c is the last one of the iteration to update its i-th information variable
    int* queue;            // work queue head (zeroed by the launcher)
    int8_t* cwd;           // [frames] or null
    int32_t* n_iter;       // [frames] or null
    unsigned long long* stats;
    int n_frames, n_groups;
    int K, M, d_info;
    int n_ite;
    float norm;
    int k8; // norm * 8 if that is an integer in 1..8, else 0
    int early_stop;
    int n_sms;
    // multi-round decoding (frame compaction): all null / 0 for a plain single-round launch
    const int* n_slots_dev;  // rounds >= 2: number of occupied slots, known only on the device (n_frames = capacity)
    const int* slot_frame;   // rounds >= 2: [slots] frame index (for cwd / n_iter) of each slot
    const int* slot_iters;   // rounds >= 2: [slots] iterations already spent; such frames carry check state
    uint32_t* surv_mask;     // [subgroups] lanes handed over to the next round (0 if none); null in the last round
    int* surv_iters;         // [slots] iterations spent by the lanes handed over
    int exit_below;          // hand the group's live frames over once fewer than this many remain (0 = never)
    int nf;                  // frames per lane of this launch (0 = ldpc_frames_per_lane(precision))
};

// Frame compaction between two rounds: survivors of round r (src slots) are packed, in slot order, into dense groups of
// fresh slots (dst) for round r+1, their var / check state copied.  After the last round one merge pass returns every
// handed-over frame's hard decisions to the bit lane of its round-1 subgroup (final_loc says where they ended up).
struct CompactLaunch
{
    const uint32_t* surv_mask; // [n_groups * nf]   of the source round
    const int* surv_iters;     // [n_groups * nf * 32]
    const int* src_frame;      // [n_groups * nf * 32] frame index of each source slot, or null (round 1: slot = frame)
    const int* n_src_slots;    // device: occupied source slots (null: all n_groups groups are occupied)
    int* dst_base;             // [n_groups]   exclusive prefix of survivors per group
    int* n_surv;               // [1]
    int* src_of_dst;           // [cap]        source slot of each destination slot
    int* dst_frame;            // [cap]        frame index of each destination slot
    int* dst_iters;            // [cap]
    int* final_loc;            // [frames]     out_hd slot (dst_slot0 + d) holding the frame's final hard decisions
    int dst_slot0;             // index of destination slot 0 in the out_hd slot space of the merge pass
    const void* var_src;
    void* var_dst;
    const uint4* cn_src;
    uint4* cn_dst;
    int state_words; // 32-bit words of the compressed check state per (check, frame or int16 pair): ldpc_state_words(d_info)
    int n_groups, nf, cap, M;
    int nf_dst;                // frames per lane of the destination groups (<= nf)
};
// merge pass: round-1 subgroups pull the final bits of the frames they handed over
struct MergeLaunch
{
    const uint32_t* surv_mask; // [n_groups * nf]  round 1
    const int* final_loc;      // [frames]
    uint32_t* out_hd;          // [slot space / 32][16200]; round-1 subgroups first
    int n_groups, nf;
};
cudaError_t launch_compact(const CompactLaunch& c, int precision, cudaStream_t s);    // scan + expand + copy (3 kernels)
cudaError_t launch_merge_back(const MergeLaunch& c, cudaStream_t s);
// frames per lane allowed for round r >= 2 given the handle's preference (0 = default)
int ldpc_later_frames_per_lane(int precision, int preferred);

int ldpc_frames_per_lane(int precision); // NF: 2 (f32, i16), 4 (i8)
cudaError_t launch_ldpc_decode(const LdpcLaunch& a, int precision, cudaStream_t s);
// float, 32-frame groups, a whole CTA per group (small batches): same results, one round
cudaError_t launch_ldpc_decode_coop(const LdpcLaunch& a, int precision, cudaStream_t s); // float or int16, one frame per lane
int ldpc_coop_state_bytes(int precision, int d_info); // check state of that kernel per (check, frame)
// smallest distance (in checks) between two checks sharing an information variable that the cooperative kernel tolerates
int ldpc_coop_min_reuse();
int ldpc_state_words(int d_info); // 32-bit words of compressed check state per (check, frame / int16 pair): 3 for degree <= 11, else 4
cudaError_t launch_pack_llr(const void* in, void* var, int precision, int n_frames, int nf, cudaStream_t s);
// int8 LLRs on the wire (compact input socket of the fixed-point decoder) into the int16 decoder's var layout
cudaError_t launch_pack_llr8_as_i16(const int8_t* in, void* var, int n_frames, int nf, cudaStream_t s);
// scr (optional): BB scrambler sequence, 8 bits per byte LSB first, XORed into the output (Scrambler_BB::descramble)
cudaError_t launch_unpack_bits(const uint32_t* words, int32_t* out, int stride, int K, int n_frames, cudaStream_t s,
                               const uint8_t* scr = nullptr);
cudaError_t launch_unpack_bytes(const uint32_t* words, const uint8_t* scr, uint8_t* out, int stride, int K, int n_frames,
                                cudaStream_t s);
cudaError_t launch_pack_bits(const int32_t* in, uint32_t* words, int stride, int N, int n_frames, cudaStream_t s);

// ---- "design B": variable nodes on chip (ldpc_onchip_kernel.cu).  Float only; frames come frame-major, bits leave packed
// per frame and are transposed into the bit-sliced out_hd layout for the BCH stage.
constexpr int ONCHIP_WINDOW = 26; // checks per window: two neighbouring windows must fit inside min_reuse_distance
struct OnchipLaunch
{
    const float* llr;        // [n_frames][16200] LLRs, natural order
    const uint32_t* idxT;    // [windows][(d_info+1)/2][32] byte offsets of the information variables, two per word
    uint4* cn;               // [n_sms * F][M] compressed check state of the frames in flight
    float* pv;               // [n_sms * F][M] parity variables of the frames in flight
    uint32_t* out_bits;      // [n_frames][ceil(K/32)] hard decisions, 32 per word
    int* queue;              // frame queue head (zeroed by the launcher)
    int8_t* cwd;             // [frames] or null
    int32_t* n_iter;         // [frames] or null
    unsigned long long* stats;
    int n_frames, K, M, d_info;
    int n_ite;
    float norm;
    int early_stop;
    int n_sms;
};
int ldpc_onchip_frames_per_cta(int d_info);
size_t ldpc_onchip_smem_bytes(int K, int M, int F);
std::vector<uint32_t> build_onchip_index(const LdpcTables& t);
cudaError_t launch_ldpc_onchip(const OnchipLaunch& a, cudaStream_t s);
cudaError_t launch_bits_to_sliced(const uint32_t* bits, uint32_t* out_hd, int K, int stride, int n_frames, cudaStream_t s);

struct DemapLaunch
{
    const float* sigma; // [frames]
    const float* sym;   // [frames][2*n_sym]
    const float* cstl;  // [2^bps][2] device
    int n_frames;
    int bps, n_sym;
    int itl_cols, itl_rev;
    int quant_dec, quant_bits; // for integer outputs
    int pl_sym;                // > 0: sym holds whole PL frames [frames][2*pl_sym]; the payload symbols are picked out in place
                               // (Framer::remove_plh fused into the read side), 0: sym holds XFEC frames
    int sym_fmt;               // wire format of sym (dvbs2_iq_format): 0 float, 1 IEEE half, 2 int16 * iq_scale, 3 int8 * iq_scale, 4 packed 12-bit * iq_scale
    float iq_scale;            // value of one LSB for the integer formats
};
// bytes per complex symbol (I and Q) of a dvbs2_iq_format
__host__ __device__ inline int
iq_symbol_bytes(int fmt)
{
    return fmt == 0 ? 8 : fmt == 3 ? 2 : fmt == 4 ? 3 : 4;
}
// PL frame geometry (src/common/Module/Framer/Framer.hxx:36-38,330-343; src/common/Factory/DVBS2/DVBS2.cpp:352-355): 90-symbol
// header, then the payload with a 36-symbol pilot block after every 1440 payload symbols (n_sym / 1440 blocks in all)
__host__ __device__ inline int
pl_frame_symbols(int n_sym)
{
    return 90 * (n_sym / 90 + 1) + 36 * (n_sym / 1440);
}
__host__ __device__ inline int
pl_symbol_pos(int k, int n_sym) // PL-frame position of payload symbol k
{
    const int blocks = k / 1440, n_pilots = n_sym / 1440;
    return 90 + k + 36 * (blocks < n_pilots ? blocks : n_pilots);
}
// Framer::remove_plh as a copy: [frames][2*pl_sym] -> [frames][2*n_sym]
cudaError_t launch_remove_plh(const float* pl, float* out, int n_sym, int n_frames, cudaStream_t s);
// Estimator_DVBS2::estimate: per frame sigma (and optionally Eb/N0, Es/N0 in dB) from the M2/M4 moments of its payload symbols
cudaError_t launch_estimate(const DemapLaunch& a, int K_bch, float* sigma, float* ebn0, float* esn0, cudaStream_t s);
// frame-major float LLRs; deinterleave = 0 gives Modem::demodulate order, 1 gives natural order
cudaError_t launch_demap_frames(const DemapLaunch& a, float* llr, int deinterleave, cudaStream_t s);
// straight into the decoder's lane-interleaved var layout (natural order, nf frames per lane), quantised for integers
cudaError_t launch_demap_to_var(const DemapLaunch& a, void* var, int precision, int nf, cudaStream_t s);
cudaError_t launch_deinterleave(const float* itl, float* nat, int n_frames, int n_sym, int itl_cols, int itl_rev,
                                cudaStream_t s);
cudaError_t launch_quantize(const float* in, void* out, int precision, size_t n, int dec, int bits, cudaStream_t s);

struct BchDevTables
{
    const uint16_t* alpha_to; // [16384]
    const int16_t* index_of;  // [16384]
    uint16_t minpoly[BCH_NPOLY];
    uint8_t poly_of[2 * BCH_T + 1];
    uint16_t synd_mat[2 * BCH_T + 1][GF_M];
};
struct BchLaunch
{
    uint32_t* words;  // [groups][stride] bit-sliced received word, natural order; corrected in place
    int stride;
    uint32_t* rem;    // [groups][BCH_NPOLY][14] scratch: remainders mod the minimal polynomials
    uint32_t* nz;     // [groups] scratch: bit f set if frame f has a non-zero syndrome
    int8_t* cwd;      // [frames] or null
    unsigned long long* stats;
    int n_frames, n_groups;
    int N, K;         // shortened BCH length and payload
    // segmented division: lane j of a warp divides bits [j * seg_len, (j + 1) * seg_len) and scales its remainder by
    // seg_mul[p][j] = x^(number of bits after the segment) mod m_p(x), low 14 coefficients (filled by bch_segment_table)
    int seg_len;
    uint16_t seg_mul[BCH_NPOLY][32];
};
void bch_segment_table(BchLaunch& a, const uint16_t* minpoly_low14); // host: seg_len / seg_mul from N
cudaError_t launch_bch_decode(const BchLaunch& a, const BchDevTables& t, cudaStream_t s);

// on-device transmitter / channel / monitor (tx_kernel.cu); all buffers bit-sliced per group of 32 frames
struct TxLaunch
{
    uint32_t* src;         // [groups][K_bch]  payload words
    uint32_t* cw;          // [groups][16200]  codeword words: scrambled payload | BCH parity | LDPC parity
    const uint32_t* scr;   // [K_bch] BB scrambler sequence, 0 or 0xffffffff per bit
    const uint32_t* gmask; // [192] BCH generator coefficients g_0..g_167 (then zeros), 0 or 0xffffffff
    const uint16_t* idx;   // [M][d_pad] LDPC information indices per check
    int K_bch, N_bch, d_info, d_pad;
    int n_frames, n_groups;
};
cudaError_t launch_tx_random_payload(uint32_t* src, int K, int n_groups, uint64_t seed, cudaStream_t s);
cudaError_t launch_tx_encode(const TxLaunch& a, cudaStream_t s); // scramble + BCH + LDPC (2 kernels)
cudaError_t launch_tx_modulate_awgn(const TxLaunch& a, const float* cstl, float* sym, float* cp, int bps, int n_sym,
                                    int itl_cols, int itl_rev, float sigma, uint64_t seed, cudaStream_t s);
// mon[0] += frames, mon[1] += bit errors, mon[2] += frame errors of descramble(dec) vs the payload
cudaError_t launch_monitor(const uint32_t* dec, int dec_stride, const TxLaunch& a, unsigned long long* mon, cudaStream_t s);
} // namespace dvbs2
