// Host-side code/modulation tables for libdvbs2_cuda: built once per handle in dvbs2_cuda_create().
//
// What the reference builds in tools::build_dvbs2/build_H (lib/aff3ct/src/Tools/Code/LDPC/Standard/DVBS2/
// DVBS2_constants.cpp:12-49), tools::Galois / BCH_polynomial_generator (lib/aff3ct/src/Tools/Math/Galois.cpp:158-183,
// lib/aff3ct/src/Tools/Code/BCH/BCH_polynomial_generator.cpp:52-139), tools::Constellation_user
// (lib/aff3ct/include/Tools/Constellation/User/Constellation_user.hxx:22-67) and Interleaver_core_column_row
// (lib/aff3ct/src/Tools/Interleaver/Column_row/Interleaver_core_column_row.cpp:59-87) — re-derived here in the
// shapes the kernels want (per-check index rows, minimal polynomials for bit-sliced syndromes, closed-form
// deinterleaver), not as sparse-matrix / LUT objects.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace dvbs2
{
constexpr int N_LDPC = 16200;
constexpr int QC = 360;       // quasi-cyclic parallelism of the DVB-S2 IRA codes
constexpr int GF_M = 14;      // GF(2^14)
constexpr int GF_N = 16383;   // 2^14 - 1, unshortened BCH length (Decoder_BCH.cpp:10-26)
constexpr int BCH_T = 12;     // correction power
constexpr int BCH_NPOLY = 12; // odd syndromes 1,3,...,23 -> 12 minimal polynomials of degree 14

struct ModcodParams
{
    int modcod;
    int bps;       // bits per symbol
    int n_sym;     // symbols per XFECFRAME = 16200 / bps
    int K_bch;     // payload bits
    int N_bch;     // = K_ldpc
    int M;         // LDPC checks = 16200 - K_ldpc
    int q;         // M / 360
    int d_info;    // information edges per check (uniform for these codes): 25 (8/9) or 9 (3/5)
    int itl_cols;  // 0 = no interleaver (QPSK), else number of columns (= bps)
    int itl_rev;   // 1 = TOP_RIGHT column order (8PSK 3/5), 0 = TOP_LEFT
    const char* name;
};

bool modcod_params(int modcod, ModcodParams& p);

// natural (deinterleaved) LDPC bit position of the LLR of bit b of symbol k
inline int
deinterleaved_pos(const ModcodParams& p, int k, int b)
{
    if (p.itl_cols == 0) return k * p.bps + b;
    const int col = p.itl_rev ? (p.itl_cols - 1 - b) : b;
    return col * p.n_sym + k;
}

struct LdpcTables
{
    int K, M, E, d_info, d_pad;     // d_pad: row stride of idx (multiple of 8 so rows are uint4-loadable)
    std::vector<uint16_t> idx;      // [M][d_pad] information variables of check c, ascending; padding = 0xFFFF
    // the same rows packed for the decoder's shared-memory copy: [M][iw] u32, two indices per word (low half first)
    int iw;
    std::vector<uint32_t> idx32;
    // [M] bit i set: check c is the LAST check (natural order) that updates information variable idx[c][i]
    std::vector<uint32_t> last;
    int min_reuse_distance;         // smallest gap (in checks, cyclically) between two checks sharing an information variable
};
// throws std::runtime_error on malformed tables
LdpcTables build_ldpc_tables(int K_ldpc);

struct BchTables
{
    std::vector<uint16_t> alpha_to; // [16384] antilog, alpha_to[16383] = 0 (unused slot, as in the reference)
    std::vector<int16_t> index_of;  // [16384] log, index_of[0] = -1
    uint16_t minpoly[BCH_NPOLY];    // low 14 coefficients of m_{2p+1}(x) (x^14 implicit), bit k = coeff of x^k
    // S_j = sum_k rem_{poly_of[j]}[k] * alpha^(j k), j = 1..24: synd_mat[j][k] = alpha^(j k mod 16383)
    uint8_t poly_of[2 * BCH_T + 1];
    uint16_t synd_mat[2 * BCH_T + 1][GF_M];
};
BchTables build_bch_tables();
// g(x) = product of the 12 minimal polynomials, coefficient of x^k at [k], degree 168 (169 entries of 0/1)
std::vector<uint8_t> build_bch_generator(const BchTables& t);
// BB scrambler sequence for K bits: 15-stage LFSR 1 + x^14 + x^15, initial state 100101010000000
// (src/common/Module/Scrambler/Scrambler_BB/Scrambler_BB.hxx:51-72)
std::vector<uint8_t> build_bb_scrambler(int K);

// constellation normalised to unit mean energy, index j carries bit b at (j >> b) & 1
std::vector<float> build_constellation(int bps); // 2 * 2^bps floats (re, im)
} // namespace dvbs2
