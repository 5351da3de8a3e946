// See code_tables.h.  Host-only C++; data tables come from dvbs2_tables.h (ETSI EN 302 307 short-frame tables).
#include "code_tables.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <stdexcept>

#include "dvbs2_tables.h"

namespace dvbs2
{
bool
modcod_params(int modcod, ModcodParams& p)
{
    // D/common/Factory/DVBS2/DVBS2.cpp:287-356 (modcod_init)
    static const ModcodParams table[5] = {
        { 0, 2, 8100, 14232, 14400, 1800, 5, 25, 0, 0, "QPSK-S_8/9" },
        { 1, 2, 8100, 9552, 9720, 6480, 18, 9, 0, 0, "QPSK-S_3/5" },
        { 2, 3, 5400, 9552, 9720, 6480, 18, 9, 3, 1, "8PSK-S_3/5" },
        { 3, 3, 5400, 14232, 14400, 1800, 5, 25, 3, 0, "8PSK-S_8/9" },
        { 4, 4, 4050, 14232, 14400, 1800, 5, 25, 4, 0, "16APSK-S_8/9" },
    };
    if (modcod < 0 || modcod > 4) return false;
    p = table[modcod];
    return true;
}

LdpcTables
build_ldpc_tables(int K)
{
    const unsigned short *rowlen, *addr;
    int nrows;
    if (K == 14400)
        rowlen = DVBS2_IRA_R8_9_ROWLEN, addr = DVBS2_IRA_R8_9_ADDR, nrows = 40;
    else if (K == 9720)
        rowlen = DVBS2_IRA_R3_5_ROWLEN, addr = DVBS2_IRA_R3_5_ADDR, nrows = 27;
    else
        throw std::runtime_error("unsupported K_ldpc");
    if (nrows * QC != K) throw std::runtime_error("IRA table does not cover K information bits");

    LdpcTables t;
    t.K = K, t.M = N_LDPC - K;
    const int q = t.M / QC;
    // Information bit 360*r + l feeds parity accumulators (a + l*q) mod M for every address a of row r.
    // Bucket by check; scanning variables in ascending order leaves every bucket sorted.
    std::vector<std::vector<uint16_t>> per_check(t.M);
    const unsigned short* a = addr;
    for (int r = 0; r < nrows; r++)
    {
        for (int l = 0; l < QC; l++)
            for (int k = 0; k < rowlen[r]; k++)
                per_check[(a[k] + l * q) % t.M].push_back((uint16_t)(QC * r + l));
        a += rowlen[r];
    }
    t.d_info = (int)per_check[0].size();
    for (auto& v : per_check)
    {
        if ((int)v.size() != t.d_info) throw std::runtime_error("non-uniform check degree");
        if (std::adjacent_find(v.begin(), v.end()) != v.end()) throw std::runtime_error("duplicate edge");
    }
    t.d_pad = (t.d_info + 7) / 8 * 8;
    t.idx.assign((size_t)t.M * t.d_pad, 0xFFFF);
    for (int c = 0; c < t.M; c++)
        std::copy(per_check[c].begin(), per_check[c].end(), t.idx.begin() + (size_t)c * t.d_pad);
    // + the dual diagonal: check c also holds parity bits K+c-1 (c>0) and K+c
    t.E = t.M * (t.d_info + 2) - 1;

    t.iw = (t.d_info + 1) / 2;
    t.idx32.assign((size_t)t.M * t.iw, 0u);
    t.last.assign(t.M, 0u);
    std::vector<int> last_check(K, -1), first_check(K, -1), prev_check(K, -1);
    t.min_reuse_distance = t.M;
    for (int c = 0; c < t.M; c++)
        for (int i = 0; i < t.d_info; i++)
        {
            const int v = per_check[c][i];
            t.idx32[(size_t)c * t.iw + i / 2] |= (uint32_t)v << ((i & 1) * 16);
            if (first_check[v] < 0) first_check[v] = c;
            if (prev_check[v] >= 0) t.min_reuse_distance = std::min(t.min_reuse_distance, c - prev_check[v]);
            prev_check[v] = last_check[v] = c;
        }
    for (int v = 0; v < K; v++)
    {
        if (last_check[v] < 0) throw std::runtime_error("information variable without a check");
        t.min_reuse_distance = std::min(t.min_reuse_distance, first_check[v] + t.M - last_check[v]);
    }
    for (int c = 0; c < t.M; c++)
        for (int i = 0; i < t.d_info; i++)
            if (last_check[per_check[c][i]] == c) t.last[c] |= 1u << i;
    return t;
}

BchTables
build_bch_tables()
{
    BchTables t;
    t.alpha_to.assign(GF_N + 1, 0);
    t.index_of.assign(GF_N + 1, -1);
    // p(x) = 1 + x + x^3 + x^5 + x^14 (D/common/Factory/DVBS2/DVBS2.hpp:55)
    const unsigned prim = (1u << 14) | (1u << 5) | (1u << 3) | (1u << 1) | 1u;
    unsigned x = 1;
    for (int i = 0; i < GF_N; i++)
    {
        t.alpha_to[i] = (uint16_t)x;
        t.index_of[x] = (int16_t)i;
        x <<= 1;
        if (x & (1u << GF_M)) x ^= prim;
    }
    auto mul = [&](unsigned a, unsigned b) -> unsigned {
        if (!a || !b) return 0;
        return t.alpha_to[(t.index_of[a] + t.index_of[b]) % GF_N];
    };
    // minimal polynomial of alpha^i: product over the conjugates alpha^(i 2^e)
    for (int p = 0; p < BCH_NPOLY; p++)
    {
        const int i = 2 * p + 1;
        std::vector<unsigned> poly{ 1 }; // coefficients, low degree first
        int z = i;
        do
        {
            std::vector<unsigned> nxt(poly.size() + 1, 0);
            const unsigned root = t.alpha_to[z];
            for (size_t k = 0; k < poly.size(); k++)
            {
                nxt[k + 1] ^= poly[k];          // * x
                nxt[k] ^= mul(poly[k], root);   // * alpha^z
            }
            poly.swap(nxt);
            z = (z * 2) % GF_N;
        } while (z != i);
        if (poly.size() != GF_M + 1) throw std::runtime_error("minimal polynomial degree != 14");
        unsigned mask = 0;
        for (int k = 0; k <= GF_M; k++)
        {
            if (poly[k] > 1) throw std::runtime_error("minimal polynomial not binary");
            if (poly[k]) mask |= 1u << k;
        }
        t.minpoly[p] = (uint16_t)(mask & 0x3FFF);
    }
    for (int j = 1; j <= 2 * BCH_T; j++)
    {
        int odd = j;
        while (!(odd & 1))
            odd >>= 1;
        t.poly_of[j] = (uint8_t)((odd - 1) / 2);
        for (int k = 0; k < GF_M; k++)
            t.synd_mat[j][k] = t.alpha_to[(j * k) % GF_N];
    }
    t.poly_of[0] = 0;
    for (int k = 0; k < GF_M; k++)
        t.synd_mat[0][k] = 0;
    return t;
}

std::vector<uint8_t>
build_bch_generator(const BchTables& t)
{
    std::vector<uint8_t> g{ 1 };
    for (int p = 0; p < BCH_NPOLY; p++)
    {
        const unsigned mp = (unsigned)t.minpoly[p] | (1u << GF_M);
        std::vector<uint8_t> nxt(g.size() + GF_M, 0);
        for (size_t i = 0; i < g.size(); i++)
            if (g[i])
                for (int k = 0; k <= GF_M; k++)
                    if ((mp >> k) & 1u) nxt[i + k] ^= 1;
        g.swap(nxt);
    }
    if (g.size() != (size_t)BCH_NPOLY * GF_M + 1 || !g.front() || !g.back()) throw std::runtime_error("bad BCH generator");
    return g;
}

std::vector<uint8_t>
build_bb_scrambler(int K)
{
    int lfsr[15] = { 1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0 };
    std::vector<uint8_t> seq(K);
    for (int i = 0; i < K; i++)
    {
        const int fb = lfsr[14] ^ lfsr[13];
        for (int k = 14; k > 0; k--)
            lfsr[k] = lfsr[k - 1];
        lfsr[0] = fb;
        seq[i] = (uint8_t)fb;
    }
    return seq;
}

std::vector<float>
build_constellation(int bps)
{
    const double(*src)[2] = bps == 2 ? DVBS2_CSTL_QPSK : bps == 3 ? DVBS2_CSTL_PSK8 : bps == 4 ? DVBS2_CSTL_APSK16 : nullptr;
    if (!src) throw std::runtime_error("unsupported bits per symbol");
    const int n = 1 << bps;
    // same float operations as Constellation_user.hxx:43-65 so the points are bit-identical to the reference's
    std::vector<std::complex<float>> c(n);
    float es = 0.f;
    for (int j = 0; j < n; j++)
    {
        c[j] = std::complex<float>((float)src[j][0], (float)src[j][1]);
        es += std::norm(c[j]);
    }
    es = std::sqrt(es / n);
    std::vector<float> out(2 * n);
    for (int j = 0; j < n; j++)
    {
        c[j] /= std::complex<float>(es);
        out[2 * j] = c[j].real(), out[2 * j + 1] = c[j].imag();
    }
    return out;
}
} // namespace dvbs2
