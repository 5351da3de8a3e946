/*
 * TEST INFRASTRUCTURE — CPU oracle for the DVB-S2 baseband decode chain.  NOT product code.
 *
 * A plain-C restatement of the algorithms aff3ct/dvbs2 runs on this path, each function citing the
 * reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load it; the product (libdvbs2_cuda) never links or calls it.
 *
 * Parity status: PINNED — tests/test_oracle_vs_reference.py checks every function below bit-for-bit
 * (demapper: to float rounding) against the compiled, unmodified reference (oracle/_ref/libdvbs2_ref.so,
 * built by oracle/build_ref.sh) and tests/golden/ holds vectors generated from that reference.
 *
 * Citations are relative to /root/reference; A/ = lib/aff3ct, D/ = src.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "dvbs2_tables_oracle.h"

#define N_LDPC 16200
#define QC 360
#define GF_N 16383
#define BCH_T 12

/* ------------------------------------------------------------------------------------------------ H matrix */
/* A/src/Tools/Code/LDPC/Standard/DVBS2/DVBS2_constants.cpp:12-49 (build_H) with the table of
 * A/include/Tools/Code/LDPC/Standard/DVBS2/DVBS2_constants_16200.hpp:17-39,132-156.
 * Variable x = 360*row + l joins checks (addr + l*q) mod M; parity variable K+j joins checks j and j+1.
 * Sparse_matrix::add_connection appends, so check c lists its variables in ascending order of insertion:
 * information variables ascending, then K+c-1, then K+c. */
typedef struct
{
    int K, M, E, q;
    int* off;  /* M+1 */
    int* cols; /* E, natural edge order */
} oracle_H;

static int
rate_tables(int K, const unsigned short** rowlen, const unsigned short** addr, int* nrows)
{
    if (K == 14400)
    {
        *rowlen = DVBS2_IRA_R8_9_ROWLEN, *addr = DVBS2_IRA_R8_9_ADDR, *nrows = 40;
        return 0;
    }
    if (K == 9720)
    {
        *rowlen = DVBS2_IRA_R3_5_ROWLEN, *addr = DVBS2_IRA_R3_5_ADDR, *nrows = 27;
        return 0;
    }
    return -1;
}

oracle_H*
oracle_H_new(int K)
{
    const unsigned short *rowlen, *addr;
    int nrows;
    if (rate_tables(K, &rowlen, &addr, &nrows)) return NULL;
    oracle_H* H = (oracle_H*)calloc(1, sizeof(oracle_H));
    H->K = K, H->M = N_LDPC - K, H->q = H->M / QC;
    int* deg = (int*)calloc(H->M, sizeof(int));
    for (int pass = 0; pass < 2; pass++)
    {
        const unsigned short* a = addr;
        int x = 0;
        for (int r = 0; r < nrows; r++)
        {
            for (int l = 0; l < QC; l++, x++)
                for (int k = 0; k < rowlen[r]; k++)
                {
                    int c = (a[k] + l * H->q) % H->M;
                    if (pass) H->cols[H->off[c] + deg[c]] = x;
                    deg[c]++;
                }
            a += rowlen[r];
        }
        for (x = K; x < N_LDPC; x++)
        {
            int c = x - K;
            if (pass) H->cols[H->off[c] + deg[c]] = x;
            deg[c]++;
            if (c + 1 < H->M)
            {
                if (pass) H->cols[H->off[c + 1] + deg[c + 1]] = x;
                deg[c + 1]++;
            }
        }
        if (!pass)
        {
            H->off = (int*)calloc(H->M + 1, sizeof(int));
            for (int c = 0; c < H->M; c++)
                H->off[c + 1] = H->off[c] + deg[c];
            H->E = H->off[H->M];
            H->cols = (int*)calloc(H->E, sizeof(int));
            memset(deg, 0, H->M * sizeof(int));
        }
    }
    free(deg);
    return H;
}
void
oracle_H_free(oracle_H* H)
{
    if (!H) return;
    free(H->off), free(H->cols), free(H);
}
int oracle_H_M(const oracle_H* H) { return H->M; }
int oracle_H_E(const oracle_H* H) { return H->E; }
const int* oracle_H_off(const oracle_H* H) { return H->off; }
const int* oracle_H_cols(const oracle_H* H) { return H->cols; }

/* ------------------------------------------------------------------------------------------------ LDPC decode */
/* Horizontal-layered (N)MS, sequential semantics:
 *   A/include/Module/Decoder/LDPC/BP/Horizontal_layered/Decoder_LDPC_BP_horizontal_layered.hxx:89-123 (_decode_siho),
 *   :156-178 (_decode: iterate, syndrome stop), :180-207 (_decode_single_ite: natural check order, two passes);
 *   A/include/Tools/Code/LDPC/Update_rule/MS/Update_rule_MS.hxx:48-92 (sign/min1/min2, output select, copysign);
 *   A/include/Tools/Code/LDPC/Update_rule/NMS/Update_rule_NMS.hxx:14-56,100-106 (normalise then clamp at 0);
 *   A/include/Tools/Code/LDPC/Syndrome/LDPC_syndrome.hxx:43-69 (soft syndrome on llr<0);
 *   A/src/Module/Decoder/Decoder.cpp auto-reset: every decode starts from zero messages / zero var nodes.
 * Integer types follow C++ promotion-then-truncate (wrap) exactly as the reference's R=int16_t / int8_t code. */
static int
norm_eighths(float f)
{ /* Update_rule_NMS.hxx:14-39: only k/8 legal for integers */
    for (int k = 1; k <= 8; k++)
        if (f == 0.125f * (float)k) return k;
    return -1;
}

#define DEFINE_LDPC(NAME, R, RMAX, IS_FLOAT)                                                                           \
    static R NAME##_norm(R v, float f, int k8)                                                                         \
    {                                                                                                                  \
        if (IS_FLOAT) return (R)((float)v * f);                                                                        \
        int x = (int)v, r = 0; /* div2/div4/div8 are arithmetic shifts of the promoted value (utils.hxx:18-78) */     \
        if (k8 == 8) return v;                                                                                         \
        if (k8 & 4) r += (R)(x >> 1);                                                                                  \
        if (k8 & 2) r += (R)(x >> 2);                                                                                  \
        if (k8 & 1) r += (R)(x >> 3);                                                                                  \
        return (R)r;                                                                                                   \
    }                                                                                                                  \
    int NAME(const oracle_H* H, int n_ite, float norm, int early_stop, int n_frames, const R* Y, int8_t* cwd,         \
             int32_t* V_K, int32_t* n_iter_out, R* var_out)                                                            \
    {                                                                                                                  \
        const int k8 = norm_eighths(norm);                                                                             \
        if (!IS_FLOAT && k8 < 0) return -1;                                                                            \
        R* var = (R*)malloc(sizeof(R) * N_LDPC);                                                                       \
        R* msg = (R*)malloc(sizeof(R) * H->E);                                                                         \
        R t[64];                                                                                                       \
        for (int f = 0; f < n_frames; f++)                                                                             \
        {                                                                                                              \
            for (int v = 0; v < N_LDPC; v++)                                                                           \
                var[v] = (R)((R)0 + Y[(size_t)f * N_LDPC + v]); /* _load: var_nodes(=0) += Y */                        \
            memset(msg, 0, sizeof(R) * H->E);                                                                          \
            int ok = 0, ite = 0;                                                                                       \
            for (ite = 0; ite < n_ite; ite++)                                                                          \
            {                                                                                                          \
                for (int c = 0; c < H->M; c++)                                                                         \
                {                                                                                                      \
                    const int e0 = H->off[c], d = H->off[c + 1] - e0;                                                  \
                    int sign = 0;                                                                                      \
                    R min1 = RMAX, min2 = RMAX;                                                                        \
                    for (int i = 0; i < d; i++)                                                                        \
                    {                                                                                                  \
                        t[i] = (R)(var[H->cols[e0 + i]] - msg[e0 + i]);                                                \
                        const R a = IS_FLOAT ? (R)fabsf((float)t[i]) : (R)abs((int)t[i]);                              \
                        sign ^= signbit((float)t[i]) ? -1 : 0;                                                         \
                        const R mx = a > min1 ? a : min1;                                                              \
                        min2 = min2 < mx ? min2 : mx;                                                                  \
                        min1 = min1 < a ? min1 : a;                                                                    \
                    }                                                                                                  \
                    R cst1 = NAME##_norm(min2, norm, k8), cst2 = NAME##_norm(min1, norm, k8);                          \
                    cst1 = cst1 > (R)0 ? cst1 : (R)0;                                                                  \
                    cst2 = cst2 > (R)0 ? cst2 : (R)0;                                                                  \
                    for (int i = 0; i < d; i++)                                                                        \
                    {                                                                                                  \
                        const R a = IS_FLOAT ? (R)fabsf((float)t[i]) : (R)abs((int)t[i]);                              \
                        const R m = (a == min1) ? cst1 : cst2;                                                         \
                        const int s = sign ^ (signbit((float)t[i]) ? -1 : 0);                                          \
                        const R out = (R)copysign((double)m, (double)s);                                               \
                        msg[e0 + i] = out;                                                                             \
                        var[H->cols[e0 + i]] = (R)(t[i] + out);                                                        \
                    }                                                                                                  \
                }                                                                                                      \
                ok = 1;                                                                                                \
                for (int c = 0; c < H->M && ok; c++)                                                                   \
                {                                                                                                      \
                    int s = 0;                                                                                         \
                    for (int e = H->off[c]; e < H->off[c + 1]; e++)                                                    \
                        s ^= (var[H->cols[e]] < 0) ? 1 : 0;                                                            \
                    if (s) ok = 0;                                                                                     \
                }                                                                                                      \
                if (ok && early_stop)                                                                                  \
                {                                                                                                      \
                    ite++;                                                                                             \
                    break;                                                                                             \
                }                                                                                                      \
            }                                                                                                          \
            for (int i = 0; i < H->K; i++)                                                                             \
                V_K[(size_t)f * H->K + i] = !(var[i] >= 0);                                                            \
            /* enable_syndrome == false: check_syndrome_soft is never consulted, _decode returns 0 and CWD = 1             \
             * (Decoder_LDPC_BP.hxx:8-20, ...horizontal_layered.hxx:120-123) */                                            \
            cwd[f] = (int8_t)(early_stop ? ok : 1);                                                                    \
            if (n_iter_out) n_iter_out[f] = ite;                                                                       \
            if (var_out) memcpy(var_out + (size_t)f * N_LDPC, var, sizeof(R) * N_LDPC);                                \
        }                                                                                                              \
        free(var), free(msg);                                                                                          \
        return 0;                                                                                                      \
    }

DEFINE_LDPC(oracle_ldpc_f32, float, 3.402823466e+38f, 1)
DEFINE_LDPC(oracle_ldpc_i16, int16_t, 32767, 0)
DEFINE_LDPC(oracle_ldpc_i8, int8_t, 127, 0)

/* A/src/Module/Encoder/LDPC/DVBS2/Encoder_LDPC_DVBS2.cpp:29-63: accumulate then running XOR over parity. */
int
oracle_ldpc_encode(int K, int n_frames, const int32_t* U, int32_t* X)
{
    const unsigned short *rowlen, *addr;
    int nrows;
    if (rate_tables(K, &rowlen, &addr, &nrows)) return -1;
    const int M = N_LDPC - K, q = M / QC;
    for (int f = 0; f < n_frames; f++)
    {
        const int32_t* u = U + (size_t)f * K;
        int32_t* x = X + (size_t)f * N_LDPC;
        memcpy(x, u, sizeof(int32_t) * K);
        memset(x + K, 0, sizeof(int32_t) * M);
        const unsigned short* a = addr;
        int pos = 0;
        for (int r = 0; r < nrows; r++)
        {
            for (int l = 0; l < QC; l++, pos++)
                if (u[pos] == 1)
                    for (int k = 0; k < rowlen[r]; k++)
                        x[K + (a[k] + l * q) % M] ^= 1;
            a += rowlen[r];
        }
        for (int i = 1; i < M; i++)
            x[K + i] ^= x[K + i - 1];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ GF(2^14), BCH */
/* A/src/Tools/Math/Galois.cpp:158-183 (generate_gf) with p(x) = 1+x+x^3+x^5+x^14 (D/common/Factory/DVBS2/DVBS2.hpp:55). */
static int g_alpha_to[GF_N + 1], g_index_of[GF_N + 1], g_gen[200], g_gen_len, g_gf_ready;
static void
gf_init(void)
{
    if (g_gf_ready) return;
    static const int p[15] = { 1, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1 };
    const int m = 14;
    int mask = 1;
    g_alpha_to[m] = 0;
    for (int i = 0; i < m; i++)
    {
        g_alpha_to[i] = mask;
        g_index_of[g_alpha_to[i]] = i;
        if (p[i]) g_alpha_to[m] ^= mask;
        mask <<= 1;
    }
    g_index_of[g_alpha_to[m]] = m;
    mask >>= 1;
    for (int i = m + 1; i < GF_N; i++)
    {
        if (g_alpha_to[i - 1] >= mask)
            g_alpha_to[i] = g_alpha_to[m] ^ ((g_alpha_to[i - 1] ^ mask) << 1);
        else
            g_alpha_to[i] = g_alpha_to[i - 1] << 1;
        g_index_of[g_alpha_to[i]] = i;
    }
    g_index_of[0] = -1;
    g_alpha_to[GF_N] = 0; /* the reference leaves alpha_to[N] value-initialised (vector of N+1) */

    /* A/src/Tools/Code/BCH/BCH_polynomial_generator.cpp:52-139: g(x) = prod over the conjugacy classes that contain
     * 1..2t of (x + alpha^z).  Restated: collect the union of cyclotomic cosets of 1..2t, multiply out. */
    static char used[GF_N];
    int zeros[200], nz = 0;
    memset(used, 0, sizeof(used));
    for (int rep = 1; rep < GF_N && nz < 200; rep++)
    { /* cosets in increasing representative order, as the reference enumerates them */
        if (used[rep]) continue;
        int cos[16], nc = 0, z = rep, hit = 0;
        do
        {
            cos[nc++] = z, used[z] = 1;
            if (z >= 1 && z <= 2 * BCH_T) hit = 1;
            z = (z * 2) % GF_N;
        } while (z != rep);
        if (hit)
            for (int i = 0; i < nc; i++)
                zeros[nz++] = cos[i];
        if (rep > 2 * BCH_T) break;
    }
    g_gen_len = nz + 1;
    g_gen[0] = g_alpha_to[zeros[0]], g_gen[1] = 1;
    for (int i = 2; i <= nz; i++)
    {
        g_gen[i] = 1;
        for (int j = i - 1; j > 0; j--)
            g_gen[j] = g_gen[j] ? g_gen[j - 1] ^ g_alpha_to[(g_index_of[g_gen[j]] + zeros[i - 1]) % GF_N] : g_gen[j - 1];
        g_gen[0] = g_alpha_to[(g_index_of[g_gen[0]] + zeros[i - 1]) % GF_N];
    }
    g_gf_ready = 1;
}
void
oracle_gf_tables(int* alpha_to, int* index_of, int* g, int* g_len)
{
    gf_init();
    memcpy(alpha_to, g_alpha_to, sizeof(int) * (GF_N + 1));
    memcpy(index_of, g_index_of, sizeof(int) * (GF_N + 1));
    memcpy(g, g_gen, sizeof(int) * g_gen_len);
    *g_len = g_gen_len;
}

/* D/common/Module/Encoder_BCH_DVBS2/Encoder_BCH_DVBS2.cpp:29-44 around A/src/Module/Encoder/BCH/Encoder_BCH.cpp:46-68:
 * systematic LFSR division by g(x) on the bit-reversed message, codeword reversed back: X = [U | parity]. */
int
oracle_bch_encode(int K, int N, int n_frames, const int32_t* U, int32_t* X)
{
    gf_init();
    const int R = N - K;
    if (R != g_gen_len - 1) return -1;
    int par[200];
    for (int f = 0; f < n_frames; f++)
    {
        const int32_t* u = U + (size_t)f * K;
        int32_t* x = X + (size_t)f * N;
        memset(par, 0, sizeof(par));
        for (int i = 0; i < K; i++) /* reversed message index K-1-i scanned from its top = natural order */
        {
            const int fb = u[i] ^ par[R - 1];
            for (int j = R - 1; j > 0; j--)
                par[j] = par[j - 1] ^ (g_gen[j] & fb);
            par[0] = fb ? (g_gen[0] && fb) : 0;
        }
        memcpy(x, u, sizeof(int32_t) * K);
        for (int j = 0; j < R; j++)
            x[K + j] = par[R - 1 - j];
    }
    return 0;
}

/* D/common/Module/Decoder_BCH_DVBS2/Decoder_BCH_DVBS2.cpp:28-40 (bit reversal in/out, CWD = !status) around
 * A/src/Module/Decoder/BCH/Standard/Decoder_BCH_std.cpp:52-221:
 *   :58-67 syndromes S_i = XOR_j r[j] alpha^(i j), i = 1..2t;  :71-167 Berlekamp iterative algorithm in Lin & Costello's
 *   tabular form (rows mu = -1..2t: discrepancy d, polynomial, degree l, mu - l), aborting once l > t;
 *   :176-201 Chien search over all 16383 field elements;  :203-217 accept iff #roots == l, flip located bits inside
 *   the shortened length only.  Returns per frame status (0 ok / 1 failure), cwd = !status. */
int
oracle_bch_decode(int K, int N, int n_frames, const int32_t* Y, int8_t* cwd, int32_t* V, int32_t* n_corrected)
{
    gf_init();
    const int t = BCH_T, t2 = 2 * BCH_T;
    int* r = (int*)malloc(sizeof(int) * N);
    static int elp[2 * BCH_T + 3][2 * BCH_T + 4]; /* rows keep the reference's mixed index/polynomial form */
    int d[2 * BCH_T + 3], l[2 * BCH_T + 3], u_lu[2 * BCH_T + 3], s[2 * BCH_T + 1], loc[BCH_T + 1], reg[BCH_T + 1];
    for (int f = 0; f < n_frames; f++)
    {
        for (int j = 0; j < N; j++)
            r[j] = Y[(size_t)f * N + (N - 1 - j)];
        int syn_error = 0, status = 0, ncorr = 0;
        for (int i = 1; i <= t2; i++)
        {
            int acc = 0;
            for (int j = 0; j < N; j++)
                if (r[j]) acc ^= g_alpha_to[(int)(((long)i * j) % GF_N)];
            if (acc) syn_error = 1;
            s[i] = g_index_of[acc];
        }
        if (syn_error)
        {
            d[0] = 0, d[1] = s[1];
            memset(elp, 0, sizeof(elp));
            elp[0][0] = 0, elp[1][0] = 1;
            for (int i = 1; i < t2; i++)
                elp[0][i] = -1, elp[1][i] = 0;
            l[0] = l[1] = 0, u_lu[0] = -1, u_lu[1] = 0;
            int u = 0, q;
            do
            {
                u++;
                if (d[u] == -1)
                {
                    l[u + 1] = l[u];
                    for (int i = 0; i <= l[u]; i++)
                        elp[u + 1][i] = elp[u][i], elp[u][i] = g_index_of[elp[u][i]];
                }
                else
                {
                    q = u - 1;
                    while (d[q] == -1 && q > 0)
                        q--;
                    if (q > 0)
                    {
                        int j = q;
                        do
                        {
                            j--;
                            if (d[j] != -1 && u_lu[q] < u_lu[j]) q = j;
                        } while (j > 0);
                    }
                    l[u + 1] = (l[u] > l[q] + u - q) ? l[u] : l[q] + u - q;
                    for (int i = 0; i < t2; i++)
                        elp[u + 1][i] = 0;
                    for (int i = 0; i <= l[q]; i++)
                        if (elp[q][i] != -1) elp[u + 1][i + u - q] = g_alpha_to[(d[u] + GF_N - d[q] + elp[q][i]) % GF_N];
                    for (int i = 0; i <= l[u]; i++)
                        elp[u + 1][i] ^= elp[u][i], elp[u][i] = g_index_of[elp[u][i]];
                }
                u_lu[u + 1] = u - l[u + 1];
                if (u < t2)
                {
                    int acc = (s[u + 1] != -1) ? g_alpha_to[s[u + 1]] : 0;
                    for (int i = 1; i <= l[u + 1]; i++)
                        if (s[u + 1 - i] != -1 && elp[u + 1][i] != 0)
                            acc ^= g_alpha_to[(s[u + 1 - i] + g_index_of[elp[u + 1][i]]) % GF_N];
                    d[u + 1] = g_index_of[acc];
                }
            } while (u < t2 && l[u + 1] <= t);
            u++;
            status = 1;
            if (l[u] <= t)
            {
                for (int i = 0; i <= l[u]; i++)
                    elp[u][i] = g_index_of[elp[u][i]];
                for (int i = 1; i <= l[u]; i++)
                    reg[i] = elp[u][i];
                int count = 0, overflow = 0;
                for (int i = 1; i <= GF_N && !overflow; i++)
                {
                    int acc = 1;
                    for (int j = 1; j <= l[u]; j++)
                        if (reg[j] != -1)
                        {
                            reg[j] = (reg[j] + j) % GF_N;
                            acc ^= g_alpha_to[reg[j]];
                        }
                    if (!acc)
                    {
                        if (count >= t + 1)
                            overflow = 1; /* the reference throws here; cannot happen for deg <= t */
                        else
                            loc[count++] = GF_N - i;
                    }
                }
                if (!overflow && count == l[u])
                {
                    for (int i = 0; i < l[u]; i++)
                        if (loc[i] < N) r[loc[i]] ^= 1, ncorr++;
                    status = 0;
                }
            }
        }
        for (int i = 0; i < K; i++)
            V[(size_t)f * K + i] = r[N - 1 - i];
        cwd[f] = (int8_t)!status;
        if (n_corrected) n_corrected[f] = ncorr;
    }
    free(r);
    return 0;
}

/* D/common/Module/Scrambler/Scrambler_BB/Scrambler_BB.hxx:51-72: 15-stage LFSR 1+x^14+x^15, init 100101010000000. */
int
oracle_bb_scramble(int K, int n_frames, const int32_t* in, int32_t* out)
{
    static const int init[15] = { 1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0 };
    for (int f = 0; f < n_frames; f++)
    {
        int lfsr[15];
        memcpy(lfsr, init, sizeof(init));
        for (int i = 0; i < K; i++)
        {
            const int fb = (lfsr[14] + lfsr[13]) % 2;
            memmove(lfsr + 1, lfsr, 14 * sizeof(int));
            lfsr[0] = fb;
            out[(size_t)f * K + i] = (in[(size_t)f * K + i] + fb) % 2;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ framer / estimator */
/* D/common/Module/Framer/Framer.hxx:330-343 (_remove_plh): drop the 90-symbol PL header, then the 36-symbol pilot block
 * that follows every 16 slots (1440 symbols) of payload; N_pilots = n_sym / 1440 (Framer.hxx:36-38),
 * pl_frame_size = 90 * (n_sym / 90 + 1) + 36 * N_pilots symbols (D/common/Factory/DVBS2/DVBS2.cpp:352-355).
 * in: [F][2 * pl_sym] floats (I/Q interleaved), out: [F][2 * n_sym]. */
int
oracle_pl_frame_symbols(int n_sym)
{
    return 90 * (n_sym / 90 + 1) + 36 * (n_sym / 1440);
}

int
oracle_remove_plh(int n_sym, int n_frames, const float* in, float* out)
{
    const int M = 90, P = 36, n_pilots = n_sym / (16 * M), pl = oracle_pl_frame_symbols(n_sym);
    for (int f = 0; f < n_frames; f++)
    {
        const float* src = in + (size_t)f * 2 * pl;
        float* dst = out + (size_t)f * 2 * n_sym;
        int pos = M; /* symbol index in the PL frame */
        for (int k = 0; k < n_sym; k++)
        {
            /* the i-th erase (i = 1..n_pilots) removes the P symbols that follow payload symbol 16*M*i - 1 */
            if (k > 0 && k % (16 * M) == 0 && k / (16 * M) <= n_pilots) pos += P;
            dst[2 * k] = src[2 * (pos + k)];
            dst[2 * k + 1] = src[2 * (pos + k) + 1];
        }
    }
    return 0;
}

/* D/common/Module/Estimator/Estimator_DVBS2.hxx:31-58 (_estimate, M2M4 moments, all float, summed in symbol order) with
 * A/src/Tools/general_utils.cpp:78-89 (esn0_to_sigma) and :91-113 (esn0_to_ebn0); code_rate = K_bch / N_ldpc as floats
 * (D/common/Factory/DVBS2/DVBS2.cpp:557).  X: [F][2 * n_sym]; outputs [F] each (Estimator.hxx:103-117). */
int
oracle_estimate(int n_sym, int K_bch, int bps, int n_frames, const float* X, float* sig, float* ebn0, float* esn0)
{
    const float code_rate = (float)K_bch / (float)16200;
    for (int f = 0; f < n_frames; f++)
    {
        const float* x = X + (size_t)f * 2 * n_sym;
        float m2 = 0, m4 = 0;
        for (int i = 0; i < n_sym; i++)
        {
            const float tmp = x[2 * i] * x[2 * i] + x[2 * i + 1] * x[2 * i + 1];
            m2 += tmp;
            m4 += tmp * tmp;
        }
        m2 /= n_sym;
        m4 /= n_sym;
        const float Se = sqrtf(fabsf(2 * m2 * m2 - m4));
        const float Ne = fabsf(m2 - Se);
        float e = 10 * log10f(Se / Ne);
        if (isinf(e)) e = 100.f;
        sig[f] = sqrtf((float)1 / ((float)2 * powf((float)10, e / (float)10)));
        if (ebn0) ebn0[f] = e - (float)10 * log10f(code_rate * (float)bps);
        if (esn0) esn0[f] = e;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ interleaver */
/* A/src/Tools/Interleaver/Column_row/Interleaver_core_column_row.cpp:59-87 (gen_lut TOP_LEFT / TOP_RIGHT) and
 * A/include/Tools/Interleaver/Interleaver_core.hxx:137-142 (pi_inv[pi[i]] = i); selection per MODCOD in
 * D/common/Factory/DVBS2/DVBS2.cpp:300-317,451-464.  order: 0 = none (QPSK), 1 = TOP_LEFT, 2 = TOP_RIGHT.
 * A/include/Module/Interleaver/Interleaver.hxx:228-258: itl[i] = nat[pi[i]], nat[i] = itl[pi_inv[i]]. */
int
oracle_itl_lut(int n_cols, int order, uint32_t* pi, uint32_t* pi_inv)
{
    const int n_rows = N_LDPC / n_cols;
    uint32_t idx = 0;
    if (order == 0)
        for (int i = 0; i < N_LDPC; i++)
            pi[i] = (uint32_t)i;
    else if (order == 1)
        for (int i = 0; i < n_cols; i++)
            for (int j = 0; j < n_rows; j++)
                pi[j * n_cols + i] = idx++;
    else if (order == 2)
        for (int i = n_cols - 1; i >= 0; i--)
            for (int j = 0; j < n_rows; j++)
                pi[j * n_cols + i] = idx++;
    else
        return -1;
    for (int i = 0; i < N_LDPC; i++)
        pi_inv[pi[i]] = (uint32_t)i;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ modem */
/* A/include/Tools/Constellation/User/Constellation_user.hxx:22-67: points parsed as float, divided by
 * sqrt(mean |c|^2) (float accumulation, complex<float> division by (sqrt_es + 0i)). */
int
oracle_constellation(int bps, float* cstl /* 2*2^bps */)
{
    const double(*src)[2] = bps == 2 ? DVBS2_CSTL_QPSK : bps == 3 ? DVBS2_CSTL_PSK8 : bps == 4 ? DVBS2_CSTL_APSK16 : NULL;
    if (!src) return -1;
    const int n = 1 << bps;
    float es = 0.f;
    for (int j = 0; j < n; j++)
    {
        const float re = (float)src[j][0], im = (float)src[j][1];
        es += re * re + im * im;
    }
    es = sqrtf(es / (float)n);
    for (int j = 0; j < n; j++)
    {
        float _Complex z = ((float)src[j][0] + (float)src[j][1] * 1.0fi) / (es + 0.0fi);
        cstl[2 * j] = __real__ z, cstl[2 * j + 1] = __imag__ z;
    }
    return 0;
}

/* A/include/Module/Modem/Generic/Modem_generic.hxx:150-174: symbol index = sum_b X[i*bps+b] << b. */
int
oracle_modulate(int bps, int n_frames, const int32_t* X, float* sym)
{
    float cstl[32];
    if (oracle_constellation(bps, cstl)) return -1;
    const int ns = N_LDPC / bps;
    for (int f = 0; f < n_frames; f++)
        for (int i = 0; i < ns; i++)
        {
            unsigned idx = 0;
            for (int b = 0; b < bps; b++)
                idx += (unsigned)(1 << b) * (unsigned)X[(size_t)f * N_LDPC + i * bps + b];
            sym[((size_t)f * ns + i) * 2] = cstl[2 * idx], sym[((size_t)f * ns + i) * 2 + 1] = cstl[2 * idx + 1];
        }
    return 0;
}

/* A/include/Module/Modem/Generic/Modem_generic.hxx:100-110 (inv_sigma2 = 1/(2 sigma^2), float) and
 * A/include/Module/Modem/Generic/Modem_generic_fast.hxx:204-224 (scalar form of the log-MAP demapper) with
 * A/include/Tools/Math/max.hxx:29-36 (max* = max + log1p(exp(-|a-b|))).  The reference's SIMD body (:72-138) evaluates
 * the same recursion with MIPP's polynomial exp/log, so parity with the reference is to float rounding (<= 1e-4 rel.),
 * as SURVEY App. A.1 states; the decoder is judged on identical LLRs. */
static float
max_star_f(float a, float b)
{
    return (a > b ? a : b) + (float)log1p(exp(-fabsf(a - b)));
}
int
oracle_demodulate(int bps, int n_frames, const float* sigma, const float* sym, float* llr)
{
    float cstl[32];
    if (oracle_constellation(bps, cstl)) return -1;
    const int ns = N_LDPC / bps, np = 1 << bps;
    for (int f = 0; f < n_frames; f++)
    {
        const float inv = (float)(1.0f / (2 * sigma[f] * sigma[f]));
        for (int n = 0; n < N_LDPC; n++)
        {
            const int b = n % bps, k = n / bps;
            const float yr = sym[((size_t)f * ns + k) * 2], yi = sym[((size_t)f * ns + k) * 2 + 1];
            float L0 = -INFINITY, L1 = -INFINITY;
            for (int j = 0; j < np; j++)
            {
                const float dr = yr - cstl[2 * j], di = yi - cstl[2 * j + 1];
                const float m = -(dr * dr + di * di) * inv;
                if (((j >> b) & 1) == 0)
                    L0 = max_star_f(L0, m);
                else
                    L1 = max_star_f(L1, m);
            }
            llr[(size_t)f * N_LDPC + n] = L0 - L1;
        }
    }
    return 0;
}

/* A/src/Module/Quantizer/Pow2/Quantizer_pow2.cpp:66-123,184-190: q = sat(round(2^p y), +-(2^(s-1)-1)), round half away. */
int
oracle_quantize_i16(size_t n, int fixed_point_pos, int saturation_pos, const float* in, int16_t* out)
{
    const float factor = (float)(1 << fixed_point_pos);
    const float vmax = (float)((1 << (saturation_pos - 1)) - 1), vmin = -vmax;
    for (size_t i = 0; i < n; i++)
    {
        float v = roundf(factor * in[i]);
        v = v < vmin ? vmin : v > vmax ? vmax : v;
        out[i] = (int16_t)v;
    }
    return 0;
}
int
oracle_quantize_i8(size_t n, int fixed_point_pos, int saturation_pos, const float* in, int8_t* out)
{
    const float factor = (float)(1 << fixed_point_pos);
    const float vmax = (float)((1 << (saturation_pos - 1)) - 1), vmin = -vmax;
    for (size_t i = 0; i < n; i++)
    {
        float v = roundf(factor * in[i]);
        v = v < vmin ? vmin : v > vmax ? vmax : v;
        out[i] = (int8_t)v;
    }
    return 0;
}
