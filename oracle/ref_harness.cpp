// TEST INFRASTRUCTURE — not product code.
//
// C-callable harness around the UNMODIFIED reference (aff3ct/dvbs2 with its vendored
// AFF3CT + StreamPU), compiled in place from /root/reference by oracle/build_ref.sh into
// oracle/_ref/libdvbs2_ref.so.  It exists to (1) pin oracle/dvbs2_oracle.c against the
// real reference classes and (2) serve as the "reference" CPU arm of bench.py.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it.
//
// Every entry point instantiates the very class the dvbs2 binaries build
// (src/common/Factory/DVBS2/DVBS2.cpp:406-488) and calls its public task method.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <aff3ct.hpp>
#include <streampu.hpp>

#include "Module/Decoder/LDPC/BP/Horizontal_layered/Decoder_LDPC_BP_horizontal_layered.hpp"
#include "Module/Decoder/LDPC/BP/Horizontal_layered/ONMS/Decoder_LDPC_BP_horizontal_layered_ONMS_inter.hpp"
#include "Module/Decoder/LDPC/BP/Horizontal_layered/Decoder_LDPC_BP_horizontal_layered_inter.hpp" // C++17 only
#include "Tools/Code/LDPC/Update_rule/NMS/Update_rule_NMS_simd.hpp"
#include "Tools/Code/LDPC/Update_rule/NMS/Update_rule_NMS.hpp"
#include "Tools/Code/LDPC/Standard/DVBS2/DVBS2_constants.hpp"

#include "Module/Decoder_BCH_DVBS2/Decoder_BCH_DVBS2.hpp"
#include "Module/Encoder_BCH_DVBS2/Encoder_BCH_DVBS2.hpp"
#include "Module/Scrambler/Scrambler_BB/Scrambler_BB.hpp"
#include "Module/Framer/Framer.hpp"
#include "Module/Estimator/Estimator_DVBS2.hpp"

using namespace aff3ct;

namespace
{
const std::vector<int> kPrimPoly = { 1, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1 }; // DVBS2.hpp:55

std::vector<unsigned>
info_pos(int K)
{
    std::vector<unsigned> p(K);
    for (int i = 0; i < K; i++)
        p[i] = (unsigned)i;
    return p;
}

template<typename B, typename R>
int
ldpc_seq(int K, int n_ite, float norm, int n_frames, const R* Y, int8_t* cwd, B* V, bool enable_syndrome = true)
{
    const int N = 16200;
    auto dv = tools::build_dvbs2(K, N);
    auto H = tools::build_H(*dv);
    auto ibp = info_pos(K);
    tools::Update_rule_NMS<R> rule(norm);
    module::Decoder_LDPC_BP_horizontal_layered<B, R, tools::Update_rule_NMS<R>> dec(K, N, n_ite, H, ibp, rule,
                                                                                   enable_syndrome, 1);
    dec.set_n_frames(1);
    for (int f = 0; f < n_frames; f++)
        dec.decode_siho(Y + (size_t)f * N, cwd + f, V + (size_t)f * K);
    return 0;
}

template<typename B, typename R>
int
ldpc_inter(int K, int n_ite, float norm, int n_frames, const R* Y, int8_t* cwd, B* V)
{
    const int N = 16200;
    auto dv = tools::build_dvbs2(K, N);
    auto H = tools::build_H(*dv);
    auto ibp = info_pos(K);
    module::Decoder_LDPC_BP_horizontal_layered_ONMS_inter<B, R> dec(K, N, n_ite, H, ibp, norm, (R)0, true, 1);
    const int W = (int)mipp::N<R>();
    if (n_frames % W) return -1;
    dec.set_n_frames(W);
    for (int f = 0; f < n_frames; f += W)
        dec.decode_siho(Y + (size_t)f * N, cwd + f, V + (size_t)f * K);
    return W;
}

// the templated inter-frame SIMD decoder AFF3CT's factory builds in C++17 builds
// (Decoder_LDPC.cpp:540-650: Update_rule_NMS_simd<Q> for float, <Q, k8> with norm = k8/8 for the integer types)
template<typename B, typename R, class Rule>
int
ldpc_tinter_rule(int K, int n_ite, float norm, int n_frames, const R* Y, int8_t* cwd, B* V)
{
    const int N = 16200;
    auto dv = tools::build_dvbs2(K, N);
    auto H = tools::build_H(*dv);
    auto ibp = info_pos(K);
    module::Decoder_LDPC_BP_horizontal_layered_inter<B, R, Rule> dec(K, N, n_ite, H, ibp, Rule(norm), true, 1);
    const int W = (int)mipp::N<R>();
    if (n_frames % W) return -1;
    dec.set_n_frames(W);
    for (int f = 0; f < n_frames; f += W)
        dec.decode_siho(Y + (size_t)f * N, cwd + f, V + (size_t)f * K);
    return W;
}
template<typename B, typename R>
int
ldpc_tinter(int K, int n_ite, float norm, int n_frames, const R* Y, int8_t* cwd, B* V)
{
    if (sizeof(R) == 4) return ldpc_tinter_rule<B, R, tools::Update_rule_NMS_simd<R>>(K, n_ite, norm, n_frames, Y, cwd, V);
#define K8(k)                                                                                                          \
    if (norm == 0.125f * k) return ldpc_tinter_rule<B, R, tools::Update_rule_NMS_simd<R, k>>(K, n_ite, norm, n_frames, Y, cwd, V)
    K8(1); K8(2); K8(3); K8(4); K8(5); K8(6); K8(7); K8(8);
#undef K8
    return -2;
}

// frames [0, n_frames) split over n_threads decoder instances (whole SIMD waves per thread)
template<typename F>
int
run_mt(int n_frames, int n_threads, int wave, F&& body)
{
    if (n_threads < 1) n_threads = 1;
    const int waves = (n_frames + wave - 1) / wave;
    n_threads = std::min(n_threads, std::max(1, waves));
    std::vector<std::thread> th;
    std::vector<int> rc(n_threads, 0);
    for (int t = 0; t < n_threads; t++)
    {
        const int f0 = (int)((long)waves * t / n_threads) * wave;
        const int f1 = std::min(n_frames, (int)((long)waves * (t + 1) / n_threads) * wave);
        th.emplace_back([&, t, f0, f1]() { rc[t] = f1 > f0 ? body(f0, f1 - f0) : 0; });
    }
    for (auto& x : th) x.join();
    int worst = 0;
    for (int r : rc) worst = std::min(worst, r);
    return worst;
}
} // namespace

extern "C"
{
    int ref_simd_lanes_f32() { return (int)mipp::N<float>(); }
    int ref_simd_lanes_i16() { return (int)mipp::N<int16_t>(); }

    // H as built by tools::build_H (DVBS2_constants.cpp:12-49): per check degree + variables in insertion order
    int ref_build_H(int K, int* n_checks, int* n_edges, int* deg, int* cols)
    {
        auto dv = tools::build_dvbs2(K, 16200);
        auto H = tools::build_H(*dv);
        const int M = (int)H.get_n_cols();
        *n_checks = M;
        *n_edges = (int)H.get_n_connections();
        if (deg == nullptr) return 0;
        size_t e = 0;
        for (int c = 0; c < M; c++)
        {
            deg[c] = (int)H[c].size();
            for (auto v : H[c])
                cols[e++] = (int)v;
        }
        return 0;
    }

    int ref_ldpc_seq_f32(int K, int n_ite, float norm, int n_frames, const float* Y, int8_t* cwd, int32_t* V)
    {
        return ldpc_seq<int32_t, float>(K, n_ite, norm, n_frames, Y, cwd, V);
    }
    int ref_ldpc_seq_i16(int K, int n_ite, float norm, int n_frames, const int16_t* Y, int8_t* cwd, int16_t* V)
    {
        return ldpc_seq<int16_t, int16_t>(K, n_ite, norm, n_frames, Y, cwd, V);
    }
    int ref_ldpc_seq_i8(int K, int n_ite, float norm, int n_frames, const int8_t* Y, int8_t* cwd, int8_t* V)
    {
        return ldpc_seq<int8_t, int8_t>(K, n_ite, norm, n_frames, Y, cwd, V);
    }
    int ref_ldpc_inter_f32(int K, int n_ite, float norm, int n_frames, const float* Y, int8_t* cwd, int32_t* V)
    {
        return ldpc_inter<int32_t, float>(K, n_ite, norm, n_frames, Y, cwd, V);
    }
    int ref_ldpc_inter_i16(int K, int n_ite, float norm, int n_frames, const int16_t* Y, int8_t* cwd, int16_t* V)
    {
        return ldpc_inter<int16_t, int16_t>(K, n_ite, norm, n_frames, Y, cwd, V);
    }

    // a9: Decoder_LDPC_BP_horizontal_layered_inter<B,R,Update_rule_NMS_simd<R[,k8]>> (…_inter.hxx:95-108,220-271)
    int ref_ldpc_tinter_f32(int K, int n_ite, float norm, int n_frames, const float* Y, int8_t* cwd, int32_t* V)
    {
        return ldpc_tinter<int32_t, float>(K, n_ite, norm, n_frames, Y, cwd, V);
    }
    int ref_ldpc_tinter_i16(int K, int n_ite, float norm, int n_frames, const int16_t* Y, int8_t* cwd, int16_t* V)
    {
        return ldpc_tinter<int16_t, int16_t>(K, n_ite, norm, n_frames, Y, cwd, V);
    }
    // enable_syndrome = false (Decoder_LDPC_BP.hxx:8-20: the syndrome is never tested, status 0, CWD = 1)
    int ref_ldpc_seq_f32_nosyn(int K, int n_ite, float norm, int n_frames, const float* Y, int8_t* cwd, int32_t* V)
    {
        return ldpc_seq<int32_t, float>(K, n_ite, norm, n_frames, Y, cwd, V, false);
    }

    // the same decoders, frames split over n_threads instances (for >= 1e4-frame parity runs on the GPU box's host cores).
    // kind: 0 sequential, 1 ONMS_inter, 2 templated inter;  prec: 0 float/int32 bits, 1 int16
    int ref_ldpc_mt(int kind, int prec, int K, int n_ite, float norm, int n_frames, int n_threads, const void* Y, int8_t* cwd,
                    void* V)
    {
        const size_t N = 16200;
        const int wave = kind == 0 ? 1 : (prec == 0 ? (int)mipp::N<float>() : (int)mipp::N<int16_t>());
        if (kind != 0 && n_frames % wave) return -1;
        return run_mt(n_frames, n_threads, wave, [&](int f0, int n) {
            if (prec == 0)
            {
                const float* y = (const float*)Y + (size_t)f0 * N;
                int32_t* v = (int32_t*)V + (size_t)f0 * K;
                return kind == 0   ? ldpc_seq<int32_t, float>(K, n_ite, norm, n, y, cwd + f0, v)
                       : kind == 1 ? ldpc_inter<int32_t, float>(K, n_ite, norm, n, y, cwd + f0, v)
                                   : ldpc_tinter<int32_t, float>(K, n_ite, norm, n, y, cwd + f0, v);
            }
            const int16_t* y = (const int16_t*)Y + (size_t)f0 * N;
            int16_t* v = (int16_t*)V + (size_t)f0 * K;
            return kind == 0   ? ldpc_seq<int16_t, int16_t>(K, n_ite, norm, n, y, cwd + f0, v)
                   : kind == 1 ? ldpc_inter<int16_t, int16_t>(K, n_ite, norm, n, y, cwd + f0, v)
                               : ldpc_tinter<int16_t, int16_t>(K, n_ite, norm, n, y, cwd + f0, v);
        });
    }

    // Galois tables + generator polynomial (Galois.cpp:158-183, BCH_polynomial_generator.cpp:52-139)
    int ref_bch_tables(int* alpha_to, int* index_of, int* g, int* g_len)
    {
        tools::BCH_polynomial_generator<int> poly(16383, 12, kPrimPoly);
        for (int i = 0; i < 16384; i++)
        {
            alpha_to[i] = poly.get_alpha_to()[i];
            index_of[i] = poly.get_index_of()[i];
        }
        *g_len = (int)poly.get_g().size();
        for (size_t i = 0; i < poly.get_g().size(); i++)
            g[i] = poly.get_g()[i];
        return 0;
    }

    struct ref_bch
    {
        tools::BCH_polynomial_generator<int> poly;
        module::Decoder_BCH_DVBS2<int, float> dec;
        ref_bch(int K, int N) : poly(16383, 12, kPrimPoly), dec(K, N, poly) { dec.set_n_frames(1); }
    };
    void* ref_bch_new(int K, int N) { return new ref_bch(K, N); } // allocates the 1.07 GB elp table (Decoder_BCH_std.cpp:18)
    void ref_bch_free(void* h) { delete (ref_bch*)h; }
    int ref_bch_decode(void* h, int n_frames, const int32_t* Y, int8_t* cwd, int32_t* V)
    {
        auto* b = (ref_bch*)h;
        const int N = b->dec.get_N(), K = b->dec.get_K();
        for (int f = 0; f < n_frames; f++)
            b->dec.decode_hiho(Y + (size_t)f * N, cwd + f, V + (size_t)f * K);
        return 0;
    }

    // Decoder_BCH_DVBS2 over n_threads instances (each allocates its own 1.07 GB elp table)
    int ref_bch_decode_mt(int K, int N, int n_frames, int n_threads, const int32_t* Y, int8_t* cwd, int32_t* V)
    {
        return run_mt(n_frames, n_threads, 1, [&](int f0, int n) {
            ref_bch b(K, N);
            for (int f = f0; f < f0 + n; f++)
                b.dec.decode_hiho(Y + (size_t)f * N, cwd + f, V + (size_t)f * K);
            return 0;
        });
    }

    // TX side, used only to make valid codewords for tests (main.cpp:75-84 order: scramble -> BCH -> LDPC -> itl -> map)
    int ref_tx(int K_bch, int N_bch, int bps, const char* mod_path, const char* read_order, int n_frames,
               const int32_t* U, int32_t* X_scr, int32_t* X_bch, int32_t* X_ldpc, int32_t* X_itl, float* sym)
    {
        const int N = 16200;
        tools::BCH_polynomial_generator<int> poly(16383, 12, kPrimPoly);
        module::Scrambler_BB<int> scr(K_bch, 1);
        module::Encoder_BCH_DVBS2<int> ebch(K_bch, N_bch, poly);
        auto dv = tools::build_dvbs2(N_bch, N);
        module::Encoder_LDPC_DVBS2<int> eldpc(*dv);
        std::unique_ptr<tools::Interleaver_core<uint32_t>> core;
        if (bps == 2)
            core.reset(new tools::Interleaver_core_NO<uint32_t>(N));
        else
            core.reset(new tools::Interleaver_core_column_row<uint32_t>(N, bps, std::string(read_order)));
        module::Interleaver<int, uint32_t> itl(*core);
        tools::Constellation_user<float> cstl{ std::string(mod_path) };
        module::Modem_generic_fast<int, float, float, tools::max_star, tools::max_star_i> mdm(N, cstl, false);
        for (int f = 0; f < n_frames; f++)
        {
            scr.scramble(const_cast<int32_t*>(U) + (size_t)f * K_bch, X_scr + (size_t)f * K_bch);
            ebch.encode(X_scr + (size_t)f * K_bch, X_bch + (size_t)f * N_bch);
            eldpc.encode(X_bch + (size_t)f * N_bch, X_ldpc + (size_t)f * N);
            itl.interleave(X_ldpc + (size_t)f * N, X_itl + (size_t)f * N);
            mdm.modulate(X_itl + (size_t)f * N, sym + (size_t)f * (2 * N / bps));
        }
        return 0;
    }

    // RX front: Modem_generic_fast::demodulate (Modem.hxx:222-240) then Interleaver::deinterleave
    int ref_demodulate(int bps, const char* mod_path, const char* read_order, int n_frames, const float* sigma,
                       const float* sym, float* llr_itl, float* llr_nat, float* cstl_out)
    {
        const int N = 16200;
        tools::Constellation_user<float> cstl{ std::string(mod_path) };
        if (cstl_out)
            for (unsigned j = 0; j < cstl.get_n_symbols(); j++)
            {
                cstl_out[2 * j] = cstl[j].real();
                cstl_out[2 * j + 1] = cstl[j].imag();
            }
        module::Modem_generic_fast<int, float, float, tools::max_star, tools::max_star_i> mdm(N, cstl, false);
        std::unique_ptr<tools::Interleaver_core<uint32_t>> core;
        if (bps == 2)
            core.reset(new tools::Interleaver_core_NO<uint32_t>(N));
        else
            core.reset(new tools::Interleaver_core_column_row<uint32_t>(N, bps, std::string(read_order)));
        module::Interleaver<float, uint32_t> itl(*core);
        for (int f = 0; f < n_frames; f++)
        {
            mdm.demodulate(sigma + f, sym + (size_t)f * (2 * N / bps), llr_itl + (size_t)f * N);
            itl.deinterleave(llr_itl + (size_t)f * N, llr_nat + (size_t)f * N);
        }
        return 0;
    }

    int ref_descramble(int K_bch, int n_frames, const int32_t* in, int32_t* out)
    {
        module::Scrambler_BB<int> scr(K_bch, 1);
        for (int f = 0; f < n_frames; f++)
            scr.descramble(const_cast<int32_t*>(in) + (size_t)f * K_bch, out + (size_t)f * K_bch);
        return 0;
    }

    // Framer::remove_plh (Framer.hxx:319-343) and Estimator_DVBS2::estimate (Estimator.hxx:103-117, Estimator_DVBS2.hxx:31-58)
    int ref_remove_plh(const char* modcod, int n_sym, int pl_sym, int n_frames, const float* in, float* out)
    {
        module::Framer<float> fr(2 * n_sym, 2 * pl_sym, std::string(modcod), 1);
        for (int f = 0; f < n_frames; f++)
            fr.remove_plh(const_cast<float*>(in) + (size_t)f * 2 * pl_sym, out + (size_t)f * 2 * n_sym);
        return 0;
    }
    int ref_framer_generate(const char* modcod, int n_sym, int pl_sym, int n_frames, const float* in, float* out)
    {
        module::Framer<float> fr(2 * n_sym, 2 * pl_sym, std::string(modcod), 1);
        for (int f = 0; f < n_frames; f++)
            fr.generate(const_cast<float*>(in) + (size_t)f * 2 * n_sym, out + (size_t)f * 2 * pl_sym);
        return 0;
    }
    int ref_estimate(int n_sym, int K_bch, int bps, int n_frames, const float* X, float* sig, float* ebn0, float* esn0)
    {
        const float code_rate = (float)K_bch / (float)16200;
        module::Estimator_DVBS2<float> est(2 * n_sym, code_rate, bps, 1);
        for (int f = 0; f < n_frames; f++)
            est.estimate(X + (size_t)f * 2 * n_sym, sig + f, ebn0 + f, esn0 + f);
        return 0;
    }

    // Quantizer_pow2 (Quantizer_pow2.cpp:184-190)
    int ref_quantize_i16(int n, int fixed_point_pos, int saturation_pos, const float* in, int16_t* out)
    {
        module::Quantizer_pow2<float, int16_t> q(n, (short)fixed_point_pos, (short)saturation_pos);
        q.process(in, out);
        return 0;
    }
    int ref_quantize_i8(int n, int fixed_point_pos, int saturation_pos, const float* in, int8_t* out)
    {
        module::Quantizer_pow2<float, int8_t> q(n, (short)fixed_point_pos, (short)saturation_pos);
        q.process(in, out);
        return 0;
    }

    // CPU arm of bench.py: the chain the dvbs2 binaries run per thread with --dec-simd INTER -F 16
    // (demodulate -> deinterleave -> ONMS_inter decode_siho -> Decoder_BCH_DVBS2 decode_hiho), one clone per thread,
    // all threads looping over the same pre-generated frames for `seconds`.  stages bit0=demod bit1=ldpc bit2=bch.
    // Returns frames decoded per second (all threads).
    double ref_bench_chain(int K_bch, int N_bch, int bps, const char* mod_path, const char* read_order, int n_ite,
                           int n_threads, double seconds, int stages, int n_frames, const float* sigma,
                           const float* sym, const float* llr_in, double* frames_done)
    {
        const int N = 16200;
        const int W = (int)mipp::N<float>();
        if (n_frames % W) return -1.0;
        std::atomic<long> total{ 0 };
        std::atomic<int> ready{ 0 };
        std::atomic<bool> go{ false }, stop{ false };
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++)
            th.emplace_back([&, t]() {
                tools::Constellation_user<float> cstl{ std::string(mod_path) };
                module::Modem_generic_fast<int, float, float, tools::max_star, tools::max_star_i> mdm(N, cstl, false);
                std::unique_ptr<tools::Interleaver_core<uint32_t>> core;
                if (bps == 2)
                    core.reset(new tools::Interleaver_core_NO<uint32_t>(N));
                else
                    core.reset(new tools::Interleaver_core_column_row<uint32_t>(N, bps, std::string(read_order)));
                        module::Interleaver<float, uint32_t> itl(*core);
                auto dv = tools::build_dvbs2(N_bch, N);
                auto H = tools::build_H(*dv);
                auto ibp = info_pos(N_bch);
                module::Decoder_LDPC_BP_horizontal_layered_ONMS_inter<int, float> ldpc(N_bch, N, n_ite, H, ibp, 1.0f, 0.f,
                                                                                      true, 1);
                ldpc.set_n_frames(W);
                std::unique_ptr<ref_bch> bch;
                if (stages & 4) bch.reset(new ref_bch(K_bch, N_bch));
                std::vector<float> l1((size_t)W * N), l2((size_t)W * N);
                std::vector<int> v1((size_t)W * N_bch), v2((size_t)W * K_bch);
                std::vector<int8_t> c1(W), c2(W);
                ready++;
                while (!go.load()) std::this_thread::yield();
                long mine = 0;
                int f0 = (t * W) % n_frames;
                while (!stop.load())
                {
                    const float* llr = llr_in + (size_t)f0 * N;
                    if (stages & 1)
                    {
                        for (int f = 0; f < W; f++)
                        {
                            mdm.demodulate(sigma + f0 + f, sym + (size_t)(f0 + f) * (2 * N / bps), l1.data() + (size_t)f * N);
                            itl.deinterleave(l1.data() + (size_t)f * N, l2.data() + (size_t)f * N);
                        }
                        llr = l2.data();
                    }
                    if (stages & 2) ldpc.decode_siho(llr, c1.data(), v1.data());
                    if (stages & 4)
                        for (int f = 0; f < W; f++)
                            bch->dec.decode_hiho(v1.data() + (size_t)f * N_bch, c2.data() + f, v2.data() + (size_t)f * K_bch);
                    mine += W;
                    f0 = (f0 + W) % n_frames;
                }
                total += mine;
            });
        while (ready.load() < n_threads) std::this_thread::yield();
        auto t0 = std::chrono::steady_clock::now();
        go = true;
        std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
        stop = true;
        for (auto& x : th) x.join();
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (frames_done) *frames_done = (double)total.load();
        return (double)total.load() / dt;
    }
}
