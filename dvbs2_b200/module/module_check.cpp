// Drop-in check of the C++ GPU modules against the reference's own CPU modules, through the StreamPU task interface.
//
// Builds the receiver part of src/mains/TX_RX_BB/main.cpp:43-94 twice — once with the reference classes
// (Modem_generic_fast, Interleaver, Decoder_LDPC_BP_horizontal_layered / ONMS_inter, Decoder_BCH_DVBS2), once with
// dvbs2_b200::{Modem,Decoder_LDPC,Decoder_BCH,Decoder_chain}_DVBS2_cuda — feeds both the same frames (reference TX chain +
// AWGN) and compares socket contents; then runs the fused GPU module inside a 2-thread spu::runtime::Sequence (clone(),
// one handle per thread) and compares the error counters with the single-threaded CPU chain.
// Links: oracle/_ref/libdvbs2_ref.so (the compiled reference, for the AFF3CT/StreamPU symbols) + libdvbs2_cuda.so.
// Usage: module_check <conf/mod dir or ''> [modcod] [n_frames] [ebn0]   -> prints one JSON line, exit code 0 iff all checks pass.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include <aff3ct.hpp>
#include <streampu.hpp>

#include "Module/Decoder/LDPC/BP/Horizontal_layered/Decoder_LDPC_BP_horizontal_layered.hpp"
#include "Module/Decoder/LDPC/BP/Horizontal_layered/ONMS/Decoder_LDPC_BP_horizontal_layered_ONMS_inter.hpp"
#include "Tools/Code/LDPC/Standard/DVBS2/DVBS2_constants.hpp"
#include "Tools/Code/LDPC/Update_rule/NMS/Update_rule_NMS.hpp"

#include "Module/Decoder_BCH_DVBS2/Decoder_BCH_DVBS2.hpp"
#include "Module/Encoder_BCH_DVBS2/Encoder_BCH_DVBS2.hpp"
#include "Module/Scrambler/Scrambler_BB/Scrambler_BB.hpp"
#include "Module/Framer/Framer.hpp"
#include "Module/Estimator/Estimator_DVBS2.hpp"

#include "Decoder_BCH_DVBS2_cuda.hpp"
#include "Decoder_LDPC_DVBS2_cuda.hpp"
#include "Adaptor_DVBS2_cuda.hpp"
#include "Decoder_chain_DVBS2_cuda.hpp"
#include "Modem_DVBS2_cuda.hpp"
#include "Receiver_chain_DVBS2_cuda.hpp"

using namespace aff3ct;

namespace
{
const std::vector<int> kPrimPoly = { 1, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1 }; // DVBS2.hpp:55

struct Modcod
{
    std::string name, cstl, order;
    int bps, K_bch, N_bch;
};
Modcod
modcod_of(const std::string& n)
{
    if (n == "QPSK-S_8/9") return { n, "QPSK", "TOP_LEFT", 2, 14232, 14400 };
    if (n == "QPSK-S_3/5") return { n, "QPSK", "TOP_LEFT", 2, 9552, 9720 };
    if (n == "8PSK-S_3/5") return { n, "8PSK", "TOP_RIGHT", 3, 9552, 9720 };
    if (n == "8PSK-S_8/9") return { n, "8PSK", "TOP_LEFT", 3, 14232, 14400 };
    if (n == "16APSK-S_8/9") return { n, "16APSK", "TOP_LEFT", 4, 14232, 14400 };
    throw std::runtime_error("unknown modcod " + n);
}

std::string
write_mod_file(const std::string& cstl)
{ // conf/mod/*.mod contents (Constellation_user only reads files)
    const char* txt = cstl == "QPSK" ? "1 1\n-1 1\n1 -1\n-1 -1\n"
                      : cstl == "8PSK"
                        ? "0.707106781 0.707106781\n0 1\n-1 0\n-0.707106781 0.707106781\n1 0\n0.707106781 -0.707106781\n"
                          "-0.707106781 -0.707106781\n0 -1\n"
                        : "0.797081141330456 0.797081141330456\n0.291751946604441 1.08883308793490\n"
                          "1.08883308793490 0.291751946604441\n0.306569669742483 0.306569669742483\n"
                          "-0.797081141330456 0.797081141330456\n-0.291751946604441 1.08883308793490\n"
                          "-1.08883308793490 0.291751946604442\n-0.306569669742483 0.306569669742483\n"
                          "0.797081141330456 -0.797081141330456\n0.291751946604441 -1.08883308793490\n"
                          "1.08883308793490 -0.291751946604441\n0.306569669742483 -0.306569669742483\n"
                          "-0.797081141330456 -0.797081141330456\n-0.291751946604441 -1.08883308793490\n"
                          "-1.08883308793490 -0.291751946604441\n-0.306569669742483 -0.306569669742483\n";
    const std::string path = "/tmp/dvbs2_b200_" + cstl + ".mod";
    std::ofstream f(path);
    f << "# " << cstl << "\n" << txt;
    return path;
}

// frame pool handed out batch by batch to the clones of a Sequence
struct Pool
{
    int F, n_batches, n_sym2, K;
    std::vector<float> sigma, sym; // [n_batches * F] / [n_batches * F][n_sym2]
    std::vector<int> U;            // [n_batches * F][K]
    std::atomic<int> next{ 0 };
};

class Pool_source : public spu::module::Stateful
{
    std::shared_ptr<Pool> pool;

  public:
    explicit Pool_source(std::shared_ptr<Pool> p)
      : pool(p)
    {
        this->set_name("Pool_source");
        this->set_short_name("Pool_source");
        auto& t = this->create_task("generate");
        auto s_cp = this->template create_socket_out<float>(t, "CP", 1);
        auto s_y = this->template create_socket_out<float>(t, "Y_N1", pool->n_sym2);
        auto s_u = this->template create_socket_out<int32_t>(t, "U_K", pool->K);
        auto s_v = this->template create_socket_out<int32_t>(t, "VALID", 1);
        this->create_codelet(t,
                             [s_cp, s_y, s_u, s_v](spu::module::Module& m, spu::runtime::Task& tk, const size_t) -> int
                             {
                                 auto& src = static_cast<Pool_source&>(m);
                                 Pool& p = *src.pool;
                                 const int b = p.next++;
                                 const bool valid = b < p.n_batches;
                                 const size_t f0 = (size_t)(valid ? b : 0) * p.F;
                                 std::copy_n(p.sigma.data() + f0, p.F, static_cast<float*>(tk[s_cp].get_dataptr()));
                                 std::copy_n(p.sym.data() + f0 * p.n_sym2, (size_t)p.F * p.n_sym2,
                                             static_cast<float*>(tk[s_y].get_dataptr()));
                                 std::copy_n(p.U.data() + f0 * p.K, (size_t)p.F * p.K, static_cast<int32_t*>(tk[s_u].get_dataptr()));
                                 std::fill_n(static_cast<int32_t*>(tk[s_v].get_dataptr()), p.F, valid ? 1 : 0);
                                 return 0;
                             });
        this->set_single_wave(true);
        this->set_n_frames(pool->F);
        for (auto& tk : this->tasks)
            tk->set_replicability(true);
    }
    virtual Pool_source* clone() const
    {
        auto m = new Pool_source(*this);
        m->deep_copy(*this);
        return m;
    }
};

struct Counters
{
    std::atomic<long> frames{ 0 }, be{ 0 }, fe{ 0 }, accepted_wrong{ 0 };
};

class Bit_counter : public spu::module::Stateful
{
    std::shared_ptr<Counters> cnt;
    int K;

  public:
    Bit_counter(std::shared_ptr<Counters> c, const int K, const int F)
      : cnt(c)
      , K(K)
    {
        this->set_name("Bit_counter");
        this->set_short_name("Bit_counter");
        auto& t = this->create_task("check");
        auto s_u = this->template create_socket_in<int32_t>(t, "U", K);
        auto s_v = this->template create_socket_in<int32_t>(t, "V", K);
        auto s_c = this->template create_socket_in<int8_t>(t, "CWD", 1);
        auto s_ok = this->template create_socket_in<int32_t>(t, "VALID", 1);
        this->create_codelet(t,
                             [s_u, s_v, s_c, s_ok](spu::module::Module& m, spu::runtime::Task& tk, const size_t) -> int
                             {
                                 auto& bc = static_cast<Bit_counter&>(m);
                                 const int F = (int)bc.get_n_frames();
                                 const int32_t* U = static_cast<int32_t*>(tk[s_u].get_dataptr());
                                 const int32_t* V = static_cast<int32_t*>(tk[s_v].get_dataptr());
                                 const int8_t* C = static_cast<int8_t*>(tk[s_c].get_dataptr());
                                 const int32_t* ok = static_cast<int32_t*>(tk[s_ok].get_dataptr());
                                 for (int f = 0; f < F; f++)
                                 {
                                     if (!ok[f]) continue;
                                     long e = 0;
                                     for (int i = 0; i < bc.K; i++)
                                         e += U[(size_t)f * bc.K + i] != V[(size_t)f * bc.K + i];
                                     bc.cnt->frames++;
                                     bc.cnt->be += e;
                                     bc.cnt->fe += e != 0;
                                     bc.cnt->accepted_wrong += (e != 0 && C[f]);
                                 }
                                 return 0;
                             });
        this->set_single_wave(true);
        this->set_n_frames(F);
        for (auto& tk : this->tasks)
            tk->set_replicability(true);
    }
    virtual Bit_counter* clone() const
    {
        auto m = new Bit_counter(*this);
        m->deep_copy(*this);
        return m;
    }
};
} // namespace

int
main(int argc, char** argv)
{
    const std::string modcod = argc > 1 ? argv[1] : "QPSK-S_8/9";
    const int F = argc > 2 ? atoi(argv[2]) : 32; // -F, multiple of 16 for the AVX-512 inter decoder
    const float ebn0 = argc > 3 ? (float)atof(argv[3]) : 3.95f;
    const int n_batches = 6;
    const Modcod mc = modcod_of(modcod);
    const int N = 16200, n_sym2 = 2 * N / mc.bps, P = F * n_batches;
    int failures = 0;
    auto expect = [&](bool ok, const char* what) {
        if (!ok) failures++, fprintf(stderr, "CHECK FAILED: %s\n", what);
    };

    try
    {
        // ------------------------------------------------------------------ reference TX chain (main.cpp:75-84)
        tools::Constellation_user<float> cstl(write_mod_file(mc.cstl));
        module::Modem_generic_fast<int, float, float, tools::max_star, tools::max_star_i> mdm(N, cstl, false);
        std::unique_ptr<tools::Interleaver_core<uint32_t>> core;
        if (mc.bps == 2)
            core.reset(new tools::Interleaver_core_NO<uint32_t>(N));
        else
            core.reset(new tools::Interleaver_core_column_row<uint32_t>(N, mc.bps, mc.order));
        module::Interleaver<int32_t, uint32_t> itl_b(*core);
        module::Interleaver<float, uint32_t> itl_f(*core);
        tools::BCH_polynomial_generator<int> poly(16383, 12, kPrimPoly);
        module::Encoder_BCH_DVBS2<int> enc_bch(mc.K_bch, mc.N_bch, poly);
        auto dv = tools::build_dvbs2(mc.N_bch, N);
        auto H = tools::build_H(*dv);
        module::Encoder_LDPC_DVBS2<int> enc_ldpc(*dv);
        module::Scrambler_BB<int> scr(mc.K_bch);
        for (spu::module::Module* m : std::vector<spu::module::Module*>{ &mdm, &itl_b, &itl_f, &enc_bch, &enc_ldpc, &scr })
            m->set_n_frames(P);

        std::mt19937 rng(0xD5B52);
        auto pool = std::make_shared<Pool>();
        pool->F = F, pool->n_batches = n_batches, pool->n_sym2 = n_sym2, pool->K = mc.K_bch;
        pool->U.resize((size_t)P * mc.K_bch);
        for (auto& b : pool->U)
            b = rng() & 1;
        std::vector<int> s1((size_t)P * mc.K_bch), x1((size_t)P * mc.N_bch), x2((size_t)P * N), x3((size_t)P * N);
        pool->sym.resize((size_t)P * n_sym2);
        scr.scramble(pool->U, s1);
        enc_bch.encode(s1, x1);
        enc_ldpc.encode(x1, x2);
        itl_b.interleave(x2, x3);
        mdm.modulate(x3, pool->sym);
        const float rate = (float)mc.K_bch / N;
        const float esn0 = ebn0 + 10.f * std::log10(rate * mc.bps);
        const float sigma = std::sqrt(1.f / (2.f * std::pow(10.f, esn0 / 10.f)));
        std::normal_distribution<float> gauss(0.f, sigma);
        for (auto& v : pool->sym)
            v += gauss(rng);
        pool->sigma.assign(P, sigma);

        // ------------------------------------------------------------------ CPU receiver, one batch at a time (-F frames)
        std::vector<unsigned> ibp(mc.N_bch);
        for (int i = 0; i < mc.N_bch; i++)
            ibp[i] = i;
        tools::Update_rule_NMS<float> rule(1.f);
        module::Decoder_LDPC_BP_horizontal_layered<int, float, tools::Update_rule_NMS<float>> ldpc_seq(mc.N_bch, N, 10, H, ibp,
                                                                                                    rule, true, 1);
        module::Decoder_LDPC_BP_horizontal_layered_ONMS_inter<int, float> ldpc_inter(mc.N_bch, N, 10, H, ibp, 1.f, 0.f, true, 1);
        module::Decoder_BCH_DVBS2<int, float> bch(mc.K_bch, mc.N_bch, poly);
        for (spu::module::Module* m : std::vector<spu::module::Module*>{ &ldpc_seq, &ldpc_inter, &bch })
            m->set_n_frames(P);
        std::vector<float> llr_i((size_t)P * N), llr_n((size_t)P * N);
        std::vector<int> v_seq((size_t)P * mc.N_bch), v_int((size_t)P * mc.N_bch), v_bch((size_t)P * mc.K_bch), desc((size_t)P * mc.K_bch);
        std::vector<int8_t> c_seq(P), c_int(P), c_bch(P);
        mdm.demodulate(pool->sigma, pool->sym, llr_i);
        itl_f.deinterleave(llr_i, llr_n);
        ldpc_seq.decode_siho(llr_n.data(), c_seq.data(), v_seq.data());
        ldpc_inter.decode_siho(llr_n.data(), c_int.data(), v_int.data());
        bch.decode_hiho(v_seq.data(), c_bch.data(), v_bch.data());
        scr.descramble(v_bch, desc);
        long cpu_be = 0, cpu_fe = 0;
        for (int f = 0; f < P; f++)
        {
            long e = 0;
            for (int i = 0; i < mc.K_bch; i++)
                e += desc[(size_t)f * mc.K_bch + i] != pool->U[(size_t)f * mc.K_bch + i];
            cpu_be += e, cpu_fe += e != 0;
        }

        // ------------------------------------------------------------------ GPU modules, same sockets, same inputs
        dvbs2_b200::Modem_DVBS2_cuda<> g_mdm(modcod, P);
        dvbs2_b200::Decoder_LDPC_DVBS2_cuda<> g_ldpc(modcod, 10, "NMS", 1.f, true, P);
        dvbs2_b200::Decoder_BCH_DVBS2_cuda<> g_bch(modcod, P);
        std::vector<float> g_llr((size_t)P * N);
        std::vector<int> gv_ldpc((size_t)P * mc.N_bch), gv_bch((size_t)P * mc.K_bch);
        std::vector<int8_t> gc_ldpc(P), gc_bch(P);
        g_mdm.demodulate(pool->sigma, pool->sym, g_llr);
        double max_rel = 0;
        for (size_t i = 0; i < g_llr.size(); i++)
            max_rel = std::max(max_rel, (double)std::fabs(g_llr[i] - llr_i[i]) / std::max(1.0, (double)std::fabs(llr_i[i])));
        expect(max_rel <= 1e-4, "Modem_DVBS2_cuda::demodulate within 1e-4 of Modem_generic_fast");
        // decoders are judged on identical LLRs: the reference chain once more, fed with the GPU demapper's LLRs
        long ref_be = 0, ref_fe = 0;
        {
            std::vector<float> n2((size_t)P * N);
            std::vector<int> v1((size_t)P * mc.N_bch), v2((size_t)P * mc.K_bch), d2((size_t)P * mc.K_bch);
            std::vector<int8_t> c1(P), c2(P);
            itl_f.deinterleave(g_llr, n2);
            ldpc_seq.decode_siho(n2.data(), c1.data(), v1.data());
            bch.decode_hiho(v1.data(), c2.data(), v2.data());
            scr.descramble(v2, d2);
            for (int f = 0; f < P; f++)
            {
                long e = 0;
                for (int i = 0; i < mc.K_bch; i++)
                    e += d2[(size_t)f * mc.K_bch + i] != pool->U[(size_t)f * mc.K_bch + i];
                ref_be += e, ref_fe += e != 0;
            }
        }
        g_ldpc.decode_siho(llr_n.data(), gc_ldpc.data(), gv_ldpc.data());
        expect(gc_ldpc == c_seq && gv_ldpc == v_seq, "Decoder_LDPC_DVBS2_cuda == sequential reference decoder (bits + CWD)");
        bool inter_ok = true;
        int n_conv = 0;
        for (int f = 0; f < P; f++)
            if (c_int[f])
            {
                n_conv++;
                inter_ok &= std::equal(v_int.begin() + (size_t)f * mc.N_bch, v_int.begin() + (size_t)(f + 1) * mc.N_bch,
                                       gv_ldpc.begin() + (size_t)f * mc.N_bch);
            }
        expect(inter_ok && n_conv > 0, "Decoder_LDPC_DVBS2_cuda == ONMS_inter on every frame that converges");
        g_bch.decode_hiho(v_seq.data(), gc_bch.data(), gv_bch.data());
        expect(gc_bch == c_bch && gv_bch == v_bch, "Decoder_BCH_DVBS2_cuda == Decoder_BCH_DVBS2 (bits + CWD)");
        // set_n_frames after construction (DVBS2.cpp:447) and a clone on its own handle
        g_ldpc.set_n_frames(F);
        std::unique_ptr<dvbs2_b200::Decoder_LDPC_DVBS2_cuda<>> g_clone(g_ldpc.clone());
        std::vector<float> y1(llr_n.begin(), llr_n.begin() + (size_t)F * N);
        std::vector<int> vk((size_t)F * mc.N_bch);
        std::vector<int8_t> ck(F);
        g_clone->decode_siho(y1.data(), ck.data(), vk.data());
        expect(std::equal(vk.begin(), vk.end(), v_seq.begin()) && std::equal(ck.begin(), ck.end(), c_seq.begin()),
               "clone() + set_n_frames(): same results on a fresh handle");
        expect(g_clone->get_handle().get() != g_ldpc.get_handle().get(), "clone owns its own handle");
        bool threw = false;
        try
        {
            dvbs2_b200::Decoder_LDPC_DVBS2_cuda<> bad(modcod, 10, "SPA");
        }
        catch (const spu::tools::invalid_argument&)
        {
            threw = true;
        }
        expect(threw, "--dec-implem SPA is rejected with spu::tools::invalid_argument");

        // ------------------------------------------------------------------ fused module inside a 2-thread Sequence
        auto cnt = std::make_shared<Counters>();
        Pool_source src(pool);
        dvbs2_b200::Decoder_chain_DVBS2_cuda chain(modcod, 10, "NMS", 1.f, true, F);
        module::Scrambler_BB<int> dscr(mc.K_bch);
        dscr.set_n_frames(F);
        Bit_counter counter(cnt, mc.K_bch, F);
        using namespace dvbs2_b200;
        chain[chn::sck::demodulate_decode::CP] = src["generate::CP"];
        chain[chn::sck::demodulate_decode::Y_N1] = src["generate::Y_N1"];
        dscr[module::scr::sck::descramble::Y_N1] = chain[chn::sck::demodulate_decode::V_K];
        counter["check::U"] = src["generate::U_K"];
        counter["check::V"] = dscr[module::scr::sck::descramble::Y_N2];
        counter["check::CWD"] = chain[chn::sck::demodulate_decode::CWD_bch];
        counter["check::VALID"] = src["generate::VALID"];
        spu::runtime::Sequence seq(src("generate"), 2);
        seq.exec([&]() { return pool->next.load() >= n_batches + 2; });
        expect(cnt->frames.load() == P, "Sequence over 2 clones processed every pool frame exactly once");
        // in-process multi-GPU: the clones StreamPU made (Sequence.cpp:1506-1535) sit on consecutive devices (Cuda_handle
        // hands devices out round-robin when the module was built with device -1)
        int seq_devices = 0; // bit d set: a clone of the fused stage runs on device d
        for (auto* m : seq.get_cloned_modules<dvbs2_b200::Decoder_chain_DVBS2_cuda>(chain))
            seq_devices |= 1 << m->get_handle().cfg().device;
        if (dvbs2_cuda_device_count() >= 2)
            expect(__builtin_popcount(seq_devices) == 2, "the two clones of the Sequence run on two different GPUs");
        expect(cnt->be.load() == ref_be && cnt->fe.load() == ref_fe,
               "fused GPU module in a 2-thread Sequence: BE/FE equal the reference decoders' on the same LLRs");
        expect(cnt->fe.load() == cpu_fe || std::labs(cnt->fe.load() - cpu_fe) <= 1 + cpu_fe / 10,
               "fused GPU module: FE agrees with the all-CPU chain (its own demapper) up to LLR rounding");
        expect(cnt->accepted_wrong.load() == 0 || cpu_fe > 0, "BCH-accepted frames carry the payload");

        // ------------------------------------------------------------------ compact sockets: IQ_packer<sc16> ->
        //        Decoder_chain_compact_DVBS2_cuda<int16_t> -> Bit_unpacker in a 2-thread Sequence; and, stand-alone, the compact
        //        stage against the float-socket stage on the same sample values (must be bit-identical)
        long cq_frames = 0, cq_be = 0, cq_fe = 0;
        size_t n_pinned = 0;
        {
            using namespace dvbs2_b200;
            const float scale = 1.f / 4096.f;
            pool->next = 0;
            auto cnt2 = std::make_shared<Counters>();
            Pool_source src2(pool);
            IQ_packer<int16_t> packer(n_sym2, scale, F);
            Decoder_chain_compact_DVBS2_cuda<int16_t> cchain(modcod, 10, "NMS", 1.f, true, F, -1, scale, true);
            Bit_unpacker unpacker(mc.K_bch, F);
            Bit_counter counter2(cnt2, mc.K_bch, F);
            packer["pack::Y_N1"] = src2["generate::Y_N1"];
            cchain[chc::sck::demodulate_decode::CP] = src2["generate::CP"];
            cchain[chc::sck::demodulate_decode::Y_N1] = packer["pack::Y_N2"];
            unpacker["unpack::X"] = cchain[chc::sck::demodulate_decode::V_K];
            counter2["check::U"] = src2["generate::U_K"];
            counter2["check::V"] = unpacker["unpack::Y"];
            counter2["check::CWD"] = cchain[chc::sck::demodulate_decode::CWD_bch];
            counter2["check::VALID"] = src2["generate::VALID"];
            spu::runtime::Sequence seq2(src2("generate"), 2);
            seq2.exec([&]() { return pool->next.load() >= n_batches + 2; });
            cq_frames = cnt2->frames.load(), cq_be = cnt2->be.load(), cq_fe = cnt2->fe.load();
            expect(cq_frames == P, "compact stage in a 2-thread Sequence processed every pool frame exactly once");
            expect(cq_fe == ref_fe || std::labs(cq_fe - ref_fe) <= 1 + ref_fe / 10,
                   "compact stage (sc16 symbols, packed payload): FE agrees with the float-socket stage up to the 16-bit samples");
            // stand-alone, whole pool at once: compact stage == float stage on the very same sample values
            std::vector<int16_t> q((size_t)P * n_sym2);
            std::vector<float> deq((size_t)P * n_sym2);
            for (size_t i = 0; i < q.size(); i++)
            {
                q[i] = iq_pack_one<int16_t>(pool->sym[i], 1.f / scale);
                deq[i] = (float)q[i] * scale;
            }
            Decoder_chain_DVBS2_cuda fchain(modcod, 10, "NMS", 1.f, true, P);
            Decoder_chain_compact_DVBS2_cuda<int16_t> c2(modcod, 10, "NMS", 1.f, true, P, -1, scale, false);
            fchain[chn::sck::demodulate_decode::CP].bind(pool->sigma);
            fchain[chn::sck::demodulate_decode::Y_N1].bind(deq);
            fchain[chn::tsk::demodulate_decode].exec();
            c2[chc::sck::demodulate_decode::CP].bind(pool->sigma);
            c2[chc::sck::demodulate_decode::Y_N1].bind(q);
            c2[chc::tsk::demodulate_decode].exec();
            const int32_t* vf = static_cast<int32_t*>(fchain[chn::sck::demodulate_decode::V_K].get_dataptr());
            const uint8_t* vc = static_cast<uint8_t*>(c2[chc::sck::demodulate_decode::V_K].get_dataptr());
            bool same = true;
            for (size_t f = 0; f < (size_t)P && same; f++)
                for (int i = 0; i < mc.K_bch && same; i++)
                    same = vf[f * mc.K_bch + i] == ((vc[f * (mc.K_bch / 8) + (i >> 3)] >> (i & 7)) & 1);
            expect(same, "compact stage == float-socket stage bit for bit on the same sample values");
            expect(std::memcmp(fchain[chn::sck::demodulate_decode::CWD_bch].get_dataptr(),
                               c2[chc::sck::demodulate_decode::CWD_bch].get_dataptr(), P) == 0,
                   "compact stage: same CWD as the float-socket stage");
            n_pinned = fchain.get_handle().n_pinned() + c2.get_handle().n_pinned();
            expect(n_pinned >= 2, "the modules page-locked their socket buffers on first use (cudaHostRegister)");
        }

        // ------------------------------------------------------------------ receiver stage from PL frames (main.cpp:86-92)
        long rx_be = 0, rx_fe = 0, rxc_be = 0, rxc_fe = 0;
        double sig_rel = 0;
        {
            const int n_sym = N / mc.bps, pl_sym = 90 * (n_sym / 90 + 1) + 36 * (n_sym / 1440);
            module::Framer<float> framer(2 * n_sym, 2 * pl_sym, modcod, P);
            module::Estimator_DVBS2<float> est(2 * n_sym, (float)mc.K_bch / (float)N, mc.bps, P);
            std::vector<float> plf((size_t)P * 2 * pl_sym), xfec((size_t)P * 2 * n_sym), sg(P), eb(P), es(P), li((size_t)P * N),
              ln((size_t)P * N);
            std::vector<int> v1((size_t)P * mc.N_bch), v2((size_t)P * mc.K_bch), d2((size_t)P * mc.K_bch);
            std::vector<int8_t> c1(P), c2(P);
            framer.generate(pool->sym, plf);
            // all-CPU: remove_plh -> estimate -> demodulate -> deinterleave -> LDPC -> BCH -> descramble
            framer.remove_plh(plf, xfec);
            est.estimate(xfec, sg, eb, es);
            mdm.demodulate(sg, xfec, li);
            itl_f.deinterleave(li, ln);
            ldpc_seq.decode_siho(ln.data(), c1.data(), v1.data());
            bch.decode_hiho(v1.data(), c2.data(), v2.data());
            scr.descramble(v2, d2);
            expect(xfec == pool->sym, "reference Framer: remove_plh(generate(x)) == x");
            dvbs2_b200::Receiver_chain_DVBS2_cuda g_rx(modcod, 10, "NMS", 1.f, true, P);
            std::vector<float> g_sig(P);
            std::vector<int> g_v((size_t)P * mc.K_bch);
            std::vector<int8_t> g_c1(P), g_c2(P);
            using namespace dvbs2_b200;
            g_rx[rcv::sck::receive::Y_N1].bind(plf);
            g_rx[rcv::tsk::receive].exec();
            std::copy_n(static_cast<float*>(g_rx[rcv::sck::receive::SIG].get_dataptr()), P, g_sig.begin());
            std::copy_n(static_cast<int32_t*>(g_rx[rcv::sck::receive::V_K].get_dataptr()), (size_t)P * mc.K_bch, g_v.begin());
            std::copy_n(static_cast<int8_t*>(g_rx[rcv::sck::receive::CWD_bch].get_dataptr()), P, g_c2.begin());
            for (int f = 0; f < P; f++)
            {
                sig_rel = std::max(sig_rel, (double)std::fabs(g_sig[f] - sg[f]) / sg[f]);
                long e = 0, ec = 0;
                for (int i = 0; i < mc.K_bch; i++)
                {
                    e += g_v[(size_t)f * mc.K_bch + i] != pool->U[(size_t)f * mc.K_bch + i];
                    ec += d2[(size_t)f * mc.K_bch + i] != pool->U[(size_t)f * mc.K_bch + i];
                }
                rx_be += e, rx_fe += e != 0, rxc_be += ec, rxc_fe += ec != 0;
                if (g_c2[f] && c2[f]) expect(e == 0 && ec == 0, "receiver stage: frames both BCH decoders accept carry the payload");
            }
            expect(sig_rel <= 1e-5, "Receiver_chain_DVBS2_cuda: sigma within 1e-5 of Estimator_DVBS2");
            expect(rx_fe == rxc_fe || std::labs(rx_fe - rxc_fe) <= 1 + rxc_fe / 10,
                   "Receiver_chain_DVBS2_cuda: FE agrees with the all-CPU receiver stage up to LLR rounding");
        }

        printf("{\"modcod\": \"%s\", \"frames\": %d, \"demod_max_rel_err\": %.3g, \"ldpc_fe_cpu\": %d, \"bch_fe_cpu\": %ld, "
               "\"seq_frames\": %ld, \"seq_be\": %ld, \"seq_fe\": %ld, \"ref_be\": %ld, \"ref_fe\": %ld, \"cpu_be\": %ld, \"cpu_fe\": %ld, "
               "\"rx_stage_sigma_rel\": %.3g, \"rx_stage_fe\": %ld, \"rx_stage_fe_cpu\": %ld, \"compact_frames\": %ld, \"compact_fe\": %ld, "
               "\"pinned_sockets\": %zu, \"gpus\": %d, \"seq_device_mask\": %d, \"failures\": %d}\n",
               modcod.c_str(), P, max_rel, (int)std::count(c_seq.begin(), c_seq.end(), 0), cpu_fe, cnt->frames.load(),
               cnt->be.load(), cnt->fe.load(), ref_be, ref_fe, cpu_be, cpu_fe, sig_rel, rx_fe, rxc_fe, cq_frames, cq_fe, n_pinned,
               dvbs2_cuda_device_count(), seq_devices, failures);
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
    return failures ? 1 : 0;
}
