/*!
 * \file
 * \brief GPU drop-in for the LDPC decoder the dvbs2 binaries build with --dec-implem {MS,NMS}:
 *        aff3ct::module::Decoder_LDPC_BP_horizontal_layered<B,R,Update_rule_NMS<R>> (sequential) and
 *        Decoder_LDPC_BP_horizontal_layered_ONMS_inter<B,R> (--dec-simd INTER)
 *        (src/common/Factory/DVBS2/DVBS2.cpp:418-449; lib/aff3ct/src/Factory/Module/Decoder/LDPC/Decoder_LDPC.cpp:298-331,692-717).
 *
 * Same Module/Task surface: task dec::tsk::decode_siho with sockets dec::sck::decode_siho::{Y_N, CWD, V_K}
 * (lib/aff3ct/include/Module/Decoder/Decoder_SIHO.hxx:57-79), clone(), set_n_frames(), replicable tasks.  Like the
 * inter-frame SIMD decoders (ONMS_inter.cpp:40) it processes a whole wave per codelet call; here the wave is the whole
 * -F batch (set_single_wave(true)).  R = float is what the dvbs2 binaries run; R = int16_t / int8_t are AFF3CT's fixed-point
 * decoders (quantised LLRs in, bit-exact against the sequential int16/int8 instantiations).
 */
#ifndef DVBS2_B200_DECODER_LDPC_DVBS2_CUDA_HPP_
#define DVBS2_B200_DECODER_LDPC_DVBS2_CUDA_HPP_

#include <cstdint>
#include <memory>
#include <string>
#include <type_traits>

#include <aff3ct.hpp>

#include "Cuda_handle.hpp"

namespace dvbs2_b200
{
template<typename B = int, typename R = float>
class Decoder_LDPC_DVBS2_cuda : public aff3ct::module::Decoder_SIHO<B, R>
{
    static_assert(std::is_same<B, int>::value, "bits are int32_t on every socket (AFF3CT_PREC=MULTI default B)");
    static_assert(std::is_same<R, float>::value || std::is_same<R, int16_t>::value || std::is_same<R, int8_t>::value,
                  "R is float, int16_t or int8_t");

  protected:
    std::shared_ptr<Cuda_handle> handle;
    const bool auto_device;

  public:
    /*!
     * \param modcod:          --mod-cod value; fixes K, N = 16200 and H (tools::build_dvbs2 + build_H).
     * \param n_ite:           --dec-ite.
     * \param implem:          --dec-implem, "MS" or "NMS" ("SPA" throws: CPU only).
     * \param norm_factor:     NMS normalisation; the dvbs2 factory leaves it at 1.0; integers accept k/8.
     * \param enable_syndrome: per-frame early termination (syndrome depth 1).
     * \param n_frames:        -F.
     * \param device:          CUDA device, -1 = round-robin per clone.
     */
    Decoder_LDPC_DVBS2_cuda(const std::string& modcod, const int n_ite, const std::string& implem = "NMS",
                            const float norm_factor = 1.f, const bool enable_syndrome = true, const int n_frames = 1,
                            const int device = -1, const int quant_bits = 6, const int quant_dec = 2)
      : aff3ct::module::Decoder_SIHO<B, R>(K_of(modcod), DVBS2_N_LDPC)
      , auto_device(device < 0)
    {
        const std::string name = "Decoder_LDPC_DVBS2_cuda";
        this->set_name(name);
        if (implem != "MS" && implem != "NMS")
            throw spu::tools::invalid_argument(__FILE__, __LINE__, __func__,
                                               "'implem' has to be MS or NMS on the GPU ('implem' = " + implem + ").");
        dvbs2_cuda_cfg cfg;
        dvbs2_cuda_default_cfg(&cfg);
        cfg.modcod = modcod_from_string(modcod);
        cfg.n_ite = n_ite;
        cfg.norm = implem == "MS" ? 1.f : norm_factor;
        cfg.precision = std::is_same<R, float>::value ? DVBS2_PREC_F32 : sizeof(R) == 2 ? DVBS2_PREC_I16 : DVBS2_PREC_I8;
        cfg.early_stop = enable_syndrome ? 1 : 0;
        cfg.max_frames = n_frames;
        cfg.device = device;
        cfg.quant_bits = quant_bits, cfg.quant_dec = quant_dec;
        this->handle = std::make_shared<Cuda_handle>(cfg);
        this->set_single_wave(true); // one codelet call sees the whole batch (lib/streampu/src/Module/Module.cpp:88-108)
        this->set_n_frames(n_frames);
        for (auto& t : this->tasks)
            t->set_replicability(true);
    }

    virtual ~Decoder_LDPC_DVBS2_cuda() = default;

    virtual Decoder_LDPC_DVBS2_cuda<B, R>* clone() const
    {
        auto m = new Decoder_LDPC_DVBS2_cuda(*this);
        m->deep_copy(*this);
        return m;
    }

    virtual void set_n_frames(const size_t n_frames)
    {
        this->handle->unpin_all(); // the sockets are about to be reallocated
        aff3ct::module::Decoder_SIHO<B, R>::set_n_frames(n_frames);
        if ((int)n_frames > this->handle->info().max_frames) // (re)allocate device buffers, as DVBS2.cpp:447 expects
            this->handle = std::make_shared<Cuda_handle>(this->handle->cfg_for((int)n_frames, this->handle->cfg().device));
    }

    const Cuda_handle& get_handle() const { return *this->handle; }

  protected:
    virtual void deep_copy(const Decoder_LDPC_DVBS2_cuda<B, R>& m)
    {
        aff3ct::module::Decoder_SIHO<B, R>::deep_copy(m);
        // a clone owns its stream and buffers, on the next device when devices are assigned automatically
        this->handle = std::make_shared<Cuda_handle>(
          m.handle->cfg_for(m.handle->info().max_frames, m.auto_device ? -1 : m.handle->cfg().device));
    }

    virtual int _decode_siho(const R* Y_N, int8_t* CWD, B* V_K, const size_t frame_id)
    {
        const int n = (int)this->get_n_frames_per_wave();
        this->handle->pin(Y_N, sizeof(R) * (size_t)n * DVBS2_N_LDPC);
        this->handle->pin(V_K, sizeof(B) * (size_t)n * this->K);
        DVBS2_B200_CHECK(this->handle, dvbs2_cuda_ldpc_decode_siho(this->handle->get(), Y_N, CWD, V_K, nullptr, n));
        int status = 0; // like the SIMD decoders: non-zero when some frame of the wave is not a codeword
        for (int f = 0; f < n; f++)
            status |= !CWD[f];
        (void)frame_id;
        return status;
    }

    static int K_of(const std::string& modcod)
    {
        const int mc = modcod_from_string(modcod);
        return (mc == DVBS2_QPSK_S_3_5 || mc == DVBS2_8PSK_S_3_5) ? 9720 : 14400; // DVBS2.cpp:336-349
    }
};
}

#endif /* DVBS2_B200_DECODER_LDPC_DVBS2_CUDA_HPP_ */
