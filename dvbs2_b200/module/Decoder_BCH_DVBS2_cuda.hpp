/*!
 * \file
 * \brief GPU drop-in for aff3ct::module::Decoder_BCH_DVBS2<B,R> (src/common/Module/Decoder_BCH_DVBS2/Decoder_BCH_DVBS2.cpp:28-40
 *        around lib/aff3ct/src/Module/Decoder/BCH/Standard/Decoder_BCH_std.cpp:52-221), built by factory::DVBS2::build_bch_decoder
 *        (src/common/Factory/DVBS2/DVBS2.cpp:406-416).  Task dec::tsk::decode_hiho, sockets dec::sck::decode_hiho::{Y_N, CWD, V_K}
 *        (lib/aff3ct/include/Module/Decoder/Decoder_HIHO.hxx:42-63).  No 1.07 GB scratch table per clone.
 */
#ifndef DVBS2_B200_DECODER_BCH_DVBS2_CUDA_HPP_
#define DVBS2_B200_DECODER_BCH_DVBS2_CUDA_HPP_

#include <cstdint>
#include <memory>
#include <string>
#include <type_traits>

#include <aff3ct.hpp>

#include "Cuda_handle.hpp"

namespace dvbs2_b200
{
template<typename B = int>
class Decoder_BCH_DVBS2_cuda : public aff3ct::module::Decoder_HIHO<B>
{
    static_assert(std::is_same<B, int>::value, "bits are int32_t on every socket");

  protected:
    std::shared_ptr<Cuda_handle> handle;
    const bool auto_device;

  public:
    Decoder_BCH_DVBS2_cuda(const std::string& modcod, const int n_frames = 1, const int device = -1)
      : aff3ct::module::Decoder_HIHO<B>(K_of(modcod), K_of(modcod) + 168)
      , auto_device(device < 0)
    {
        const std::string name = "Decoder_BCH_DVBS2_cuda";
        this->set_name(name);
        dvbs2_cuda_cfg cfg;
        dvbs2_cuda_default_cfg(&cfg);
        cfg.modcod = modcod_from_string(modcod);
        cfg.max_frames = n_frames;
        cfg.device = device;
        this->handle = std::make_shared<Cuda_handle>(cfg);
        this->set_single_wave(true);
        this->set_n_frames(n_frames);
        for (auto& t : this->tasks)
            t->set_replicability(true);
    }

    virtual ~Decoder_BCH_DVBS2_cuda() = default;

    virtual Decoder_BCH_DVBS2_cuda<B>* clone() const
    {
        auto m = new Decoder_BCH_DVBS2_cuda(*this);
        m->deep_copy(*this);
        return m;
    }

    virtual void set_n_frames(const size_t n_frames)
    {
        this->handle->unpin_all(); // the sockets are about to be reallocated
        aff3ct::module::Decoder_HIHO<B>::set_n_frames(n_frames);
        if ((int)n_frames > this->handle->info().max_frames)
            this->handle = std::make_shared<Cuda_handle>(this->handle->cfg_for((int)n_frames, this->handle->cfg().device));
    }

  protected:
    virtual void deep_copy(const Decoder_BCH_DVBS2_cuda<B>& m)
    {
        aff3ct::module::Decoder_HIHO<B>::deep_copy(m);
        this->handle = std::make_shared<Cuda_handle>(
          m.handle->cfg_for(m.handle->info().max_frames, m.auto_device ? -1 : m.handle->cfg().device));
    }

    virtual int _decode_hiho(const B* Y_N, int8_t* CWD, B* V_K, const size_t frame_id)
    {
        const int n = (int)this->get_n_frames_per_wave();
        this->handle->pin(Y_N, sizeof(B) * (size_t)n * this->N);
        this->handle->pin(V_K, sizeof(B) * (size_t)n * this->K);
        DVBS2_B200_CHECK(this->handle, dvbs2_cuda_bch_decode_hiho(this->handle->get(), Y_N, CWD, V_K, n));
        int status = 0;
        for (int f = 0; f < n; f++)
            status |= !CWD[f];
        (void)frame_id;
        return status;
    }

    static int K_of(const std::string& modcod)
    {
        const int mc = modcod_from_string(modcod);
        return (mc == DVBS2_QPSK_S_3_5 || mc == DVBS2_8PSK_S_3_5) ? 9552 : 14232; // DVBS2.cpp:336-349
    }
};
}

#endif /* DVBS2_B200_DECODER_BCH_DVBS2_CUDA_HPP_ */
