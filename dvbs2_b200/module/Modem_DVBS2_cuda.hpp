/*!
 * \file
 * \brief GPU drop-in for the demodulate task of aff3ct::module::Modem_generic_fast<B,R,Q,max_star,max_star_i> as built by
 *        factory::DVBS2::build_modem (src/common/Factory/DVBS2/DVBS2.cpp:478-488): log-MAP soft demapper over the MODCOD's
 *        constellation, per-frame sigma on the CP socket.  Task mdm::tsk::demodulate, sockets mdm::sck::demodulate::{CP, Y_N1, Y_N2}
 *        (lib/aff3ct/include/Module/Modem/Modem.hxx:222-240).  The modulate task stays on the CPU modem (TX side).
 */
#ifndef DVBS2_B200_MODEM_DVBS2_CUDA_HPP_
#define DVBS2_B200_MODEM_DVBS2_CUDA_HPP_

#include <memory>
#include <string>
#include <type_traits>

#include <aff3ct.hpp>

#include "Cuda_handle.hpp"

namespace dvbs2_b200
{
template<typename B = int, typename R = float, typename Q = R>
class Modem_DVBS2_cuda : public aff3ct::module::Modem<B, R, Q>
{
    static_assert(std::is_same<R, float>::value && std::is_same<Q, float>::value, "LLRs and symbols are float");

  protected:
    std::shared_ptr<Cuda_handle> handle;
    const bool auto_device;

  public:
    Modem_DVBS2_cuda(const std::string& modcod, const int n_frames = 1, const int device = -1)
      : aff3ct::module::Modem<B, R, Q>(DVBS2_N_LDPC, 2 * (DVBS2_N_LDPC / bps_of(modcod)))
      , auto_device(device < 0)
    {
        const std::string name = "Modem_DVBS2_cuda";
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

    virtual ~Modem_DVBS2_cuda() = default;

    virtual Modem_DVBS2_cuda<B, R, Q>* clone() const
    {
        auto m = new Modem_DVBS2_cuda(*this);
        m->deep_copy(*this);
        return m;
    }

    virtual void set_n_frames(const size_t n_frames)
    {
        this->handle->unpin_all(); // the sockets are about to be reallocated
        aff3ct::module::Modem<B, R, Q>::set_n_frames(n_frames);
        if ((int)n_frames > this->handle->info().max_frames)
            this->handle = std::make_shared<Cuda_handle>(this->handle->cfg_for((int)n_frames, this->handle->cfg().device));
    }

  protected:
    virtual void deep_copy(const Modem_DVBS2_cuda<B, R, Q>& m)
    {
        aff3ct::module::Modem<B, R, Q>::deep_copy(m);
        this->handle = std::make_shared<Cuda_handle>(
          m.handle->cfg_for(m.handle->info().max_frames, m.auto_device ? -1 : m.handle->cfg().device));
    }

    virtual void _demodulate(const float* CP, const Q* Y_N1, Q* Y_N2, const size_t frame_id)
    {
        const int n = (int)this->get_n_frames_per_wave();
        this->handle->pin(Y_N1, sizeof(Q) * (size_t)n * this->N_fil);
        this->handle->pin(Y_N2, sizeof(Q) * (size_t)n * this->N);
        DVBS2_B200_CHECK(this->handle, dvbs2_cuda_demodulate(this->handle->get(), CP, Y_N1, Y_N2, n));
        (void)frame_id;
    }

    static int bps_of(const std::string& modcod)
    {
        const int mc = modcod_from_string(modcod);
        return mc <= DVBS2_QPSK_S_3_5 ? 2 : mc == DVBS2_16APSK_S_8_9 ? 4 : 3;
    }
};
}

#endif /* DVBS2_B200_MODEM_DVBS2_CUDA_HPP_ */
