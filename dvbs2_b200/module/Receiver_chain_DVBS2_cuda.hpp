/*!
 * \file
 * \brief The receiver stage from PL frames as ONE StreamPU module: what src/mains/TX_RX_BB/main.cpp:86-92
 *        binds as framer.remove_plh -> estimator.estimate -> modem.demodulate -> itl.deinterleave -> LDPC.decode_siho ->
 *        BCH.decode_hiho -> bb_scrambler.descramble (pipeline stages 5-7 of src/mains/RX/main_sched.cpp:204-211) runs on the
 *        device back to back: only PL frames go up and payload bits come down.
 *        Task "receive": in Y_N1[pl_frame_size]; out SIG[1], CWD_ldpc[1], CWD_bch[1], V_K[K_bch] (descrambled payload).
 */
#ifndef DVBS2_B200_RECEIVER_CHAIN_DVBS2_CUDA_HPP_
#define DVBS2_B200_RECEIVER_CHAIN_DVBS2_CUDA_HPP_

#include <cstdint>
#include <memory>
#include <string>

#include <streampu.hpp>

#include "Cuda_handle.hpp"

namespace dvbs2_b200
{
namespace rcv
{
enum class tsk : size_t
{
    receive,
    SIZE
};
namespace sck
{
enum class receive : size_t
{
    Y_N1,
    SIG,
    CWD_ldpc,
    CWD_bch,
    V_K,
    status
};
}
}

class Receiver_chain_DVBS2_cuda : public spu::module::Stateful
{
  protected:
    std::shared_ptr<Cuda_handle> handle;
    const bool auto_device;

  public:
    inline spu::runtime::Task& operator[](const rcv::tsk t) { return spu::module::Module::operator[]((size_t)t); }
    inline spu::runtime::Socket& operator[](const rcv::sck::receive s)
    {
        return spu::module::Module::operator[]((size_t)rcv::tsk::receive)[(size_t)s];
    }

    Receiver_chain_DVBS2_cuda(const std::string& modcod, const int n_ite, const std::string& implem = "NMS",
                              const float norm_factor = 1.f, const bool enable_syndrome = true, const int n_frames = 1,
                              const int device = -1)
      : spu::module::Stateful()
      , auto_device(device < 0)
    {
        const std::string name = "Receiver_chain_DVBS2_cuda";
        this->set_name(name);
        this->set_short_name(name);
        if (implem != "MS" && implem != "NMS")
            throw spu::tools::invalid_argument(__FILE__, __LINE__, __func__,
                                               "'implem' has to be MS or NMS on the GPU ('implem' = " + implem + ").");
        dvbs2_cuda_cfg cfg;
        dvbs2_cuda_default_cfg(&cfg);
        cfg.modcod = modcod_from_string(modcod);
        cfg.n_ite = n_ite;
        cfg.norm = implem == "MS" ? 1.f : norm_factor;
        cfg.early_stop = enable_syndrome ? 1 : 0;
        cfg.max_frames = n_frames;
        cfg.device = device;
        this->handle = std::make_shared<Cuda_handle>(cfg);
        const auto& info = this->handle->info();

        auto& p = this->create_task("receive");
        auto ps_Y_N1 = this->template create_socket_in<float>(p, "Y_N1", dvbs2_cuda_pl_frame_size(this->handle->get()));
        auto ps_SIG = this->template create_socket_out<float>(p, "SIG", 1);
        auto ps_CWD1 = this->template create_socket_out<int8_t>(p, "CWD_ldpc", 1);
        auto ps_CWD2 = this->template create_socket_out<int8_t>(p, "CWD_bch", 1);
        auto ps_V_K = this->template create_socket_out<int32_t>(p, "V_K", info.K_bch);
        this->create_codelet(
          p,
          [ps_Y_N1, ps_SIG, ps_CWD1, ps_CWD2, ps_V_K](spu::module::Module& m, spu::runtime::Task& t, const size_t frame_id) -> int
          {
              auto& c = static_cast<Receiver_chain_DVBS2_cuda&>(m);
              const int n = (int)c.get_n_frames_per_wave();
              int8_t* cwd = static_cast<int8_t*>(t[ps_CWD2].get_dataptr());
              c.handle->pin(t[ps_Y_N1].get_dataptr(), sizeof(float) * (size_t)n * dvbs2_cuda_pl_frame_size(c.handle->get()));
              c.handle->pin(t[ps_V_K].get_dataptr(), sizeof(int32_t) * (size_t)n * c.handle->info().K_bch);
              DVBS2_B200_CHECK(c.handle,
                               dvbs2_cuda_rx_chain(c.handle->get(), static_cast<float*>(t[ps_Y_N1].get_dataptr()), nullptr,
                                                   static_cast<float*>(t[ps_SIG].get_dataptr()),
                                                   static_cast<int8_t*>(t[ps_CWD1].get_dataptr()), cwd,
                                                   static_cast<int32_t*>(t[ps_V_K].get_dataptr()), nullptr, 1, nullptr, n));
              int status = 0;
              for (int f = 0; f < n; f++)
                  status |= !cwd[f];
              (void)frame_id;
              return status;
          });
        this->set_single_wave(true);
        this->set_n_frames(n_frames);
        for (auto& t : this->tasks)
            t->set_replicability(true);
    }

    virtual ~Receiver_chain_DVBS2_cuda() = default;

    virtual Receiver_chain_DVBS2_cuda* clone() const
    {
        auto m = new Receiver_chain_DVBS2_cuda(*this);
        m->deep_copy(*this);
        return m;
    }

    virtual void set_n_frames(const size_t n_frames)
    {
        this->handle->unpin_all(); // the sockets are about to be reallocated
        spu::module::Stateful::set_n_frames(n_frames);
        if ((int)n_frames > this->handle->info().max_frames)
            this->handle = std::make_shared<Cuda_handle>(this->handle->cfg_for((int)n_frames, this->handle->cfg().device));
    }

    const Cuda_handle& get_handle() const { return *this->handle; }

  protected:
    virtual void deep_copy(const Receiver_chain_DVBS2_cuda& m)
    {
        spu::module::Stateful::deep_copy(m);
        this->handle = std::make_shared<Cuda_handle>(
          m.handle->cfg_for(m.handle->info().max_frames, m.auto_device ? -1 : m.handle->cfg().device));
    }
};
}

#endif /* DVBS2_B200_RECEIVER_CHAIN_DVBS2_CUDA_HPP_ */
