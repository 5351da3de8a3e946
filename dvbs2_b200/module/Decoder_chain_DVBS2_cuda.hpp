/*!
 * \file
 * \brief The fused receiver stage as ONE StreamPU module: what src/mains/TX_RX_BB/main.cpp:87-91 binds as
 *        modem.demodulate -> itl.deinterleave -> LDPC.decode_siho -> BCH.decode_hiho runs on the device back to back, so LLRs
 *        and LDPC bits never cross PCIe.  Task "demodulate_decode": in CP[1], Y_N1[2*n_sym]; out CWD_ldpc[1], CWD_bch[1], V_K[K_bch].
 *        Bind it in place of lines 87-91.
 */
#ifndef DVBS2_B200_DECODER_CHAIN_DVBS2_CUDA_HPP_
#define DVBS2_B200_DECODER_CHAIN_DVBS2_CUDA_HPP_

#include <cstdint>
#include <memory>
#include <string>

#include <streampu.hpp>

#include "Cuda_handle.hpp"

namespace dvbs2_b200
{
namespace chn
{
enum class tsk : size_t
{
    demodulate_decode,
    SIZE
};
namespace sck
{
enum class demodulate_decode : size_t
{
    CP,
    Y_N1,
    CWD_ldpc,
    CWD_bch,
    V_K,
    status
};
}
}

class Decoder_chain_DVBS2_cuda : public spu::module::Stateful
{
  protected:
    std::shared_ptr<Cuda_handle> handle;
    const bool auto_device;

  public:
    inline spu::runtime::Task& operator[](const chn::tsk t) { return spu::module::Module::operator[]((size_t)t); }
    inline spu::runtime::Socket& operator[](const chn::sck::demodulate_decode s)
    {
        return spu::module::Module::operator[]((size_t)chn::tsk::demodulate_decode)[(size_t)s];
    }

    Decoder_chain_DVBS2_cuda(const std::string& modcod, const int n_ite, const std::string& implem = "NMS",
                             const float norm_factor = 1.f, const bool enable_syndrome = true, const int n_frames = 1,
                             const int device = -1)
      : spu::module::Stateful()
      , auto_device(device < 0)
    {
        const std::string name = "Decoder_chain_DVBS2_cuda";
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

        auto& p = this->create_task("demodulate_decode");
        auto ps_CP = this->template create_socket_in<float>(p, "CP", 1);
        auto ps_Y_N1 = this->template create_socket_in<float>(p, "Y_N1", 2 * info.n_sym);
        auto ps_CWD1 = this->template create_socket_out<int8_t>(p, "CWD_ldpc", 1);
        auto ps_CWD2 = this->template create_socket_out<int8_t>(p, "CWD_bch", 1);
        auto ps_V_K = this->template create_socket_out<int32_t>(p, "V_K", info.K_bch);
        this->create_codelet(
          p,
          [ps_CP, ps_Y_N1, ps_CWD1, ps_CWD2, ps_V_K](spu::module::Module& m, spu::runtime::Task& t, const size_t frame_id) -> int
          {
              auto& c = static_cast<Decoder_chain_DVBS2_cuda&>(m);
              const int n = (int)c.get_n_frames_per_wave();
              int8_t* cwd = static_cast<int8_t*>(t[ps_CWD2].get_dataptr());
              c.handle->pin(t[ps_Y_N1].get_dataptr(), sizeof(float) * (size_t)n * 2 * c.handle->info().n_sym);
              c.handle->pin(t[ps_V_K].get_dataptr(), sizeof(int32_t) * (size_t)n * c.handle->info().K_bch);
              DVBS2_B200_CHECK(c.handle,
                               dvbs2_cuda_chain(c.handle->get(), static_cast<float*>(t[ps_CP].get_dataptr()),
                                                static_cast<float*>(t[ps_Y_N1].get_dataptr()),
                                                static_cast<int8_t*>(t[ps_CWD1].get_dataptr()), cwd,
                                                static_cast<int32_t*>(t[ps_V_K].get_dataptr()), nullptr, n));
              c.handle->stage_timers(t);
              int status = 0;
              for (int f = 0; f < n; f++)
                  status |= !cwd[f];
              (void)frame_id;
              return status;
          });
        for (const char* key : { "demap", "ldpc", "bch", "unpack" }) // device time per stage (Cuda_handle::stage_timers)
            this->register_timer(p, key);
        this->set_single_wave(true);
        this->set_n_frames(n_frames);
        for (auto& t : this->tasks)
            t->set_replicability(true);
    }

    virtual ~Decoder_chain_DVBS2_cuda() = default;

    virtual Decoder_chain_DVBS2_cuda* clone() const
    {
        auto m = new Decoder_chain_DVBS2_cuda(*this);
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
    virtual void deep_copy(const Decoder_chain_DVBS2_cuda& m)
    {
        spu::module::Stateful::deep_copy(m);
        this->handle = std::make_shared<Cuda_handle>(
          m.handle->cfg_for(m.handle->info().max_frames, m.auto_device ? -1 : m.handle->cfg().device));
    }
};
}

#endif /* DVBS2_B200_DECODER_CHAIN_DVBS2_CUDA_HPP_ */
