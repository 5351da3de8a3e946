/*!
 * \file
 * \brief Compact socket formats for the GPU stage and the StreamPU adaptor modules at its edges.
 *
 * AFF3CT's sockets carry one float per I/Q component and one int32_t per bit (lib/aff3ct/include/Tools/types.h:11-27,
 * socket buffers: lib/streampu/include/Runtime/Socket/Socket.hxx:1007-1011).  Behind PCIe that is 64.8 KB in + 56.9 KB out per
 * QPSK-S_8/9 frame and caps one GPU near 11 Gbit/s whatever the kernels do.  This header provides
 *
 *   IQ_packer<Q>                      task `pack`   : Y_N1 float[2 n_sym]  -> Y_N2 Q[2 n_sym]   (Q = int16_t "sc16", int8_t "sc8",
 *                                                     uint16_t = IEEE binary16 bits); runs on the CPU in the producer's stage —
 *                                                     or is skipped when the radio already delivers sc16 (UHD's wire format)
 *   Decoder_chain_compact_DVBS2_cuda<Q>  task `demodulate_decode` : CP float[1], Y_N1 Q[2 n_sym] -> CWD_ldpc, CWD_bch int8[1],
 *                                                     V_K uint8[K_bch / 8] (spu::tools::Bit_packer layout, LSB first; optionally
 *                                                     BB-descrambled) — dvbs2_cuda_chain_compact
 *   Bit_unpacker                      task `unpack` : X uint8[K/8] -> Y int32[K], for consumers that want AFF3CT's one-int-per-bit
 *                                                     (Monitor_BFER, Scrambler_BB); Sink_user_binary wants packed bytes anyway
 *
 * For the same sample values the stage's results are bit-identical to the float-socket modules (the device converts
 * (float)x * scale with one rounded multiply, exactly as IQ_unpack below / a host adaptor would).
 */
#ifndef DVBS2_B200_ADAPTOR_DVBS2_CUDA_HPP_
#define DVBS2_B200_ADAPTOR_DVBS2_CUDA_HPP_

#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <type_traits>

#include <streampu.hpp>

#include "Cuda_handle.hpp"

namespace dvbs2_b200
{
template<typename Q>
struct iq_wire
{
    static_assert(std::is_same<Q, float>::value || std::is_same<Q, uint16_t>::value || std::is_same<Q, int16_t>::value ||
                    std::is_same<Q, int8_t>::value,
                  "wire formats: float, uint16_t (IEEE half bits), int16_t (sc16), int8_t (sc8)");
    static constexpr int format = std::is_same<Q, float>::value      ? DVBS2_IQ_F32
                                  : std::is_same<Q, uint16_t>::value ? DVBS2_IQ_F16
                                  : std::is_same<Q, int16_t>::value  ? DVBS2_IQ_SC16
                                                                     : DVBS2_IQ_SC8;
};

// float -> IEEE binary16 bits, round to nearest even (what the device's __half2float inverts exactly)
inline uint16_t
float_to_half_bits(const float f)
{
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0u)); // inf / nan
    if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                                    // rounds to inf
    if (x < 0x33000001u) return (uint16_t)sign;                                                 // rounds to zero
    const int e = (int)(x >> 23) - 127;
    uint32_t m = (x & 0x7fffffu) | 0x800000u;
    int shift = e < -14 ? 13 + (-14 - e) : 13; // subnormal halves lose more bits
    const uint32_t half = m >> shift, rest = m & ((1u << shift) - 1u), mid = 1u << (shift - 1);
    uint32_t h = half + (rest > mid || (rest == mid && (half & 1u)) ? 1u : 0u);
    if (e >= -14)
        h += (uint32_t)(e + 15 - 1) << 10; // the implicit one of `half` carries into the exponent field
    return (uint16_t)(sign | h);
}

template<typename Q>
inline Q
iq_pack_one(const float v, const float inv_scale)
{
    const float q = std::round(v * inv_scale); // half away from zero, as Quantizer_pow2 rounds
    const float lim = sizeof(Q) == 2 ? 32767.f : 127.f;
    return (Q)std::fmin(std::fmax(q, -lim), lim);
}
template<>
inline float
iq_pack_one<float>(const float v, const float)
{
    return v;
}
template<>
inline uint16_t
iq_pack_one<uint16_t>(const float v, const float)
{
    return float_to_half_bits(v);
}

/*! float I/Q -> wire format (CPU).  scale = value of one LSB of the integer formats. */
template<typename Q = int16_t>
class IQ_packer : public spu::module::Stateful
{
    const float scale;

  public:
    IQ_packer(const int n_elmts, const float scale = 1.f / 8192.f, const int n_frames = 1)
      : spu::module::Stateful()
      , scale(scale)
    {
        const std::string name = "IQ_packer";
        this->set_name(name);
        this->set_short_name(name);
        if (n_elmts <= 0 || !(scale > 0.f))
            throw spu::tools::invalid_argument(__FILE__, __LINE__, __func__, "'n_elmts' and 'scale' have to be positive.");
        auto& p = this->create_task("pack");
        auto ps_in = this->template create_socket_in<float>(p, "Y_N1", n_elmts);
        auto ps_out = this->template create_socket_out<Q>(p, "Y_N2", n_elmts);
        this->create_codelet(p,
                             [ps_in, ps_out, n_elmts](spu::module::Module& m, spu::runtime::Task& t, const size_t) -> int
                             {
                                 auto& c = static_cast<IQ_packer<Q>&>(m);
                                 const float* in = static_cast<const float*>(t[ps_in].get_dataptr());
                                 Q* out = static_cast<Q*>(t[ps_out].get_dataptr());
                                 const float inv = 1.f / c.scale;
                                 for (int i = 0; i < n_elmts; i++)
                                     out[i] = iq_pack_one<Q>(in[i], inv);
                                 return 0;
                             });
        this->set_n_frames(n_frames);
        for (auto& t : this->tasks)
            t->set_replicability(true);
    }
    virtual IQ_packer<Q>* clone() const
    {
        auto m = new IQ_packer<Q>(*this);
        m->deep_copy(*this);
        return m;
    }
    float get_scale() const { return scale; }
};

/*! packed payload bytes (Bit_packer layout, LSB first) -> one int32 per bit */
class Bit_unpacker : public spu::module::Stateful
{
  public:
    Bit_unpacker(const int K, const int n_frames = 1)
      : spu::module::Stateful()
    {
        const std::string name = "Bit_unpacker";
        this->set_name(name);
        this->set_short_name(name);
        if (K <= 0 || K % 8) throw spu::tools::invalid_argument(__FILE__, __LINE__, __func__, "'K' has to be a positive multiple of 8.");
        auto& p = this->create_task("unpack");
        auto ps_in = this->template create_socket_in<uint8_t>(p, "X", K / 8);
        auto ps_out = this->template create_socket_out<int32_t>(p, "Y", K);
        this->create_codelet(p,
                             [ps_in, ps_out, K](spu::module::Module&, spu::runtime::Task& t, const size_t) -> int
                             {
                                 const uint8_t* in = static_cast<const uint8_t*>(t[ps_in].get_dataptr());
                                 int32_t* out = static_cast<int32_t*>(t[ps_out].get_dataptr());
                                 for (int i = 0; i < K; i++)
                                     out[i] = (in[i >> 3] >> (i & 7)) & 1;
                                 return 0;
                             });
        this->set_n_frames(n_frames);
        for (auto& t : this->tasks)
            t->set_replicability(true);
    }
    virtual Bit_unpacker* clone() const
    {
        auto m = new Bit_unpacker(*this);
        m->deep_copy(*this);
        return m;
    }
};

namespace chc
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

/*! The fused GPU stage behind compact sockets: Q symbols in, packed payload bytes out. */
template<typename Q = int16_t>
class Decoder_chain_compact_DVBS2_cuda : public spu::module::Stateful
{
  protected:
    std::shared_ptr<Cuda_handle> handle;
    const bool auto_device;
    const float scale;
    const bool descramble;

  public:
    inline spu::runtime::Task& operator[](const chc::tsk t) { return spu::module::Module::operator[]((size_t)t); }
    inline spu::runtime::Socket& operator[](const chc::sck::demodulate_decode s)
    {
        return spu::module::Module::operator[]((size_t)chc::tsk::demodulate_decode)[(size_t)s];
    }

    /*!
     * \param scale:      value of one LSB for the integer wire formats (ignored for float / half).
     * \param descramble: also apply Scrambler_BB::descramble (the task bound behind decode_hiho, main.cpp:92), so that V_K is
     *                    the payload a Sink_user_binary stores.
     */
    Decoder_chain_compact_DVBS2_cuda(const std::string& modcod, const int n_ite, const std::string& implem = "NMS",
                                     const float norm_factor = 1.f, const bool enable_syndrome = true, const int n_frames = 1,
                                     const int device = -1, const float scale = 1.f / 8192.f, const bool descramble = true)
      : spu::module::Stateful()
      , auto_device(device < 0)
      , scale(scale)
      , descramble(descramble)
    {
        const std::string name = "Decoder_chain_compact_DVBS2_cuda";
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
        auto ps_Y_N1 = this->template create_socket_in<Q>(p, "Y_N1", 2 * info.n_sym);
        auto ps_CWD1 = this->template create_socket_out<int8_t>(p, "CWD_ldpc", 1);
        auto ps_CWD2 = this->template create_socket_out<int8_t>(p, "CWD_bch", 1);
        auto ps_V_K = this->template create_socket_out<uint8_t>(p, "V_K", info.K_bch / 8);
        this->create_codelet(
          p,
          [ps_CP, ps_Y_N1, ps_CWD1, ps_CWD2, ps_V_K](spu::module::Module& m, spu::runtime::Task& t, const size_t frame_id) -> int
          {
              auto& c = static_cast<Decoder_chain_compact_DVBS2_cuda<Q>&>(m);
              const int n = (int)c.get_n_frames_per_wave();
              const auto& inf = c.handle->info();
              int8_t* cwd = static_cast<int8_t*>(t[ps_CWD2].get_dataptr());
              c.handle->pin(t[ps_Y_N1].get_dataptr(), sizeof(Q) * (size_t)n * 2 * inf.n_sym);
              c.handle->pin(t[ps_V_K].get_dataptr(), (size_t)n * (inf.K_bch / 8));
              DVBS2_B200_CHECK(c.handle,
                               dvbs2_cuda_chain_compact(c.handle->get(), iq_wire<Q>::format, c.scale,
                                                        static_cast<float*>(t[ps_CP].get_dataptr()), t[ps_Y_N1].get_dataptr(),
                                                        nullptr, static_cast<int8_t*>(t[ps_CWD1].get_dataptr()), cwd, nullptr,
                                                        static_cast<uint8_t*>(t[ps_V_K].get_dataptr()), c.descramble ? 1 : 0,
                                                        nullptr, 0, n));
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

    virtual ~Decoder_chain_compact_DVBS2_cuda() = default;

    virtual Decoder_chain_compact_DVBS2_cuda<Q>* clone() const
    {
        auto m = new Decoder_chain_compact_DVBS2_cuda<Q>(*this);
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
    virtual void deep_copy(const Decoder_chain_compact_DVBS2_cuda<Q>& m)
    {
        spu::module::Stateful::deep_copy(m);
        this->handle = std::make_shared<Cuda_handle>(
          m.handle->cfg_for(m.handle->info().max_frames, m.auto_device ? -1 : m.handle->cfg().device));
    }
};
}

#endif /* DVBS2_B200_ADAPTOR_DVBS2_CUDA_HPP_ */
