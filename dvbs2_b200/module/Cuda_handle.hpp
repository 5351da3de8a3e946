/*!
 * \file
 * \brief RAII owner of one libdvbs2_cuda handle (one device + streams + device buffers) for the StreamPU/AFF3CT modules.
 *
 * One Cuda_handle per module clone: StreamPU replicates a chain by cloning every module once per thread
 * (lib/streampu/src/Runtime/Sequence/Sequence.cpp:1506-1535); each clone gets its own stream and buffers and, when the
 * configuration asks for device -1 ("auto"), the next GPU round-robin.  C-ABI status codes become the exceptions the
 * reference modules throw (spu::tools::runtime_error / invalid_argument), so Sequence's error path is unchanged.
 */
#ifndef DVBS2_B200_CUDA_HANDLE_HPP_
#define DVBS2_B200_CUDA_HANDLE_HPP_

#include <atomic>
#include <sstream>
#include <string>

#include <streampu.hpp>

#include "dvbs2_cuda.h"

namespace dvbs2_b200
{
inline int
modcod_from_string(const std::string& modcod)
{
    // the --mod-cod values of src/common/Factory/DVBS2/DVBS2.cpp:287-317
    if (modcod == "QPSK-S_8/9") return DVBS2_QPSK_S_8_9;
    if (modcod == "QPSK-S_3/5") return DVBS2_QPSK_S_3_5;
    if (modcod == "8PSK-S_3/5") return DVBS2_8PSK_S_3_5;
    if (modcod == "8PSK-S_8/9") return DVBS2_8PSK_S_8_9;
    if (modcod == "16APSK-S_8/9") return DVBS2_16APSK_S_8_9;
    throw spu::tools::invalid_argument(__FILE__, __LINE__, __func__, modcod + " mod-cod scheme not supported.");
}

class Cuda_handle
{
  private:
    dvbs2_cuda_cfg cfg_;
    dvbs2_cuda_handle* h_;
    dvbs2_cuda_info info_;

    static int next_device()
    {
        static std::atomic<int> counter{ 0 };
        const int n = dvbs2_cuda_device_count();
        return n > 0 ? counter++ % n : 0;
    }

  public:
    /*! \param cfg: C-ABI configuration; cfg.device < 0 picks the next GPU round-robin (per created handle). */
    explicit Cuda_handle(const dvbs2_cuda_cfg& cfg)
      : cfg_(cfg)
      , h_(nullptr)
    {
        if (cfg_.device < 0) cfg_.device = next_device();
        const int rc = dvbs2_cuda_create(&cfg_, &h_);
        if (rc != DVBS2_OK)
        {
            std::stringstream message;
            message << "libdvbs2_cuda: " << dvbs2_cuda_strerror(rc) << " (device " << cfg_.device
                    << "); there is no CPU fallback.";
            if (rc == DVBS2_ERR_INVALID) throw spu::tools::invalid_argument(__FILE__, __LINE__, __func__, message.str());
            throw spu::tools::runtime_error(__FILE__, __LINE__, __func__, message.str());
        }
        dvbs2_cuda_get_info(h_, &info_);
    }
    Cuda_handle(const Cuda_handle&) = delete;
    Cuda_handle& operator=(const Cuda_handle&) = delete;
    ~Cuda_handle() { dvbs2_cuda_destroy(h_); }

    dvbs2_cuda_handle* get() const { return h_; }
    const dvbs2_cuda_cfg& cfg() const { return cfg_; }
    const dvbs2_cuda_info& info() const { return info_; }
    /*! Same configuration, other batch capacity / device (used by clone() and set_n_frames()). */
    dvbs2_cuda_cfg cfg_for(const int max_frames, const int device) const
    {
        dvbs2_cuda_cfg c = cfg_;
        c.max_frames = max_frames;
        c.device = device;
        return c;
    }

    void check(const int rc, const char* file, const int line, const char* func) const
    {
        if (rc == DVBS2_OK) return;
        std::stringstream message;
        message << "libdvbs2_cuda: " << dvbs2_cuda_strerror(rc);
        if (rc == DVBS2_ERR_CUDA) message << ": " << dvbs2_cuda_last_error(h_);
        if (rc == DVBS2_ERR_INVALID || rc == DVBS2_ERR_TOO_MANY)
            throw spu::tools::invalid_argument(std::string(file), int(line), std::string(func), message.str());
        throw spu::tools::runtime_error(std::string(file), int(line), std::string(func), message.str());
    }
};
}

#define DVBS2_B200_CHECK(handle, call) (handle)->check((call), __FILE__, __LINE__, __func__)

#endif /* DVBS2_B200_CUDA_HANDLE_HPP_ */
