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
#include <chrono>
#include <cstdlib>
#include <map>
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
    std::map<const void*, size_t> pinned_; // socket buffers registered by this handle (0 bytes: registration refused, not retried)
    bool pin_enabled_;
    bool profiling_ = false;
    dvbs2_cuda_timers timers_{};

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
      , pin_enabled_(std::getenv("DVBS2_CUDA_NO_PIN") == nullptr)
    {
        if (dvbs2_cuda_abi_check(DVBS2_CUDA_ABI_VERSION, sizeof(dvbs2_cuda_cfg), sizeof(dvbs2_cuda_stats)) != DVBS2_OK)
            throw spu::tools::runtime_error(__FILE__, __LINE__, __func__,
                                            "libdvbs2_cuda: this module was built against another version of dvbs2_cuda.h; rebuild it.");
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
    ~Cuda_handle()
    {
        unpin_all();
        dvbs2_cuda_destroy(h_);
    }

    /*!
     * Page-locks a socket buffer the first time a codelet sees it (cudaHostRegister through the C ABI), so that the
     * library's chunked cudaMemcpyAsync calls run at PCIe rate and overlap with the kernels instead of being staged
     * through the driver's bounce buffer.  Socket buffers are std::vector<uint8_t> owned by the producing task and stay
     * put after binding (lib/streampu/include/Runtime/Socket/Socket.hxx:1007-1011); set_n_frames() reallocates them, so
     * the modules call unpin_all() there.  A refused registration (e.g. the neighbouring GPU module already pinned the
     * shared buffer) is not an error: the copy takes the pageable path.  DVBS2_CUDA_NO_PIN=1 disables pinning.
     */
    void pin(const void* ptr, const size_t bytes)
    {
        if (!pin_enabled_ || ptr == nullptr || bytes < (size_t)1 << 16) return;
        auto it = pinned_.find(ptr);
        if (it != pinned_.end() && (it->second == 0 || it->second >= bytes)) return;
        if (it != pinned_.end()) dvbs2_cuda_host_unregister(h_, const_cast<void*>(ptr));
        pinned_[ptr] = dvbs2_cuda_host_register(h_, const_cast<void*>(ptr), bytes) == DVBS2_OK ? bytes : 0;
    }
    void unpin_all()
    {
        for (auto& kv : pinned_)
            if (kv.second) dvbs2_cuda_host_unregister(h_, const_cast<void*>(kv.first));
        pinned_.clear();
    }
    size_t n_pinned() const
    {
        size_t n = 0;
        for (auto& kv : pinned_)
            n += kv.second != 0;
        return n;
    }

    /*!
     * Per-stage device times as sub-timers of a task, like the "load / decode / store" timers of the CPU decoders
     * (lib/aff3ct/include/Module/Decoder/Decoder_SIHO.hxx:76-79): a module registers "demap", "ldpc", "bch", "unpack" (in this
     * order) after creating its codelet and calls stage_timers() once per codelet run.  While the task collects statistics
     * (--sim-stats, Task::set_stats) the library records CUDA events at the stage boundaries on the launching stream
     * (dvbs2_cuda_set_profiling); the time since the previous call goes to Task::update_timer, so the stage split of the GPU
     * task shows up in StreamPU's statistics table.  Costs nothing when statistics are off.
     */
    void stage_timers(spu::runtime::Task& t)
    {
        if (!t.is_stats())
        {
            if (profiling_) dvbs2_cuda_set_profiling(h_, 0), profiling_ = false;
            return;
        }
        if (!profiling_)
        { // first call with statistics on: this call was not timed, start with the next one
            dvbs2_cuda_set_profiling(h_, 1);
            dvbs2_cuda_get_timers(h_, &timers_);
            profiling_ = true;
            return;
        }
        dvbs2_cuda_timers now;
        if (dvbs2_cuda_get_timers(h_, &now) != DVBS2_OK) return;
        const double d[4] = { now.demap_ms - timers_.demap_ms, now.ldpc_ms - timers_.ldpc_ms, now.bch_ms - timers_.bch_ms,
                              (now.unpack_ms - timers_.unpack_ms) + (now.pack_ms - timers_.pack_ms) };
        timers_ = now;
        if (t.get_timers_name().size() >= 4)
            for (size_t i = 0; i < 4; i++)
                t.update_timer(i, std::chrono::nanoseconds((long long)(d[i] * 1e6)));
    }

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
