/*
 * libdvbs2_cuda — C ABI of the B200-native DVB-S2 baseband decode chain
 * (soft demapper -> deinterleaver -> horizontal-layered (N)MS LDPC decoder -> BCH t=12 hard decoder).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point replaces the codelet
 * body of one AFF3CT/StreamPU task of aff3ct/dvbs2 (citations relative to the reference tree; A/ = lib/aff3ct,
 * D/ = src, S/ = lib/streampu):
 *
 *   dvbs2_cuda_demodulate        <- Modem<B,R,Q>::demodulate codelet, A/include/Module/Modem/Modem.hxx:222-240
 *                                   (-> Modem_generic_fast::_demodulate_complex, .../Generic/Modem_generic_fast.hxx:193-225)
 *   dvbs2_cuda_deinterleave      <- Interleaver<D,T>::deinterleave, A/include/Module/Interleaver/Interleaver.hxx:228-258
 *   dvbs2_cuda_ldpc_decode_siho  <- Decoder_SIHO<B,R>::decode_siho codelet, A/include/Module/Decoder/Decoder_SIHO.hxx:57-79
 *                                   (-> Decoder_LDPC_BP_horizontal_layered<...,Update_rule_NMS>::_decode_siho,
 *                                    A/include/Module/Decoder/LDPC/BP/Horizontal_layered/Decoder_LDPC_BP_horizontal_layered.hxx:89-123)
 *   dvbs2_cuda_bch_decode_hiho   <- Decoder_HIHO<B>::decode_hiho codelet, A/include/Module/Decoder/Decoder_HIHO.hxx:42-63
 *                                   (-> Decoder_BCH_DVBS2::_decode_hiho, D/common/Module/Decoder_BCH_DVBS2/Decoder_BCH_DVBS2.cpp:28-40)
 *   dvbs2_cuda_chain             <- the four tasks above bound back to back as in D/mains/TX_RX_BB/main.cpp:87-91
 *                                   (pipeline stage 6 of D/mains/RX/main_8-sta.cpp:291-296), fused on the device
 *   dvbs2_cuda_quantize          <- Quantizer_pow2::_process, A/src/Module/Quantizer/Pow2/Quantizer_pow2.cpp:184-190
 *   dvbs2_cuda_remove_plh        <- Framer<B>::remove_plh, D/common/Module/Framer/Framer.hxx:319-343
 *   dvbs2_cuda_estimate          <- Estimator<R>::estimate codelet, D/common/Module/Estimator/Estimator.hxx:103-117
 *                                   (-> Estimator_DVBS2::_estimate, D/common/Module/Estimator/Estimator_DVBS2.hxx:31-58)
 *   dvbs2_cuda_rx_chain          <- remove_plh, estimate, the four decode tasks and Scrambler_BB::descramble
 *                                   (D/common/Module/Scrambler/Scrambler_BB/Scrambler_BB.hxx:51-72) as bound in
 *                                   D/mains/TX_RX_BB/main.cpp:86-92, fused on the device
 *   dvbs2_cuda_chain_packed      <- dvbs2_cuda_chain with the payload packed like spu::tools::Bit_packer::pack
 *                                   (S/include/Tools/Algo/Bit_packer/Bit_packer.hxx, LSB first), optional descramble
 *
 * Conventions (same as the reference sockets): buffers are frame-major and contiguous,
 * LLR/symbols/sigma are float, bits are int32_t (0/1), CWD is int8_t (1 = decoder holds a valid codeword).
 * Every function returns DVBS2_OK (0) or a negative dvbs2_status; nothing throws across the ABI.
 * One handle = one device + one stream + its device buffers; a handle is NOT thread-safe (one per module clone).
 * There is no CPU fallback: if no CUDA device is usable, dvbs2_cuda_create fails with DVBS2_ERR_CUDA.
 */
#ifndef DVBS2_CUDA_H
#define DVBS2_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define DVBS2_N_LDPC 16200 /* D/common/Factory/DVBS2/DVBS2.hpp:49 — short FECFRAME */

    typedef enum
    {
        DVBS2_OK = 0,
        DVBS2_ERR_INVALID = -1,   /* bad argument / unsupported configuration                */
        DVBS2_ERR_CUDA = -2,      /* CUDA runtime error (see dvbs2_cuda_last_error)          */
        DVBS2_ERR_TOO_MANY = -3,  /* n_frames > cfg.max_frames                               */
        DVBS2_ERR_NOMEM = -4      /* device or host allocation failed                        */
    } dvbs2_status;

    /* --mod-cod values of D/common/Factory/DVBS2/DVBS2.cpp:287-317 */
    typedef enum
    {
        DVBS2_QPSK_S_8_9 = 0,
        DVBS2_QPSK_S_3_5 = 1,
        DVBS2_8PSK_S_3_5 = 2,
        DVBS2_8PSK_S_8_9 = 3,
        DVBS2_16APSK_S_8_9 = 4
    } dvbs2_modcod;

    /* arithmetic of the LDPC decoder: AFF3CT's R type (A/include/Tools/types.h:11-49) */
    typedef enum
    {
        DVBS2_PREC_F32 = 0, /* R = float   (what the dvbs2 binaries run)                                  */
        DVBS2_PREC_I16 = 1, /* R = int16_t (AFF3CT Q_16 fixed point, sequential wrap semantics)           */
        DVBS2_PREC_I8 = 2   /* R = int8_t  (AFF3CT Q_8, sequential wrap semantics)         */
    } dvbs2_prec;

    typedef struct
    {
        int modcod;      /* dvbs2_modcod                                                                   */
        int n_ite;       /* --dec-ite  (DVBS2.cpp:135), >= 1                                               */
        float norm;      /* NMS normalisation factor; dvbs2 runs 1.0; ints: k/8 only           */
        int precision;   /* dvbs2_prec                                                                     */
        int early_stop;  /* 1 = syndrome early termination per frame (enable_syndrome, depth 1)            */
        int max_frames;  /* largest -F this handle will be asked to process in one call                    */
        int device;      /* CUDA device ordinal                                                            */
        int quant_bits;  /* Quantizer_pow2 saturation_pos (AFF3CT default 6), used by int precisions       */
        int quant_dec;   /* Quantizer_pow2 fixed_point_pos (AFF3CT default 2)                              */
        int no_compaction; /* 0 (default): frames still iterating when most of their group has converged are packed
                              into dense groups for the remaining iterations (same results, fewer bytes moved);
                              1: every frame stays in its group                                                  */
        int compaction_rounds;       /* decoding rounds incl. the first, 2..6 (0 = 2, or 3 from two waves of groups on).  With
                                      * this and the next two fields all 0 the schedule is ADAPTIVE: the handle reads back its own
                                      * iteration statistics after every call (no synchronisation) and, once the mean iteration
                                      * count of the traffic lies between 0.45 and 0.9 of n_ite, large calls run 4 rounds with a
                                      * hand-over threshold of 6/8 (float) / 7/8 (int16).  Results never depend on the schedule. */
        int compaction_exit_eighths; /* a group hands its live frames over once fewer than this many eighths of its
                                        frames still iterate, 1..7 (0 = default 4)                                   */
        int compaction_nf_later;     /* frames per lane in rounds >= 2: f32 1 or 2 (0 = 1), int16 2 or 4 (0 = 2)   */
        int ldpc_frames_per_lane;    /* f32 only: frames per lane of round 1, 1 (32-frame groups: lowest latency) or 2
                                      * (64-frame groups, 256-byte chunks: highest throughput); 0 = chosen per call from
                                      * the batch size (1 while the call fits in one wave of 32-frame groups)           */
        int ldpc_coop_max_frames;    /* f32 and int16: calls of up to this many frames (with one frame per lane) use the
                                      * cooperative small-batch LDPC kernel — a CTA per 32-frame group, ~10x shorter calls
                                      * below a few thousand frames, one round, same results; 0 = default 16384, < 0 = never */
        int host_pipeline_chunks;    /* host-buffer entry points: how many chunks a call is split into over the handle's
                                      * pipeline streams (H2D / kernels / D2H of different chunks overlap).  0 = automatic (up
                                      * to 4, never below four decoder groups per SM per chunk); 1 = one chunk per call — the
                                      * right choice when several module clones keep the GPU busy, since the copies of one
                                      * clone then overlap the kernels of the others and every launch fills the machine      */
        int ldpc_onchip;             /* f32 only: the on-chip LDPC kernel ("design B", DESIGN.md §4: variable nodes in shared
                                      * memory, lanes = checks, a chain warp per CTA, no compaction rounds; same results).
                                      * 0 = automatic (per MODCOD and batch size, from measurements), 1 = always, -1 = never
                                      * (added in round 2 in place of a reserved word: the struct keeps its size and layout, a
                                      * caller built against the older header passes 0 = automatic)                         */
        int reserved[1];             /* must be 0 */
    } dvbs2_cuda_cfg;

    typedef struct
    {
        int K_bch, N_bch, K_ldpc, N_ldpc, bps, n_sym; /* code/modulation sizes of cfg.modcod               */
        int M_ldpc, E_ldpc;                           /* checks / edges of H                               */
        int max_frames;
    } dvbs2_cuda_info;

    typedef struct
    {
        uint64_t frames;          /* frames decoded by LDPC since create/reset                              */
        uint64_t ldpc_iterations; /* sum over frames of iterations executed                                 */
        uint64_t ldpc_failures;   /* frames whose syndrome was non-zero after the last iteration            */
        uint64_t bch_nonzero;     /* frames entering Berlekamp (non-zero BCH syndrome)                      */
        uint64_t bch_failures;    /* frames BCH could not correct                                           */
        uint64_t kernel_launches; /* kernels launched by this handle (counted at every launch site)         */
        uint64_t iter_hist[33];   /* early-stop histogram: frames that left the LDPC decoder after i iterations
                                   * (i = 1 .. 31; bin 32 collects >= 32; bin 0 stays 0)                    */
    } dvbs2_cuda_stats;

    /* device time per stage, measured with CUDA events recorded on the launching stream around each kernel of the
     * chain while profiling is enabled (the reference's per-task timers: Task::set_stats / register_timer,
     * S/src/Runtime/Task/Task.cpp:348-365).  *_ms are sums over the *_calls launches since the last reset_stats. */
    typedef struct
    {
        double demap_ms, ldpc_ms, bch_ms, unpack_ms, pack_ms;
        uint64_t demap_calls, ldpc_calls, bch_calls, unpack_calls, pack_calls;
    } dvbs2_cuda_timers;

    typedef struct dvbs2_cuda_handle dvbs2_cuda_handle;

    void dvbs2_cuda_default_cfg(dvbs2_cuda_cfg* cfg); /* QPSK-S_8/9, 10 it, norm 1.0, f32, early stop, 1024 frames */
    int dvbs2_cuda_create(const dvbs2_cuda_cfg* cfg, dvbs2_cuda_handle** out);
    int dvbs2_cuda_destroy(dvbs2_cuda_handle* h);
    int dvbs2_cuda_get_info(const dvbs2_cuda_handle* h, dvbs2_cuda_info* info);
    int dvbs2_cuda_get_stats(dvbs2_cuda_handle* h, dvbs2_cuda_stats* stats); /* synchronises the stream */
    int dvbs2_cuda_reset_stats(dvbs2_cuda_handle* h);
    int dvbs2_cuda_set_profiling(dvbs2_cuda_handle* h, int enable);            /* default off */
    int dvbs2_cuda_get_timers(dvbs2_cuda_handle* h, dvbs2_cuda_timers* timers); /* synchronises the stream */
    const char* dvbs2_cuda_strerror(int status);
    const char* dvbs2_cuda_last_error(const dvbs2_cuda_handle* h); /* text of the last CUDA failure */
    int dvbs2_cuda_device_count(void);
    /* Layout version of the structs above.  A caller compiled against this header passes DVBS2_CUDA_ABI_VERSION to
     * dvbs2_cuda_abi_check(); the library answers DVBS2_ERR_INVALID when its own structs differ (a module built against an
     * older header must be rebuilt instead of corrupting memory through dvbs2_cuda_get_stats).  Cuda_handle.hpp checks it
     * before the first create. */
#define DVBS2_CUDA_ABI_VERSION 2
    int dvbs2_cuda_abi_version(void);
    int dvbs2_cuda_abi_check(int caller_version, size_t sizeof_cfg, size_t sizeof_stats);

    /* ---- host-buffer entry points (what the StreamPU modules call; H2D/D2H inside, synchronous on return) ------- */

    /* CP[F] = sigma per frame; Y_N1[F][2*n_sym] interleaved I/Q; Y_N2[F][16200] LLRs in transmitted (interleaved)
     * bit order, exactly Modem::demodulate.  Log-MAP (max*) over the MODCOD's constellation. */
    int dvbs2_cuda_demodulate(dvbs2_cuda_handle* h, const float* CP, const float* Y_N1, float* Y_N2, int n_frames);
    /* nat[f][i] = itl[f][pi_inv[i]] (identity for QPSK) */
    int dvbs2_cuda_deinterleave(dvbs2_cuda_handle* h, const float* itl, float* nat, int n_frames);
    /* Y_N[F][16200]: float (F32) / int16_t (I16) / int8_t (I8) LLRs in natural order; CWD[F]; V_K[F][K_ldpc].
     * n_iter (optional, may be NULL) receives the iterations executed per frame. */
    int dvbs2_cuda_ldpc_decode_siho(dvbs2_cuda_handle* h, const void* Y_N, int8_t* CWD, int32_t* V_K, int32_t* n_iter,
                                    int n_frames);
    /* Y_N[F][N_bch] hard bits; CWD[F]; V_K[F][K_bch] */
    int dvbs2_cuda_bch_decode_hiho(dvbs2_cuda_handle* h, const int32_t* Y_N, int8_t* CWD, int32_t* V_K, int n_frames);
    /* float LLR -> fixed point with cfg.quant_dec / cfg.quant_bits; out is int16_t or int8_t per cfg.precision */
    int dvbs2_cuda_quantize(dvbs2_cuda_handle* h, const float* in, void* out, size_t n);
    /* Fused stage: demodulate -> deinterleave -> (quantize) -> LDPC -> BCH; LLRs and LDPC bits never leave the device.
     * CWD_ldpc / CWD_bch / n_iter may be NULL.  Copies are chunked and overlapped with compute; pass pinned memory
     * (dvbs2_cuda_host_register) for full PCIe rate. */
    int dvbs2_cuda_chain(dvbs2_cuda_handle* h, const float* CP, const float* Y_N1, int8_t* CWD_ldpc, int8_t* CWD_bch,
                         int32_t* V_K, int32_t* n_iter, int n_frames);

    /* ---- compact socket formats --------------------------------------------------------------------
     * AFF3CT's sockets carry float symbols / LLRs and one int32 per bit (A/include/Tools/types.h:11-27; buffers are
     * std::vector<uint8_t> per socket, S/include/Runtime/Socket/Socket.hxx:1007-1011): 64.8 KB in and 56.9 KB out per
     * QPSK-S_8/9 frame, which is what bounds the stage once it sits behind PCIe.  The entry points below move the same
     * information in fewer bytes; dvbs2_b200/module/Adaptor_DVBS2_cuda.hpp holds the StreamPU modules that convert at the
     * stage's edges (float I/Q -> wire format upstream, packed bytes -> int32 bits downstream). */
    typedef enum
    {
        DVBS2_IQ_F32 = 0,  /* float I, float Q: the reference socket (8 B / symbol)                                    */
        DVBS2_IQ_F16 = 1,  /* IEEE binary16 I/Q (4 B / symbol); converted exactly to float on the device               */
        DVBS2_IQ_SC16 = 2, /* int16 I/Q x iq_scale (4 B / symbol), the USRP/UHD "sc16" wire format                      */
        DVBS2_IQ_SC8 = 3,  /* int8 I/Q x iq_scale (2 B / symbol), UHD "sc8"                                             */
        DVBS2_IQ_SC12 = 4  /* 12-bit two's-complement I and Q x iq_scale in 3 bytes per symbol (the resolution of UHD's "sc12"
                            * wire format): little-endian 24-bit word, I = bits 0..11, Q = bits 12..23                    */
    } dvbs2_iq_format;
    /* bytes of one frame's symbols in that format (XFEC frame, or PL frame when pl_frames != 0) */
    size_t dvbs2_cuda_iq_frame_bytes(const dvbs2_cuda_handle* h, int iq_format, int pl_frames);

    /* dvbs2_cuda_chain / dvbs2_cuda_rx_chain with the symbols in a compact wire format.  Y_N1[F][2*n_sym] (or whole PL
     * frames when pl_frames != 0) holds iq_format elements; integer samples are converted as (float)x * iq_scale — one
     * rounded multiply, exactly what a host adaptor in front of Modem::demodulate would compute — so for the same sample
     * values the results are bit-identical to the float entry points.  CP == NULL (PL frames only): sigma from the
     * on-device Estimator_DVBS2, returned in SIG if given.  Exactly one of V_K (int32 per bit) and V_K_packed
     * ([F][K_bch/8], Bit_packer layout) must be given. */
    int dvbs2_cuda_chain_compact(dvbs2_cuda_handle* h, int iq_format, float iq_scale, const float* CP, const void* Y_N1,
                                 float* SIG, int8_t* CWD_ldpc, int8_t* CWD_bch, int32_t* V_K, uint8_t* V_K_packed,
                                 int descramble, int32_t* n_iter, int pl_frames, int n_frames);
    /* Asynchronous form of dvbs2_cuda_chain_compact (host-pointer *_async entry points): queues the H2D copies,
     * the kernels and the D2H copies of the call on the handle's pipeline streams and returns; dvbs2_cuda_wait() blocks until
     * they have finished (and returns the first CUDA error).  The host buffers must be page-locked (dvbs2_cuda_host_register,
     * otherwise the copies are synchronous anyway) and must stay untouched until dvbs2_cuda_wait returns; one call in flight
     * per handle (a second chain call on the handle first waits for the pending one).  One host thread can thus keep several handles — several GPUs — busy: queue on each, then wait on each.
     * A failure while queueing waits for what was already queued before it returns, like the synchronous entry points. */
    int dvbs2_cuda_chain_compact_async(dvbs2_cuda_handle* h, int iq_format, float iq_scale, const float* CP, const void* Y_N1,
                                       float* SIG, int8_t* CWD_ldpc, int8_t* CWD_bch, int32_t* V_K, uint8_t* V_K_packed,
                                       int descramble, int32_t* n_iter, int pl_frames, int n_frames);
    int dvbs2_cuda_wait(dvbs2_cuda_handle* h);
    int dvbs2_cuda_chain_compact_dev(dvbs2_cuda_handle* h, int iq_format, float iq_scale, const float* d_CP,
                                     const void* d_Y_N1, float* d_SIG, int8_t* d_CWD_ldpc, int8_t* d_CWD_bch, int32_t* d_V_K,
                                     uint8_t* d_V_K_packed, int descramble, int32_t* d_n_iter, int pl_frames, int n_frames);

    /* Fixed-point decode stage from int8 LLRs: Decoder_SIHO::decode_siho (R = int16_t arithmetic) -> Decoder_HIHO::decode_hiho
     * (-> Scrambler_BB::descramble).  Y_N[F][16200] int8 are Quantizer_pow2 outputs (the AFF3CT default, 6 bits with 2
     * fractional, fits int8: A/src/Launcher/Code/LDPC/LDPC.cpp:58-62), a quarter of the float LLR bytes; they are widened
     * to int16 on the device, so the results equal the reference's int16 decoder on the same values bit for bit.  The handle
     * must have been created with DVBS2_PREC_I16.  Exactly one of V_K and V_K_packed must be given. */
    int dvbs2_cuda_decode_llr8(dvbs2_cuda_handle* h, const int8_t* Y_N, int8_t* CWD_ldpc, int8_t* CWD_bch, int32_t* V_K,
                               uint8_t* V_K_packed, int descramble, int32_t* n_iter, int n_frames);
    int dvbs2_cuda_decode_llr8_dev(dvbs2_cuda_handle* h, const int8_t* d_Y_N, int8_t* d_CWD_ldpc, int8_t* d_CWD_bch,
                                   int32_t* d_V_K, uint8_t* d_V_K_packed, int descramble, int32_t* d_n_iter, int n_frames);

    /* The same stage with the bit-packed output format: V_K_packed[F][K_bch/8] bytes, bit i of a frame at
     * byte i/8, bit i%8 — the layout of spu::tools::Bit_packer::pack(msb_to_lsb = false) that Sink_user_binary writes
     * (lib/streampu/src/Module/Stateful/Sink/User/Sink_user_binary.cpp:70-77) — 32x fewer D2H bytes than int32 per bit.
     * descramble != 0 also applies Scrambler_BB::descramble (src/common/Module/Scrambler/Scrambler_BB/Scrambler_BB.hxx:51-72),
     * the task bound behind decode_hiho in src/mains/TX_RX_BB/main.cpp:92 and src/mains/RX/main_sched.cpp:211, so the bytes
     * are the payload the sink stores. */
    int dvbs2_cuda_chain_packed(dvbs2_cuda_handle* h, const float* CP, const float* Y_N1, int8_t* CWD_ldpc, int8_t* CWD_bch,
                                uint8_t* V_K_packed, int32_t* n_iter, int descramble, int n_frames);

    /* ---- receiver stage from PL frames: the tasks bound in front of and behind the decode chain in
     * src/mains/TX_RX_BB/main.cpp:86-92 and src/mains/RX/main_sched.cpp:204-211 ------------------------------------------- */

    /* floats per PL frame = 2 * (90 * (n_sym / 90 + 1) + 36 * (n_sym / 1440)) (src/common/Factory/DVBS2/DVBS2.cpp:352-355) */
    int dvbs2_cuda_pl_frame_size(const dvbs2_cuda_handle* h);
    /* Framer::remove_plh (src/common/Module/Framer/Framer.hxx:319-343): Y_N1[F][pl_frame_size] -> Y_N2[F][2*n_sym]; drops the
     * 90-symbol PL header and the 36-symbol pilot block after every 1440 payload symbols */
    int dvbs2_cuda_remove_plh(dvbs2_cuda_handle* h, const float* Y_N1, float* Y_N2, int n_frames);
    /* Estimator_DVBS2::estimate (src/common/Module/Estimator/Estimator.hxx:103-117, Estimator_DVBS2.hxx:31-58): X_N[F][2*n_sym]
     * -> SIG[F], Eb_N0[F], Es_N0[F] (dB; the last two may be NULL).  M2/M4 moments summed in float in the reference's order. */
    int dvbs2_cuda_estimate(dvbs2_cuda_handle* h, const float* X_N, float* SIG, float* Eb_N0, float* Es_N0, int n_frames);
    /* Fused receiver stage: remove_plh -> estimate (when CP == NULL; SIG, optional, receives the estimates) or the caller's
     * CP[F] -> demodulate -> deinterleave -> LDPC -> BCH -> (Scrambler_BB::descramble when descramble != 0).  Exactly one of
     * V_K (int32 per bit, [F][K_bch]) and V_K_packed ([F][K_bch/8], Bit_packer layout) must be given.  The PL header and
     * pilots are skipped on the read side of the demapper; only PL frames cross PCIe. */
    int dvbs2_cuda_rx_chain(dvbs2_cuda_handle* h, const float* PL_frames, const float* CP, float* SIG, int8_t* CWD_ldpc,
                            int8_t* CWD_bch, int32_t* V_K, uint8_t* V_K_packed, int descramble, int32_t* n_iter, int n_frames);

    /* pin / unpin a caller-owned host buffer (StreamPU socket buffers are stable after binding) */
    int dvbs2_cuda_host_register(dvbs2_cuda_handle* h, void* ptr, size_t bytes);
    int dvbs2_cuda_host_unregister(dvbs2_cuda_handle* h, void* ptr);

    /* ---- device-buffer entry points (asynchronous on the handle's stream; for benchmarks and device-resident use) -- */
    void* dvbs2_cuda_stream(dvbs2_cuda_handle* h); /* cudaStream_t */
    int dvbs2_cuda_sync(dvbs2_cuda_handle* h);
    int dvbs2_cuda_demodulate_dev(dvbs2_cuda_handle* h, const float* d_CP, const float* d_Y_N1, float* d_Y_N2,
                                  int deinterleave, int n_frames);
    int dvbs2_cuda_ldpc_decode_siho_dev(dvbs2_cuda_handle* h, const void* d_Y_N, int8_t* d_CWD, int32_t* d_V_K,
                                        int32_t* d_n_iter, int n_frames);
    int dvbs2_cuda_bch_decode_hiho_dev(dvbs2_cuda_handle* h, const int32_t* d_Y_N, int8_t* d_CWD, int32_t* d_V_K,
                                       int n_frames);
    int dvbs2_cuda_chain_dev(dvbs2_cuda_handle* h, const float* d_CP, const float* d_Y_N1, int8_t* d_CWD_ldpc,
                             int8_t* d_CWD_bch, int32_t* d_V_K, int32_t* d_n_iter, int n_frames);

    int dvbs2_cuda_chain_packed_dev(dvbs2_cuda_handle* h, const float* d_CP, const float* d_Y_N1, int8_t* d_CWD_ldpc,
                                    int8_t* d_CWD_bch, uint8_t* d_V_K_packed, int32_t* d_n_iter, int descramble, int n_frames);

    int dvbs2_cuda_remove_plh_dev(dvbs2_cuda_handle* h, const float* d_Y_N1, float* d_Y_N2, int n_frames);
    int dvbs2_cuda_estimate_dev(dvbs2_cuda_handle* h, const float* d_X_N, float* d_SIG, float* d_Eb_N0, float* d_Es_N0,
                                int n_frames);
    /* d_CP == NULL: estimate on the device into d_SIG (required then) */
    int dvbs2_cuda_rx_chain_dev(dvbs2_cuda_handle* h, const float* d_PL_frames, const float* d_CP, float* d_SIG,
                                int8_t* d_CWD_ldpc, int8_t* d_CWD_bch, int32_t* d_V_K, uint8_t* d_V_K_packed, int descramble,
                                int32_t* d_n_iter, int n_frames);

    /* ---- on-device transmitter, AWGN channel and BER/FER monitor ------------------------------------------------
     * Monte-Carlo / benchmark harness around the decode path: what D/mains/TX_RX_BB/main.cpp:75-84
     * chains in front of the receiver (Source -> Scrambler_BB::scramble -> Encoder_BCH_DVBS2 -> Encoder_LDPC_DVBS2 ->
     * Interleaver::interleave -> Modem::modulate -> Channel_AWGN_LLR::add_noise) and what it chains behind it
     * (Scrambler_BB::descramble -> Monitor_BFER::check_errors, main.cpp:92-94), so frames can be produced and scored in
     * HBM.  The payload of the last tx call stays in the handle for the monitor. */
    typedef struct
    {
        uint64_t n_frames, n_bit_errors, n_frame_errors; /* Monitor_BFER counters (FRA, BE, FE) */
    } dvbs2_cuda_monitor;

    /* payload from the host: U[F][K_bch] int32 0/1 */
    int dvbs2_cuda_tx_set_payload(dvbs2_cuda_handle* h, const int32_t* U, int n_frames);
    /* payload generated on the device (counter-based hash of seed, frame, bit) */
    int dvbs2_cuda_tx_random_payload(dvbs2_cuda_handle* h, uint64_t seed, int n_frames);
    /* scramble + BCH + LDPC encode the payload, interleave, map, add AWGN of std-dev sigma per dimension (sigma = 0: no
     * noise): d_CP[F] <- sigma, d_Y_N1[F][2*n_sym] <- received symbols.  Device pointers, asynchronous. */
    int dvbs2_cuda_tx_modulate_awgn_dev(dvbs2_cuda_handle* h, float sigma, uint64_t seed, float* d_CP, float* d_Y_N1,
                                        int n_frames);
    /* bits of the last tx call, to the host: stage 0 payload [F][K_bch], 1 scrambled [F][K_bch], 2 BCH codeword
     * [F][N_bch], 3 LDPC codeword [F][16200] (natural order) */
    int dvbs2_cuda_tx_get_bits(dvbs2_cuda_handle* h, int stage, int32_t* out, int n_frames);
    /* descramble the BCH output of the last chain call on this handle and compare with the payload (asynchronous) */
    int dvbs2_cuda_monitor_check(dvbs2_cuda_handle* h, int n_frames);
    int dvbs2_cuda_monitor_get(dvbs2_cuda_handle* h, dvbs2_cuda_monitor* m); /* synchronises; reset_stats clears it */

#ifdef __cplusplus
}
#endif
#endif
