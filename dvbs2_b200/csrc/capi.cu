// C ABI of libdvbs2_cuda (include/dvbs2_cuda.h): handle, device buffers, stream plumbing.  No compute here.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only; ranges cost nothing unless a profiler is attached

#include "../../include/dvbs2_cuda.h"
#include "kernels.h"

namespace dvbs2
{
bool bch_minpoly_matches(const BchTables& t); // bch_kernel.cu
uint64_t&
launch_counter()
{
    static thread_local uint64_t n = 0;
    return n;
}
}
using namespace dvbs2;

namespace
{
constexpr int N_PIPE = 4; // streams used to overlap H2D / compute / D2H in the host-buffer entry points
constexpr int MAX_ROUNDS = 6; // decoding rounds of one LDPC call (frame compaction between them)

struct DevBuf
{
    void* p = nullptr;
    size_t bytes = 0;
};

// per-stage device timers (dvbs2_cuda_set_profiling): an event is recorded on the launching stream at every stage
// boundary; consecutive marks of one stream delimit the stage named by the first of the two
enum ProfStage
{
    P_END = -1,
    P_DEMAP = 0,
    P_LDPC,
    P_BCH,
    P_UNPACK,
    P_PACK,
    P_COUNT
};
struct ProfMark
{
    cudaEvent_t ev;
    cudaStream_t s;
    int stage;
};
constexpr size_t PROF_MAX_MARKS = 1 << 15;
} // namespace

struct dvbs2_cuda_handle
{
    dvbs2_cuda_cfg cfg{};
    ModcodParams mc{};
    LdpcTables lt;
    int max_frames = 0, max_groups = 0; // max_groups counts 32-frame subgroups
    int nf = 1, n_sms = 148;             // frames per lane of the LDPC kernel (group = 32 * nf frames)
    int k8 = 0;
    size_t elt = 4; // sizeof(R)
    int coop_max = 16384; // float calls of up to this many frames use the cooperative small-batch LDPC kernel (0 = never)
    int nf_cfg = 0; // float only: frames per lane forced by the configuration (0 = chosen per call from the batch size)
    int cn_div = 1; // check state: 4 * state_words / cn_div bytes per frame and check (int16 packs two frames into one element)
    int state_words = 4; // ldpc_state_words(d_info): 3 for the degree-11 checks of the rate-3/5 codes
    size_t cn_bytes = 16; // = 4 * state_words / cn_div
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[N_PIPE] = {};
    // code tables
    uint16_t* d_idx = nullptr;
    uint32_t* d_idx32 = nullptr;
    uint32_t* d_last = nullptr;
    int* d_queue = nullptr;
    uint16_t* d_alpha_to = nullptr;
    int16_t* d_index_of = nullptr;
    float* d_cstl = nullptr;
    uint8_t* d_scr_bytes = nullptr; // BB scrambler sequence, 8 bits per byte, LSB first
    BchDevTables bch{};
    BchLaunch bch_seg{}; // seg_len / seg_mul of the segmented syndrome division (per MODCOD)
    // decoder state
    DevBuf var, cn, hd, out_hd, rem, nz, cwd_ldpc, cwd_bch, nit, stats;
    // frame compaction (multi-round decoding).  var/cn: region A = slots [0, max_frames) (rounds 1, 3), region B =
    // [max_frames, max_frames + ext_cap) (rounds 2, 4).  hd/out_hd and the per-slot bookkeeping live in one slot space:
    // round 1 at [0, max_frames), later rounds in [max_frames, max_frames + ext_slots).
    DevBuf surv_mask, surv_iters, dst_base, src_of_dst, slot_frame, slot_iters, final_loc, n_surv;
    int ext_cap = 0, ext_slots = 0;
    int rounds = 1;       // decoding rounds (1 = no compaction)
    bool auto_rounds = false; // rounds chosen per call from the batch size (cfg.compaction_rounds == 0)
    int exit_eighths = 4; // a group hands its live frames over once fewer than exit_eighths/8 of its frames still iterate
    int nf_later = 1;     // frames per lane in rounds >= 2
    // Adaptive schedule (all three compaction knobs of the configuration left at 0): the best (rounds, hand-over threshold,
    // later-round width) depends on how the iteration counts of the frames spread, which the decoder learns from its own
    // statistics.  After a call the counters {frames, iterations} are copied to page-locked host memory on the call's stream;
    // the next call polls the event (no synchronisation) and picks its schedule from the mean iteration count of the calls
    // completed so far: see pick_schedule().  The memory regions are laid out for the widest hand-over threshold any
    // schedule uses (layout_eighths).
    bool async_pending = false; // a dvbs2_cuda_chain_compact_async call has been queued and not yet collected by dvbs2_cuda_wait
    bool adaptive = false;
    int layout_eighths = 4;
    unsigned long long* stat_mirror = nullptr; // pinned: {STAT_FRAMES, STAT_LDPC_ITERS, STAT_LDPC_FAIL} as of the last read-back
    cudaEvent_t stat_event = nullptr;
    bool stat_pending = false;
    unsigned long long seen_frames = 0, seen_iters = 0, seen_fails = 0;
    double mean_it = -1.0; // mean iterations per frame of the most recent read-back interval (-1: nothing known yet)
    double fail_frac = 0.0; // fraction of frames of that interval whose syndrome was still non-zero after the last iteration
    // staging for the host-buffer entry points (grown on demand, never beyond max_frames)
    DevBuf in_a, out_a, sigma;
    // transmitter / monitor (allocated on first use)
    DevBuf tx_src, tx_cw, tx_scr, tx_gmask, mon;
    // on-chip LDPC kernel ("design B"): transposed index table, state of the frames in flight (one region per stream), packed bits
    uint32_t* d_idxT = nullptr;
    DevBuf oc_cn, oc_pv, oc_bits;
    int oc_F = 0;      // frames per CTA
    int onchip = 0;    // cfg.ldpc_onchip
    int tx_encoded_groups = 0; // 32-frame groups of the current payload that launch_tx_encode has encoded
    uint64_t launches = 0;
    std::string last_error;
    // profiling
    bool profiling = false;
    std::vector<ProfMark> marks;
    size_t n_marks = 0;
    double t_ms[P_COUNT] = {};
    uint64_t t_calls[P_COUNT] = {};
};

namespace
{
int
fail_cuda(dvbs2_cuda_handle* h, cudaError_t e, const char* what)
{
    if (h)
    {
        char buf[512];
        snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
        h->last_error = buf;
    }
    cudaGetLastError(); // clear sticky-less errors
    return e == cudaErrorMemoryAllocation ? DVBS2_ERR_NOMEM : DVBS2_ERR_CUDA;
}

// NVTX range around a C-ABI call / a stage of the chain as it is queued: the hook a timeline profiler (Nsight Systems) needs to
// attribute kernels and copies to the StreamPU task that caused them (the reference's tasks carry timers; these are
// the device-side counterpart next to dvbs2_cuda_get_timers)
struct NvtxRange
{
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// credits the kernels launched by this thread during an API call to the handle the call works on
struct LaunchScope
{
    dvbs2_cuda_handle* h;
    uint64_t at;
    explicit LaunchScope(dvbs2_cuda_handle* hh)
      : h(hh)
      , at(launch_counter())
    {
    }
    ~LaunchScope()
    {
        if (h) h->launches += launch_counter() - at;
    }
};

#define CU(h, call)                                                                                                    \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess) return fail_cuda((h), e__, #call);                                                     \
    } while (0)

#define TRY(expr)                                                                                                      \
    do                                                                                                                 \
    {                                                                                                                  \
        int rc__ = (expr);                                                                                             \
        if (rc__ != DVBS2_OK) return rc__;                                                                             \
    } while (0)

int
ensure(dvbs2_cuda_handle* h, DevBuf& b, size_t bytes)
{
    if (b.bytes >= bytes) return DVBS2_OK;
    if (b.p) CU(h, cudaFree(b.p));
    b.p = nullptr, b.bytes = 0;
    CU(h, cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
    return DVBS2_OK;
}


int
check_frames(dvbs2_cuda_handle* h, int n_frames)
{
    if (!h || n_frames < 0) return DVBS2_ERR_INVALID;
    if (n_frames > h->max_frames) return DVBS2_ERR_TOO_MANY;
    return DVBS2_OK;
}

int
prof_collect(dvbs2_cuda_handle* h)
{
    if (h->n_marks == 0) return DVBS2_OK;
    CU(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < N_PIPE; i++)
        CU(h, cudaStreamSynchronize(h->pipe[i]));
    for (size_t i = 0; i < h->n_marks; i++)
    {
        const ProfMark& a = h->marks[i];
        if (a.stage < 0) continue;
        for (size_t j = i + 1; j < h->n_marks; j++)
            if (h->marks[j].s == a.s)
            {
                float ms = 0.f;
                CU(h, cudaEventElapsedTime(&ms, a.ev, h->marks[j].ev));
                h->t_ms[a.stage] += ms;
                h->t_calls[a.stage] += 1;
                break;
            }
    }
    h->n_marks = 0;
    return DVBS2_OK;
}

int
prof_mark(dvbs2_cuda_handle* h, cudaStream_t s, int stage)
{
    static const char* const names[] = { "dvbs2:demap", "dvbs2:ldpc", "dvbs2:bch", "dvbs2:unpack", "dvbs2:pack" };
    nvtxMarkA(stage >= 0 && stage < P_COUNT ? names[stage] : "dvbs2:end"); // where the stage's launches begin on the host timeline
    if (!h->profiling) return DVBS2_OK;
    if (h->n_marks == PROF_MAX_MARKS) TRY(prof_collect(h));
    if (h->n_marks == h->marks.size())
    {
        ProfMark m{};
        CU(h, cudaEventCreate(&m.ev));
        h->marks.push_back(m);
    }
    ProfMark& m = h->marks[h->n_marks++];
    m.s = s, m.stage = stage;
    CU(h, cudaEventRecord(m.ev, s));
    return DVBS2_OK;
}

// Frames per lane of one call.  Float: 32-frame groups (one frame per lane) while the call fits in one wave of them — a warp
// then needs about half the time per check, which is what bounds a small batch (one call is a chain of n_ite * M dependent
// checks: 30 ms at -F <= 16384 with 64-frame groups, measured) — and 64-frame groups (256-byte chunks, the bandwidth-efficient
// shape) above that.  The integer precisions keep their fixed four frames per lane.
int
call_nf(const dvbs2_cuda_handle* h, int n_frames)
{
    // int16: four frames per lane (packed pairs), except for the small calls that go to the cooperative kernel (one per lane)
    if (h->cfg.precision == DVBS2_PREC_I16 && h->coop_max > 0 && n_frames <= h->coop_max) return 1;
    if (h->cfg.precision != DVBS2_PREC_F32) return h->nf;
    if (h->nf_cfg) return h->nf_cfg;
    return n_frames <= h->n_sms * LDPC_WARPS * 32 ? 1 : 2;
}

// Small float calls go to the cooperative kernel (a CTA per 32-frame group, ldpc_kernel.cu): below ~32k frames a call is bound
// by the length of its dependency chain, which that kernel shortens by an order of magnitude; above, the multi-round
// lane = frame kernel moves fewer bytes per frame (frame compaction) in larger chunks.
bool
call_coop(const dvbs2_cuda_handle* h, int n_frames, int nf)
{
    return (h->cfg.precision == DVBS2_PREC_F32 || h->cfg.precision == DVBS2_PREC_I16) && nf == 1 && h->coop_max > 0 &&
           n_frames <= h->coop_max;
}

// `first_frame` is a multiple of the group size 32 * nf; per-frame offsets follow from the layouts in kernels.h
LdpcLaunch
ldpc_args(dvbs2_cuda_handle* h, cudaStream_t s, int first_frame, int n_frames, int8_t* cwd, int32_t* n_iter, int nf)
{
    LdpcLaunch a{};
    a.var = (char*)h->var.p + (size_t)first_frame * N_LDPC * h->elt;
    a.cn = (uint4*)((char*)h->cn.p + (size_t)first_frame * h->mc.M * h->cn_bytes);
    a.hd = (uint32_t*)h->hd.p + (size_t)(first_frame / 32) * N_LDPC;
    a.out_hd = (uint32_t*)h->out_hd.p + (size_t)(first_frame / 32) * N_LDPC;
    a.idx32 = h->d_idx32, a.last = h->d_last;
    int slot = 0; // launches of one stream are ordered, so one queue head per stream is enough
    for (int i = 0; i < N_PIPE; i++)
        if (s == h->pipe[i]) slot = 1 + i;
    a.queue = h->d_queue + slot;
    a.cwd = cwd;
    a.n_iter = n_iter;
    a.stats = (unsigned long long*)h->stats.p;
    a.n_frames = n_frames;
    a.n_groups = (n_frames + 32 * nf - 1) / (32 * nf);
    a.nf = nf;
    a.K = h->mc.N_bch, a.M = h->mc.M, a.d_info = h->mc.d_info;
    a.n_ite = h->cfg.n_ite, a.norm = h->cfg.norm, a.k8 = h->k8, a.early_stop = h->cfg.early_stop;
    a.n_sms = h->n_sms;
    return a;
}

BchLaunch
bch_args(dvbs2_cuda_handle* h, int first_frame, int n_frames, int8_t* cwd)
{
    const int sg = first_frame / 32;
    BchLaunch b{};
    b.words = (uint32_t*)h->out_hd.p + (size_t)sg * N_LDPC;
    b.stride = N_LDPC;
    b.rem = (uint32_t*)h->rem.p + (size_t)sg * BCH_NPOLY * GF_M;
    b.nz = (uint32_t*)h->nz.p + sg;
    b.cwd = cwd;
    b.stats = (unsigned long long*)h->stats.p;
    b.n_frames = n_frames, b.n_groups = (n_frames + 31) / 32;
    b.N = h->mc.N_bch, b.K = h->mc.K_bch;
    b.seg_len = h->bch_seg.seg_len;
    memcpy(b.seg_mul, h->bch_seg.seg_mul, sizeof(b.seg_mul));
    return b;
}

DemapLaunch
demap_args(dvbs2_cuda_handle* h, const float* d_sigma, const float* d_sym, int n_frames, bool pl_frames = false,
           int iq_fmt = DVBS2_IQ_F32, float iq_scale = 1.f)
{
    DemapLaunch d{};
    d.sym_fmt = iq_fmt, d.iq_scale = iq_scale;
    d.pl_sym = pl_frames ? pl_frame_symbols(h->mc.n_sym) : 0;
    d.sigma = d_sigma, d.sym = d_sym, d.cstl = h->d_cstl, d.n_frames = n_frames;
    d.bps = h->mc.bps, d.n_sym = h->mc.n_sym, d.itl_cols = h->mc.itl_cols, d.itl_rev = h->mc.itl_rev;
    d.quant_dec = h->cfg.quant_dec, d.quant_bits = h->cfg.quant_bits;
    return d;
}

// Slot-space offsets of a chunk that starts at frame f0 (a multiple of 256; chunks are >= 2048 frames long except the
// last): the regions of different chunks must be disjoint because chunks run concurrently on the pipeline streams.
// A chunk of n frames needs < n * e + 128 slots of region B (e = exit_eighths / 8) and < n * (e + e^2 + e^3) + 384 slots
// of the later rounds' hd / out_hd / bookkeeping space (up to MAX_ROUNDS rounds).
size_t
region_b_offset(const dvbs2_cuda_handle* h, int f0)
{
    // exit_eighths = 4 (the default): n / 2 slots suffice for a chunk of n frames (n a multiple of 256), no padding
    if (h->layout_eighths == 4) return (size_t)f0 / 2;
    return ((size_t)f0 * h->layout_eighths / 8 + 255) / 256 * 256 + (size_t)384 * (f0 / 2048);
}
size_t
ext_slot_offset(const dvbs2_cuda_handle* h, int f0)
{
    // n * (e + e^2 + ... + e^(MAX_ROUNDS - 1)) with e = exit_eighths / 8, rounded up, + one group of padding per round and chunk
    double sum = 0, pw = 1;
    for (int r = 2; r <= MAX_ROUNDS; r++)
        pw *= h->layout_eighths / 8.0, sum += pw;
    return ((size_t)std::ceil((double)f0 * sum) + 255) / 256 * 256 + (size_t)256 * (MAX_ROUNDS - 1) * (f0 / 2048);
}

// Which float calls go to the on-chip kernel when the configuration leaves it open (cfg.ldpc_onchip == 0): see DESIGN.md §4
// for the measurements behind the rule.
bool
call_onchip(const dvbs2_cuda_handle* h, int n_frames)
{
    if (h->onchip < 0) return false;
    if (h->onchip > 0) return true;
    (void)n_frames;
    return false; // not the default until it wins on the workload at hand (profiles/README.md)
}

// on-chip LDPC decode of frame-major LLRs; hard decisions end up in out_hd (bit-sliced), like after run_ldpc_core
int
run_ldpc_onchip(dvbs2_cuda_handle* h, cudaStream_t s, int first_frame, const float* d_llr, int8_t* d_cwd, int32_t* d_nit, int n_frames)
{
    int slot = 0;
    for (int i = 0; i < N_PIPE; i++)
        if (s == h->pipe[i]) slot = 1 + i;
    const size_t region = (size_t)h->n_sms * h->oc_F * slot;
    OnchipLaunch a{};
    a.llr = d_llr;
    a.idxT = h->d_idxT;
    a.cn = (uint4*)h->oc_cn.p + region * h->mc.M;
    a.pv = (float*)h->oc_pv.p + region * h->mc.M;
    const int KW = (h->mc.N_bch + 31) / 32;
    a.out_bits = (uint32_t*)h->oc_bits.p + (size_t)first_frame * KW;
    a.queue = h->d_queue + 40 + slot;
    a.cwd = d_cwd, a.n_iter = d_nit, a.stats = (unsigned long long*)h->stats.p;
    a.n_frames = n_frames, a.K = h->mc.N_bch, a.M = h->mc.M, a.d_info = h->mc.d_info;
    a.n_ite = h->cfg.n_ite, a.norm = h->cfg.norm, a.early_stop = h->cfg.early_stop, a.n_sms = h->n_sms;
    CU(h, launch_ldpc_onchip(a, s));
    CU(h, launch_bits_to_sliced(a.out_bits, (uint32_t*)h->out_hd.p + (size_t)(first_frame / 32) * N_LDPC, h->mc.N_bch, N_LDPC,
                                n_frames, s));
    return DVBS2_OK;
}

// The schedule of one multi-round call: decoding rounds, hand-over threshold (eighths of a group), frames per lane of the
// later rounds.  Fixed by the configuration, or — h->adaptive — picked from what the previous calls measured:
//   * mean iterations between 0.45 and 0.9 of the budget (frames converge at widely different iterations — the
//     waterfall edge, e.g. QPSK-S_8/9 at 4 dB: 6.1 of 10), few frames that never converge (<= 20 % float, <= 8 % int16) and a
//     launch of at least two waves: hand over earlier and more often (4 rounds, threshold 6/8 with 64-frame later groups for
//     float, 7/8 for int16);
//   * otherwise (everything converges early and together — 8PSK-S_8/9 at 7.5 dB, 2.9 of 10 — or hardly anything converges
//     — QPSK-S_3/5 at 2.2 dB, 9.7 of 10 — or nothing is known yet): 3 rounds at 4/8 from two waves on, else 2 rounds.
// Measured on one box (profiles/r2_schedule_sweep.txt): QPSK-S_8/9 float 19.0 -> 19.75 Gbit/s, int16 28.8 -> 29.9; the same
// schedule costs 5 % (float) / 10 % (int16) on 8PSK-S_8/9 at 7.5 dB and 2 - 4 % on QPSK-S_3/5, hence the rule.  The schedule
// never changes a result (tests: test_multi_round_compaction_is_transparent).
struct Schedule
{
    int rounds, eighths, nf_later;
};

Schedule
pick_schedule(dvbs2_cuda_handle* h, int n_groups0)
{
    const bool waves2 = n_groups0 >= 2 * h->n_sms * LDPC_WARPS;
    Schedule sc{ (h->auto_rounds && !waves2) ? 2 : h->rounds, h->exit_eighths, h->nf_later };
    if (!h->adaptive) return sc;
    if (h->stat_pending && cudaEventQuery(h->stat_event) == cudaSuccess)
    {
        h->stat_pending = false;
        unsigned long long f = h->stat_mirror[0], it = h->stat_mirror[1], fl = h->stat_mirror[2];
        if (f < h->seen_frames || it < h->seen_iters || fl < h->seen_fails) // the counters were reset in between
            h->seen_frames = h->seen_iters = h->seen_fails = 0;
        if (f - h->seen_frames >= 4096)
        {
            h->mean_it = (double)(it - h->seen_iters) / (double)(f - h->seen_frames);
            h->fail_frac = (double)(fl - h->seen_fails) / (double)(f - h->seen_frames);
            h->seen_frames = f, h->seen_iters = it, h->seen_fails = fl;
        }
    }
    else
        cudaGetLastError(); // cudaErrorNotReady is not an error
    const double x = h->mean_it / h->cfg.n_ite;
    // frames that never converge are carried through every compaction: the wider the threshold, the fewer of them it tolerates
    // (int16 with 7/8: +3.6 % at 4 % failures, -4.2 % at 13 %; float with 6/8: +4.9 % at 3 %, +3.2 % at 11 %)
    const double max_fail = h->cfg.precision == DVBS2_PREC_F32 ? 0.2 : 0.08;
    if (waves2 && h->mean_it > 0 && x >= 0.45 && x <= 0.9 && h->fail_frac <= max_fail)
    {
        sc.rounds = 4;
        sc.eighths = h->layout_eighths; // 6 (float) / 7 (int16)
        sc.nf_later = h->cfg.precision == DVBS2_PREC_F32 ? 2 : h->nf_later;
    }
    return sc;
}

// LDPC decode of slots [first_frame, first_frame + n_frames): one round, or several rounds with frame compaction in
// between (at most h->rounds; a round is worth launching only while the previous one had at least two groups)
int
run_ldpc_core(dvbs2_cuda_handle* h, cudaStream_t s, int first_frame, int n_frames, int8_t* d_cwd, int32_t* d_nit, int nf0)
{
    LdpcLaunch a = ldpc_args(h, s, first_frame, n_frames, d_cwd, d_nit, nf0);
    if (call_coop(h, n_frames, nf0))
    {
        // the cooperative kernel keeps its own check-state layout (one round, nothing else reads it): per-chunk offset in its units
        a.cn = (uint4*)((char*)h->cn.p + (size_t)first_frame * h->mc.M * ldpc_coop_state_bytes(h->cfg.precision, h->mc.d_info));
        CU(h, launch_ldpc_decode_coop(a, h->cfg.precision, s));
        return DVBS2_OK;
    }
    const bool multi = h->rounds > 1 && h->ext_cap > 0 && h->cfg.early_stop && h->cfg.n_ite >= 4 && a.n_groups >= 2;
    if (!multi)
    {
        CU(h, launch_ldpc_decode(a, h->cfg.precision, s));
        return DVBS2_OK;
    }
    int slot = 0;
    for (int i = 0; i < N_PIPE; i++)
        if (s == h->pipe[i]) slot = 1 + i;
    const size_t F = (size_t)h->max_frames;
    const size_t vb = F + region_b_offset(h, first_frame);     // region B of this chunk (var / cn slots)
    size_t hs = F + ext_slot_offset(h, first_frame);           // next free slot of the later rounds' slot space
    uint32_t* const out_hd0 = a.out_hd;                        // round-1 subgroups of this chunk: the merge target
    uint32_t* const mask0 = (uint32_t*)h->surv_mask.p + first_frame / 32;
    int* const final_loc = (int*)h->final_loc.p + first_frame;
    const int n_groups0 = a.n_groups;

    // round 1 (in place), then compaction + round r for r = 2..rounds
    const Schedule sc = pick_schedule(h, n_groups0);
    a.exit_below = 32 * nf0 * sc.eighths / 8;
    a.surv_mask = mask0;
    a.surv_iters = (int*)h->surv_iters.p + first_frame;
    CU(h, launch_ldpc_decode(a, h->cfg.precision, s));
    LdpcLaunch prev = a;
    size_t prev_hs = (size_t)first_frame; // slot-space index of prev's slot 0
    const int* prev_frame = nullptr;
    const int* prev_n_slots = nullptr;
    // measured (profiles/README.md): a third round pays once a launch holds two or more waves of groups (the queue then
    // has work to balance); with a single wave the extra compaction and the short tail kernel cost more than they save
    const int rounds = sc.rounds;
    for (int r = 2; r <= rounds && prev.n_groups >= 2; r++)
    {
        const int nf = std::min(nf0, ldpc_later_frames_per_lane(h->cfg.precision, sc.nf_later)); // groups never widen
        const int gsz = 32 * nf;
        // every group of the previous round hands over fewer than exit_below frames
        const int cap = (prev.n_groups * (prev.exit_below - 1) + gsz - 1) / gsz * gsz;
        const bool to_b = (r % 2) == 0;
        const size_t vslot = to_b ? vb : (size_t)first_frame;
        if ((to_b ? (vb - F + cap > (size_t)h->ext_cap) : (cap > n_frames + 32 * nf0)) || hs - F + cap > (size_t)h->ext_slots)
        { // cannot happen by construction; frames already handed over would be lost, so fail loudly
            h->last_error = "frame compaction: region overflow";
            return DVBS2_ERR_INVALID;
        }
        CompactLaunch c{};
        c.surv_mask = prev.surv_mask, c.surv_iters = prev.surv_iters, c.src_frame = prev_frame, c.n_src_slots = prev_n_slots;
        c.dst_base = (int*)h->dst_base.p + prev_hs / 32;
        c.n_surv = (int*)h->n_surv.p + 8 * (r - 2) + slot;
        c.src_of_dst = (int*)h->src_of_dst.p + hs;
        c.dst_frame = (int*)h->slot_frame.p + hs;
        c.dst_iters = (int*)h->slot_iters.p + hs;
        c.final_loc = final_loc;
        c.dst_slot0 = (int)(hs - first_frame);
        c.var_src = prev.var, c.var_dst = (char*)h->var.p + vslot * N_LDPC * h->elt;
        c.cn_src = prev.cn, c.cn_dst = (uint4*)((char*)h->cn.p + vslot * h->mc.M * h->cn_bytes);
        c.state_words = h->state_words;
        c.n_groups = prev.n_groups, c.nf = prev.nf, c.cap = cap, c.M = h->mc.M, c.nf_dst = nf;
        CU(h, launch_compact(c, h->cfg.precision, s));

        LdpcLaunch b = a;
        b.var = c.var_dst, b.cn = c.cn_dst;
        b.hd = (uint32_t*)h->hd.p + (hs / 32) * N_LDPC;
        b.out_hd = (uint32_t*)h->out_hd.p + (hs / 32) * N_LDPC;
        b.queue = h->d_queue + 8 * (r - 1) + slot;
        b.n_frames = cap, b.n_groups = cap / gsz, b.nf = nf;
        b.n_slots_dev = c.n_surv, b.slot_frame = c.dst_frame, b.slot_iters = c.dst_iters;
        const bool last = r == rounds || b.n_groups < 2;
        b.exit_below = last ? 0 : gsz * sc.eighths / 8;
        b.surv_mask = last ? nullptr : (uint32_t*)h->surv_mask.p + hs / 32;
        b.surv_iters = last ? nullptr : (int*)h->surv_iters.p + hs;
        CU(h, launch_ldpc_decode(b, h->cfg.precision, s));
        prev = b, prev_hs = hs, prev_frame = c.dst_frame, prev_n_slots = c.n_surv;
        hs += cap;
        if (last) break;
    }
    MergeLaunch m{};
    m.surv_mask = mask0, m.final_loc = final_loc, m.out_hd = out_hd0, m.n_groups = n_groups0, m.nf = nf0;
    CU(h, launch_merge_back(m, s));
    if (h->adaptive && !h->stat_pending)
    { // what this call (and everything before it on this stream) measured, for the calls after it: no host synchronisation
        static_assert(STAT_FRAMES == 0 && STAT_LDPC_ITERS == 1 && STAT_LDPC_FAIL == 2, "the read-back copies the first three counters");
        CU(h, cudaMemcpyAsync(h->stat_mirror, h->stats.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CU(h, cudaEventRecord(h->stat_event, s));
        h->stat_pending = true;
    }
    return DVBS2_OK;
}

// device-side pieces, all asynchronous on `s`; frames occupy device slots [first_frame, ...), first_frame a multiple of the group size
int
run_ldpc(dvbs2_cuda_handle* h, cudaStream_t s, int first_frame, const void* d_Y, int8_t* d_cwd, int32_t* d_VK,
         int32_t* d_nit, int n_frames)
{
    if (n_frames == 0) return DVBS2_OK;
    void* var = (char*)h->var.p + (size_t)first_frame * N_LDPC * h->elt;
    const int nf = call_nf(h, n_frames);
    if (call_onchip(h, n_frames))
    { // the on-chip kernel reads the socket layout directly: no repacking
        TRY(prof_mark(h, s, P_LDPC));
        TRY(run_ldpc_onchip(h, s, first_frame, (const float*)d_Y, d_cwd, d_nit, n_frames));
    }
    else
    {
        TRY(prof_mark(h, s, P_PACK));
        CU(h, launch_pack_llr(d_Y, var, h->cfg.precision, n_frames, nf, s));
        TRY(prof_mark(h, s, P_LDPC));
        TRY(run_ldpc_core(h, s, first_frame, n_frames, d_cwd, d_nit, nf));
    }
    TRY(prof_mark(h, s, P_UNPACK));
    CU(h, launch_unpack_bits((uint32_t*)h->out_hd.p + (size_t)(first_frame / 32) * N_LDPC, d_VK, N_LDPC, h->mc.N_bch, n_frames, s));
    TRY(prof_mark(h, s, P_END));
    return DVBS2_OK;
}

int
run_bch(dvbs2_cuda_handle* h, cudaStream_t s, int first_frame, const int32_t* d_Y, int8_t* d_cwd, int32_t* d_VK,
        int n_frames)
{
    if (n_frames == 0) return DVBS2_OK;
    uint32_t* words = (uint32_t*)h->out_hd.p + (size_t)(first_frame / 32) * N_LDPC;
    TRY(prof_mark(h, s, P_PACK));
    CU(h, launch_pack_bits(d_Y, words, N_LDPC, h->mc.N_bch, n_frames, s));
    TRY(prof_mark(h, s, P_BCH));
    CU(h, launch_bch_decode(bch_args(h, first_frame, n_frames, d_cwd), h->bch, s));
    TRY(prof_mark(h, s, P_UNPACK));
    CU(h, launch_unpack_bits(words, d_VK, N_LDPC, h->mc.K_bch, n_frames, s));
    TRY(prof_mark(h, s, P_END));
    return DVBS2_OK;
}

int
run_chain(dvbs2_cuda_handle* h, cudaStream_t s, int first_frame, const float* d_sigma, const float* d_sym,
          int8_t* d_cwd_ldpc, int8_t* d_cwd_bch, int32_t* d_VK, int32_t* d_nit, int n_frames, uint8_t* d_packed = nullptr,
          int descramble = 0, bool pl_frames = false, float* d_sigma_est = nullptr, int iq_fmt = DVBS2_IQ_F32,
          float iq_scale = 1.f)
{
    if (n_frames == 0) return DVBS2_OK;
    void* var = (char*)h->var.p + (size_t)first_frame * N_LDPC * h->elt;
    TRY(prof_mark(h, s, P_DEMAP));
    if (d_sigma_est)
    { // Estimator_DVBS2 in front of the demapper: sigma per frame from the frame's own payload symbols
        CU(h, launch_estimate(demap_args(h, nullptr, d_sym, n_frames, pl_frames, iq_fmt, iq_scale), h->mc.K_bch, d_sigma_est,
                              nullptr, nullptr, s));
        d_sigma = d_sigma_est;
    }
    const int nf = call_nf(h, n_frames);
    if (call_onchip(h, n_frames))
    { // frame-major LLRs (natural order) into the var buffer, which the on-chip kernel reads as its input
        CU(h, launch_demap_frames(demap_args(h, d_sigma, d_sym, n_frames, pl_frames, iq_fmt, iq_scale), (float*)var, 1, s));
        TRY(prof_mark(h, s, P_LDPC));
        TRY(run_ldpc_onchip(h, s, first_frame, (const float*)var, d_cwd_ldpc, d_nit, n_frames));
    }
    else
    {
        CU(h, launch_demap_to_var(demap_args(h, d_sigma, d_sym, n_frames, pl_frames, iq_fmt, iq_scale), var, h->cfg.precision, nf, s));
        TRY(prof_mark(h, s, P_LDPC));
        TRY(run_ldpc_core(h, s, first_frame, n_frames, d_cwd_ldpc, d_nit, nf));
    }
    TRY(prof_mark(h, s, P_BCH));
    CU(h, launch_bch_decode(bch_args(h, first_frame, n_frames, d_cwd_bch), h->bch, s));
    TRY(prof_mark(h, s, P_UNPACK));
    const uint32_t* words = (uint32_t*)h->out_hd.p + (size_t)(first_frame / 32) * N_LDPC;
    if (d_packed)
        CU(h, launch_unpack_bytes(words, descramble ? h->d_scr_bytes : nullptr, d_packed, N_LDPC, h->mc.K_bch, n_frames, s));
    else
        CU(h, launch_unpack_bits(words, d_VK, N_LDPC, h->mc.K_bch, n_frames, s, descramble ? h->d_scr_bytes : nullptr));
    TRY(prof_mark(h, s, P_END));
    return DVBS2_OK;
}

// int8 LLRs (compact input socket) -> int16 decoder -> BCH -> bits / packed bytes
int
run_llr8_chain(dvbs2_cuda_handle* h, cudaStream_t s, int first_frame, const int8_t* d_llr, int8_t* d_cwd_ldpc, int8_t* d_cwd_bch,
               int32_t* d_VK, uint8_t* d_packed, int descramble, int32_t* d_nit, int n_frames)
{
    if (n_frames == 0) return DVBS2_OK;
    void* var = (char*)h->var.p + (size_t)first_frame * N_LDPC * h->elt;
    const int nf = call_nf(h, n_frames);
    TRY(prof_mark(h, s, P_PACK));
    CU(h, launch_pack_llr8_as_i16(d_llr, var, n_frames, nf, s));
    TRY(prof_mark(h, s, P_LDPC));
    TRY(run_ldpc_core(h, s, first_frame, n_frames, d_cwd_ldpc, d_nit, nf));
    TRY(prof_mark(h, s, P_BCH));
    CU(h, launch_bch_decode(bch_args(h, first_frame, n_frames, d_cwd_bch), h->bch, s));
    TRY(prof_mark(h, s, P_UNPACK));
    const uint32_t* words = (uint32_t*)h->out_hd.p + (size_t)(first_frame / 32) * N_LDPC;
    if (d_packed)
        CU(h, launch_unpack_bytes(words, descramble ? h->d_scr_bytes : nullptr, d_packed, N_LDPC, h->mc.K_bch, n_frames, s));
    else
        CU(h, launch_unpack_bits(words, d_VK, N_LDPC, h->mc.K_bch, n_frames, s, descramble ? h->d_scr_bytes : nullptr));
    TRY(prof_mark(h, s, P_END));
    return DVBS2_OK;
}

// waits for every pipeline stream (also on the error path of the chunked host entry points: copies already queued on the
// other streams must not outlive the call, the caller may free or reuse its buffers as soon as we return)
int
sync_pipes(dvbs2_cuda_handle* h)
{
    int rc = DVBS2_OK;
    for (int i = 0; i < N_PIPE; i++)
    {
        const cudaError_t e = cudaStreamSynchronize(h->pipe[i]);
        if (e != cudaSuccess && rc == DVBS2_OK) rc = fail_cuda(h, e, "cudaStreamSynchronize(pipe)");
    }
    return rc;
}

// transmitter state: tables + payload/codeword words for max_frames
int
tx_init(dvbs2_cuda_handle* h)
{
    if (h->tx_src.p) return DVBS2_OK;
    std::vector<uint32_t> scr, gm(192, 0u);
    try
    {
        const std::vector<uint8_t> g = build_bch_generator(build_bch_tables());
        if ((int)g.size() - 1 != h->mc.N_bch - h->mc.K_bch) return DVBS2_ERR_INVALID;
        for (size_t k = 0; k + 1 < g.size(); k++)
            gm[k] = g[k] ? 0xffffffffu : 0u;
        for (uint8_t b : build_bb_scrambler(h->mc.K_bch))
            scr.push_back(b ? 0xffffffffu : 0u);
    }
    catch (const std::exception& e)
    {
        h->last_error = e.what();
        return DVBS2_ERR_INVALID;
    }
    const size_t G = (size_t)h->max_groups;
    TRY(ensure(h, h->tx_scr, scr.size() * 4));
    TRY(ensure(h, h->tx_gmask, gm.size() * 4));
    TRY(ensure(h, h->mon, 3 * sizeof(unsigned long long)));
    TRY(ensure(h, h->tx_cw, G * N_LDPC * 4));
    CU(h, cudaMemcpy(h->tx_scr.p, scr.data(), scr.size() * 4, cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->tx_gmask.p, gm.data(), gm.size() * 4, cudaMemcpyHostToDevice));
    CU(h, cudaMemset(h->mon.p, 0, h->mon.bytes));
    TRY(ensure(h, h->tx_src, G * h->mc.K_bch * 4)); // last: its presence marks the transmitter as initialised
    return DVBS2_OK;
}

TxLaunch
tx_args(dvbs2_cuda_handle* h, int n_frames)
{
    TxLaunch t{};
    t.src = (uint32_t*)h->tx_src.p, t.cw = (uint32_t*)h->tx_cw.p;
    t.scr = (const uint32_t*)h->tx_scr.p, t.gmask = (const uint32_t*)h->tx_gmask.p, t.idx = h->d_idx;
    t.K_bch = h->mc.K_bch, t.N_bch = h->mc.N_bch, t.d_info = h->lt.d_info, t.d_pad = h->lt.d_pad;
    t.n_frames = n_frames, t.n_groups = (n_frames + 31) / 32;
    return t;
}

// chunking of a host batch over the pipeline streams: chunks are whole groups so their device regions are disjoint
int
chunk_frames(const dvbs2_cuda_handle* h, int n_frames)
{
    // one chunk per pipeline stream: the LDPC kernel's duration hardly depends on the chunk size (a warp walks its group
    // through every check either way), so fewer, larger chunks keep more SMs busy while copies of the other chunks — and
    // of other module clones — proceed
    // ... but not below four groups per SM: a launch with fewer groups runs one or two warps per SM at the latency of a
    // full one (measured: 12 chunks of 148 groups each took as long as the PCIe traffic they were meant to hide), so calls
    // of up to n_sms * 4 groups are not split at all and rely on the other module clones for overlap
    const int want = h->cfg.host_pipeline_chunks > 0 ? std::min(h->cfg.host_pipeline_chunks, N_PIPE) : N_PIPE;
    int per = (n_frames + want - 1) / want;
    if (h->cfg.host_pipeline_chunks <= 0) per = std::max(per, h->n_sms * 4 * 32 * h->nf);
    per = (per + 255) / 256 * 256; // whole decoder groups for every precision
    return per;
}
} // namespace

extern "C"
{
    void dvbs2_cuda_default_cfg(dvbs2_cuda_cfg* cfg)
    {
        if (!cfg) return;
        memset(cfg, 0, sizeof(*cfg));
        cfg->modcod = DVBS2_QPSK_S_8_9;
        cfg->n_ite = 10;
        cfg->norm = 1.0f;
        cfg->precision = DVBS2_PREC_F32;
        cfg->early_stop = 1;
        cfg->max_frames = 1024;
        cfg->device = 0;
        cfg->quant_bits = 6;
        cfg->quant_dec = 2;
    }

    int dvbs2_cuda_device_count(void)
    {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess)
        {
            cudaGetLastError();
            return 0;
        }
        return n;
    }

    int dvbs2_cuda_abi_version(void) { return DVBS2_CUDA_ABI_VERSION; }

    int dvbs2_cuda_abi_check(int caller_version, size_t sizeof_cfg, size_t sizeof_stats)
    {
        static_assert(sizeof(dvbs2_cuda_cfg) == 18 * sizeof(int), "dvbs2_cuda_cfg changed size: bump DVBS2_CUDA_ABI_VERSION");
        static_assert(sizeof(dvbs2_cuda_stats) == (6 + 33) * sizeof(uint64_t), "dvbs2_cuda_stats changed size: bump DVBS2_CUDA_ABI_VERSION");
        return (caller_version == DVBS2_CUDA_ABI_VERSION && sizeof_cfg == sizeof(dvbs2_cuda_cfg) && sizeof_stats == sizeof(dvbs2_cuda_stats))
                 ? DVBS2_OK
                 : DVBS2_ERR_INVALID;
    }

    const char* dvbs2_cuda_strerror(int status)
    {
        switch (status)
        {
            case DVBS2_OK: return "success";
            case DVBS2_ERR_INVALID: return "invalid argument or unsupported configuration";
            case DVBS2_ERR_CUDA: return "CUDA runtime error";
            case DVBS2_ERR_TOO_MANY: return "n_frames exceeds the handle's max_frames";
            case DVBS2_ERR_NOMEM: return "out of memory";
        }
        return "unknown status";
    }

    const char* dvbs2_cuda_last_error(const dvbs2_cuda_handle* h) { return h ? h->last_error.c_str() : ""; }

    int dvbs2_cuda_destroy(dvbs2_cuda_handle* h)
    {
        if (!h) return DVBS2_OK;
        cudaSetDevice(h->cfg.device);
        for (DevBuf* b : { &h->var, &h->cn, &h->hd, &h->out_hd, &h->rem, &h->nz, &h->cwd_ldpc, &h->cwd_bch, &h->nit,
                           &h->stats, &h->surv_mask, &h->surv_iters, &h->dst_base, &h->src_of_dst, &h->slot_frame, &h->slot_iters,
                           &h->final_loc, &h->n_surv,
                           &h->oc_cn, &h->oc_pv, &h->oc_bits, &h->in_a, &h->out_a, &h->sigma, &h->tx_src, &h->tx_cw, &h->tx_scr,
                           &h->tx_gmask, &h->mon })
            if (b->p) cudaFree(b->p);
        for (void* p : { (void*)h->d_idx, (void*)h->d_idx32, (void*)h->d_last, (void*)h->d_queue, (void*)h->d_alpha_to,
                         (void*)h->d_index_of, (void*)h->d_cstl, (void*)h->d_scr_bytes, (void*)h->d_idxT })
            if (p) cudaFree(p);
        for (int i = 0; i < N_PIPE; i++)
        {
            if (h->pipe[i]) cudaStreamDestroy(h->pipe[i]);
        }
        if (h->stream) cudaStreamDestroy(h->stream);
        for (auto& m : h->marks)
            cudaEventDestroy(m.ev);
        if (h->stat_event) cudaEventDestroy(h->stat_event);
        if (h->stat_mirror) cudaFreeHost(h->stat_mirror);
        delete h;
        return DVBS2_OK;
    }

    int dvbs2_cuda_create(const dvbs2_cuda_cfg* cfg, dvbs2_cuda_handle** out)
    {
        if (!cfg || !out) return DVBS2_ERR_INVALID;
        *out = nullptr;
        ModcodParams mc;
        if (!modcod_params(cfg->modcod, mc)) return DVBS2_ERR_INVALID;
        if (cfg->n_ite < 1 || cfg->max_frames < 1 || cfg->precision < 0 || cfg->precision > 2) return DVBS2_ERR_INVALID;
        if (!(cfg->norm > 0.f) || cfg->norm > 1.f) return DVBS2_ERR_INVALID;
        int k8 = 0;
        for (int k = 1; k <= 8; k++)
            if (cfg->norm == 0.125f * (float)k) k8 = k;
        // integer decoders only know k/8 factors (Update_rule_NMS.hxx:14-39 throws otherwise)
        if (cfg->precision != DVBS2_PREC_F32 && k8 == 0) return DVBS2_ERR_INVALID;
        if (cfg->precision != DVBS2_PREC_F32)
        {
            const int type_bits = cfg->precision == DVBS2_PREC_I16 ? 16 : 8;
            if (cfg->quant_dec <= 0 || cfg->quant_bits < 2 || cfg->quant_dec > cfg->quant_bits || cfg->quant_bits > type_bits)
                return DVBS2_ERR_INVALID; // Quantizer_pow2.cpp:81-122
        }
        if (cfg->host_pipeline_chunks < 0 || cfg->host_pipeline_chunks > N_PIPE) return DVBS2_ERR_INVALID;
        if (cfg->ldpc_onchip < -1 || cfg->ldpc_onchip > 1 || (cfg->ldpc_onchip == 1 && cfg->precision != DVBS2_PREC_F32))
            return DVBS2_ERR_INVALID;
        if (cfg->reserved[0]) return DVBS2_ERR_INVALID;
        if (cfg->ldpc_frames_per_lane != 0 &&
            (cfg->precision != DVBS2_PREC_F32 || (cfg->ldpc_frames_per_lane != 1 && cfg->ldpc_frames_per_lane != 2)))
            return DVBS2_ERR_INVALID;
        int ndev = dvbs2_cuda_device_count();
        if (ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) return DVBS2_ERR_CUDA; // no CPU fallback

        dvbs2_cuda_handle* h = new (std::nothrow) dvbs2_cuda_handle();
        if (!h) return DVBS2_ERR_NOMEM;
        h->cfg = *cfg;
        h->mc = mc;
        h->k8 = k8;
        h->elt = cfg->precision == DVBS2_PREC_F32 ? 4 : cfg->precision == DVBS2_PREC_I16 ? 2 : 1;
        h->nf = ldpc_frames_per_lane(cfg->precision);
        h->coop_max = cfg->ldpc_coop_max_frames < 0 ? 0 : cfg->ldpc_coop_max_frames > 0 ? cfg->ldpc_coop_max_frames : 16384;
        h->nf_cfg = cfg->ldpc_frames_per_lane; // float only (validated above): 1 or 2 in round 1; 0 = per call from the batch size
        h->cn_div = cfg->precision == DVBS2_PREC_I16 ? 2 : 1;
        h->state_words = ldpc_state_words(mc.d_info);
        h->cn_bytes = (size_t)4 * h->state_words / h->cn_div;
        h->max_frames = (cfg->max_frames + 32 * h->nf - 1) / (32 * h->nf) * (32 * h->nf);
        h->max_groups = h->max_frames / 32;

        auto bail = [&](int rc) {
            dvbs2_cuda_destroy(h);
            return rc;
        };
#define CUB(call)                                                                                                      \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess) return bail(fail_cuda(h, e__, #call));                                                 \
    } while (0)

        CUB(cudaSetDevice(cfg->device));
        BchTables bt;
        std::vector<float> cstl;
        try
        {
            h->lt = build_ldpc_tables(mc.N_bch);
            bt = build_bch_tables();
            cstl = build_constellation(mc.bps);
        }
        catch (const std::exception& e)
        {
            h->last_error = e.what();
            return bail(DVBS2_ERR_INVALID);
        }
        if (h->lt.d_info != mc.d_info || h->lt.M != mc.M || !bch_minpoly_matches(bt)) return bail(DVBS2_ERR_INVALID);

        CUB(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        for (int i = 0; i < N_PIPE; i++)
        {
            CUB(cudaStreamCreateWithFlags(&h->pipe[i], cudaStreamNonBlocking));
        }
        CUB(cudaMalloc(&h->d_idx, h->lt.idx.size() * sizeof(uint16_t)));
        CUB(cudaMemcpy(h->d_idx, h->lt.idx.data(), h->lt.idx.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        CUB(cudaMalloc(&h->d_idx32, h->lt.idx32.size() * sizeof(uint32_t)));
        CUB(cudaMemcpy(h->d_idx32, h->lt.idx32.data(), h->lt.idx32.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUB(cudaMalloc(&h->d_last, h->lt.last.size() * sizeof(uint32_t)));
        CUB(cudaMemcpy(h->d_last, h->lt.last.data(), h->lt.last.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUB(cudaMalloc(&h->d_queue, 64 * sizeof(int)));
        CUB(cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, cfg->device));
        // the lane = frame decoder prefetches one check ahead: legal only while no two checks that close share an information
        // variable; the cooperative kernel's worker pipeline runs 3 * COOP_WORKERS checks deep (ldpc_kernel.cu) and is
        // switched off for a table that is denser than that
        if (h->lt.min_reuse_distance < 4) return bail(DVBS2_ERR_INVALID);
        if (h->lt.min_reuse_distance < ldpc_coop_min_reuse()) h->coop_max = 0;
        CUB(cudaMalloc(&h->d_alpha_to, bt.alpha_to.size() * sizeof(uint16_t)));
        CUB(cudaMemcpy(h->d_alpha_to, bt.alpha_to.data(), bt.alpha_to.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        CUB(cudaMalloc(&h->d_index_of, bt.index_of.size() * sizeof(int16_t)));
        CUB(cudaMemcpy(h->d_index_of, bt.index_of.data(), bt.index_of.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
        CUB(cudaMalloc(&h->d_cstl, cstl.size() * sizeof(float)));
        CUB(cudaMemcpy(h->d_cstl, cstl.data(), cstl.size() * sizeof(float), cudaMemcpyHostToDevice));
        {
            std::vector<uint8_t> sb(mc.K_bch / 8, 0);
            const std::vector<uint8_t> bits = build_bb_scrambler(mc.K_bch);
            for (int i = 0; i < mc.K_bch; i++)
                sb[i / 8] |= (uint8_t)((bits[i] & 1) << (i % 8));
            CUB(cudaMalloc(&h->d_scr_bytes, sb.size()));
            CUB(cudaMemcpy(h->d_scr_bytes, sb.data(), sb.size(), cudaMemcpyHostToDevice));
        }
        h->bch.alpha_to = h->d_alpha_to, h->bch.index_of = h->d_index_of;
        memcpy(h->bch.minpoly, bt.minpoly, sizeof(bt.minpoly));
        h->bch_seg.N = mc.N_bch;
        bch_segment_table(h->bch_seg, bt.minpoly);
        memcpy(h->bch.poly_of, bt.poly_of, sizeof(bt.poly_of));
        memcpy(h->bch.synd_mat, bt.synd_mat, sizeof(bt.synd_mat));

        const size_t G = (size_t)h->max_groups, F = (size_t)h->max_frames;
        // multi-round decoding: rounds / hand-over threshold / later-round group width (0 = defaults)
        h->rounds = cfg->no_compaction ? 1 : (cfg->compaction_rounds > 0 ? std::min(cfg->compaction_rounds, MAX_ROUNDS) : 3);
        h->auto_rounds = !cfg->no_compaction && cfg->compaction_rounds <= 0;
        h->exit_eighths = cfg->compaction_exit_eighths > 0 ? std::min(cfg->compaction_exit_eighths, 7) : 4;
        h->nf_later = ldpc_later_frames_per_lane(cfg->precision, cfg->compaction_nf_later);
        h->adaptive = !cfg->no_compaction && cfg->compaction_rounds <= 0 && cfg->compaction_exit_eighths <= 0 &&
                      cfg->compaction_nf_later <= 0 && cfg->precision != DVBS2_PREC_I8 && cfg->early_stop;
        h->layout_eighths = !h->adaptive ? h->exit_eighths : (cfg->precision == DVBS2_PREC_F32 ? 6 : 7);
        if (h->adaptive)
        {
            CUB(cudaHostAlloc((void**)&h->stat_mirror, 3 * sizeof(unsigned long long), cudaHostAllocDefault));
            h->stat_mirror[0] = h->stat_mirror[1] = h->stat_mirror[2] = 0;
            CUB(cudaEventCreateWithFlags(&h->stat_event, cudaEventDisableTiming));
        }
        h->ext_cap = h->rounds > 1 ? (int)region_b_offset(h, (h->max_frames + 255) / 256 * 256) + 512 : 0;
        h->ext_slots = h->rounds > 1 ? (int)ext_slot_offset(h, (h->max_frames + 255) / 256 * 256) + 256 * MAX_ROUNDS : 0;
        const size_t FE = F + h->ext_cap;                    // var / cn slots
        const size_t FS = F + h->ext_slots, GS = FS / 32;    // hd / out_hd / bookkeeping slots
        int rc;
        if ((rc = ensure(h, h->var, FE * N_LDPC * h->elt)) || (rc = ensure(h, h->cn, std::max(FE * mc.M * h->cn_bytes,
                                          // small calls of the cooperative kernel: its own (wider for int16) state per frame
                                          std::min(F, (size_t)std::max(h->coop_max, 0) + 128) * mc.M *
                                            (cfg->precision == DVBS2_PREC_I8 ? 0 : ldpc_coop_state_bytes(cfg->precision, mc.d_info))))) ||
            (rc = ensure(h, h->hd, GS * N_LDPC * 4)) || (rc = ensure(h, h->out_hd, GS * N_LDPC * 4)) ||
            (rc = ensure(h, h->surv_mask, GS * 4 + 4)) || (rc = ensure(h, h->surv_iters, FS * 4 + 4)) ||
            (rc = ensure(h, h->dst_base, GS * 4 + 4)) || (rc = ensure(h, h->src_of_dst, FS * 4 + 4)) ||
            (rc = ensure(h, h->slot_frame, FS * 4 + 4)) || (rc = ensure(h, h->slot_iters, FS * 4 + 4)) ||
            (rc = ensure(h, h->final_loc, F * 4 + 4)) || (rc = ensure(h, h->n_surv, 64 * sizeof(int))) ||
            (rc = ensure(h, h->rem, G * BCH_NPOLY * GF_M * 4)) || (rc = ensure(h, h->nz, G * 4)) ||
            (rc = ensure(h, h->cwd_ldpc, F)) || (rc = ensure(h, h->cwd_bch, F)) || (rc = ensure(h, h->nit, F * 4)) ||
            (rc = ensure(h, h->stats, STAT_COUNT * sizeof(unsigned long long))))
            return bail(rc);
        // on-chip LDPC kernel (float): legal while two neighbouring windows of checks share no information variable
        h->onchip = cfg->precision == DVBS2_PREC_F32 ? cfg->ldpc_onchip : -1;
        if (h->lt.min_reuse_distance < 2 * ONCHIP_WINDOW)
        {
            if (h->onchip == 1) return bail(DVBS2_ERR_INVALID);
            h->onchip = -1;
        }
        if (h->onchip >= 0)
        {
            h->oc_F = ldpc_onchip_frames_per_cta(mc.d_info);
            const std::vector<uint32_t> it = build_onchip_index(h->lt);
            CUB(cudaMalloc(&h->d_idxT, it.size() * sizeof(uint32_t)));
            CUB(cudaMemcpy(h->d_idxT, it.data(), it.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            const size_t slots = (size_t)h->n_sms * h->oc_F * (N_PIPE + 1);
            if ((rc = ensure(h, h->oc_cn, slots * mc.M * sizeof(uint4))) || (rc = ensure(h, h->oc_pv, slots * mc.M * sizeof(float))) ||
                (rc = ensure(h, h->oc_bits, F * ((mc.N_bch + 31) / 32) * sizeof(uint32_t))))
                return bail(rc);
        }
        CUB(cudaMemset(h->stats.p, 0, h->stats.bytes));
        CUB(cudaMemset(h->out_hd.p, 0, h->out_hd.bytes));
        CUB(cudaDeviceSynchronize());
#undef CUB
        *out = h;
        return DVBS2_OK;
    }

    int dvbs2_cuda_get_info(const dvbs2_cuda_handle* h, dvbs2_cuda_info* info)
    {
        if (!h || !info) return DVBS2_ERR_INVALID;
        info->K_bch = h->mc.K_bch, info->N_bch = h->mc.N_bch, info->K_ldpc = h->mc.N_bch, info->N_ldpc = N_LDPC;
        info->bps = h->mc.bps, info->n_sym = h->mc.n_sym, info->M_ldpc = h->mc.M, info->E_ldpc = h->lt.E;
        info->max_frames = h->max_frames;
        return DVBS2_OK;
    }

    void* dvbs2_cuda_stream(dvbs2_cuda_handle* h) { return h ? (void*)h->stream : nullptr; }

    int dvbs2_cuda_sync(dvbs2_cuda_handle* h)
    {
        if (!h) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        CU(h, cudaStreamSynchronize(h->stream));
        for (int i = 0; i < N_PIPE; i++)
            CU(h, cudaStreamSynchronize(h->pipe[i]));
        return DVBS2_OK;
    }

    int dvbs2_cuda_get_stats(dvbs2_cuda_handle* h, dvbs2_cuda_stats* st)
    {
        if (!h || !st) return DVBS2_ERR_INVALID;
        TRY(dvbs2_cuda_sync(h));
        unsigned long long v[STAT_COUNT];
        CU(h, cudaMemcpy(v, h->stats.p, sizeof(v), cudaMemcpyDeviceToHost));
        st->frames = v[STAT_FRAMES], st->ldpc_iterations = v[STAT_LDPC_ITERS], st->ldpc_failures = v[STAT_LDPC_FAIL];
        st->bch_nonzero = v[STAT_BCH_NONZERO], st->bch_failures = v[STAT_BCH_FAIL];
        st->kernel_launches = h->launches;
        for (int i = 0; i < LDPC_HIST_BINS; i++)
            st->iter_hist[i] = v[STAT_ITER_HIST + i];
        return DVBS2_OK;
    }

    int dvbs2_cuda_reset_stats(dvbs2_cuda_handle* h)
    {
        if (!h) return DVBS2_ERR_INVALID;
        TRY(dvbs2_cuda_sync(h));
        CU(h, cudaMemset(h->stats.p, 0, h->stats.bytes));
        h->seen_frames = h->seen_iters = h->seen_fails = 0; // the schedule statistics restart with the counters (everything queued
        h->stat_pending = false;                            // has finished: dvbs2_cuda_sync above); what was learnt so far stays
        if (h->stat_mirror) h->stat_mirror[0] = h->stat_mirror[1] = h->stat_mirror[2] = 0;
        h->launches = 0;
        if (h->mon.p) CU(h, cudaMemset(h->mon.p, 0, h->mon.bytes));
        TRY(prof_collect(h));
        for (int i = 0; i < P_COUNT; i++)
            h->t_ms[i] = 0., h->t_calls[i] = 0;
        return DVBS2_OK;
    }

    int dvbs2_cuda_set_profiling(dvbs2_cuda_handle* h, int enable)
    {
        if (!h) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        if (!enable) TRY(prof_collect(h));
        h->profiling = enable != 0;
        return DVBS2_OK;
    }

    int dvbs2_cuda_get_timers(dvbs2_cuda_handle* h, dvbs2_cuda_timers* t)
    {
        if (!h || !t) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        TRY(prof_collect(h));
        t->demap_ms = h->t_ms[P_DEMAP], t->ldpc_ms = h->t_ms[P_LDPC], t->bch_ms = h->t_ms[P_BCH];
        t->unpack_ms = h->t_ms[P_UNPACK], t->pack_ms = h->t_ms[P_PACK];
        t->demap_calls = h->t_calls[P_DEMAP], t->ldpc_calls = h->t_calls[P_LDPC], t->bch_calls = h->t_calls[P_BCH];
        t->unpack_calls = h->t_calls[P_UNPACK], t->pack_calls = h->t_calls[P_PACK];
        return DVBS2_OK;
    }

    int dvbs2_cuda_host_register(dvbs2_cuda_handle* h, void* ptr, size_t bytes)
    {
        if (!h || !ptr) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        CU(h, cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
        return DVBS2_OK;
    }
    int dvbs2_cuda_host_unregister(dvbs2_cuda_handle* h, void* ptr)
    {
        if (!h || !ptr) return DVBS2_ERR_INVALID;
        CU(h, cudaHostUnregister(ptr));
        return DVBS2_OK;
    }

    // ------------------------------------------------------------------------------------------ device-pointer API
    int dvbs2_cuda_demodulate_dev(dvbs2_cuda_handle* h, const float* d_CP, const float* d_Y_N1, float* d_Y_N2,
                                  int deinterleave, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!d_CP || !d_Y_N1 || !d_Y_N2) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        CU(h, launch_demap_frames(demap_args(h, d_CP, d_Y_N1, n_frames), d_Y_N2, deinterleave, h->stream));
        return DVBS2_OK;
    }

    int dvbs2_cuda_ldpc_decode_siho_dev(dvbs2_cuda_handle* h, const void* d_Y_N, int8_t* d_CWD, int32_t* d_V_K,
                                        int32_t* d_n_iter, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!d_Y_N || !d_V_K) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        return run_ldpc(h, h->stream, 0, d_Y_N, d_CWD, d_V_K, d_n_iter, n_frames);
    }

    int dvbs2_cuda_bch_decode_hiho_dev(dvbs2_cuda_handle* h, const int32_t* d_Y_N, int8_t* d_CWD, int32_t* d_V_K,
                                       int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!d_Y_N || !d_V_K) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        return run_bch(h, h->stream, 0, d_Y_N, d_CWD, d_V_K, n_frames);
    }

    int dvbs2_cuda_chain_dev(dvbs2_cuda_handle* h, const float* d_CP, const float* d_Y_N1, int8_t* d_CWD_ldpc,
                             int8_t* d_CWD_bch, int32_t* d_V_K, int32_t* d_n_iter, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!d_CP || !d_Y_N1 || !d_V_K) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        return run_chain(h, h->stream, 0, d_CP, d_Y_N1, d_CWD_ldpc, d_CWD_bch, d_V_K, d_n_iter, n_frames);
    }

    int dvbs2_cuda_chain_packed_dev(dvbs2_cuda_handle* h, const float* d_CP, const float* d_Y_N1, int8_t* d_CWD_ldpc,
                                    int8_t* d_CWD_bch, uint8_t* d_V_K_packed, int32_t* d_n_iter, int descramble, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!d_CP || !d_Y_N1 || !d_V_K_packed) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        return run_chain(h, h->stream, 0, d_CP, d_Y_N1, d_CWD_ldpc, d_CWD_bch, nullptr, d_n_iter, n_frames, d_V_K_packed,
                         descramble);
    }

    // ------------------------------------------------------------------------------------------ host-pointer API
    int dvbs2_cuda_demodulate(dvbs2_cuda_handle* h, const float* CP, const float* Y_N1, float* Y_N2, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!CP || !Y_N1 || !Y_N2) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        const size_t symf = (size_t)2 * h->mc.n_sym * sizeof(float), llrf = (size_t)N_LDPC * sizeof(float);
        TRY(ensure(h, h->in_a, symf * n_frames));
        TRY(ensure(h, h->out_a, llrf * n_frames));
        TRY(ensure(h, h->sigma, sizeof(float) * n_frames));
        cudaStream_t s = h->stream;
        CU(h, cudaMemcpyAsync(h->sigma.p, CP, sizeof(float) * n_frames, cudaMemcpyHostToDevice, s));
        CU(h, cudaMemcpyAsync(h->in_a.p, Y_N1, symf * n_frames, cudaMemcpyHostToDevice, s));
        CU(h, launch_demap_frames(demap_args(h, (float*)h->sigma.p, (float*)h->in_a.p, n_frames), (float*)h->out_a.p, 0, s));
        CU(h, cudaMemcpyAsync(Y_N2, h->out_a.p, llrf * n_frames, cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
        return DVBS2_OK;
    }

    int dvbs2_cuda_deinterleave(dvbs2_cuda_handle* h, const float* itl, float* nat, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!itl || !nat) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        const size_t llrf = (size_t)N_LDPC * sizeof(float);
        TRY(ensure(h, h->in_a, llrf * n_frames));
        TRY(ensure(h, h->out_a, llrf * n_frames));
        cudaStream_t s = h->stream;
        CU(h, cudaMemcpyAsync(h->in_a.p, itl, llrf * n_frames, cudaMemcpyHostToDevice, s));
        CU(h, launch_deinterleave((float*)h->in_a.p, (float*)h->out_a.p, n_frames, h->mc.n_sym, h->mc.itl_cols,
                                  h->mc.itl_rev, s));
        CU(h, cudaMemcpyAsync(nat, h->out_a.p, llrf * n_frames, cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
        return DVBS2_OK;
    }

    int dvbs2_cuda_quantize(dvbs2_cuda_handle* h, const float* in, void* out, size_t n)
    {
        if (!h || !in || !out || h->cfg.precision == DVBS2_PREC_F32) return DVBS2_ERR_INVALID;
        if (n == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        TRY(ensure(h, h->in_a, n * sizeof(float)));
        TRY(ensure(h, h->out_a, n * h->elt));
        cudaStream_t s = h->stream;
        CU(h, cudaMemcpyAsync(h->in_a.p, in, n * sizeof(float), cudaMemcpyHostToDevice, s));
        CU(h, launch_quantize((float*)h->in_a.p, h->out_a.p, h->cfg.precision, n, h->cfg.quant_dec, h->cfg.quant_bits, s));
        CU(h, cudaMemcpyAsync(out, h->out_a.p, n * h->elt, cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
        return DVBS2_OK;
    }

    int dvbs2_cuda_ldpc_decode_siho(dvbs2_cuda_handle* h, const void* Y_N, int8_t* CWD, int32_t* V_K, int32_t* n_iter,
                                    int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!Y_N || !V_K) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        NvtxRange nvtx_range("dvbs2_cuda_ldpc_decode_siho");
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        const size_t inf = (size_t)N_LDPC * h->elt, outf = (size_t)h->mc.N_bch * sizeof(int32_t);
        TRY(ensure(h, h->in_a, inf * n_frames));
        TRY(ensure(h, h->out_a, outf * n_frames));
        const int per = chunk_frames(h, n_frames);
        int ci = 0, rc = DVBS2_OK;
        for (int f0 = 0; f0 < n_frames && rc == DVBS2_OK; f0 += per, ci++)
            rc = [&]() -> int {
                const int nf = std::min(per, n_frames - f0);
                cudaStream_t s = h->pipe[ci % N_PIPE];
                char* d_in = (char*)h->in_a.p + inf * f0;
                int32_t* d_out = (int32_t*)((char*)h->out_a.p + outf * f0);
                int8_t* d_cwd = (int8_t*)h->cwd_ldpc.p + f0;
                int32_t* d_nit = (int32_t*)h->nit.p + f0;
                CU(h, cudaMemcpyAsync(d_in, (const char*)Y_N + inf * f0, inf * nf, cudaMemcpyHostToDevice, s));
                TRY(run_ldpc(h, s, f0, d_in, d_cwd, d_out, d_nit, nf));
                CU(h, cudaMemcpyAsync((char*)V_K + outf * f0, d_out, outf * nf, cudaMemcpyDeviceToHost, s));
                if (CWD) CU(h, cudaMemcpyAsync(CWD + f0, d_cwd, nf, cudaMemcpyDeviceToHost, s));
                if (n_iter) CU(h, cudaMemcpyAsync(n_iter + f0, d_nit, sizeof(int32_t) * nf, cudaMemcpyDeviceToHost, s));
                return DVBS2_OK;
            }();
        const int rs = sync_pipes(h); // also after a failure: nothing queued may outlive the call
        return rc != DVBS2_OK ? rc : rs;
    }

    int dvbs2_cuda_bch_decode_hiho(dvbs2_cuda_handle* h, const int32_t* Y_N, int8_t* CWD, int32_t* V_K, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!Y_N || !V_K) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        NvtxRange nvtx_range("dvbs2_cuda_bch_decode_hiho");
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        const size_t inf = (size_t)h->mc.N_bch * sizeof(int32_t), outf = (size_t)h->mc.K_bch * sizeof(int32_t);
        TRY(ensure(h, h->in_a, inf * n_frames));
        TRY(ensure(h, h->out_a, outf * n_frames));
        const int per = chunk_frames(h, n_frames);
        int ci = 0, rc = DVBS2_OK;
        for (int f0 = 0; f0 < n_frames && rc == DVBS2_OK; f0 += per, ci++)
            rc = [&]() -> int {
                const int nf = std::min(per, n_frames - f0);
                cudaStream_t s = h->pipe[ci % N_PIPE];
                int32_t* d_in = (int32_t*)((char*)h->in_a.p + inf * f0);
                int32_t* d_out = (int32_t*)((char*)h->out_a.p + outf * f0);
                int8_t* d_cwd = (int8_t*)h->cwd_bch.p + f0;
                CU(h, cudaMemcpyAsync(d_in, (const char*)Y_N + inf * f0, inf * nf, cudaMemcpyHostToDevice, s));
                TRY(run_bch(h, s, f0, d_in, d_cwd, d_out, nf));
                CU(h, cudaMemcpyAsync((char*)V_K + outf * f0, d_out, outf * nf, cudaMemcpyDeviceToHost, s));
                if (CWD) CU(h, cudaMemcpyAsync(CWD + f0, d_cwd, nf, cudaMemcpyDeviceToHost, s));
                return DVBS2_OK;
            }();
        const int rs = sync_pipes(h);
        return rc != DVBS2_OK ? rc : rs;
    }

    // The fused stage on host buffers, chunked over the pipeline streams (H2D of chunk i+1 and D2H of chunk i-1 overlap the
    // kernels of chunk i).  Input: symbols in any dvbs2_iq_format (XFEC or whole PL frames; CP == nullptr: sigma from the
    // on-device Estimator_DVBS2, returned in SIG), or — llr8 — int8 LLRs for the int16 decoder.  Output: int32 per bit
    // (AFF3CT socket type) when packed == nullptr, else packed bytes.
    static int chain_host(dvbs2_cuda_handle* h, const float* CP, const void* Y_N1, int8_t* CWD_ldpc, int8_t* CWD_bch,
                          int32_t* V_K, uint8_t* packed, int descramble, int32_t* n_iter, int n_frames, bool pl_frames = false,
                          float* SIG = nullptr, int iq_fmt = DVBS2_IQ_F32, float iq_scale = 1.f, bool llr8 = false,
                          bool wait = true)
    {
        TRY(check_frames(h, n_frames));
        if (!Y_N1 || (V_K != nullptr) == (packed != nullptr)) return DVBS2_ERR_INVALID; // exactly one output format
        if (iq_fmt < DVBS2_IQ_F32 || iq_fmt > DVBS2_IQ_SC12 || !(iq_scale > 0.f)) return DVBS2_ERR_INVALID;
        if (llr8 && h->cfg.precision != DVBS2_PREC_I16) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        NvtxRange nvtx_range(llr8 ? "dvbs2_cuda_decode_llr8" : pl_frames ? "dvbs2_cuda_rx_chain" : "dvbs2_cuda_chain");
        CU(h, cudaSetDevice(h->cfg.device));
        if (h->async_pending)
        { // one call in flight per handle (its staging buffers are about to be reused): collect the previous one first
            h->async_pending = false;
            TRY(sync_pipes(h));
        }
        LaunchScope launch_scope(h);
        const size_t inf = llr8 ? (size_t)N_LDPC
                                : (size_t)(pl_frames ? pl_frame_symbols(h->mc.n_sym) : h->mc.n_sym) * iq_symbol_bytes(iq_fmt);
        const size_t outf = packed ? (size_t)h->mc.K_bch / 8 : (size_t)h->mc.K_bch * sizeof(int32_t);
        TRY(ensure(h, h->in_a, inf * n_frames));
        TRY(ensure(h, h->out_a, outf * n_frames));
        TRY(ensure(h, h->sigma, sizeof(float) * n_frames));
        const int per = chunk_frames(h, n_frames);
        char* const host_out = packed ? (char*)packed : (char*)V_K;
        int ci = 0, rc = DVBS2_OK;
        for (int f0 = 0; f0 < n_frames && rc == DVBS2_OK; f0 += per, ci++)
            rc = [&]() -> int {
                const int nf = std::min(per, n_frames - f0);
                cudaStream_t s = h->pipe[ci % N_PIPE];
                char* d_in = (char*)h->in_a.p + inf * f0;
                float* d_sig = (float*)h->sigma.p + f0;
                char* d_out = (char*)h->out_a.p + outf * f0;
                int8_t* d_c1 = (int8_t*)h->cwd_ldpc.p + f0;
                int8_t* d_c2 = (int8_t*)h->cwd_bch.p + f0;
                int32_t* d_nit = (int32_t*)h->nit.p + f0;
                if (CP) CU(h, cudaMemcpyAsync(d_sig, CP + f0, sizeof(float) * nf, cudaMemcpyHostToDevice, s));
                CU(h, cudaMemcpyAsync(d_in, (const char*)Y_N1 + inf * f0, inf * nf, cudaMemcpyHostToDevice, s));
                if (llr8)
                    TRY(run_llr8_chain(h, s, f0, (const int8_t*)d_in, d_c1, d_c2, packed ? nullptr : (int32_t*)d_out,
                                       packed ? (uint8_t*)d_out : nullptr, descramble, d_nit, nf));
                else
                    TRY(run_chain(h, s, f0, d_sig, (const float*)d_in, d_c1, d_c2, packed ? nullptr : (int32_t*)d_out, d_nit, nf,
                                  packed ? (uint8_t*)d_out : nullptr, descramble, pl_frames, CP ? nullptr : d_sig, iq_fmt,
                                  iq_scale));
                if (SIG) CU(h, cudaMemcpyAsync(SIG + f0, d_sig, sizeof(float) * nf, cudaMemcpyDeviceToHost, s));
                CU(h, cudaMemcpyAsync(host_out + outf * f0, d_out, outf * nf, cudaMemcpyDeviceToHost, s));
                if (CWD_ldpc) CU(h, cudaMemcpyAsync(CWD_ldpc + f0, d_c1, nf, cudaMemcpyDeviceToHost, s));
                if (CWD_bch) CU(h, cudaMemcpyAsync(CWD_bch + f0, d_c2, nf, cudaMemcpyDeviceToHost, s));
                if (n_iter) CU(h, cudaMemcpyAsync(n_iter + f0, d_nit, sizeof(int32_t) * nf, cudaMemcpyDeviceToHost, s));
                return DVBS2_OK;
            }();
        if (rc == DVBS2_OK && !wait)
        { // the *_async entry point: dvbs2_cuda_wait() collects the call
            h->async_pending = true;
            return DVBS2_OK;
        }
        const int rs = sync_pipes(h); // also after a failure: nothing queued may outlive the call
        return rc != DVBS2_OK ? rc : rs;
    }

    int dvbs2_cuda_chain(dvbs2_cuda_handle* h, const float* CP, const float* Y_N1, int8_t* CWD_ldpc, int8_t* CWD_bch,
                         int32_t* V_K, int32_t* n_iter, int n_frames)
    {
        if (!V_K || !CP) return DVBS2_ERR_INVALID;
        return chain_host(h, CP, Y_N1, CWD_ldpc, CWD_bch, V_K, nullptr, 0, n_iter, n_frames);
    }

    int dvbs2_cuda_chain_packed(dvbs2_cuda_handle* h, const float* CP, const float* Y_N1, int8_t* CWD_ldpc, int8_t* CWD_bch,
                                uint8_t* V_K_packed, int32_t* n_iter, int descramble, int n_frames)
    {
        if (!V_K_packed || !CP) return DVBS2_ERR_INVALID;
        return chain_host(h, CP, Y_N1, CWD_ldpc, CWD_bch, nullptr, V_K_packed, descramble, n_iter, n_frames);
    }

    // -------------------------------------------------------- compact socket formats
    size_t dvbs2_cuda_iq_frame_bytes(const dvbs2_cuda_handle* h, int iq_format, int pl_frames)
    {
        if (!h || iq_format < DVBS2_IQ_F32 || iq_format > DVBS2_IQ_SC12) return 0;
        return (size_t)(pl_frames ? pl_frame_symbols(h->mc.n_sym) : h->mc.n_sym) * iq_symbol_bytes(iq_format);
    }

    int dvbs2_cuda_chain_compact(dvbs2_cuda_handle* h, int iq_format, float iq_scale, const float* CP, const void* Y_N1,
                                 float* SIG, int8_t* CWD_ldpc, int8_t* CWD_bch, int32_t* V_K, uint8_t* V_K_packed,
                                 int descramble, int32_t* n_iter, int pl_frames, int n_frames)
    {
        if (!CP && !pl_frames) return DVBS2_ERR_INVALID; // the estimator works on PL frames' payload; XFEC callers pass sigma
        return chain_host(h, CP, Y_N1, CWD_ldpc, CWD_bch, V_K, V_K_packed, descramble, n_iter, n_frames, pl_frames != 0, SIG,
                          iq_format, iq_scale);
    }

    int dvbs2_cuda_chain_compact_async(dvbs2_cuda_handle* h, int iq_format, float iq_scale, const float* CP, const void* Y_N1,
                                       float* SIG, int8_t* CWD_ldpc, int8_t* CWD_bch, int32_t* V_K, uint8_t* V_K_packed,
                                       int descramble, int32_t* n_iter, int pl_frames, int n_frames)
    {
        if (!CP && !pl_frames) return DVBS2_ERR_INVALID;
        return chain_host(h, CP, Y_N1, CWD_ldpc, CWD_bch, V_K, V_K_packed, descramble, n_iter, n_frames, pl_frames != 0, SIG,
                          iq_format, iq_scale, false, /*wait*/ false);
    }

    int dvbs2_cuda_wait(dvbs2_cuda_handle* h)
    {
        if (!h) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        h->async_pending = false;
        return sync_pipes(h);
    }

    int dvbs2_cuda_chain_compact_dev(dvbs2_cuda_handle* h, int iq_format, float iq_scale, const float* d_CP,
                                     const void* d_Y_N1, float* d_SIG, int8_t* d_CWD_ldpc, int8_t* d_CWD_bch, int32_t* d_V_K,
                                     uint8_t* d_V_K_packed, int descramble, int32_t* d_n_iter, int pl_frames, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!d_Y_N1 || (d_V_K != nullptr) == (d_V_K_packed != nullptr) || (!d_CP && (!d_SIG || !pl_frames))) return DVBS2_ERR_INVALID;
        if (iq_format < DVBS2_IQ_F32 || iq_format > DVBS2_IQ_SC12 || !(iq_scale > 0.f)) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        return run_chain(h, h->stream, 0, d_CP, (const float*)d_Y_N1, d_CWD_ldpc, d_CWD_bch, d_V_K, d_n_iter, n_frames,
                         d_V_K_packed, descramble, pl_frames != 0, d_CP ? nullptr : d_SIG, iq_format, iq_scale);
    }

    int dvbs2_cuda_decode_llr8(dvbs2_cuda_handle* h, const int8_t* Y_N, int8_t* CWD_ldpc, int8_t* CWD_bch, int32_t* V_K,
                               uint8_t* V_K_packed, int descramble, int32_t* n_iter, int n_frames)
    {
        return chain_host(h, nullptr, Y_N, CWD_ldpc, CWD_bch, V_K, V_K_packed, descramble, n_iter, n_frames, false, nullptr,
                          DVBS2_IQ_F32, 1.f, true);
    }

    int dvbs2_cuda_decode_llr8_dev(dvbs2_cuda_handle* h, const int8_t* d_Y_N, int8_t* d_CWD_ldpc, int8_t* d_CWD_bch,
                                   int32_t* d_V_K, uint8_t* d_V_K_packed, int descramble, int32_t* d_n_iter, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!d_Y_N || (d_V_K != nullptr) == (d_V_K_packed != nullptr) || h->cfg.precision != DVBS2_PREC_I16) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        return run_llr8_chain(h, h->stream, 0, d_Y_N, d_CWD_ldpc, d_CWD_bch, d_V_K, d_V_K_packed, descramble, d_n_iter, n_frames);
    }

    // -------------------------------------------------------- receiver stage from PL frames
    int dvbs2_cuda_pl_frame_size(const dvbs2_cuda_handle* h)
    {
        return h ? 2 * pl_frame_symbols(h->mc.n_sym) : DVBS2_ERR_INVALID;
    }

    int dvbs2_cuda_remove_plh(dvbs2_cuda_handle* h, const float* Y_N1, float* Y_N2, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!Y_N1 || !Y_N2) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        const size_t inf = (size_t)2 * pl_frame_symbols(h->mc.n_sym) * sizeof(float), outf = (size_t)2 * h->mc.n_sym * sizeof(float);
        TRY(ensure(h, h->in_a, inf * n_frames));
        TRY(ensure(h, h->out_a, outf * n_frames));
        cudaStream_t s = h->stream;
        CU(h, cudaMemcpyAsync(h->in_a.p, Y_N1, inf * n_frames, cudaMemcpyHostToDevice, s));
        CU(h, launch_remove_plh((const float*)h->in_a.p, (float*)h->out_a.p, h->mc.n_sym, n_frames, s));
        CU(h, cudaMemcpyAsync(Y_N2, h->out_a.p, outf * n_frames, cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
        return DVBS2_OK;
    }

    int dvbs2_cuda_remove_plh_dev(dvbs2_cuda_handle* h, const float* d_Y_N1, float* d_Y_N2, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!d_Y_N1 || !d_Y_N2) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        CU(h, launch_remove_plh(d_Y_N1, d_Y_N2, h->mc.n_sym, n_frames, h->stream));
        return DVBS2_OK;
    }

    int dvbs2_cuda_estimate(dvbs2_cuda_handle* h, const float* X_N, float* SIG, float* Eb_N0, float* Es_N0, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!X_N || !SIG) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        const size_t inf = (size_t)2 * h->mc.n_sym * sizeof(float);
        TRY(ensure(h, h->in_a, inf * n_frames));
        TRY(ensure(h, h->out_a, 3 * sizeof(float) * (size_t)n_frames));
        cudaStream_t s = h->stream;
        float* d_sig = (float*)h->out_a.p;
        float *d_eb = d_sig + n_frames, *d_es = d_sig + 2 * (size_t)n_frames;
        CU(h, cudaMemcpyAsync(h->in_a.p, X_N, inf * n_frames, cudaMemcpyHostToDevice, s));
        CU(h, launch_estimate(demap_args(h, nullptr, (const float*)h->in_a.p, n_frames), h->mc.K_bch, d_sig, d_eb, d_es, s));
        CU(h, cudaMemcpyAsync(SIG, d_sig, sizeof(float) * n_frames, cudaMemcpyDeviceToHost, s));
        if (Eb_N0) CU(h, cudaMemcpyAsync(Eb_N0, d_eb, sizeof(float) * n_frames, cudaMemcpyDeviceToHost, s));
        if (Es_N0) CU(h, cudaMemcpyAsync(Es_N0, d_es, sizeof(float) * n_frames, cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
        return DVBS2_OK;
    }

    int dvbs2_cuda_estimate_dev(dvbs2_cuda_handle* h, const float* d_X_N, float* d_SIG, float* d_Eb_N0, float* d_Es_N0,
                                int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!d_X_N || !d_SIG) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        CU(h, launch_estimate(demap_args(h, nullptr, d_X_N, n_frames), h->mc.K_bch, d_SIG, d_Eb_N0, d_Es_N0, h->stream));
        return DVBS2_OK;
    }

    int dvbs2_cuda_rx_chain(dvbs2_cuda_handle* h, const float* PL_frames, const float* CP, float* SIG, int8_t* CWD_ldpc,
                            int8_t* CWD_bch, int32_t* V_K, uint8_t* V_K_packed, int descramble, int32_t* n_iter, int n_frames)
    {
        return chain_host(h, CP, PL_frames, CWD_ldpc, CWD_bch, V_K, V_K_packed, descramble, n_iter, n_frames, true, SIG);
    }

    int dvbs2_cuda_rx_chain_dev(dvbs2_cuda_handle* h, const float* d_PL_frames, const float* d_CP, float* d_SIG,
                                int8_t* d_CWD_ldpc, int8_t* d_CWD_bch, int32_t* d_V_K, uint8_t* d_V_K_packed, int descramble,
                                int32_t* d_n_iter, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!d_PL_frames || (d_V_K != nullptr) == (d_V_K_packed != nullptr) || (!d_CP && !d_SIG)) return DVBS2_ERR_INVALID;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        return run_chain(h, h->stream, 0, d_CP, d_PL_frames, d_CWD_ldpc, d_CWD_bch, d_V_K, d_n_iter, n_frames, d_V_K_packed,
                         descramble, true, d_CP ? nullptr : d_SIG);
    }

    // ------------------------------------------------------------------------------- transmitter / channel / monitor
    int dvbs2_cuda_tx_set_payload(dvbs2_cuda_handle* h, const int32_t* U, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!U) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        TRY(tx_init(h));
        const size_t bytes = (size_t)n_frames * h->mc.K_bch * sizeof(int32_t);
        TRY(ensure(h, h->in_a, bytes));
        cudaStream_t s = h->stream;
        CU(h, cudaMemcpyAsync(h->in_a.p, U, bytes, cudaMemcpyHostToDevice, s));
        CU(h, launch_pack_bits((const int32_t*)h->in_a.p, (uint32_t*)h->tx_src.p, h->mc.K_bch, h->mc.K_bch, n_frames, s));
        h->tx_encoded_groups = 0;
        CU(h, cudaStreamSynchronize(s));
        return DVBS2_OK;
    }

    int dvbs2_cuda_tx_random_payload(dvbs2_cuda_handle* h, uint64_t seed, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        TRY(tx_init(h));
        CU(h, launch_tx_random_payload((uint32_t*)h->tx_src.p, h->mc.K_bch, (n_frames + 31) / 32, seed, h->stream));
        h->tx_encoded_groups = 0;
        return DVBS2_OK;
    }

    int dvbs2_cuda_tx_modulate_awgn_dev(dvbs2_cuda_handle* h, float sigma, uint64_t seed, float* d_CP, float* d_Y_N1,
                                        int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!d_CP || !d_Y_N1 || !(sigma >= 0.f) || !h->tx_src.p) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        const TxLaunch t = tx_args(h, n_frames);
        if (h->tx_encoded_groups < t.n_groups)
        { // a later call may ask for more frames of the same payload than the one that triggered the encoder
            CU(h, launch_tx_encode(t, h->stream));
            h->tx_encoded_groups = t.n_groups;
        }
        CU(h, launch_tx_modulate_awgn(t, h->d_cstl, d_Y_N1, d_CP, h->mc.bps, h->mc.n_sym, h->mc.itl_cols, h->mc.itl_rev,
                                      sigma, seed, h->stream));
        return DVBS2_OK;
    }

    int dvbs2_cuda_tx_get_bits(dvbs2_cuda_handle* h, int stage, int32_t* out, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!out || stage < 0 || stage > 3 || !h->tx_src.p) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        if (stage > 0 && h->tx_encoded_groups < (n_frames + 31) / 32)
        {
            CU(h, launch_tx_encode(tx_args(h, n_frames), h->stream));
            h->tx_encoded_groups = (n_frames + 31) / 32;
        }
        const int len = stage <= 1 ? h->mc.K_bch : stage == 2 ? h->mc.N_bch : N_LDPC;
        const uint32_t* words = stage == 0 ? (const uint32_t*)h->tx_src.p : (const uint32_t*)h->tx_cw.p;
        const int stride = stage == 0 ? h->mc.K_bch : N_LDPC;
        const size_t bytes = (size_t)n_frames * len * sizeof(int32_t);
        TRY(ensure(h, h->out_a, bytes));
        CU(h, launch_unpack_bits(words, (int32_t*)h->out_a.p, stride, len, n_frames, h->stream));
        CU(h, cudaMemcpyAsync(out, h->out_a.p, bytes, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        return DVBS2_OK;
    }

    int dvbs2_cuda_monitor_check(dvbs2_cuda_handle* h, int n_frames)
    {
        TRY(check_frames(h, n_frames));
        if (!h->tx_src.p) return DVBS2_ERR_INVALID;
        if (n_frames == 0) return DVBS2_OK;
        CU(h, cudaSetDevice(h->cfg.device));
        LaunchScope launch_scope(h);
        CU(h, launch_monitor((const uint32_t*)h->out_hd.p, N_LDPC, tx_args(h, n_frames), (unsigned long long*)h->mon.p,
                             h->stream));
        return DVBS2_OK;
    }

    int dvbs2_cuda_monitor_get(dvbs2_cuda_handle* h, dvbs2_cuda_monitor* m)
    {
        if (!h || !m) return DVBS2_ERR_INVALID;
        m->n_frames = m->n_bit_errors = m->n_frame_errors = 0;
        if (!h->mon.p) return DVBS2_OK;
        TRY(dvbs2_cuda_sync(h));
        unsigned long long v[3];
        CU(h, cudaMemcpy(v, h->mon.p, sizeof(v), cudaMemcpyDeviceToHost));
        m->n_frames = v[0], m->n_bit_errors = v[1], m->n_frame_errors = v[2];
        return DVBS2_OK;
    }
}
