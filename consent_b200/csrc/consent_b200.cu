// consent_b200.cu — C ABI (include/consent_b200.h) of the B200-native CONSENT per-window correction path.
//
// Host side = the batched equivalent of the reference's per-window call
//   computeConsensusReadCorrection / computeConsensusAssemblyPolishing   (src/correctionMSA.cpp:29-70)
// as driven by processRead (src/CONSENT-correction.cpp:34-44): upload a batch of piles, run every stage as
// CUDA kernels over chunks of windows, hand back consensus + solid k-mers + status per window.
// There is no CPU implementation behind this ABI: without a CUDA device cg_create fails (CG_ERR_NO_DEVICE).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "consent_b200.h"
#include "cg_common.cuh"
#include "k_prep.cuh"
#include "k_index.cuh"
#include "k_chain.cuh"
#include "k_split.cuh"
#include "k_poa.cuh"
#include "k_poa2.cuh"
#ifndef CG_EMU
#include <nvtx3/nvToolsExt.h>          // header-only NVTX 3: ranges per ABI call and per stage / lane (nsys, ncu --nvtx)
struct CgNvtxRange {
    explicit CgNvtxRange(const char* name) { nvtxRangePushA(name); }
    ~CgNvtxRange() { nvtxRangePop(); }
};
#define CG_NVTX(name) CgNvtxRange cg_nvtx_range_##__LINE__(name)
#define CG_NVTX_PUSH(name) nvtxRangePushA(name)
#define CG_NVTX_POP() nvtxRangePop()
#else
#define CG_NVTX(name)
#define CG_NVTX_PUSH(name)
#define CG_NVTX_POP()
#endif
#include "k_polish.cuh"
#include "k_reanchor.cuh"
#include "k_extract.cuh"
#include "k_ingest.cuh"

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes, bool keep = false, cudaStream_t st = nullptr) {
        if (bytes <= cap) return cudaSuccess;
        size_t ncap = bytes + bytes / 4 + 256;
        void* np = nullptr;
        cudaError_t e = cudaMalloc(&np, ncap);
        if (e != cudaSuccess) return e;
        if (keep && p && cap) {
            e = cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFree(np); return e; }
        }
        if (p) cudaFree(p);
        p = np; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

struct ChunkPlan {
    u32 w0, nwin;
    u64 pword_base, nwords;
    u64 solid_tot, slot_tot, pos_tot, reg_tot, arena_tot;
    u32 max_tk, max_n;
    u64 max_occ;                        // largest pile (bases: an upper bound of its k-mer occurrences)
};

struct PoaTier {
    u32 vcap = 0, ecap = 0, lcap = 0;   // graph arrays in global memory (0: the kernel keeps them in shared memory)
    u64 hcap = 0;                       // score matrix + alignment in global memory (0: shared memory)
    u32 warps = 0;
    DevBuf mem, desc;
    bool ready = false;
};

// Host copy of the results of one batch, in pinned memory (so the per-chunk D2H copies overlap the kernels of the
// next chunk).  Pooled per handle: cg_free_results() hands the buffers back, the next call reuses them.
struct HostResults {
    u64* cons_off = nullptr; u64* solid_off = nullptr; u8* status = nullptr;
    char* cons = nullptr; u32* sk = nullptr; u32* sc = nullptr;
    size_t capW = 0, capC = 0, capS = 0;
    struct cg_handle* owner = nullptr;
    bool in_use = false, orphan = false;
    u64 gen = 0;            // cg_handle::run_gen when these results were produced (re-anchoring reuses the device copies)
};

// One timed stage = a pair of events on the lane's stream; collected after the final sync of cg_run.
struct StageSpan { int stage; cudaEvent_t a, b; };
// One timed kernel launch (or a short run of launches of one kind): events on the stream the kernel runs on.
struct KernelSpan { int kind; cudaEvent_t a, b; u32 launches; };

// A lane = everything one chunk needs while it is in flight: its stream(s), workspaces and pinned control words.
// Two lanes alternate over the chunks of a batch (chunk i on lane i % 2, each driven by its own host thread), so the
// tail of one chunk's POA — a few long jobs on a few warps — runs under the next chunk's kernels.
#ifndef CG_EMU
#define CG_NLANES 2
#else
#define CG_NLANES 1          // the SIMT emulator runs launches synchronously on the calling thread
#endif
struct Lane {
    cudaStream_t stream = nullptr;       // the bulk of a chunk: pack .. first pass of the POA tiers
    cudaStream_t s_tail = nullptr;       // the rest of it (re-queued POA jobs, polish, gather), high priority: it is latency, not work
    cudaStream_t s_poa[3] = {nullptr, nullptr, nullptr};   // tiers G and W1 run beside C1
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr}, ev_gather = nullptr, ev_end = nullptr, ev_tail = nullptr;
    bool tail_recorded = false;
    DevBuf pwords, ptags, win, offs, solid_k, solid_c, slot_tpos, slot_kmer, anchors, chain, rel, pos, regions, arena, fin, visited;
    DevBuf jobs_s, jobs_m, jobs_3, jobs_w, jobs_ws, jobs_r, jobs_x, ctl, off_fin, out_off;
    DevBuf g_mem, w1_mem, w2_mem; // per-warp global scratch of the POA tiers G (matrix only), W1 and W2
    DevBuf idx_keys, idx_counts;  // k_index, k > 9: one open-addressing count table per SM
    u32* h_ctl = nullptr;        // pinned: flags + queue control + totals
    std::string err;
    int rc = 0;
    float stage_ms[CG_N_STAGES]{};
    u32 stage_launches[CG_N_STAGES]{};
    std::vector<StageSpan> spans;
    std::vector<KernelSpan> kspans;
    std::vector<cudaEvent_t> evpool;
    size_t pool_at = 0;
};

struct cg_handle {
    int device = 0;
    cg_params p{};
    Lane lane[CG_NLANES];
    int n_lanes = CG_NLANES;
    cudaStream_t s_h2d = nullptr;       // batch upload, chunk by chunk
    cudaStream_t s_d2h = nullptr;       // results download, chunk by chunk
    std::vector<cudaEvent_t> ev_h2d;    // [chunk] bases of the chunk are resident
    std::vector<cudaEvent_t> ev_bulk;   // [chunk] the bulk of the chunk's work (everything up to the POA tiers' first pass) is done
    size_t bulk_enqueued = 0;           // chunks whose ev_bulk has been recorded (guarded by commit_mu)
    bool h2d_pending = false;           // cg_correct_windows: the kernels of chunk i wait for ev_h2d[i]
    HostResults* stream_out = nullptr;  // cg_correct_windows: results are downloaded chunk by chunk into this
    std::mutex pool_mu;
    std::vector<HostResults*> pool;
    std::string err;
    int sms = 148;
    int smem_optin = 0;
    // options
    // Workspaces of one chunk (per lane).  Bigger chunks amortise every kernel's drain (persistent warps idle while the last jobs of a queue
    // finish) and the host round trips: 8 GB / 16 384 windows -> 16 GB / 25 000 measured +4.5 % on the config-3 shape (tools/chunk_sweep.py).
    size_t chunk_budget = (size_t)16 << 30;
    u32 chunk_max_windows = 25000;
    // batch (device) + host copies of the offsets for planning
    u32 W = 0;
    u64 n_seqs = 0, n_bases = 0;
    DevBuf d_bases, d_seq_off, d_wsb;
    std::vector<u32> h_wsb;             // [W + 1]
    std::vector<u64> h_wbase;           // [W + 1] seq_off[win_seq_begin[w]]
    std::vector<u32> h_tlen;            // [W] template length
    std::vector<ChunkPlan> chunks;
    bool uploaded = false, ran = false, planned_ramp = false;
    // POA tiers
    cg_kernel_stats kstats{};
    bool input_2bit = false;             // cg_upload / cg_correct_windows: `bases` is 2 bits per base (cg_pack_bases_2bit), unpacked on the device
    bool results_with_solid = true;      // false: the solid k-mer lists stay on the device (cg_reanchor_reads / cg_finish_reads read them there)
    DevBuf d_packed;
    bool store_resident = false;         // cg_set_read_store: the read store of cg_upload_piles stays on the device across batches
    u32 store_n = 0;
    std::mutex tier_mu;                  // the last-resort scratch is shared by the lanes
    PoaTier tier[3];                     // k_poa.cuh: global-memory tiers without an in-degree limit (the last resort); [0] unused
    u32 c1_warps = 0, g_warps = 0, w1_warps = 0, w2_warps = 0;   // resident warps of the k_poa2.cuh tiers
    // batch outputs (device, dense), appended chunk by chunk IN ORDER: commit_mu/commit_cv serialise the tail of run_chunk
    DevBuf o_cons, o_sk, o_sc, o_status, o_len, o_nsol;
    u64 o_cons_n = 0, o_solid_n = 0;
    std::mutex commit_mu;
    std::condition_variable commit_cv;
    size_t next_commit = 0;
    bool abort_run = false;
    // instrumentation
    float stage_ms[CG_N_STAGES]{};
    float run_ms = 0;             // whole cg_run (CUDA events: first launch of lane 0 to the last result of any lane)
    cudaEvent_t ev_run0 = nullptr;
    u32 stage_launches[CG_N_STAGES]{};
    cg_counters counters{};
    // re-anchoring (cg_reanchor_reads)
    u64 run_gen = 0;              // bumped by every upload: results of an older batch are no longer resident
    DevBuf ra_cons, ra_cons_off, ra_solid_off, ra_sk, ra_tpl, ra_tpl_off;           // copies, when the results are not resident
    DevBuf ra_rwb, ra_roff, ra_rbases, ra_wpos, ra_order, ra_head_off, ra_head, ra_len, ra_out_off, ra_out, ra_scratch, ra_ctl;
    float ra_ms = 0;
    u64 ra_cells = 0;
    // window extraction (cg_upload_piles): device arrays + what cg_download_windows hands back
    DevBuf ex_store, ex_store_off, ex_pile_read, ex_pile_qlen, ex_pile_ovb, ex_ov, ex_cov, ex_cov_off, ex_cap_off, ex_cap_beg, ex_cap_end,
           ex_nwin, ex_win_pile, ex_win_beg, ex_win_end, ex_slot_base, ex_slot_len, ex_slot_src, ex_slot_loc, ex_win_nseq, ex_win_nbytes,
           ex_win_base, ex_flags;
    bool ex_valid = false;              // the resident batch was produced by cg_upload_piles
    std::vector<u32> ex_rwb, ex_wpos, ex_wend, ex_pile_read_h;
    std::vector<u64> ex_store_off_h;
    u32 ex_ws = 0, ex_ovl = 0;
    float ex_ms = 0, ex_copy_ms = 0;
    u64 ex_bytes = 0;
    // PAF ingest (cg_ingest_paf) and post-filters (cg_finish_reads)
    DevBuf in_text, in_tile, in_nl, in_names, in_name_off, in_slots, in_rec, in_head, in_pfirst, in_plast, in_keep, in_scratch,
           in_pread, in_pqlen, in_ov, in_res, in_ctl, in_htile, ra_skip;
    float in_ms = 0, in_parse_ms = 0, fin_ms = 0;
    u64 in_bytes = 0;
};

struct HostWindowSet {
    std::vector<u32> wsb, rwb, wpos, wend;
    std::vector<u64> soff, roff;
    std::vector<char> bases, rbases;
};

// Host copy of the corrected reads (pinned).
struct HostCorrected { u64* off = nullptr; char* bases = nullptr; size_t off_cap = 0, bases_cap = 0; };
// One pair of pinned buffers is kept across calls (cudaMallocHost / cudaFreeHost of tens of MB per call cost 5-10 ms and spike).
static std::mutex g_hc_mu;
static HostCorrected* g_hc_spare = nullptr;
static HostCorrected* hc_acquire() {
    std::lock_guard<std::mutex> lk(g_hc_mu);
    HostCorrected* hc = g_hc_spare ? g_hc_spare : new HostCorrected();
    g_hc_spare = nullptr;
    return hc;
}
static void hc_destroy(HostCorrected* hc) { if (hc->off) cudaFreeHost(hc->off); if (hc->bases) cudaFreeHost(hc->bases); delete hc; }
static void hc_release(HostCorrected* hc) {
    {
        std::lock_guard<std::mutex> lk(g_hc_mu);
        if (!g_hc_spare) { g_hc_spare = hc; return; }
    }
    hc_destroy(hc);
}
template <class T> static bool hc_ensure(T*& p, size_t& cap, size_t bytes) {
    if (bytes <= cap) return true;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    if (cudaMallocHost(&p, want) != cudaSuccess) { p = nullptr; return false; }
    cap = want;
    return true;
}

namespace {

std::string g_create_err;

#define CK_TO(errstr, call)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            (errstr) = std::string(#call) + ": " + cudaGetErrorString(e__);                          \
            return e__ == cudaErrorMemoryAllocation ? CG_ERR_OUT_OF_MEMORY : CG_ERR_CUDA;            \
        }                                                                                            \
    } while (0)
#define CK(call) CK_TO(h->err, call)      /* single-threaded parts */
#define CKL(call) CK_TO(L.err, call)      /* inside a lane */

inline u64 round_up(u64 v, u64 m) { return (v + m - 1) / m * m; }

// ctl layout (u32 words, device): [0] flags, [1] upload-validation flags, [4 + 4t ..] queue t {front jobs, next, back jobs, capacity}:
// t = 0 tier C1, 1 tier G, 2 tier W1 (filled by k_split); re-queued: 3 -> G again, 4 -> W1 again, 5 -> W2, 6 -> k_poa tier 1,
// 7 -> k_poa tier 2
enum { CTL_FLAGS = 0, CTL_VFLAGS = 1, CTL_Q = 4, CTL_NQ = 10, CTL_WORDS = 48, HCTL_STAGE = 64, HCTL_VFLAGS = 120, HCTL_WORDS = 128 };
// offs layout (u64 arrays of nwin+1): solid, slot, pos, reg, arena
// out_off layout: cons_off[nwin+1], solid_off[nwin+1]

// Chunk plan from per-window summaries only (O(W) on the host): every capacity below is an upper bound of what
// k_plan computes on the device from the exact per-sequence lengths (occurrences <= bases).
// ramp: the first chunk is a quarter of the others, so that a pipelined call (cg_correct_windows) starts computing after a
// quarter of the first upload.
int plan_chunks(cg_handle* h, bool ramp = false) {
    h->chunks.clear();
    h->planned_ramp = ramp;
    const u32 k = h->p.mer_size;
    u32 w = 0;
    while (w < h->W) {
        ChunkPlan c{};
        c.w0 = w;
        c.pword_base = (h->h_wbase[w] >> 4) + h->h_wsb[w];
        size_t bytes = 0;
        const u32 max_windows = (ramp && w == 0 && h->W > h->chunk_max_windows) ? std::max<u32>(1u, h->chunk_max_windows / 4) : h->chunk_max_windows;
        while (w < h->W && c.nwin < max_windows) {
            const u32 s0 = h->h_wsb[w], s1 = h->h_wsb[w + 1];
            u32 N = s1 - s0;
            u64 nb = h->h_wbase[w + 1] - h->h_wbase[w];
            const u64 tlen = h->h_tlen[w];
            u64 tk = tlen >= k ? tlen - k + 1 : 0;
            if (N > CG_N_MAX || tlen > CG_LEN_MAX || tk > CG_TK_MAX) { N = 1; tk = 0; nb = tlen; }     // k_plan: CG_WINDOW_ERROR, template only
            const u64 nocc = nb;
            const u64 a_solid = round_up(nocc / h->p.solid_thresh, 4), a_slot = round_up(tk, 8), a_pos = round_up(tk * N, 8),
                      a_reg = tk + 2, a_arena = round_up(nb + N, 16);
            const u64 words = ((h->h_wbase[w + 1] >> 4) + s1) - ((h->h_wbase[w] >> 4) + s0);
            const size_t wbytes = a_solid * 8 + a_solid / 8 + 8 + a_slot * 14 + a_pos * 2 + a_reg * (sizeof(CgRegion) + 8) + a_arena +
                                  words * 8 + sizeof(CgWin) + 64 + 6 * tlen + 256;
            if (c.nwin > 0 && bytes + wbytes > h->chunk_budget) break;
            bytes += wbytes;
            c.solid_tot += a_solid; c.slot_tot += a_slot; c.pos_tot += a_pos; c.reg_tot += a_reg; c.arena_tot += a_arena;
            c.max_tk = std::max<u32>(c.max_tk, (u32)tk); c.max_n = std::max<u32>(c.max_n, N);
            c.max_occ = std::max<u64>(c.max_occ, nocc);
            c.nwin++; ++w;
        }
        c.nwords = ((h->h_wbase[w] >> 4) + h->h_wsb[w]) - c.pword_base;
        h->chunks.push_back(c);
    }
    return CG_OK;
}

int ensure_tier(Lane& L, PoaTier& T) {
    if (T.ready) return CG_OK;
    const u32 ncap = CG_N_MAX + 1;
    const u32 scap = T.vcap ? 2 * (T.ecap + 4 * T.vcap) : 0;
    const u32 alncap = T.vcap + T.lcap + 2;
    // per-warp layout (all sub-arrays 16-byte aligned)
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += round_up(bytes, 16); return at; };
    const size_t o_letter = take(T.vcap), o_in0 = take(T.vcap), o_nal = take(T.vcap), o_leader = take(T.vcap), o_marks = take(T.vcap),
                 o_check = take(T.vcap), o_nseq = take(2 * (size_t)T.vcap), o_aligned = take(6 * (size_t)T.vcap),
                 o_rank = take(2 * (size_t)T.vcap), o_r2n = take(2 * (size_t)T.vcap), o_ih = take(4 * (size_t)T.vcap),
                 o_it = take(4 * (size_t)T.vcap), o_rd = take(4 * (size_t)T.vcap), o_ep = take(2 * (size_t)T.ecap), o_en = take(4 * (size_t)T.ecap),
                 o_stack = take(4 * (size_t)scap), o_an = take(4 * (size_t)alncap), o_ap = take(4 * (size_t)alncap),
                 o_sr = take(2 * (size_t)ncap), o_ss = take(2 * (size_t)ncap), o_sl = take(2 * (size_t)ncap), o_H = take(2 * T.hcap);
    const size_t per_warp = round_up(o, 256);
    CKL(T.mem.ensure(per_warp * T.warps));
    CKL(T.desc.ensure(sizeof(CgPoaScratch) * T.warps));
    std::vector<CgPoaScratch> d(T.warps);
    for (u32 i = 0; i < T.warps; ++i) {
        u8* b = T.mem.as<u8>() + per_warp * i;
        CgPoaScratch& s = d[i];
        s.vcap = T.vcap; s.ecap = T.ecap; s.scap = scap; s.alncap = alncap; s.ncap = ncap; s.hcap = T.hcap;
        s.letter = b + o_letter; s.in0 = b + o_in0; s.nal = b + o_nal; s.leader = b + o_leader; s.marks = b + o_marks; s.check = b + o_check;
        s.nseq = (u16*)(b + o_nseq); s.aligned = (u16*)(b + o_aligned); s.rank_of = (u16*)(b + o_rank); s.r2n = (u16*)(b + o_r2n);
        s.in_head = (u32*)(b + o_ih); s.in_tail = (u32*)(b + o_it); s.rdesc = (u32*)(b + o_rd);
        s.e_pred = (u16*)(b + o_ep); s.e_next = (u32*)(b + o_en);
        s.stack = (u32*)(b + o_stack); s.aln_node = (i32*)(b + o_an); s.aln_pos = (i32*)(b + o_ap);
        s.seg_read = (u16*)(b + o_sr); s.seg_start = (u16*)(b + o_ss); s.seg_len = (u16*)(b + o_sl); s.H = (i16*)(b + o_H);
    }
    CKL(cudaMemcpyAsync(T.desc.p, d.data(), sizeof(CgPoaScratch) * T.warps, cudaMemcpyHostToDevice, L.stream));
    CKL(cudaStreamSynchronize(L.stream));
    T.ready = true;
    return CG_OK;
}

int stream_out_chunk(cg_handle* h, Lane& L, cudaStream_t st, const ChunkPlan& cp, u64 cons_n, u64 solid_n);

// Every stage of the path for the windows of chunk ci, on lane L.  The ordered tail (dense outputs appended after those
// of chunk ci - 1) waits for its turn.
int run_chunk(cg_handle* h, Lane& L, size_t ci) {
    const ChunkPlan& cp = h->chunks[ci];
    cudaStream_t st = L.stream;
    const u32 nwin = cp.nwin;
    if (L.tail_recorded) CKL(cudaStreamWaitEvent(st, L.ev_tail, 0));        // the lane's workspaces are free again
    auto ev = [&]() -> cudaEvent_t {
        if (L.pool_at == L.evpool.size()) { cudaEvent_t e; cudaEventCreate(&e); L.evpool.push_back(e); }
        return L.evpool[L.pool_at++];
    };
    static const char* const stage_names[CG_N_STAGES] = {"cg:pack", "cg:index", "cg:chain", "cg:split", "cg:poa", "cg:stitch", "cg:polish"};
    auto span_begin = [&](int stage) { CG_NVTX_PUSH(stage_names[stage]); StageSpan sp{stage, ev(), ev()}; cudaEventRecord(sp.a, st); L.spans.push_back(sp); };
    auto span_end = [&]() { cudaEventRecord(L.spans.back().b, st); CG_NVTX_POP(); };
    auto kbegin = [&](int kind, cudaStream_t ks, u32 n = 1) { KernelSpan sp{kind, ev(), ev(), n}; cudaEventRecord(sp.a, ks); L.kspans.push_back(sp); };
    auto kend = [&](cudaStream_t ks) { cudaEventRecord(L.kspans.back().b, ks); };

    // ---- workspaces
    CKL(L.pwords.ensure((cp.nwords + 16) * 4)); CKL(L.ptags.ensure((cp.nwords + 16) * 4));
    CKL(L.win.ensure(sizeof(CgWin) * nwin));
    CKL(L.offs.ensure(sizeof(u64) * 5 * (nwin + 1)));
    CKL(L.solid_k.ensure(cp.solid_tot * 4 + 16)); CKL(L.solid_c.ensure(cp.solid_tot * 4 + 16));
    CKL(L.slot_tpos.ensure(cp.slot_tot * 2 + 16)); CKL(L.slot_kmer.ensure(cp.slot_tot * 4 + 16));
    CKL(L.anchors.ensure(cp.slot_tot * 2 + 16)); CKL(L.chain.ensure(cp.slot_tot * 2 + 16)); CKL(L.rel.ensure(cp.slot_tot * 4 + 16));
    CKL(L.pos.ensure(cp.pos_tot * 2 + 16));
    CKL(L.regions.ensure(cp.reg_tot * sizeof(CgRegion) + 16));
    CKL(L.arena.ensure(cp.arena_tot + 16));
    CKL(L.visited.ensure((cp.solid_tot / 32 + nwin + 2) * 4));
    for (DevBuf* jb : {&L.jobs_s, &L.jobs_m, &L.jobs_3, &L.jobs_w, &L.jobs_ws, &L.jobs_r, &L.jobs_x}) CKL(jb->ensure(cp.reg_tot * sizeof(uint2) + 16));
    CKL(L.off_fin.ensure(sizeof(u64) * (nwin + 1)));
    CKL(L.out_off.ensure(sizeof(u64) * 2 * (nwin + 1)));

    CgChunk c{};
    c.bases = h->d_bases.as<char>(); c.seq_off = h->d_seq_off.as<u64>(); c.win_seq_begin = h->d_wsb.as<u32>();
    c.w0 = cp.w0; c.nwin = nwin;
    c.k = h->p.mer_size; c.solid = h->p.solid_thresh; c.common = h->p.common_kmers; c.min_anchors = h->p.min_anchors;
    c.pwords = L.pwords.as<u32>(); c.ptags = L.ptags.as<u32>(); c.pword_base = cp.pword_base;
    c.win = L.win.as<CgWin>();
    u64* offs = L.offs.as<u64>();
    c.off_solid = offs; c.off_slot = offs + (nwin + 1); c.off_pos = offs + 2 * (nwin + 1); c.off_reg = offs + 3 * (nwin + 1);
    c.off_arena = offs + 4 * (nwin + 1);
    c.solid_k = L.solid_k.as<u32>(); c.solid_c = L.solid_c.as<u32>();
    c.slot_tpos = L.slot_tpos.as<u16>(); c.slot_kmer = L.slot_kmer.as<u32>(); c.anchors = L.anchors.as<u16>();
    c.chain = L.chain.as<u16>(); c.rel = L.rel.as<u32>(); c.pos = L.pos.as<u16>();
    c.regions = L.regions.as<CgRegion>(); c.arena = L.arena.as<u8>(); c.fin = nullptr; c.visited = L.visited.as<u32>();
    u32* ctl = L.ctl.as<u32>();
    c.qctl = ctl + CTL_Q; c.jobs_s = L.jobs_s.as<uint2>(); c.jobs_m = L.jobs_m.as<uint2>(); c.jobs_w = L.jobs_w.as<uint2>();
    c.flags = ctl + CTL_FLAGS;
    c.counters = (CgCountersDev*)(ctl + CTL_WORDS);
    u64* off_fin = L.off_fin.as<u64>();
    u64* cons_off = L.out_off.as<u64>();
    u64* solid_off = cons_off + (nwin + 1);

    if (h->h2d_pending) CKL(cudaStreamWaitEvent(st, h->ev_h2d[ci], 0));
    // Stagger the lanes: the bulk of chunk ci starts when the bulk of chunk ci - 1 is done, so that the latency-bound rest of
    // a chunk (re-queued POA jobs, polish, the host round trips for sizes) always runs under the next chunk's bulk.
    if (ci > 0) {
        {
            std::unique_lock<std::mutex> lk(h->commit_mu);
            h->commit_cv.wait(lk, [&] { return h->bulk_enqueued >= ci || h->abort_run; });
            if (h->abort_run) return CG_ERR_STATE;
        }
        CKL(cudaStreamWaitEvent(st, h->ev_bulk[ci - 1], 0));
    }

    // ---- stage 0: plan, offsets, pack
    span_begin(CG_STAGE_PACK);
    {
        u32 qinit[4 * CTL_NQ] = {0};
        for (int t = 0; t < CTL_NQ; ++t) qinit[4 * t + 3] = (u32)cp.reg_tot;
        memcpy(L.h_ctl + HCTL_STAGE, qinit, sizeof qinit);              // pinned staging, consumed before the lane's next sync
        CKL(cudaMemcpyAsync(ctl + CTL_Q, L.h_ctl + HCTL_STAGE, sizeof qinit, cudaMemcpyHostToDevice, st));
        CKL(cudaMemsetAsync(ctl + CTL_FLAGS, 0, sizeof(u32), st));
    }
    CKL(cudaMemsetAsync(L.pwords.as<u32>() + cp.nwords, 0, 16 * 4, st));
    CKL(cudaMemsetAsync(L.ptags.as<u32>() + cp.nwords, 0xff, 16 * 4, st));
    kbegin(CG_K_PACK, st, 3);
    CG_LAUNCH(k_plan, (nwin + 127) / 128, 128, 0, st, c);
    CG_LAUNCH(k_scan, 5, 1024, 1024 * sizeof(u64), st, c.off_solid, c.off_slot, c.off_pos, c.off_reg, c.off_arena, nwin);
    CG_LAUNCH(k_pack, nwin, CG_PACK_THREADS, CG_PACK_SMEM_BYTES, st, c);
    kend(st);
    L.stage_launches[CG_STAGE_PACK] += 3;
    span_end();

    if (h->p.mer_size > CG_KMAX) {                          // hashed k-mer counting: tables of >= 2 x the deepest pile's occurrences
        u64 cap = 1024;
        while (cap < 2 * cp.max_occ) cap <<= 1;
        if (cap > (1ull << 26)) cap = 1ull << 26;           // deeper piles are flagged by the kernel (CG_WINDOW_ERROR)
#ifndef CG_EMU
        const u32 slots = 192;                              // indexed by %smid (B200: 148 SMs enabled, ids below 160)
#else
        const u32 slots = nwin;                             // the emulator has no SM ids: one table per window
#endif
        CKL(L.idx_keys.ensure((size_t)slots * cap * 4)); CKL(L.idx_counts.ensure((size_t)slots * cap * 4));
        c.idx_keys = L.idx_keys.as<u32>(); c.idx_counts = L.idx_counts.as<u32>(); c.idx_cap = (u32)cap; c.idx_slots = slots;
    }
    span_begin(CG_STAGE_INDEX);
    kbegin(CG_K_INDEX, st);
    CG_LAUNCH(k_index, nwin, CG_IDX_THREADS, CG_IDX_SMEM_BYTES, st, c);
    kend(st);
    L.stage_launches[CG_STAGE_INDEX] += 1;
    span_end();

    span_begin(CG_STAGE_CHAIN);
    const size_t smem_cap = (size_t)h->smem_optin;
    const size_t chain_full = std::min(cg_chain_smem(cp.max_tk, cp.max_n), smem_cap);     // windows that need more are flagged by the kernel
    const size_t chain_small = std::min(cg_chain_smem(std::min<u32>(cp.max_tk, 192u), cp.max_n), chain_full);
    kbegin(CG_K_CHAIN, st, chain_full > chain_small ? 2 : 1);
    CG_LAUNCH(k_chain, nwin, CG_CHAIN_THREADS, chain_small, st, c, (u32)chain_small, 0u);
    if (chain_full > chain_small) CG_LAUNCH(k_chain, nwin, CG_CHAIN_THREADS, chain_full, st, c, (u32)chain_full, 1u);
    kend(st);
    L.stage_launches[CG_STAGE_CHAIN] += chain_full > chain_small ? 2 : 1;
    span_end();

    span_begin(CG_STAGE_SPLIT);
    kbegin(CG_K_SPLIT, st);
    CG_LAUNCH(k_split, nwin, CG_SPLIT_THREADS, 0, st, c);
    kend(st);
    L.stage_launches[CG_STAGE_SPLIT] += 1;
    span_end();

    // ---- POA (k_poa2.cuh).  k_split routed every region to the tier its predicted size fits; the three tiers run side by
    // side (the wide, long-running jobs are launched first so that they overlap the bulk of the small ones).  What a tier
    // cannot hold after all is re-queued: C1 -> G -> W1 -> W2 -> k_poa.
    span_begin(CG_STAGE_POA);
    u32* q = ctl + CTL_Q;
    uint2* jobs_q3 = L.jobs_r.as<uint2>(); uint2* jobs_q4 = L.jobs_x.as<uint2>(); uint2* jobs_q5 = L.jobs_3.as<uint2>();
    uint2* jobs_q6 = c.jobs_s; uint2* jobs_q7 = c.jobs_m;                                    // re-used once their first life is over
    CKL(L.g_mem.ensure(CgPoa2Lay<CgPoa2GT>::scratch_per_warp * (size_t)h->g_warps));
    CKL(L.w1_mem.ensure(CgPoa2Lay<CgPoa2W1>::scratch_per_warp * (size_t)h->w1_warps));
    CKL(L.w2_mem.ensure(CgPoa2Lay<CgPoa2W2>::scratch_per_warp * (size_t)h->w2_warps));
#define CG_POA2_LAUNCH(TIER, mem, warps, stream, jin, qin, jout, qout)                                                          \
    kbegin(TIER::VCAP <= 128 ? CG_K_POA_C1 : TIER::VCAP <= 254 ? CG_K_POA_G : TIER::VCAP <= 1024 ? CG_K_POA_W1 : CG_K_POA_W2, stream); \
    CG_LAUNCH(k_poa2<TIER>, ((warps) + TIER::WARPS - 1) / TIER::WARPS, TIER::WARPS * 32, CgPoa2Lay<TIER>::cta_bytes, stream, c, \
              (mem), (warps), (const uint2*)(jin), q + 4 * (qin), (jout), q + 4 * (qout));                                      \
    kend(stream)
    // The three tiers side by side (the wide, long-running jobs are launched first so that they overlap the bulk of the small ones).
    // (Measured and dropped: G first and ONE wide launch afterwards over its own queue plus G's overflow — the wide tier is latency
    // bound and G's work hides inside it: 165 k -> 158 k windows/s at 20 sequences per window.)
    // The wide tier's few long jobs go longest (predicted) first: k_split.cuh: k_poa_sort_queue.  On the main stream, before the fork: the
    // wide launch has to reach the SMs ahead of the bulk tiers (sorted on its own stream it started behind them: 72 -> 81 ms per 16 384 windows).
    CG_LAUNCH(k_poa_sort_queue, 1, CG_QSORT_THREADS, (CG_QSORT_BUCKETS + 64) * sizeof(u32), st, c, (const uint2*)c.jobs_w, q + 4 * 2, L.jobs_ws.as<uint2>());
    CKL(cudaEventRecord(L.ev_fork, st));
    for (int i = 0; i < 2; ++i) CKL(cudaStreamWaitEvent(L.s_poa[i], L.ev_fork, 0));
    CG_POA2_LAUNCH(CgPoa2W1, L.w1_mem.as<u8>(), h->w1_warps, L.s_poa[1], L.jobs_ws.as<uint2>(), 2, jobs_q5, 5);
    CG_POA2_LAUNCH(CgPoa2GT, L.g_mem.as<u8>(), h->g_warps, L.s_poa[0], c.jobs_m, 1, jobs_q4, 4);
    CG_POA2_LAUNCH(CgPoa2C1, (u8*)nullptr, h->c1_warps, st, c.jobs_s, 0, jobs_q3, 3);
    for (int i = 0; i < 2; ++i) CKL(cudaEventRecord(L.ev_join[i], L.s_poa[i]));
    CKL(cudaStreamWaitEvent(st, L.ev_join[0], 0));
    CKL(cudaEventRecord(h->ev_bulk[ci], st));              // C1 and G are through: the next chunk's bulk may start (W1's few long jobs go on)
    {
        std::lock_guard<std::mutex> lk(h->commit_mu);
        h->bulk_enqueued = ci + 1;
    }
    h->commit_cv.notify_all();
    st = L.s_tail;                                          // from here on: the chunk's tail
    CKL(cudaStreamWaitEvent(st, h->ev_bulk[ci], 0));
    CG_POA2_LAUNCH(CgPoa2GT, L.g_mem.as<u8>(), h->g_warps, st, jobs_q3, 3, jobs_q4, 4);
    CKL(cudaStreamWaitEvent(st, L.ev_join[1], 0));
    CG_POA2_LAUNCH(CgPoa2W1, L.w1_mem.as<u8>(), h->w1_warps, st, jobs_q4, 4, jobs_q5, 5);
    CG_POA2_LAUNCH(CgPoa2W2, L.w2_mem.as<u8>(), h->w2_warps, st, jobs_q5, 5, jobs_q6, 6);
#undef CG_POA2_LAUNCH
    L.stage_launches[CG_STAGE_POA] += 6;
    span_end();

    // the last resort (in-degree > 8, > 4096 nodes, > 2048-base segments): only if something got that far (count on the host)
    CKL(cudaMemcpyAsync(L.h_ctl, ctl, CTL_WORDS * sizeof(u32), cudaMemcpyDeviceToHost, st));
    CKL(cudaStreamSynchronize(st));
    if (getenv("CG_DEBUG"))
        fprintf(stderr, "[consent_b200] chunk w0=%u nwin=%u POA jobs: C1 %u+%u, G %u+%u, W1 %u+%u | re-queued: ->G %u, ->W1 %u, ->W2 %u, ->k_poa %u\n",
                cp.w0, nwin, L.h_ctl[CTL_Q + 0], L.h_ctl[CTL_Q + 2], L.h_ctl[CTL_Q + 4], L.h_ctl[CTL_Q + 6], L.h_ctl[CTL_Q + 8], L.h_ctl[CTL_Q + 10],
                L.h_ctl[CTL_Q + 12], L.h_ctl[CTL_Q + 16], L.h_ctl[CTL_Q + 20], L.h_ctl[CTL_Q + 24]);
    if (L.h_ctl[CTL_Q + 24]) {
        std::lock_guard<std::mutex> lk(h->tier_mu);         // one lane at a time on the shared last-resort scratch
        uint2* q_in = jobs_q6; uint2* q_out = jobs_q7;
        for (int t = 1; t <= 2; ++t) {
            const u32 over = L.h_ctl[CTL_Q + 4 * (t + 5)];
            if (!over) break;
            if (h->tier[t].warps == 0) { L.err = "a POA job outgrew the largest enabled scratch tier"; return CG_ERR_CAPACITY; }
            { int rc = ensure_tier(L, h->tier[t]); if (rc) return rc; }
            span_begin(CG_STAGE_POA);
            kbegin(CG_K_POA_LAST, st);
            CG_LAUNCH(k_poa, (h->tier[t].warps + CG_POA_WARPS_PER_CTA - 1) / CG_POA_WARPS_PER_CTA, CG_POA_THREADS, 0, st, c,
                      h->tier[t].desc.as<CgPoaScratch>(), h->tier[t].warps, (const uint2*)q_in, q + 4 * (t + 5), t < 2 ? q_out : (uint2*)nullptr,
                      q + 4 * (t + 6));
            kend(st);
            L.stage_launches[CG_STAGE_POA] += 1;
            span_end();
            CKL(cudaMemcpyAsync(L.h_ctl, ctl, CTL_WORDS * sizeof(u32), cudaMemcpyDeviceToHost, st));
            CKL(cudaStreamSynchronize(st));
            std::swap(q_in, q_out);
        }
    }

    // ---- stitched lengths -> work slices
    span_begin(CG_STAGE_STITCH);
    kbegin(CG_K_OUT, st, 2);
    CG_LAUNCH(k_stitch_len, (nwin + 127) / 128, 128, 0, st, c, off_fin);
    CG_LAUNCH(k_scan, 1, 1024, 1024 * sizeof(u64), st, off_fin, (u64*)nullptr, (u64*)nullptr, (u64*)nullptr, (u64*)nullptr, nwin);
    kend(st);
    L.stage_launches[CG_STAGE_STITCH] += 2;
    span_end();
    u64 fin_total = 0;
    CKL(cudaMemcpyAsync(&fin_total, off_fin + nwin, sizeof(u64), cudaMemcpyDeviceToHost, st));
    CKL(cudaStreamSynchronize(st));
    CKL(L.fin.ensure(fin_total + 16));
    c.fin = L.fin.as<u8>();

    span_begin(CG_STAGE_POLISH);
    kbegin(CG_K_POLISH, st);
    CG_LAUNCH(k_polish, (nwin + CG_POLISH_WARPS_PER_CTA - 1) / CG_POLISH_WARPS_PER_CTA, CG_POLISH_THREADS, CG_POLISH_SMEM_BYTES, st, c, (const u64*)off_fin);
    kend(st);
    L.stage_launches[CG_STAGE_POLISH] += 1;
    span_end();

    span_begin(CG_STAGE_STITCH);
    kbegin(CG_K_OUT, st, 2);
    CG_LAUNCH(k_out_sizes, (nwin + 127) / 128, 128, 0, st, c, cons_off, solid_off);
    CG_LAUNCH(k_scan, 2, 1024, 1024 * sizeof(u64), st, cons_off, solid_off, (u64*)nullptr, (u64*)nullptr, (u64*)nullptr, nwin);
    kend(st);
    L.stage_launches[CG_STAGE_STITCH] += 2;
    span_end();
    u64 tot[2] = {0, 0};
    CKL(cudaMemcpyAsync(&tot[0], cons_off + nwin, sizeof(u64), cudaMemcpyDeviceToHost, st));
    CKL(cudaMemcpyAsync(&tot[1], solid_off + nwin, sizeof(u64), cudaMemcpyDeviceToHost, st));
    CKL(cudaMemcpyAsync(L.h_ctl, ctl, CTL_WORDS * sizeof(u32), cudaMemcpyDeviceToHost, st));
    CKL(cudaStreamSynchronize(st));
    if (L.h_ctl[CTL_FLAGS] & CG_FLAG_BAD_BASE) { L.err = "a base outside {A,C,G,T} in the batch"; return CG_ERR_BAD_BASE; }
    if (L.h_ctl[CTL_FLAGS] & CG_FLAG_CAPACITY) { L.err = "a per-window capacity limit of this build was exceeded"; return CG_ERR_CAPACITY; }
    if (L.h_ctl[CTL_FLAGS] & CG_FLAG_INTERNAL) { L.err = "internal invariant violated (anchor pair without common read)"; return CG_ERR_CUDA; }

    // ---- ordered tail: append this chunk's dense outputs after those of chunk ci - 1
    std::unique_lock<std::mutex> lk(h->commit_mu);
    h->commit_cv.wait(lk, [&] { return h->next_commit == ci || h->abort_run; });
    if (h->abort_run) return CG_ERR_STATE;                 // another lane failed; its error is the one reported
    {   // dense batch outputs: sized from the density of the chunks so far, so that they rarely move
        const double frac = (double)(cp.w0 + nwin) / (double)h->W;
        const u64 need_c = h->o_cons_n + tot[0] + 16, need_s = (h->o_solid_n + tot[1]) * 4 + 16;
        if (need_c > h->o_cons.cap || need_s > h->o_sk.cap) {
            CKL(cudaDeviceSynchronize());                   // gathers of earlier chunks and downloads may still use the old buffers
            const u64 est_c = (u64)((double)need_c / frac * 1.05) + 4096, est_s = (u64)((double)need_s / frac * 1.05) + 4096;
            CKL(h->o_cons.ensure(std::max(need_c, est_c), true, st));
            CKL(h->o_sk.ensure(std::max(need_s, est_s), true, st));
            CKL(h->o_sc.ensure(std::max(need_s, est_s), true, st));
        }
    }
    span_begin(CG_STAGE_STITCH);
    kbegin(CG_K_OUT, st);
    CG_LAUNCH(k_gather, nwin, 256, 0, st, c, (const u64*)off_fin, (const u64*)cons_off, (const u64*)solid_off,
              h->o_cons.as<u8>() + h->o_cons_n, h->o_sk.as<u32>() + h->o_solid_n, h->o_sc.as<u32>() + h->o_solid_n,
              h->o_status.as<u8>() + cp.w0, h->o_cons_n, h->o_solid_n, h->o_len.as<u64>() + cp.w0, h->o_nsol.as<u64>() + cp.w0);
    kend(st);
    L.stage_launches[CG_STAGE_STITCH] += 1;
    span_end();
    if (h->stream_out) { int rc = stream_out_chunk(h, L, st, cp, tot[0], tot[1]); if (rc) return rc; }
    CKL(cudaEventRecord(L.ev_tail, st));
    L.tail_recorded = true;
    h->o_cons_n += tot[0];
    h->o_solid_n += tot[1];
    h->next_commit = ci + 1;
    lk.unlock();
    h->commit_cv.notify_all();
    return CG_OK;
}

// ---- pinned result buffers
template <class T> cudaError_t pinned_ensure(T*& p, size_t& cap, size_t need, size_t keep) {
    if (need <= cap && p) return cudaSuccess;
    const size_t ncap = need + need / 8 + 64;
    T* np = nullptr;
    cudaError_t e = cudaMallocHost((void**)&np, ncap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p && keep) memcpy(np, p, keep * sizeof(T));
    if (p) cudaFreeHost(p);
    p = np; cap = ncap;
    return cudaSuccess;
}

void host_results_release(HostResults* r) {
    if (r->cons_off) cudaFreeHost(r->cons_off);
    if (r->solid_off) cudaFreeHost(r->solid_off);
    if (r->status) cudaFreeHost(r->status);
    if (r->cons) cudaFreeHost(r->cons);
    if (r->sk) cudaFreeHost(r->sk);
    if (r->sc) cudaFreeHost(r->sc);
    delete r;
}

// A free HostResults of the handle's pool (or a new one) with room for W windows.
int host_results_acquire(cg_handle* h, u32 W, HostResults** out) {
    HostResults* r = nullptr;
    {
        std::lock_guard<std::mutex> lk(h->pool_mu);
        for (HostResults* c : h->pool) if (!c->in_use) { r = c; break; }
        if (!r) { r = new HostResults(); r->owner = h; h->pool.push_back(r); }
        r->in_use = true;
    }
    size_t capW2 = r->capW, capW3 = r->capW;
    CK(pinned_ensure(r->cons_off, r->capW, (size_t)W + 1, 0));
    CK(pinned_ensure(r->solid_off, capW2, (size_t)W + 1, 0));
    CK(pinned_ensure(r->status, capW3, (size_t)W + 1, 0));
    r->capW = std::min(r->capW, std::min(capW2, capW3));
    *out = r;
    return CG_OK;
}

// Make room for `need_c` consensus bytes and `need_s` solid k-mers, keeping what has already been downloaded.
int host_results_reserve(cg_handle* h, std::string& err, HostResults* r, u64 need_c, u64 need_s, u64 keep_c, u64 keep_s) {
    if (need_c > r->capC || !r->cons) {
        CK_TO(err, cudaStreamSynchronize(h->s_d2h));
        CK_TO(err, pinned_ensure(r->cons, r->capC, (size_t)need_c, (size_t)keep_c));
    }
    if (need_s > r->capS || !r->sk) {
        CK_TO(err, cudaStreamSynchronize(h->s_d2h));
        size_t cap2 = r->capS;
        CK_TO(err, pinned_ensure(r->sk, r->capS, (size_t)need_s, (size_t)keep_s));
        CK_TO(err, pinned_ensure(r->sc, cap2, (size_t)need_s, (size_t)keep_s));
        r->capS = std::min(r->capS, cap2);
    }
    return CG_OK;
}

// Download the dense results of one chunk on the D2H stream while the next chunk computes.
int stream_out_chunk(cg_handle* h, Lane& L, cudaStream_t st, const ChunkPlan& cp, u64 cons_n, u64 solid_n) {
    HostResults* r = h->stream_out;
    const double frac = (double)(cp.w0 + cp.nwin) / (double)h->W;
    const u64 need_c = h->o_cons_n + cons_n + 1, need_s = h->o_solid_n + solid_n + 1;
    if (need_c > r->capC || need_s > r->capS || !r->cons || !r->sk) {
        const u64 est_c = (u64)((double)need_c / frac * 1.05) + 4096, est_s = (u64)((double)need_s / frac * 1.05) + 4096;
        int rc = host_results_reserve(h, L.err, r, std::max(need_c, est_c), std::max(need_s, est_s), h->o_cons_n, h->o_solid_n);
        if (rc) return rc;
    }
    CKL(cudaEventRecord(L.ev_gather, st));
    CKL(cudaStreamWaitEvent(h->s_d2h, L.ev_gather, 0));
    cudaStream_t sd = h->s_d2h;
    if (cons_n) CKL(cudaMemcpyAsync(r->cons + h->o_cons_n, h->o_cons.as<u8>() + h->o_cons_n, cons_n, cudaMemcpyDeviceToHost, sd));
    if (solid_n && h->results_with_solid) {
        CKL(cudaMemcpyAsync(r->sk + h->o_solid_n, h->o_sk.as<u32>() + h->o_solid_n, solid_n * 4, cudaMemcpyDeviceToHost, sd));
        CKL(cudaMemcpyAsync(r->sc + h->o_solid_n, h->o_sc.as<u32>() + h->o_solid_n, solid_n * 4, cudaMemcpyDeviceToHost, sd));
    }
    CKL(cudaMemcpyAsync(r->cons_off + cp.w0, h->o_len.as<u64>() + cp.w0, cp.nwin * sizeof(u64), cudaMemcpyDeviceToHost, sd));
    CKL(cudaMemcpyAsync(r->solid_off + cp.w0, h->o_nsol.as<u64>() + cp.w0, cp.nwin * sizeof(u64), cudaMemcpyDeviceToHost, sd));
    CKL(cudaMemcpyAsync(r->status + cp.w0, h->o_status.as<u8>() + cp.w0, cp.nwin, cudaMemcpyDeviceToHost, sd));
    return CG_OK;
}

}  // namespace

extern "C" {

int cg_abi_version(void) { return CG_ABI_VERSION; }

#ifdef CG_POA_TIMING
// debug builds only: per-job phase clocks of the POA tiers, as text
int cg_debug_dump_jobs(const char* path) {
    u32 n = 0;
    cudaMemcpyFromSymbol(&n, cg_dbg_njobs, sizeof n);
    n = std::min<u32>(n, 1u << 16);
    std::vector<CgJobTiming> J(n);
    if (n) cudaMemcpyFromSymbol(J.data(), cg_dbg_jobs, sizeof(CgJobTiming) * n);
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "tier w rg nseg V maxL t0 t1 dp maxtie traceback update splice dfs setup vote\n");
    for (const CgJobTiming& j : J) {
        fprintf(f, "%u %u %u %u %u %u %lld %lld", j.tier, j.w, j.rg, j.nseg, j.V, j.maxL, j.t0, j.t1);
        for (int q = 0; q < 8; ++q) fprintf(f, " %lld", j.ph[q]);
        fprintf(f, "\n");
    }
    fclose(f);
    u32 zero = 0;
    cudaMemcpyToSymbol(cg_dbg_njobs, &zero, sizeof zero);
    return (int)n;
}
#endif

int cg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* cg_last_error(const cg_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int cg_create(int device, const cg_params* params, cg_handle** out) {
    if (!params || !out) { g_create_err = "null argument"; return CG_ERR_INVALID_ARG; }
    *out = nullptr;
    if (params->mer_size < 2 || params->mer_size > 15 || params->solid_thresh < 1) { g_create_err = "mer_size must be 2..15, solid_thresh >= 1"; return CG_ERR_INVALID_ARG; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        g_create_err = "no usable CUDA device (this library has no CPU path)";
        return CG_ERR_NO_DEVICE;
    }
    cg_handle* h = new cg_handle();
    h->device = device; h->p = *params;
    if (cudaSetDevice(device) != cudaSuccess) { g_create_err = "cudaSetDevice failed"; delete h; return CG_ERR_NO_DEVICE; }
    cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if ((size_t)h->smem_optin < CG_IDX_SMEM_BYTES) {
        g_create_err = "device has too little shared memory per block for k_index (needs sm_100-class 227 KB)";
        delete h; return CG_ERR_NO_DEVICE;
    }
    bool ok = cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&h->ev_run0) == cudaSuccess;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    for (Lane& L : h->lane) {
        ok = ok && cudaStreamCreateWithPriority(&L.stream, cudaStreamNonBlocking, prio_lo) == cudaSuccess;
        ok = ok && cudaStreamCreateWithPriority(&L.s_tail, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&L.ev_tail, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < 3; ++i)
            ok = ok && cudaStreamCreateWithFlags(&L.s_poa[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&L.ev_join[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&L.ev_fork, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&L.ev_gather, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreate(&L.ev_end) == cudaSuccess;
        ok = ok && cudaMallocHost(&L.h_ctl, HCTL_WORDS * sizeof(u32)) == cudaSuccess;
        ok = ok && L.ctl.ensure(CTL_WORDS * sizeof(u32) + sizeof(CgCountersDev)) == cudaSuccess;
    }
    ok = ok && cudaFuncSetAttribute(k_index, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CG_IDX_SMEM_BYTES) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin) == cudaSuccess;
    if (!ok) { g_create_err = std::string("CUDA initialisation failed: ") + cudaGetErrorString(cudaGetLastError()); cg_destroy(h); return CG_ERR_CUDA; }
    ok = cudaFuncSetAttribute(k_poa2<CgPoa2C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CgPoa2Lay<CgPoa2C1>::cta_bytes) == cudaSuccess &&
         cudaFuncSetAttribute(k_poa2<CgPoa2GT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CgPoa2Lay<CgPoa2GT>::cta_bytes) == cudaSuccess &&
         cudaFuncSetAttribute(k_poa2<CgPoa2W1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CgPoa2Lay<CgPoa2W1>::cta_bytes) == cudaSuccess &&
         cudaFuncSetAttribute(k_poa2<CgPoa2W2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CgPoa2Lay<CgPoa2W2>::cta_bytes) == cudaSuccess;
    if (!ok) { g_create_err = std::string("CUDA initialisation failed: ") + cudaGetErrorString(cudaGetLastError()); cg_destroy(h); return CG_ERR_CUDA; }
    // POA tiers: resident warps of the k_poa2.cuh tiers; k_poa.cuh's global-memory tiers {nodes, edges, max segment length,
    // matrix cells, resident warps} are the last resort.
    h->c1_warps = (u32)h->sms * CgPoa2C1::CTAS_PER_SM * CgPoa2C1::WARPS;
    h->g_warps = (u32)h->sms * CgPoa2GT::CTAS_PER_SM * CgPoa2GT::WARPS;
    h->w1_warps = (u32)h->sms * CgPoa2W1::CTAS_PER_SM * CgPoa2W1::WARPS;
    h->w2_warps = (u32)h->sms * 2;
    h->tier[1].vcap = 16384; h->tier[1].ecap = 65536;  h->tier[1].lcap = CG_LEN_MAX; h->tier[1].hcap = 32u << 20;
    h->tier[2].vcap = 65535; h->tier[2].ecap = 262144; h->tier[2].lcap = CG_LEN_MAX; h->tier[2].hcap = 400u << 20;
    h->tier[1].warps = (u32)h->sms;
    h->tier[2].warps = 8;
    *out = h;
    return CG_OK;
}

void cg_destroy(cg_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    {
        std::lock_guard<std::mutex> lk(h->pool_mu);
        for (HostResults* r : h->pool) { if (r->in_use) { r->orphan = true; r->owner = nullptr; } else host_results_release(r); }
        h->pool.clear();
    }
    for (cudaEvent_t e : h->ev_h2d) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_bulk) cudaEventDestroy(e);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    if (h->ev_run0) cudaEventDestroy(h->ev_run0);
    for (Lane& L : h->lane) {
        DevBuf* lb[] = {&L.pwords, &L.ptags, &L.win, &L.offs, &L.solid_k, &L.solid_c, &L.slot_tpos, &L.slot_kmer, &L.anchors, &L.chain, &L.rel,
                        &L.pos, &L.regions, &L.arena, &L.fin, &L.visited, &L.jobs_s, &L.jobs_m, &L.jobs_3, &L.jobs_w, &L.jobs_ws, &L.jobs_r, &L.jobs_x,
                        &L.ctl, &L.off_fin, &L.out_off, &L.g_mem, &L.w1_mem, &L.w2_mem, &L.idx_keys, &L.idx_counts};
        for (DevBuf* b : lb) b->release();
        for (cudaEvent_t e : L.evpool) cudaEventDestroy(e);
        for (cudaEvent_t e : {L.ev_fork, L.ev_gather, L.ev_end, L.ev_tail, L.ev_join[0], L.ev_join[1], L.ev_join[2]}) if (e) cudaEventDestroy(e);
        for (cudaStream_t st : {L.stream, L.s_tail, L.s_poa[0], L.s_poa[1], L.s_poa[2]}) if (st) cudaStreamDestroy(st);
        if (L.h_ctl) cudaFreeHost(L.h_ctl);
    }
    DevBuf* bufs[] = {&h->d_bases, &h->d_seq_off, &h->d_wsb, &h->o_cons, &h->o_sk, &h->o_sc, &h->o_status, &h->o_len, &h->o_nsol,
                      &h->ra_cons, &h->ra_cons_off, &h->ra_solid_off, &h->ra_sk, &h->ra_tpl, &h->ra_tpl_off, &h->ra_rwb, &h->ra_roff,
                      &h->ra_rbases, &h->ra_wpos, &h->ra_order, &h->ra_head_off, &h->ra_head, &h->ra_len, &h->ra_out_off, &h->ra_out,
                      &h->ra_scratch, &h->ra_ctl, &h->ex_store, &h->ex_store_off, &h->ex_pile_read, &h->ex_pile_qlen, &h->ex_pile_ovb, &h->ex_ov,
                      &h->ex_cov, &h->ex_cov_off, &h->ex_cap_off, &h->ex_cap_beg, &h->ex_cap_end, &h->ex_nwin, &h->ex_win_pile, &h->ex_win_beg,
                      &h->ex_win_end, &h->ex_slot_base, &h->ex_slot_len, &h->ex_slot_src, &h->ex_slot_loc, &h->ex_win_nseq, &h->ex_win_nbytes,
                      &h->ex_win_base, &h->ex_flags, &h->in_text, &h->in_tile, &h->in_nl, &h->in_names, &h->in_name_off, &h->in_slots, &h->in_rec,
                      &h->in_head, &h->in_pfirst, &h->in_plast, &h->in_keep, &h->in_scratch, &h->in_pread, &h->in_pqlen, &h->in_ov, &h->in_res,
                      &h->in_ctl, &h->in_htile, &h->ra_skip, &h->d_packed};
    for (DevBuf* b : bufs) b->release();
    for (auto& t : h->tier) { t.mem.release(); t.desc.release(); }
    delete h;
}

// Host helper for "input_2bit": ASCII -> 2 bits per base (A 0, C 1, G 2, anything else 3 = T, upper or lower case: the mapping of the
// reference's read index, src/utils.cpp:21-32,189), base i at bits 2 (i & 3) of byte i >> 2.  `out` holds (n + 3) / 4 bytes.
void cg_pack_bases_2bit(const char* ascii, uint64_t n, uint8_t* out, int threads) {
    if (!ascii || !out || !n) return;
    const u64 nbytes = (n + 3) / 4;
    const int T = std::max(1, std::min(threads, 64));
    auto work = [&](u64 q0, u64 q1) {
        for (u64 q = q0; q < q1; ++q) {
            u32 v = 0;
            for (u32 j = 0; j < 4; ++j) {
                const u64 i = 4 * q + j;
                if (i >= n) break;
                const char c = ascii[i] & ~0x20;
                v |= (c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u) << (2 * j);
            }
            out[q] = (u8)v;
        }
    };
#ifndef CG_EMU
    std::vector<std::thread> th;
    const u64 per = (nbytes + T - 1) / T;
    for (int t = 1; t < T; ++t) { const u64 a = std::min(nbytes, per * t), b = std::min(nbytes, per * (t + 1)); if (b > a) th.emplace_back(work, a, b); }
    work(0, std::min(nbytes, per));
    for (std::thread& x : th) x.join();
#else
    work(0, nbytes);
#endif
}

int cg_set_option(cg_handle* h, const char* key, long long value) {
    if (!h || !key) return CG_ERR_INVALID_ARG;
    const std::string k(key);
    if (k == "chunk_budget_bytes") h->chunk_budget = (size_t)value;
    else if (k == "chunk_max_windows") h->chunk_max_windows = (u32)std::max<long long>(1, value);
    else if (k == "input_2bit") h->input_2bit = value != 0;
    else if (k == "results_with_solid") h->results_with_solid = value != 0;
    else if (k == "lanes") h->n_lanes = (int)std::max<long long>(1, std::min<long long>(value, CG_NLANES));
    else if (k == "poa_c1_warps") h->c1_warps = (u32)std::max<long long>(1, value);
    else if (k == "poa_g_warps") h->g_warps = (u32)std::max<long long>(1, value);
    else if (k == "poa_wide1_warps") h->w1_warps = (u32)std::max<long long>(1, value);
    else if (k == "poa_wide2_warps") h->w2_warps = (u32)std::max<long long>(1, value);
    else if (k == "poa_tier1_warps") { h->tier[1].warps = (u32)value; h->tier[1].ready = false; }
    else if (k == "poa_tier2_warps") { h->tier[2].warps = (u32)value; h->tier[2].ready = false; }
    else if (k == "poa_tier1_nodes") { h->tier[1].vcap = (u32)value; h->tier[1].ecap = 4 * (u32)value; h->tier[1].ready = false; }
    else if (k == "poa_tier1_cells") { h->tier[1].hcap = (u64)value; h->tier[1].ready = false; }
    else if (k == "poa_tier2_nodes") { h->tier[2].vcap = (u32)std::min<long long>(value, 65535); h->tier[2].ecap = 4 * h->tier[2].vcap; h->tier[2].ready = false; }
    else if (k == "poa_tier2_cells") { h->tier[2].hcap = (u64)value; h->tier[2].ready = false; }
    else { h->err = "unknown option " + k; return CG_ERR_INVALID_ARG; }
    if (h->uploaded) plan_chunks(h, h->planned_ramp);
    return CG_OK;
}

}  // extern "C"

namespace {

// Upload = offsets first (validated on the device), then the bases chunk by chunk on the H2D stream, one event per
// chunk.  wait = true: return when everything is resident (cg_upload); false: the kernels of chunk i wait for
// ev_h2d[i] (cg_correct_windows).
int upload_impl(cg_handle* h, const cg_batch* in, bool wait) {
    if (!in || !in->win_seq_begin || !in->seq_off || (!in->bases && in->n_windows)) { h->err = "null batch pointer"; return CG_ERR_INVALID_ARG; }
    cudaSetDevice(h->device);
    h->uploaded = h->ran = false;
    h->h2d_pending = false;
    h->ex_valid = false;
    h->run_gen++;
    const u32 W = in->n_windows;
    const u64 n_seqs = in->win_seq_begin[W];
    const u32 k = h->p.mer_size;
    if (in->win_seq_begin[0] != 0) { h->err = "win_seq_begin[0] must be 0"; return CG_ERR_INVALID_ARG; }
    h->h_wsb.assign(in->win_seq_begin, in->win_seq_begin + W + 1);
    h->h_wbase.resize((size_t)W + 1);
    h->h_tlen.resize(W);
    for (u32 w = 0; w < W; ++w) {
        const u32 s0 = in->win_seq_begin[w], s1 = in->win_seq_begin[w + 1];
        if (s1 <= s0) { h->err = "empty pile (window without a template)"; return CG_ERR_INVALID_ARG; }
        h->h_wbase[w] = in->seq_off[s0];
        const u64 t1 = in->seq_off[s0 + 1];
        if (t1 < h->h_wbase[w]) { h->err = "seq_off must be non-decreasing"; return CG_ERR_INVALID_ARG; }
        const u64 tlen = t1 - h->h_wbase[w];
        if (tlen >= (1ull << 31)) { h->err = "a template longer than 2^31 bases"; return CG_ERR_INVALID_ARG; }
        h->h_tlen[w] = (u32)tlen;                   // windows over a limit of this build are flagged by k_plan (CG_WINDOW_ERROR), not refused
    }
    const u64 n_bases = n_seqs ? in->seq_off[n_seqs] : 0;
    h->h_wbase[W] = n_bases;
    for (u32 w = 0; w < W; ++w)
        if (h->h_wbase[w + 1] < h->h_wbase[w]) { h->err = "seq_off must be non-decreasing"; return CG_ERR_INVALID_ARG; }
    h->W = W; h->n_seqs = n_seqs; h->n_bases = n_bases;
    CK(h->d_bases.ensure(n_bases + 64));
    if (h->input_2bit) CK(h->d_packed.ensure(n_bases / 4 + 64));
    CK(h->d_seq_off.ensure((n_seqs + 1) * sizeof(u64)));
    CK(h->d_wsb.ensure((W + 1) * sizeof(u32)));
    cudaStream_t sc = h->s_h2d;
    CK(cudaMemcpyAsync(h->d_seq_off.p, in->seq_off, (n_seqs + 1) * sizeof(u64), cudaMemcpyHostToDevice, sc));
    CK(cudaMemcpyAsync(h->d_wsb.p, in->win_seq_begin, (W + 1) * sizeof(u32), cudaMemcpyHostToDevice, sc));
    // every sequence: offsets non-decreasing, length within the build's limit (checked on the device, one word back)
    u32* vflags = h->lane[0].ctl.as<u32>() + CTL_VFLAGS;
    CK(cudaMemsetAsync(vflags, 0, sizeof(u32), sc));
    if (n_seqs) CG_LAUNCH(k_validate, (u32)std::min<u64>((n_seqs + 255) / 256, 4096), 256, 0, sc, h->d_seq_off.as<u64>(), n_seqs, vflags);
    CK(cudaMemcpyAsync(h->lane[0].h_ctl + HCTL_VFLAGS, vflags, sizeof(u32), cudaMemcpyDeviceToHost, sc));
    plan_chunks(h, !wait);
    while (h->ev_h2d.size() < h->chunks.size()) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->ev_h2d.push_back(e);
    }
    CK(cudaStreamSynchronize(sc));
    if (h->lane[0].h_ctl[HCTL_VFLAGS] & CG_FLAG_BAD_OFFSETS) { h->err = "seq_off must be non-decreasing"; return CG_ERR_INVALID_ARG; }
    for (size_t ci = 0; ci < h->chunks.size(); ++ci) {
        const ChunkPlan& cp = h->chunks[ci];
        const u64 b0 = h->h_wbase[cp.w0], b1 = h->h_wbase[cp.w0 + cp.nwin];
        if (b1 > b0 && !h->input_2bit) CK(cudaMemcpyAsync(h->d_bases.as<char>() + b0, in->bases + b0, b1 - b0, cudaMemcpyHostToDevice, sc));
        if (b1 > b0 && h->input_2bit) {                     // a quarter of the bytes over the bus, expanded to the ASCII the kernels read
            const u64 p0 = b0 >> 2, p1 = (b1 + 3) >> 2;
            CK(cudaMemcpyAsync(h->d_packed.as<u8>() + p0, (const u8*)in->bases + p0, p1 - p0, cudaMemcpyHostToDevice, sc));
            CG_LAUNCH(k_unpack_2bit, (u32)std::min<u64>((b1 - b0 + 1023) / 1024, (u64)h->sms * 8), 256, 0, sc, (const u8*)h->d_packed.as<u8>(), h->d_bases.as<char>(), b0, b1);
        }
        CK(cudaEventRecord(h->ev_h2d[ci], sc));
    }
    if (wait) CK(cudaStreamSynchronize(sc));
    else h->h2d_pending = true;
    h->uploaded = true;
    return CG_OK;
}

// Chunks ci = lane, lane + n_lanes, ... on lane `li` (its own host thread when there are two lanes).
void lane_main(cg_handle* h, int li, int n_lanes) {
    CG_NVTX(li == 0 ? "cg:lane0" : "cg:lane1");
    Lane& L = h->lane[li];
    cudaSetDevice(h->device);
    L.rc = CG_OK;
    for (size_t ci = (size_t)li; ci < h->chunks.size(); ci += (size_t)n_lanes) {
        {
            std::lock_guard<std::mutex> lk(h->commit_mu);
            if (h->abort_run) break;
        }
        const int rc = run_chunk(h, L, ci);
        if (rc != CG_OK) {
            std::lock_guard<std::mutex> lk(h->commit_mu);
            if (!h->abort_run) L.rc = rc;                  // the first failure is the one reported
            h->abort_run = true;
            h->commit_cv.notify_all();
            break;
        }
    }
    cudaStreamWaitEvent(L.s_tail, h->ev_run0, 0);
    if (L.tail_recorded) cudaStreamWaitEvent(L.s_tail, L.ev_tail, 0);
    cudaEventRecord(L.ev_end, L.s_tail);
}

int run_impl(cg_handle* h) {
    cudaSetDevice(h->device);
    h->ran = false;
    h->o_cons_n = h->o_solid_n = 0;
    h->next_commit = 0;
    h->bulk_enqueued = 0;
    h->abort_run = false;
    while (h->ev_bulk.size() < h->chunks.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        h->ev_bulk.push_back(e);
    }
    memset(h->stage_ms, 0, sizeof h->stage_ms);
    memset(h->stage_launches, 0, sizeof h->stage_launches);
    CK(h->o_status.ensure(h->W + 16)); CK(h->o_len.ensure((h->W + 1) * sizeof(u64))); CK(h->o_nsol.ensure((h->W + 1) * sizeof(u64)));
    const int n_lanes = std::max(1, std::min<int>(h->n_lanes, (int)std::min<size_t>(h->chunks.size(), CG_NLANES)));
    CK(cudaEventRecord(h->ev_run0, h->lane[0].stream));
    for (int li = 0; li < n_lanes; ++li) {
        Lane& L = h->lane[li];
        L.spans.clear(); L.kspans.clear(); L.pool_at = 0; L.err.clear(); L.rc = CG_OK; L.tail_recorded = false;
        memset(L.stage_ms, 0, sizeof L.stage_ms); memset(L.stage_launches, 0, sizeof L.stage_launches);
        CK(cudaMemsetAsync(L.ctl.p, 0, CTL_WORDS * sizeof(u32) + sizeof(CgCountersDev), L.stream));
        if (li) CK(cudaStreamWaitEvent(L.stream, h->ev_run0, 0));      // nothing of this run starts before its first event
    }
#ifndef CG_EMU
    std::vector<std::thread> workers;
    for (int li = 1; li < n_lanes; ++li) workers.emplace_back(lane_main, h, li, n_lanes);
    lane_main(h, 0, n_lanes);
    for (std::thread& t : workers) t.join();
#else
    lane_main(h, 0, 1);
#endif
    int rc = CG_OK;
    for (int li = 0; li < n_lanes; ++li) {
        cudaStreamSynchronize(h->lane[li].stream);
        cudaStreamSynchronize(h->lane[li].s_tail);
        for (int i = 0; i < 3; ++i) cudaStreamSynchronize(h->lane[li].s_poa[i]);
        if (rc == CG_OK && h->lane[li].rc != CG_OK) { rc = h->lane[li].rc; if (!h->lane[li].err.empty()) h->err = h->lane[li].err; }
    }
    if (rc != CG_OK) { cudaStreamSynchronize(h->s_h2d); cudaStreamSynchronize(h->s_d2h); return rc; }
    if (getenv("CG_TIMELINE")) {
        static const char* names[CG_N_STAGES] = {"pack", "index", "chain", "split", "poa", "stitch", "polish"};
        for (int li = 0; li < n_lanes; ++li)
            for (const StageSpan& sp : h->lane[li].spans) {
                float t0 = 0, t1 = 0;
                cudaEventElapsedTime(&t0, h->ev_run0, sp.a); cudaEventElapsedTime(&t1, h->ev_run0, sp.b);
                fprintf(stderr, "[timeline] lane %d %-6s %9.3f -> %9.3f  (%.3f ms)\n", li, names[sp.stage], t0, t1, t1 - t0);
            }
        for (size_t ci = 0; ci < h->chunks.size(); ++ci) {
            float t = 0;
            cudaEventElapsedTime(&t, h->ev_run0, h->ev_bulk[ci]);
            fprintf(stderr, "[timeline] chunk %zu bulk done at %9.3f\n", ci, t);
        }
    }
    h->run_ms = 0;
    memset(&h->kstats, 0, sizeof h->kstats);
    CgCountersDev sum{};
    for (int li = 0; li < n_lanes; ++li) {
        Lane& L = h->lane[li];
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev_run0, L.ev_end);
        h->run_ms = std::max(h->run_ms, ms);
        for (const StageSpan& sp : L.spans) { float m2 = 0; cudaEventElapsedTime(&m2, sp.a, sp.b); h->stage_ms[sp.stage] += m2; }
        for (int i = 0; i < CG_N_STAGES; ++i) h->stage_launches[i] += L.stage_launches[i];
        for (const KernelSpan& sp : L.kspans) { float m2 = 0; cudaEventElapsedTime(&m2, sp.a, sp.b); h->kstats.ms[sp.kind] += m2; h->kstats.launches[sp.kind] += sp.launches; }
        CgCountersDev cd{};
        CK(cudaMemcpy(&cd, L.ctl.as<u32>() + CTL_WORDS, sizeof cd, cudaMemcpyDeviceToHost));
        for (int t = 0; t < 4; ++t) { h->kstats.poa_cells[t] += cd.tier_cells[t]; h->kstats.poa_pred_cells[t] += cd.tier_pred[t]; }
        sum.anchors += cd.anchors; sum.regions += cd.regions; sum.poa_graphs += cd.poa_graphs; sum.alignments += cd.alignments;
        sum.dp_cells += cd.dp_cells; sum.dp_pred_cells += cd.dp_pred_cells; sum.solid_kmers += cd.solid_kmers;
        sum.consensus_bytes += cd.consensus_bytes; sum.fallback_windows += cd.fallback_windows; sum.error_windows += cd.error_windows;
    }
    cg_counters& o = h->counters;
    o.windows = h->W; o.sequences = h->n_seqs; o.bases = h->n_bases;
    o.anchors = sum.anchors; o.regions = sum.regions; o.poa_graphs = sum.poa_graphs; o.alignments = sum.alignments;
    o.dp_cells = sum.dp_cells; o.dp_pred_cells = sum.dp_pred_cells; o.solid_kmers = sum.solid_kmers;
    o.consensus_bytes = sum.consensus_bytes; o.fallback_windows = sum.fallback_windows; o.error_windows = sum.error_windows;
    h->ran = true;
    return CG_OK;
}

void fill_results(cg_handle* h, HostResults* r, cg_results* out) {
    const u32 W = h->W;
    r->cons_off[W] = h->o_cons_n; r->solid_off[W] = h->o_solid_n;
    out->n_windows = W; out->cons_off = r->cons_off; out->cons = r->cons; out->status = r->status;
    if (!h->results_with_solid) memset(r->solid_off, 0, ((size_t)W + 1) * sizeof(u64));     // the lists stayed on the device
    out->solid_off = r->solid_off; out->solid_kmer = r->sk; out->solid_count = r->sc; out->owner_ = r;
    r->gen = h->run_gen;
}

void give_back(cg_handle* h, HostResults* r) {
    std::lock_guard<std::mutex> lk(h->pool_mu);
    r->in_use = false;
}

}  // namespace

extern "C" {

int cg_upload(cg_handle* h, const cg_batch* in) {
    if (!h) return CG_ERR_INVALID_ARG;
    CG_NVTX("cg_upload");
    return upload_impl(h, in, true);
}

int cg_run(cg_handle* h) {
    if (!h) return CG_ERR_INVALID_ARG;
    CG_NVTX("cg_run");
    if (!h->uploaded) { h->err = "cg_run before cg_upload"; return CG_ERR_STATE; }
    h->stream_out = nullptr;
    return run_impl(h);
}

int cg_download(cg_handle* h, cg_results* out) {
    if (!h || !out) return CG_ERR_INVALID_ARG;
    CG_NVTX("cg_download");
    if (!h->ran) { h->err = "cg_download before a successful cg_run"; return CG_ERR_STATE; }
    cudaSetDevice(h->device);
    const u32 W = h->W;
    HostResults* r = nullptr;
    { int rc = host_results_acquire(h, W, &r); if (rc) { if (r) give_back(h, r); return rc; } }
    { int rc = host_results_reserve(h, h->err, r, h->o_cons_n + 1, h->o_solid_n + 1, 0, 0); if (rc) { give_back(h, r); return rc; } }
    cudaError_t e = cudaSuccess;
    cudaStream_t sd = h->s_d2h;
    if (W) {
        e = cudaMemcpyAsync(r->cons_off, h->o_len.p, W * sizeof(u64), cudaMemcpyDeviceToHost, sd);
        if (e == cudaSuccess) e = cudaMemcpyAsync(r->solid_off, h->o_nsol.p, W * sizeof(u64), cudaMemcpyDeviceToHost, sd);
        if (e == cudaSuccess) e = cudaMemcpyAsync(r->status, h->o_status.p, W, cudaMemcpyDeviceToHost, sd);
        if (e == cudaSuccess && h->o_cons_n) e = cudaMemcpyAsync(r->cons, h->o_cons.p, h->o_cons_n, cudaMemcpyDeviceToHost, sd);
        if (e == cudaSuccess && h->o_solid_n && h->results_with_solid) e = cudaMemcpyAsync(r->sk, h->o_sk.p, h->o_solid_n * 4, cudaMemcpyDeviceToHost, sd);
        if (e == cudaSuccess && h->o_solid_n && h->results_with_solid) e = cudaMemcpyAsync(r->sc, h->o_sc.p, h->o_solid_n * 4, cudaMemcpyDeviceToHost, sd);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(sd);
    if (e != cudaSuccess) { h->err = std::string("download: ") + cudaGetErrorString(e); give_back(h, r); return CG_ERR_CUDA; }
    fill_results(h, r, out);
    return CG_OK;
}

void cg_free_results(cg_results* r) {
    if (!r || !r->owner_) return;
    HostResults* hr = static_cast<HostResults*>(r->owner_);
    r->owner_ = nullptr;
    if (hr->orphan) host_results_release(hr);          // its handle is gone
    else give_back(hr->owner, hr);
}

// H2D of chunk i+1, the kernels of chunk i and D2H of chunk i-1 overlap (three streams).
int cg_correct_windows(cg_handle* h, const cg_batch* in, cg_results* out) {
    if (!h || !out) return CG_ERR_INVALID_ARG;
    CG_NVTX("cg_correct_windows");
    int rc = upload_impl(h, in, false);
    if (rc != CG_OK) return rc;
    HostResults* r = nullptr;
    rc = host_results_acquire(h, h->W, &r);
    if (rc != CG_OK) { if (r) give_back(h, r); cudaStreamSynchronize(h->s_h2d); h->h2d_pending = false; return rc; }
    h->stream_out = r;
    rc = run_impl(h);
    h->stream_out = nullptr;
    h->h2d_pending = false;
    cudaError_t e1 = cudaStreamSynchronize(h->s_h2d), e2 = cudaStreamSynchronize(h->s_d2h);
    if (rc == CG_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) { h->err = "transfer stream failed"; rc = CG_ERR_CUDA; }
    if (rc == CG_OK && (!r->cons || !r->sk)) rc = host_results_reserve(h, h->err, r, 1, 1, 0, 0);      // zero windows
    if (rc != CG_OK) { give_back(h, r); return rc; }
    fill_results(h, r, out);
    return CG_OK;
}

// ---- consensus re-anchoring: alignConsensus for every read of the batch ------------------------------
// on_device: cg_finish_resident — the results, the templates and the reads are taken where cg_upload_piles / cg_run left them in HBM
// (template lengths from the host copy h_tlen, the reads gathered from the device store); windows->seq_off / reads->read_bases unused.
static int reanchor_impl(cg_handle* h, const cg_batch* windows, const cg_results* cons, const cg_reads* reads, u32 trim_mer, cg_corrected* out,
                         bool on_device = false) {
    if (!h || !out) return CG_ERR_INVALID_ARG;
    CG_NVTX("cg_reanchor_reads");
    if (!windows || !cons || !reads || !reads->read_win_begin || !reads->read_off || !windows->win_seq_begin || (!on_device && !windows->seq_off)) {
        h->err = "null argument"; return CG_ERR_INVALID_ARG;
    }
    cudaSetDevice(h->device);
    const u32 W = cons->n_windows, R = reads->n_reads;
    if (R >> 31) { h->err = "more than 2^31 reads"; return CG_ERR_CAPACITY; }          // (k_reanchor also takes its opaque zero from this)
    if (windows->n_windows != W || reads->read_win_begin[0] != 0 || reads->read_win_begin[R] != W) {
        h->err = "windows, results and reads do not describe the same windows"; return CG_ERR_INVALID_ARG;
    }
    if (W && (!cons->cons_off || !cons->solid_off || !reads->win_pos)) { h->err = "null argument"; return CG_ERR_INVALID_ARG; }
    if (R && !reads->read_bases && !on_device) { h->err = "null argument"; return CG_ERR_INVALID_ARG; }
    const u32 k = h->p.mer_size;
    // are these the results of the batch still resident on this handle?
    bool resident = on_device;
    if (!on_device && cons->owner_ && h->ran && h->W == W) {
        std::lock_guard<std::mutex> lk(h->pool_mu);
        for (HostResults* c : h->pool)
            if ((void*)c == cons->owner_ && c->gen == h->run_gen && c->cons_off == cons->cons_off) resident = true;
    }
    // sizes: longest query, per-read capacity of the corrected read
    u32 maxL = 16;
    std::vector<u64> head_off((size_t)R + 1, 0);
    std::vector<u64> tpl_off;
    u64 tpl_bytes = 0;
    if (!resident) tpl_off.assign((size_t)W + 1, 0);
    for (u32 r = 0; r < R; ++r) {
        const u32 w0 = reads->read_win_begin[r], w1 = reads->read_win_begin[r + 1];
        if (w1 < w0 || reads->read_off[r + 1] < reads->read_off[r]) { h->err = "read offsets must be non-decreasing"; return CG_ERR_INVALID_ARG; }
        u64 sum = 0;
        for (u32 w = w0; w < w1; ++w) {
            u64 len = cons->cons_off[w + 1] - cons->cons_off[w];
            if (len < k) {                                                     // the raw template stands in (correctionAlignment.cpp:72-74)
                const u32 s0 = windows->win_seq_begin[w];
                len = on_device ? (u64)h->h_tlen[w] : windows->seq_off[s0 + 1] - windows->seq_off[s0];
                if (!resident) { tpl_off[w + 1] = len; tpl_bytes += len; }
            }
            if (len > (u64)CG_RA_QMAX) { h->err = "a consensus is longer than 8000 bases"; return CG_ERR_CAPACITY; }
            maxL = std::max<u32>(maxL, (u32)len);
            sum += len;
        }
        const u64 rawLen = reads->read_off[r + 1] - reads->read_off[r];
        if (rawLen >= (1ull << 31)) { h->err = "a read is longer than 2^31 bases"; return CG_ERR_CAPACITY; }
        head_off[r + 1] = head_off[r] + round_up(rawLen + 2 * sum + 64, 16);
    }
    const u32 rmax = std::max<u32>(maxL, reads->window_size + 2 * reads->window_overlap);
    if (rmax > 48 * 1024) { h->err = "window_size + 2 * window_overlap too large for the shared-memory reference buffer"; return CG_ERR_CAPACITY; }
    std::vector<u32> order(R);
    for (u32 r = 0; r < R; ++r) order[r] = r;
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) {
        return reads->read_win_begin[a + 1] - reads->read_win_begin[a] > reads->read_win_begin[b + 1] - reads->read_win_begin[b];
    });
    cudaStream_t st = h->lane[0].stream;
    const u64 n_rbases = R ? reads->read_off[R] : 0;
    // ---- uploads
    CK(h->ra_rwb.ensure(((size_t)R + 1) * 4)); CK(h->ra_roff.ensure(((size_t)R + 1) * 8)); CK(h->ra_rbases.ensure(n_rbases + 16));
    CK(h->ra_wpos.ensure(((size_t)W + 1) * 4)); CK(h->ra_order.ensure(((size_t)R + 1) * 4)); CK(h->ra_head_off.ensure(((size_t)R + 1) * 8));
    CK(h->ra_head.ensure(head_off[R] + 64)); CK(h->ra_len.ensure(((size_t)R + 1) * 4)); CK(h->ra_out_off.ensure(((size_t)R + 1) * 8));
    CK(h->ra_ctl.ensure(64));
    CK(cudaMemcpyAsync(h->ra_rwb.p, reads->read_win_begin, ((size_t)R + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->ra_roff.p, reads->read_off, ((size_t)R + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_rbases && !on_device) CK(cudaMemcpyAsync(h->ra_rbases.p, reads->read_bases, n_rbases, cudaMemcpyHostToDevice, st));
    if (n_rbases && on_device)                                     // read r = store[pile_read[r]], already normalised by cg_upload_piles
        CG_LAUNCH(k_reanchor_reads_from_store, std::min<u32>(R, (u32)h->sms * 8), 256, 0, st, (const char*)h->ex_store.as<char>(),
                  (const u64*)h->ex_store_off.as<u64>(), (const u32*)h->ex_pile_read.as<u32>(), (const u64*)h->ra_roff.as<u64>(), h->ra_rbases.as<char>(), R);
    if (W) CK(cudaMemcpyAsync(h->ra_wpos.p, reads->win_pos, (size_t)W * 4, cudaMemcpyHostToDevice, st));
    if (R) CK(cudaMemcpyAsync(h->ra_order.p, order.data(), (size_t)R * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->ra_head_off.p, head_off.data(), ((size_t)R + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(h->ra_ctl.p, 0, 64, st));
    CgReanchorArgs A{};
    std::vector<char> tpl;
    if (resident) {
        const u64 tot_c = h->o_cons_n, tot_s = h->o_solid_n;                 // o_len / o_nsol hold the W start offsets: close them
        CK(cudaMemcpyAsync(h->o_len.as<u64>() + W, &tot_c, 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->o_nsol.as<u64>() + W, &tot_s, 8, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));                                       // tot_c / tot_s live on this stack frame
        A.cons = h->o_cons.as<char>(); A.cons_off = h->o_len.as<u64>(); A.solid_off = h->o_nsol.as<u64>(); A.solid_kmer = h->o_sk.as<u32>();
        A.bases = h->d_bases.as<char>(); A.seq_off = h->d_seq_off.as<u64>(); A.win_seq_begin = h->d_wsb.as<u32>();
    } else {
        const u64 nc = cons->cons_off[W], ns = cons->solid_off[W];
        for (u32 w = 0; w < W; ++w) tpl_off[w + 1] += tpl_off[w];
        tpl.resize(tpl_bytes + 1);
        for (u32 w = 0; w < W; ++w)
            if (tpl_off[w + 1] > tpl_off[w])
                memcpy(tpl.data() + tpl_off[w], windows->bases + windows->seq_off[windows->win_seq_begin[w]], tpl_off[w + 1] - tpl_off[w]);
        CK(h->ra_cons.ensure(nc + 16)); CK(h->ra_cons_off.ensure(((size_t)W + 1) * 8)); CK(h->ra_solid_off.ensure(((size_t)W + 1) * 8));
        CK(h->ra_sk.ensure(ns * 4 + 16)); CK(h->ra_tpl.ensure(tpl_bytes + 16)); CK(h->ra_tpl_off.ensure(((size_t)W + 1) * 8));
        if (nc) CK(cudaMemcpyAsync(h->ra_cons.p, cons->cons, nc, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->ra_cons_off.p, cons->cons_off, ((size_t)W + 1) * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->ra_solid_off.p, cons->solid_off, ((size_t)W + 1) * 8, cudaMemcpyHostToDevice, st));
        if (ns) CK(cudaMemcpyAsync(h->ra_sk.p, cons->solid_kmer, ns * 4, cudaMemcpyHostToDevice, st));
        if (tpl_bytes) CK(cudaMemcpyAsync(h->ra_tpl.p, tpl.data(), tpl_bytes, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->ra_tpl_off.p, tpl_off.data(), ((size_t)W + 1) * 8, cudaMemcpyHostToDevice, st));
        A.cons = h->ra_cons.as<char>(); A.cons_off = h->ra_cons_off.as<u64>(); A.solid_off = h->ra_solid_off.as<u64>(); A.solid_kmer = h->ra_sk.as<u32>();
        A.tpl = h->ra_tpl.as<char>(); A.tpl_off = h->ra_tpl_off.as<u64>();
    }
    // ---- scratch: resident warps, each with its buffers and a direction matrix for the banded sub-alignment
    u32 ctas = std::min<u32>((R + CG_RA_WARPS - 1) / CG_RA_WARPS, (u32)h->sms * CG_RA_CTAS_PER_SM);
    if (ctas == 0) ctas = 1;
    u64 dir_cap = std::min<u64>((u64)maxL * maxL, 4ull << 20);
    const u64 budget = 6ull << 30;
    while ((cg_ra_fixed_bytes(maxL, rmax) + dir_cap) * ctas * CG_RA_WARPS > budget && ctas > (u32)h->sms) ctas = (ctas + 1) / 2;
    const u64 stride = round_up(cg_ra_fixed_bytes(maxL, rmax) + dir_cap, 256);
    CK(h->ra_scratch.ensure(stride * ctas * CG_RA_WARPS));
    A.n_reads = R; A.order = h->ra_order.as<u32>(); A.read_win_begin = h->ra_rwb.as<u32>(); A.read_off = h->ra_roff.as<u64>();
    A.read_bases = h->ra_rbases.as<char>(); A.win_pos = h->ra_wpos.as<u32>();
    A.ws = reads->window_size; A.ov = reads->window_overlap; A.k = k;
    A.head = h->ra_head.as<char>(); A.head_off = h->ra_head_off.as<u64>(); A.out_len = h->ra_len.as<u32>();
    A.scratch = h->ra_scratch.as<u8>(); A.scratch_stride = stride; A.maxL = maxL; A.rmax = rmax; A.dir_cap = dir_cap;
    A.ctl = h->ra_ctl.as<u32>();
    const size_t smem_ref = (size_t)CG_RA_WARPS * cg_ra_smem_per_warp(rmax, maxL, false);
    size_t smem = (size_t)CG_RA_WARPS * cg_ra_smem_per_warp(rmax, maxL, true);
    A.lines_in_smem = smem <= (224 / CG_RA_CTAS_PER_SM) * 1024 ? 1u : 0u;   // CG_RA_CTAS_PER_SM CTAs per SM keep their lines on chip
    if (!A.lines_in_smem) smem = smem_ref;
    CK(cudaFuncSetAttribute(k_reanchor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    struct Events {                                         // destroyed on every return path, CK's included
        cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
        ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
    } evs;
    for (cudaEvent_t& x : evs.e) CK(cudaEventCreate(&x));
    cudaEvent_t e0 = evs.e[0], e1 = evs.e[1], f0 = evs.e[2], f1 = evs.e[3];
    CK(cudaEventRecord(e0, st));
    if (R) CG_LAUNCH(k_reanchor, ctas, CG_RA_WARPS * 32, smem, st, A);
    CK(cudaEventRecord(e1, st));
    CK(cudaGetLastError());
    // ---- post-filters in place: trimRead + dropRead rewrite the lengths and give the first kept base
    h->fin_ms = 0;
    if (trim_mer) {
        CK(h->ra_skip.ensure(((size_t)R + 1) * 4));
        CK(cudaEventRecord(f0, st));
        if (R) CG_LAUNCH(k_finish_reads, std::min<u32>(R, (u32)h->sms * 8), 256, 128, st, (const char*)h->ra_head.as<char>(), (const u64*)h->ra_head_off.as<u64>(),
                         h->ra_len.as<u32>(), h->ra_skip.as<u32>(), R, trim_mer);
        CK(cudaEventRecord(f1, st));
        CK(cudaEventSynchronize(f1));
        CK(cudaEventElapsedTime(&h->fin_ms, f0, f1));
    }
    // ---- lengths -> dense offsets -> gather -> host
    std::vector<u32> len((size_t)R + 1, 0);
    u32 ctl[4] = {0, 0, 0, 0};
    if (R) CK(cudaMemcpyAsync(len.data(), h->ra_len.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctl, h->ra_ctl.p, sizeof ctl, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&h->ra_ms, e0, e1));
    memcpy(&h->ra_cells, ctl + 2, 8);
    if (ctl[1]) {
        h->err = (ctl[1] & CG_RA_FLAG_CAPACITY) ? "re-anchoring: a window exceeds a capacity limit of this build (alignment region, consensus length or banded sub-alignment)"
               : (ctl[1] & CG_RA_FLAG_TRACEBACK) ? "re-anchoring: the banded traceback left its matrix (undefined behaviour in the reference)"
               : "re-anchoring: an alignment region is empty or nothing aligns (the reference does not survive this input either)";
        return (ctl[1] & CG_RA_FLAG_CAPACITY) ? CG_ERR_CAPACITY : CG_ERR_INVALID_ARG;
    }
    HostCorrected* hc = hc_acquire();
    auto fail = [&](int rc) { hc_destroy(hc); return rc; };
    if (!hc_ensure(hc->off, hc->off_cap, ((size_t)R + 1) * 8)) { h->err = "out of pinned memory"; return fail(CG_ERR_OUT_OF_MEMORY); }
    hc->off[0] = 0;
    for (u32 r = 0; r < R; ++r) hc->off[r + 1] = hc->off[r] + len[r];
    const u64 tot = hc->off[R];
    if (!hc_ensure(hc->bases, hc->bases_cap, tot + 1)) { h->err = "out of pinned memory"; return fail(CG_ERR_OUT_OF_MEMORY); }
    cudaError_t e = h->ra_out.ensure(tot + 16);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->ra_out_off.p, hc->off, ((size_t)R + 1) * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && R && tot) {
        CG_LAUNCH(k_reanchor_gather, std::min<u32>(R, (u32)h->sms * 8), 256, 0, st, (const char*)h->ra_head.as<char>(), (const u64*)h->ra_head_off.as<u64>(),
                  (const u32*)(trim_mer ? h->ra_skip.as<u32>() : nullptr), (const u64*)h->ra_out_off.as<u64>(), h->ra_out.as<char>(), R);
        e = cudaMemcpyAsync(hc->bases, h->ra_out.p, tot, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { h->err = std::string("re-anchoring download: ") + cudaGetErrorString(e); return fail(CG_ERR_CUDA); }
    hc->bases[tot] = 0;
    out->n_reads = R; out->read_off = hc->off; out->bases = hc->bases; out->owner_ = hc;
    return CG_OK;
}

int cg_reanchor_reads(cg_handle* h, const cg_batch* windows, const cg_results* cons, const cg_reads* reads, cg_corrected* out) {
    return reanchor_impl(h, windows, cons, reads, 0u, out);
}

// ---- post-filters: alignConsensus + trimRead + dropRead = the sequence line of every FASTA record --------
int cg_finish_reads(cg_handle* h, const cg_batch* windows, const cg_results* cons, const cg_reads* reads, uint32_t trim_mer, cg_corrected* out) {
    return reanchor_impl(h, windows, cons, reads, trim_mer, out);
}

// The tail of the chain cg_upload_piles -> cg_run on one handle without taking the windows, the results or the reads through the host.
int cg_finish_resident(cg_handle* h, uint32_t trim_mer, cg_corrected* out) {
    if (!h || !out) return CG_ERR_INVALID_ARG;
    CG_NVTX("cg_finish_resident");
    if (!h->uploaded || !h->ran || !h->ex_valid) { h->err = "cg_finish_resident needs cg_upload_piles and cg_run on this handle"; return CG_ERR_STATE; }
    cudaSetDevice(h->device);
    cudaStream_t st = h->lane[0].stream;
    const u32 W = h->W, R = (u32)h->ex_pile_read_h.size();
    std::vector<u64> cons_off((size_t)W + 1, 0), roff((size_t)R + 1, 0);
    if (W) CK(cudaMemcpyAsync(cons_off.data(), h->o_len.p, (size_t)W * 8, cudaMemcpyDeviceToHost, st));   // start offsets of the consensuses
    CK(cudaStreamSynchronize(st));
    cons_off[W] = h->o_cons_n;
    for (u32 r = 0; r < R; ++r) {
        const u32 q = h->ex_pile_read_h[r];
        roff[r + 1] = roff[r] + (h->ex_store_off_h[q + 1] - h->ex_store_off_h[q]);
    }
    u32 zero = 0;
    cg_batch wb = {W, h->h_wsb.data(), nullptr, nullptr};
    cg_results cr{};
    cr.n_windows = W; cr.cons_off = cons_off.data(); cr.solid_off = cons_off.data();   // solid_off: only read when the results are not resident
    cg_reads rd = {R, h->ex_rwb.data(), roff.data(), nullptr, W ? h->ex_wpos.data() : &zero, h->ex_ws, h->ex_ovl};
    return reanchor_impl(h, &wb, &cr, &rd, trim_mer, out, true);
}

int cg_finish_stats(const cg_handle* h, float* kernel_ms) {
    if (!h) return CG_ERR_INVALID_ARG;
    if (kernel_ms) *kernel_ms = h->fin_ms;
    return CG_OK;
}

void cg_free_corrected(cg_corrected* c) {
    if (!c || !c->owner_) return;
    HostCorrected* hc = static_cast<HostCorrected*>(c->owner_);
    c->owner_ = nullptr;
    hc_release(hc);
}

int cg_reanchor_stats(const cg_handle* h, float* kernel_ms, uint64_t* dp_cells) {
    if (!h) return CG_ERR_INVALID_ARG;
    if (kernel_ms) *kernel_ms = h->ra_ms;
    if (dp_cells) *dp_cells = h->ra_cells;
    return CG_OK;
}

// ---- PAF ingest: every getNextReadPile of a PAF text on the device ------------------------------------------
struct HostPileSet { std::vector<u32> pile_read, pile_qlen, ovb, res; std::vector<cg_overlap> ov; };

int cg_ingest_paf(cg_handle* h, const char* paf, uint64_t nbytes, const cg_read_names* names, uint32_t max_support, cg_pile_set* out) {
    if (!h || !out) return CG_ERR_INVALID_ARG;
    CG_NVTX("cg_ingest_paf");
    if (!names || !names->name_off || (nbytes && !paf) || (names->n_reads && !names->names)) { h->err = "null argument"; return CG_ERR_INVALID_ARG; }
    if (max_support == 0) { h->err = "max_support must be at least 1"; return CG_ERR_INVALID_ARG; }
    if (nbytes && paf[nbytes - 1] != '\n') {
        h->err = "the PAF text must end with a newline (getNextReadPile never returns on a last line without one, src/alignmentPiles.cpp:29-35)";
        return CG_ERR_INVALID_ARG;
    }
    cudaSetDevice(h->device);
    cudaStream_t st = h->lane[0].stream;
    const u32 NN = names->n_reads;
    const u64 name_bytes = names->name_off[NN];
    const u32 n_tiles = (u32)((nbytes + CG_IN_TILE - 1) / CG_IN_TILE);
    if ((nbytes + CG_IN_TILE - 1) / CG_IN_TILE >= (1ull << 31)) { h->err = "PAF text too large for one call"; return CG_ERR_CAPACITY; }
    u32 slots = 64;
    while (slots < 2 * NN + 2) slots <<= 1;
    CK(h->in_text.ensure((size_t)n_tiles * CG_IN_TILE + 16)); CK(h->in_tile.ensure(((size_t)n_tiles + 1) * 8));
    CK(h->in_names.ensure(name_bytes + 16)); CK(h->in_name_off.ensure(((size_t)NN + 1) * 8)); CK(h->in_slots.ensure((size_t)slots * 4));
    CK(h->in_ctl.ensure(64));
    if (nbytes) {
        CK(cudaMemcpyAsync(h->in_text.p, paf, nbytes, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(h->in_text.as<char>() + nbytes, 0, (size_t)n_tiles * CG_IN_TILE - nbytes, st));
    }
    if (name_bytes) CK(cudaMemcpyAsync(h->in_names.p, names->names, name_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->in_name_off.p, names->name_off, ((size_t)NN + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(h->in_slots.p, 0, (size_t)slots * 4, st));
    CK(cudaMemsetAsync(h->in_ctl.p, 0, 64, st));
    CgIngestArgs A{};
    A.text = h->in_text.as<char>(); A.nbytes = nbytes; A.tile_cnt = h->in_tile.as<u64>(); A.n_tiles = n_tiles;
    A.names = h->in_names.as<char>(); A.name_off = h->in_name_off.as<u64>(); A.n_names = NN; A.slots = h->in_slots.as<u32>(); A.slot_mask = slots - 1;
    A.max_support = max_support; A.ctl = h->in_ctl.as<u32>();
    struct Events {                                                 // destroyed on every return path, CK's included
        cudaEvent_t e[8] = {};
        ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
    } evs;
    cudaEvent_t* ev = evs.e;
    for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&ev[i]));
    auto done = [&](int rc) { return rc; };
    auto scan = [&](u64* a, u32 n) { CG_LAUNCH(k_in_scan, 1, 1024, 40 * sizeof(u64), st, a, n); };
    // ---- lines
    CK(cudaEventRecord(ev[0], st));
    if (NN) CG_LAUNCH(k_names_build, (NN + 63) / 64, 64, 0, st, A);
    if (n_tiles) CG_LAUNCH(k_paf_count, n_tiles, 256, 256, st, A);
    scan(A.tile_cnt, n_tiles);
    CK(cudaEventRecord(ev[1], st));
    u64 n_lines = 0;
    CK(cudaMemcpyAsync(&n_lines, A.tile_cnt + n_tiles, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_lines >= (1ull << 31)) { h->err = "more than 2^31 PAF lines in one call"; return done(CG_ERR_CAPACITY); }
    CK(h->in_nl.ensure((n_lines + 1) * 8)); CK(h->in_rec.ensure((n_lines + 1) * sizeof(CgPafRec))); CK(h->in_head.ensure((n_lines + 2) * 4)); CK(h->in_htile.ensure((n_lines / 256 + 3) * 8));
    CK(h->in_scratch.ensure((n_lines + 1) * 8));
    A.nl_pos = h->in_nl.as<u64>(); A.n_lines = n_lines; A.rec = h->in_rec.as<CgPafRec>(); A.head = h->in_head.as<u32>(); A.head_tile = h->in_htile.as<u64>(); A.sort_scratch = h->in_scratch.as<u64>();
    // ---- records, pile boundaries
    CK(cudaEventRecord(ev[2], st));
    if (n_tiles) CG_LAUNCH(k_paf_lines, n_tiles, 256, 256, st, A);
    CK(cudaEventRecord(ev[3], st));
    if (n_lines) CG_LAUNCH(k_paf_parse, (u32)((n_lines + 7) / 8), 256, 8 * 12 * sizeof(u32), st, A);
    CK(cudaEventRecord(ev[4], st));
    const u32 n_hblocks = (u32)((n_lines + 255) / 256);
    if (n_lines) CG_LAUNCH(k_paf_heads, n_hblocks, 256, 256, st, A);
    scan(A.head_tile, n_hblocks);
    u64 n_piles64 = 0;
    u32 ctl[2] = {0, 0};
    CK(cudaMemcpyAsync(&n_piles64, A.head_tile + n_hblocks, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctl, A.ctl, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ev[5], st));
    CK(cudaStreamSynchronize(st));
    if (ctl[0]) {
        h->err = (ctl[0] & CG_IN_FLAG_COLUMNS) ? "PAF ingest: a line has fewer than 12 columns"
               : (ctl[0] & CG_IN_FLAG_NUMBER) ? "PAF ingest: a numeric column is not a non-negative int (stoi throws in the reference)"
               : "PAF ingest: a read name is not in the name table (the reference then works on an empty read)";
        return done(CG_ERR_INVALID_ARG);
    }
    const u32 NP = (u32)n_piles64;
    CK(h->in_pfirst.ensure(((size_t)NP + 1) * 4)); CK(h->in_plast.ensure(((size_t)NP + 1) * 4)); CK(h->in_keep.ensure(((size_t)NP + 2) * 8));
    CK(h->in_pread.ensure(((size_t)NP + 1) * 4)); CK(h->in_pqlen.ensure(((size_t)NP + 1) * 4));
    A.n_piles = NP; A.pile_first = h->in_pfirst.as<u32>(); A.pile_last = h->in_plast.as<u32>(); A.keep = h->in_keep.as<u64>();
    A.pile_read = h->in_pread.as<u32>(); A.pile_qlen = h->in_pqlen.as<u32>();
    CK(cudaEventRecord(ev[6], st));
    if (n_lines) CG_LAUNCH(k_paf_piles, (u32)((n_lines + 255) / 256), 256, 0, st, A);
    if (NP) CG_LAUNCH(k_paf_sizes, (NP + 255) / 256, 256, 0, st, A);
    scan(A.keep, NP);
    u64 n_keep = 0;
    CK(cudaMemcpyAsync(&n_keep, A.keep + NP, 8, cudaMemcpyDeviceToHost, st));
    u32 ctl3[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(ctl3, A.ctl, 12, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctl[1] = ctl3[1];
    if (n_keep >= (1ull << 32)) { h->err = "too many overlaps"; return done(CG_ERR_CAPACITY); }
    CK(h->in_ov.ensure((n_keep + 1) * sizeof(CgOverlapDev))); CK(h->in_res.ensure((n_keep + 1) * 4));
    A.ov = h->in_ov.as<CgOverlapDev>(); A.res = h->in_res.as<u32>();
    // ---- sort + cut: keys in shared memory when the longest pile fits (16 warps per SM up to 1 700 lines)
    const u32 longest = ctl[1];
    size_t smem = (size_t)longest * 8;
    const size_t smem_max = h->smem_optin > 0 ? (size_t)h->smem_optin : 48 * 1024;
    if (smem > smem_max) smem = smem_max;
    if (smem < 1024) smem = 1024;
    A.smem_cap = (u32)(smem / 8);
    CK(cudaFuncSetAttribute(k_paf_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    if (NP) CG_LAUNCH(k_paf_select, std::min<u32>(NP, (u32)h->sms * 32), 32, smem, st, A);
    CK(cudaEventRecord(ev[7], st));
    CK(cudaGetLastError());
    // ---- host copy
    HostPileSet* ps = new HostPileSet();
    ps->pile_read.resize((size_t)NP + 1); ps->pile_qlen.resize((size_t)NP + 1); ps->ovb.resize((size_t)NP + 1);
    ps->res.resize(n_keep + 1); ps->ov.resize(n_keep + 1);
    std::vector<u64> ovb64((size_t)NP + 1, 0);
    cudaError_t e = cudaSuccess;
    if (NP) {
        e = cudaMemcpyAsync(ps->pile_read.data(), A.pile_read, (size_t)NP * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ps->pile_qlen.data(), A.pile_qlen, (size_t)NP * 4, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(ovb64.data(), A.keep, ((size_t)NP + 1) * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && n_keep) e = cudaMemcpyAsync(ps->ov.data(), A.ov, n_keep * sizeof(cg_overlap), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && n_keep) e = cudaMemcpyAsync(ps->res.data(), A.res, n_keep * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { delete ps; h->err = std::string("PAF ingest: ") + cudaGetErrorString(e); return done(CG_ERR_CUDA); }
    for (u32 p = 0; p <= NP; ++p) ps->ovb[p] = (u32)ovb64[p];
    float m[4] = {0, 0, 0, 0}, mp = 0;
    cudaEventElapsedTime(&m[0], ev[0], ev[1]); cudaEventElapsedTime(&m[1], ev[2], ev[5]); cudaEventElapsedTime(&m[2], ev[6], ev[7]);
    cudaEventElapsedTime(&mp, ev[3], ev[4]);
    h->in_ms = m[0] + m[1] + m[2]; h->in_parse_ms = mp; h->in_bytes = nbytes;
    out->n_piles = NP; out->pile_read = ps->pile_read.data(); out->pile_qlen = ps->pile_qlen.data(); out->pile_ov_begin = ps->ovb.data();
    out->overlaps = ps->ov.data(); out->res_matches = ps->res.data(); out->owner_ = ps;
    out->n_lines = ctl3[2];
    return done(CG_OK);
}

void cg_free_pile_set(cg_pile_set* s) {
    if (!s || !s->owner_) return;
    delete static_cast<HostPileSet*>(s->owner_);
    s->owner_ = nullptr;
}

int cg_ingest_stats(const cg_handle* h, float* kernel_ms, float* parse_ms, uint64_t* paf_bytes) {
    if (!h) return CG_ERR_INVALID_ARG;
    if (kernel_ms) *kernel_ms = h->in_ms;
    if (parse_ms) *parse_ms = h->in_parse_ms;
    if (paf_bytes) *paf_bytes = h->in_bytes;
    return CG_OK;
}

// ---- window extraction: phase A of processRead on the device, into the resident batch ----------------------
int cg_set_read_store(cg_handle* h, uint32_t n_store, const uint64_t* store_off, const char* store_bases) {
    if (!h) return CG_ERR_INVALID_ARG;
    if (!store_off || (n_store && store_off[n_store] && !store_bases)) { h->err = "null argument"; return CG_ERR_INVALID_ARG; }
    for (u32 i = 0; i < n_store; ++i) if (store_off[i + 1] < store_off[i]) { h->err = "store_off must be non-decreasing"; return CG_ERR_INVALID_ARG; }
    cudaSetDevice(h->device);
    h->store_resident = false; h->ex_valid = false;
    const u64 nb = store_off[n_store];
    cudaStream_t st = h->lane[0].stream;
    CK(h->ex_store.ensure(nb + 16)); CK(h->ex_store_off.ensure(((size_t)n_store + 1) * 8));
    if (nb) CK(cudaMemcpyAsync(h->ex_store.p, store_bases, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->ex_store_off.p, store_off, ((size_t)n_store + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nb) CG_LAUNCH(k_ex_normalise, (u32)std::min<u64>((nb + 255) / 256, (u64)h->sms * 16), 256, 0, st, h->ex_store.as<char>(), nb);
    CK(cudaStreamSynchronize(st));
    h->ex_store_off_h.assign(store_off, store_off + n_store + 1);
    h->store_n = n_store;
    h->store_resident = true;
    return CG_OK;
}

int cg_upload_piles(cg_handle* h, const cg_piles* P) {
    if (!h) return CG_ERR_INVALID_ARG;
    CG_NVTX("cg_upload_piles");
    const bool resident = P && !P->store_off && !P->store_bases;     // the store cg_set_read_store left on the device
    if (!P || !P->pile_ov_begin || (P->n_piles && (!P->pile_read || !P->pile_qlen)) ||
        (!resident && (!P->store_off || (P->n_store && !P->store_bases)))) {
        h->err = "null argument"; return CG_ERR_INVALID_ARG;
    }
    if (resident && (!h->store_resident || P->n_store != h->store_n)) { h->err = "cg_upload_piles without a store: cg_set_read_store first (same n_store)"; return CG_ERR_STATE; }
    if (!resident) h->store_resident = false;               // the call's own store replaces whatever was there
    if (P->window_size == 0 || P->window_overlap >= P->window_size) { h->err = "window_overlap must be smaller than window_size"; return CG_ERR_INVALID_ARG; }
    if (P->window_size > CG_LEN_MAX) { h->err = "window_size above 6000"; return CG_ERR_CAPACITY; }
    cudaSetDevice(h->device);
    h->uploaded = h->ran = false; h->h2d_pending = false; h->ex_valid = false;
    h->run_gen++;
    const u32 NP = P->n_piles, NS = P->n_store;
    const u64* store_off_h = resident ? h->ex_store_off_h.data() : P->store_off;
    const u64 n_store_bases = store_off_h[NS];
    const u64 n_ov = P->pile_ov_begin[NP];
    if (n_ov && !P->overlaps) { h->err = "null argument"; return CG_ERR_INVALID_ARG; }
    // per pile: coverage scratch and an upper bound of its windows (a window every ws - ovl bases, plus the last one)
    std::vector<u64> cov_off((size_t)NP + 1, 0), cap_off((size_t)NP + 1, 0);
    const u32 step = P->window_size - P->window_overlap;
    for (u32 p = 0; p < NP; ++p) {
        if (P->pile_read[p] >= NS) { h->err = "pile_read out of range"; return CG_ERR_INVALID_ARG; }
        if (P->pile_ov_begin[p + 1] < P->pile_ov_begin[p]) { h->err = "pile_ov_begin must be non-decreasing"; return CG_ERR_INVALID_ARG; }
        cov_off[p + 1] = cov_off[p] + round_up((u64)P->pile_qlen[p] + 2, 4);
        cap_off[p + 1] = cap_off[p] + (u64)P->pile_qlen[p] / step + 2;
    }
    cudaStream_t st = h->lane[0].stream;
    if (!resident) { CK(h->ex_store.ensure(n_store_bases + 16)); CK(h->ex_store_off.ensure(((size_t)NS + 1) * 8)); }
    CK(h->ex_pile_read.ensure(((size_t)NP + 1) * 4)); CK(h->ex_pile_qlen.ensure(((size_t)NP + 1) * 4)); CK(h->ex_pile_ovb.ensure(((size_t)NP + 1) * 4));
    CK(h->ex_ov.ensure((n_ov + 1) * sizeof(CgOverlapDev)));
    CK(h->ex_cov.ensure((cov_off[NP] + 4) * 4)); CK(h->ex_cov_off.ensure(((size_t)NP + 1) * 8)); CK(h->ex_cap_off.ensure(((size_t)NP + 1) * 8));
    CK(h->ex_cap_beg.ensure((cap_off[NP] + 1) * 4)); CK(h->ex_cap_end.ensure((cap_off[NP] + 1) * 4)); CK(h->ex_nwin.ensure(((size_t)NP + 1) * 4));
    CK(h->ex_flags.ensure(16));
    if (!resident) {
        if (n_store_bases) CK(cudaMemcpyAsync(h->ex_store.p, P->store_bases, n_store_bases, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->ex_store_off.p, P->store_off, ((size_t)NS + 1) * 8, cudaMemcpyHostToDevice, st));
    }
    if (NP) {
        CK(cudaMemcpyAsync(h->ex_pile_read.p, P->pile_read, (size_t)NP * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->ex_pile_qlen.p, P->pile_qlen, (size_t)NP * 4, cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemcpyAsync(h->ex_pile_ovb.p, P->pile_ov_begin, ((size_t)NP + 1) * 4, cudaMemcpyHostToDevice, st));
    if (n_ov) CK(cudaMemcpyAsync(h->ex_ov.p, P->overlaps, n_ov * sizeof(CgOverlapDev), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->ex_cov_off.p, cov_off.data(), ((size_t)NP + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->ex_cap_off.p, cap_off.data(), ((size_t)NP + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(h->ex_flags.p, 0, 16, st));
    CgExtractArgs A{};
    A.store_off = h->ex_store_off.as<u64>(); A.store = h->ex_store.as<char>(); A.n_store = NS;
    A.n_piles = NP; A.pile_read = h->ex_pile_read.as<u32>(); A.pile_qlen = h->ex_pile_qlen.as<u32>(); A.pile_ov_begin = h->ex_pile_ovb.as<u32>();
    A.ov = h->ex_ov.as<CgOverlapDev>();
    A.min_support = P->min_support; A.ws = P->window_size; A.ovl = P->window_overlap; A.k = h->p.mer_size;
    A.cov = h->ex_cov.as<u32>(); A.cov_off = h->ex_cov_off.as<u64>(); A.win_cap_off = h->ex_cap_off.as<u64>();
    A.cap_beg = h->ex_cap_beg.as<u32>(); A.cap_end = h->ex_cap_end.as<u32>(); A.n_win = h->ex_nwin.as<u32>();
    A.flags = h->ex_flags.as<u32>();
    struct Events {                                         // destroyed on every return path, CK's included
        cudaEvent_t e[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
    } evs;
    for (cudaEvent_t& x : evs.e) CK(cudaEventCreate(&x));
    cudaEvent_t e0 = evs.e[0], e1 = evs.e[1], e2 = evs.e[2], e2b = evs.e[3], e2c = evs.e[4], e3 = evs.e[5], e2d = evs.e[6];
    CK(cudaEventRecord(e0, st));
    if (n_store_bases && !resident) CG_LAUNCH(k_ex_normalise, (u32)std::min<u64>((n_store_bases + 255) / 256, (u64)h->sms * 16), 256, 0, st, h->ex_store.as<char>(), n_store_bases);
    if (NP) CG_LAUNCH(k_ex_positions, (NP + 3) / 4, 128, 0, st, A);
    CK(cudaEventRecord(e1, st));
    // ---- window counts -> dense windows (host prefix; a few bytes per window)
    std::vector<u32> nwin((size_t)NP + 1, 0), cap_beg(cap_off[NP] + 1), cap_end(cap_off[NP] + 1);
    u32 flags = 0;
    if (NP) CK(cudaMemcpyAsync(nwin.data(), h->ex_nwin.p, (size_t)NP * 4, cudaMemcpyDeviceToHost, st));
    if (cap_off[NP]) {
        CK(cudaMemcpyAsync(cap_beg.data(), h->ex_cap_beg.p, cap_off[NP] * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(cap_end.data(), h->ex_cap_end.p, cap_off[NP] * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaMemcpyAsync(&flags, h->ex_flags.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    auto flag_error = [&](u32 f) {
        if (f & CG_EX_FLAG_CAPACITY) { h->err = "window extraction: a capacity limit of this build was exceeded (sequence over 6000 bases)"; return (int)CG_ERR_CAPACITY; }
        h->err = (f & CG_EX_FLAG_EMPTY_PILE) ? "window extraction: a window lies beyond its stored read (the reference dereferences an empty pile here)"
               : (f & CG_EX_FLAG_SUBSTR) ? "window extraction: an overlap points outside its target read (std::out_of_range in the reference)"
               : "window extraction: an overlap ends beyond qLength, names a read outside the store, or qLength is 0";
        return (int)CG_ERR_INVALID_ARG;
    };
    if (flags) return flag_error(flags);
    h->ex_rwb.assign((size_t)NP + 1, 0);
    u64 Wtot = 0;
    for (u32 p = 0; p < NP; ++p) { Wtot += nwin[p]; h->ex_rwb[p + 1] = (u32)Wtot; }
    if (Wtot >= (1ull << 31)) { h->err = "too many windows"; return CG_ERR_CAPACITY; }
    const u32 W = (u32)Wtot;
    h->ex_wpos.resize(W); h->ex_wend.resize(W);
    std::vector<u32> win_pile(W);
    for (u32 p = 0, w = 0; p < NP; ++p)
        for (u32 i = 0; i < nwin[p]; ++i, ++w) {
            h->ex_wpos[w] = cap_beg[cap_off[p] + i]; h->ex_wend[w] = cap_end[cap_off[p] + i]; win_pile[w] = p;
        }
    CK(h->ex_win_pile.ensure(((size_t)W + 1) * 4)); CK(h->ex_win_beg.ensure(((size_t)W + 1) * 4)); CK(h->ex_win_end.ensure(((size_t)W + 1) * 4));
    CK(h->ex_win_nseq.ensure(((size_t)W + 1) * 4)); CK(h->ex_win_nbytes.ensure(((size_t)W + 1) * 4));
    CK(h->ex_win_base.ensure(((size_t)W + 1) * 8));
    if (W) {
        CK(cudaMemcpyAsync(h->ex_win_pile.p, win_pile.data(), (size_t)W * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->ex_win_beg.p, h->ex_wpos.data(), (size_t)W * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->ex_win_end.p, h->ex_wend.data(), (size_t)W * 4, cudaMemcpyHostToDevice, st));
    }
    A.n_windows = W; A.win_pile = h->ex_win_pile.as<u32>(); A.win_beg = h->ex_win_beg.as<u32>(); A.win_end = h->ex_win_end.as<u32>();
    A.win_nseq = h->ex_win_nseq.as<u32>(); A.win_nbytes = h->ex_win_nbytes.as<u32>();
    A.mode = 0;                                             // count the kept pieces of every window
    CK(cudaEventRecord(e2, st));
    if (W) CG_LAUNCH(k_ex_sizes, (W + 3) / 4, 128, 0, st, A);
    CK(cudaEventRecord(e2b, st));
    // ---- per-window totals -> win_seq_begin / window base offsets (host prefix: the planner needs them on the host anyway)
    std::vector<u32> nseq((size_t)W + 1, 0), nbytes((size_t)W + 1, 0);
    if (W) {
        CK(cudaMemcpyAsync(nseq.data(), h->ex_win_nseq.p, (size_t)W * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(nbytes.data(), h->ex_win_nbytes.p, (size_t)W * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaMemcpyAsync(&flags, h->ex_flags.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (flags) return flag_error(flags);
    h->h_wsb.assign((size_t)W + 1, 0); h->h_wbase.assign((size_t)W + 1, 0); h->h_tlen.assign(W, 0);
    const u32 k = h->p.mer_size;
    for (u32 w = 0; w < W; ++w) {
        const u64 ns = (u64)h->h_wsb[w] + nseq[w];
        if (ns >= (1ull << 32)) { h->err = "too many sequences"; return CG_ERR_CAPACITY; }
        h->h_wsb[w + 1] = (u32)ns;
        h->h_wbase[w + 1] = h->h_wbase[w] + nbytes[w];
        h->h_tlen[w] = h->ex_wend[w] - h->ex_wpos[w] + 1;
    }
    const u64 n_seqs = h->h_wsb[W], n_bases = h->h_wbase[W];
    h->W = W; h->n_seqs = n_seqs; h->n_bases = n_bases;
    CK(h->d_bases.ensure(n_bases + 64)); CK(h->d_seq_off.ensure((n_seqs + 1) * sizeof(u64))); CK(h->d_wsb.ensure(((size_t)W + 1) * sizeof(u32)));
    CK(cudaMemcpyAsync(h->d_wsb.p, h->h_wsb.data(), ((size_t)W + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->ex_win_base.p, h->h_wbase.data(), ((size_t)W + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_seq_off.as<u64>() + n_seqs, &h->h_wbase[W], 8, cudaMemcpyHostToDevice, st));
    CK(h->ex_slot_len.ensure((n_seqs + 1) * 4)); CK(h->ex_slot_src.ensure((n_seqs + 1) * 8)); CK(h->ex_slot_loc.ensure((n_seqs + 1) * 4));
    A.slot_len = h->ex_slot_len.as<u32>(); A.slot_src = h->ex_slot_src.as<u64>(); A.slot_loc = h->ex_slot_loc.as<u32>();
    A.win_seq_begin = h->d_wsb.as<u32>(); A.win_base = h->ex_win_base.as<u64>(); A.seq_off = h->d_seq_off.as<u64>(); A.bases = h->d_bases.as<char>();
    A.mode = 1;                                             // the pieces themselves, one record per kept piece
    CK(cudaEventRecord(e2d, st));
    if (W) CG_LAUNCH(k_ex_sizes, (W + 3) / 4, 128, 0, st, A);
    CK(cudaEventRecord(e2c, st));
    if (W) CG_LAUNCH(k_ex_copy, std::min<u32>(W, (u32)h->sms * 16), 256, 0, st, A);
    CK(cudaEventRecord(e3, st));
    CK(cudaGetLastError());
    plan_chunks(h, false);
    while (h->ev_h2d.size() < h->chunks.size()) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->ev_h2d.push_back(e);
    }
    CK(cudaStreamSynchronize(st));
    float m1 = 0, m2 = 0, m2w = 0, m3 = 0;
    CK(cudaEventElapsedTime(&m1, e0, e1)); CK(cudaEventElapsedTime(&m2, e2, e2b)); CK(cudaEventElapsedTime(&m2w, e2d, e2c));
    CK(cudaEventElapsedTime(&m3, e2c, e3));
    h->ex_ms = m1 + m2 + m2w + m3; h->ex_copy_ms = m3; h->ex_bytes = n_bases;
    h->ex_pile_read_h.assign(P->pile_read, P->pile_read + NP);
    if (!resident) h->ex_store_off_h.assign(P->store_off, P->store_off + NS + 1);
    h->ex_ws = P->window_size; h->ex_ovl = P->window_overlap;
    h->ex_valid = true;
    h->uploaded = true;
    return CG_OK;
}

int cg_download_windows(cg_handle* h, int with_bases, cg_window_set* out) {
    if (!h || !out) return CG_ERR_INVALID_ARG;
    if (!h->uploaded || !h->ex_valid) { h->err = "cg_download_windows needs a batch produced by cg_upload_piles"; return CG_ERR_STATE; }
    cudaSetDevice(h->device);
    cudaStream_t st = h->lane[0].stream;
    HostWindowSet* ws = new HostWindowSet();
    const u32 W = h->W, NP = (u32)h->ex_pile_read_h.size();
    ws->wsb = h->h_wsb; ws->rwb = h->ex_rwb; ws->wpos = h->ex_wpos; ws->wend = h->ex_wend;
    if (ws->wpos.empty()) { ws->wpos.push_back(0); ws->wend.push_back(0); }
    ws->soff.resize(h->n_seqs + 1);
    cudaError_t e = cudaMemcpyAsync(ws->soff.data(), h->d_seq_off.p, (h->n_seqs + 1) * 8, cudaMemcpyDeviceToHost, st);
    if (with_bases) {
        ws->bases.resize(h->n_bases + 1);
        if (e == cudaSuccess && h->n_bases) e = cudaMemcpyAsync(ws->bases.data(), h->d_bases.p, h->n_bases, cudaMemcpyDeviceToHost, st);
    }
    // the reads of the piles, as stored (normalised) on the device
    std::vector<char> store(h->ex_store_off_h.back() + 1);
    if (e == cudaSuccess && h->ex_store_off_h.back()) e = cudaMemcpyAsync(store.data(), h->ex_store.p, h->ex_store_off_h.back(), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { delete ws; h->err = std::string("cg_download_windows: ") + cudaGetErrorString(e); return CG_ERR_CUDA; }
    ws->roff.assign((size_t)NP + 1, 0);
    for (u32 p = 0; p < NP; ++p) {
        const u32 r = h->ex_pile_read_h[p];
        ws->roff[p + 1] = ws->roff[p] + (h->ex_store_off_h[r + 1] - h->ex_store_off_h[r]);
    }
    ws->rbases.resize(ws->roff[NP] + 1);
    for (u32 p = 0; p < NP; ++p) {
        const u32 r = h->ex_pile_read_h[p];
        memcpy(ws->rbases.data() + ws->roff[p], store.data() + h->ex_store_off_h[r], h->ex_store_off_h[r + 1] - h->ex_store_off_h[r]);
    }
    out->batch.n_windows = W; out->batch.win_seq_begin = ws->wsb.data(); out->batch.seq_off = ws->soff.data();
    out->batch.bases = with_bases ? ws->bases.data() : nullptr;
    out->reads.n_reads = NP; out->reads.read_win_begin = ws->rwb.data(); out->reads.read_off = ws->roff.data();
    out->reads.read_bases = ws->rbases.data(); out->reads.win_pos = ws->wpos.data();
    out->reads.window_size = h->ex_ws; out->reads.window_overlap = h->ex_ovl;
    out->win_end = ws->wend.data();
    out->owner_ = ws;
    return CG_OK;
}

void cg_free_window_set(cg_window_set* s) {
    if (s && s->owner_) { delete static_cast<HostWindowSet*>(s->owner_); s->owner_ = nullptr; }
}

int cg_extract_stats(const cg_handle* h, float* kernel_ms, float* copy_ms, uint64_t* pile_bytes) {
    if (!h) return CG_ERR_INVALID_ARG;
    if (kernel_ms) *kernel_ms = h->ex_ms;
    if (copy_ms) *copy_ms = h->ex_copy_ms;
    if (pile_bytes) *pile_bytes = h->ex_bytes;
    return CG_OK;
}

int cg_stage_ms(const cg_handle* h, float ms[CG_N_STAGES], uint32_t launches[CG_N_STAGES]) {
    if (!h) return CG_ERR_INVALID_ARG;
    for (int i = 0; i < CG_N_STAGES; ++i) { if (ms) ms[i] = h->stage_ms[i]; if (launches) launches[i] = h->stage_launches[i]; }
    return CG_OK;
}

// Per-stage text dump of window w of the batch cg_run just processed, in the format of the oracle's / the reference harness's
// dump (the per-stage text the test oracle prints): S, M (solid list), T (template k-mers that survive fill + filter), A (chain), R (mean distances),
// G (regions), g (their segments), c (their consensuses), C (the stitched consensus).  Test instrumentation: the workspaces of a
// chunk are reused by the next one, so the batch must have been a single chunk.  *text is malloc'ed (free() it).
int cg_debug_dump_window(cg_handle* h, uint32_t w, char** text) {
    if (!h || !text) return CG_ERR_INVALID_ARG;
    if (!h->ran || h->chunks.size() != 1 || w >= h->W) { h->err = "cg_debug_dump_window: needs a finished single-chunk run and a window of it"; return CG_ERR_STATE; }
    cudaSetDevice(h->device);
    Lane& L = h->lane[0];
    const u32 nwin = h->chunks[0].nwin;
    auto get = [&](void* dst, const void* src, size_t n) { return n ? cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost) : cudaSuccess; };
    CgWin W;
    CK(get(&W, L.win.as<CgWin>() + w, sizeof W));
    u64 off[5][2];
    for (int a = 0; a < 5; ++a) CK(get(off[a], L.offs.as<u64>() + (size_t)a * (nwin + 1) + w, 16));
    const u64 o_solid = off[0][0], o_slot = off[1][0], o_pos = off[2][0], o_reg = off[3][0], o_arena = off[4][0];
    const u32 N = W.n_seqs, C = W.n_cand, nA = W.n_chain, k = h->p.mer_size;
    std::vector<u32> sk(W.n_solid), sc(W.n_solid), slot_kmer(C), rel(nA ? nA : 1);
    std::vector<u16> anchors(W.n_alive), chain(nA), pos((size_t)N * C);
    CK(get(sk.data(), L.solid_k.as<u32>() + o_solid, 4 * (size_t)W.n_solid)); CK(get(sc.data(), L.solid_c.as<u32>() + o_solid, 4 * (size_t)W.n_solid));
    CK(get(slot_kmer.data(), L.slot_kmer.as<u32>() + o_slot, 4 * (size_t)C));
    CK(get(anchors.data(), L.anchors.as<u16>() + o_slot, 2 * (size_t)W.n_alive)); CK(get(chain.data(), L.chain.as<u16>() + o_slot, 2 * (size_t)nA));
    CK(get(rel.data(), L.rel.as<u32>() + o_slot, 4 * (size_t)(nA ? nA - 1 : 0)));
    CK(get(pos.data(), L.pos.as<u16>() + o_pos, 2 * (size_t)N * C));
    std::vector<u64> soff(N + 1);
    CK(get(soff.data(), h->d_seq_off.as<u64>() + W.seq_begin, 8 * (size_t)(N + 1)));
    std::vector<char> bases((size_t)(soff[N] - soff[0]) + 1);
    CK(get(bases.data(), h->d_bases.as<char>() + soff[0], (size_t)(soff[N] - soff[0])));
    const u64 base0 = soff[0];
    for (u64& x : soff) x -= base0;
    const u32 n_regions = nA ? nA + 1 : 2;
    std::vector<CgRegion> regs(W.n_regions);
    CK(get(regs.data(), L.regions.as<CgRegion>() + o_reg, sizeof(CgRegion) * (size_t)W.n_regions));
    std::vector<char> arena((size_t)(off[4][1] - off[4][0]) + 1);
    CK(get(arena.data(), L.arena.as<u8>() + o_arena, (size_t)(off[4][1] - off[4][0])));
    std::string t;
    char line[64];
    const int S = (int)h->p.common_kmers < (int)N / 2 ? (int)h->p.common_kmers : (int)N / 2;
    snprintf(line, sizeof line, "S %d\nM %u", S, W.n_solid); t += line;
    for (u32 i = 0; i < W.n_solid; ++i) { snprintf(line, sizeof line, " %u:%u", sk[i], sc[i]); t += line; }
    snprintf(line, sizeof line, "\nT %u", W.n_alive); t += line;
    for (u32 i = 0; i < W.n_alive; ++i) { snprintf(line, sizeof line, " %u", slot_kmer[anchors[i]]); t += line; }
    snprintf(line, sizeof line, "\nA %u", nA); t += line;
    for (u32 i = 0; i < nA; ++i) { snprintf(line, sizeof line, " %u", slot_kmer[chain[i]]); t += line; }
    snprintf(line, sizeof line, "\nR %u", nA ? nA - 1 : 0u); t += line;
    for (u32 i = 0; i + 1 < nA; ++i) { snprintf(line, sizeof line, " %lld", (long long)rel[i]); t += line; }
    snprintf(line, sizeof line, "\nG %u\n", n_regions); t += line;
    std::string stacked;
    if (W.n_regions) {                                      // 0: MSABMAAC bailed out (regions < minAnchors), nothing is dumped per region
        CgWinView v;
        v.seq_off = soff.data(); v.pos = pos.data(); v.chain = chain.data(); v.rel = rel.data(); v.N = N; v.C = C; v.nA = nA;
        for (u32 g = 0; g < W.n_regions; ++g) {
            std::string segs;
            u32 n = 0;
            for (u32 r = 0; r < N; ++r) {
                u32 st = 0, ln = 0;
                if (cg_eval_segment(v, g, r, &st, &ln)) { segs += ' '; segs.append(bases.data() + soff[r] + st, ln); ++n; }
            }
            snprintf(line, sizeof line, "g %u %u", g, n); t += line; t += segs; t += '\n';
            const CgRegion& R = regs[g];
            if (R.kind == CG_REG_EMPTY) continue;
            std::string cons = R.kind == CG_REG_COPY ? std::string(bases.data() + soff[R.read] + R.start, R.len) : std::string(arena.data() + R.arena_off, R.cons_len);
            snprintf(line, sizeof line, "c %u ", g); t += line; t += cons; t += '\n';
            stacked += cons;
        }
    }
    t += "C "; t += stacked; t += '\n';
    (void)k;
    *text = (char*)malloc(t.size() + 1);
    if (!*text) return CG_ERR_OUT_OF_MEMORY;
    memcpy(*text, t.c_str(), t.size() + 1);
    return CG_OK;
}

int cg_get_kernel_stats(const cg_handle* h, cg_kernel_stats* out) {
    if (!h || !out) return CG_ERR_INVALID_ARG;
    *out = h->kstats;
    return CG_OK;
}

int cg_run_ms(const cg_handle* h, float* ms) {
    if (!h || !ms) return CG_ERR_INVALID_ARG;
    *ms = h->run_ms;
    return CG_OK;
}

int cg_chunk_count(const cg_handle* h) { return h && h->uploaded ? (int)h->chunks.size() : 0; }

int cg_get_counters(const cg_handle* h, cg_counters* out) {
    if (!h || !out) return CG_ERR_INVALID_ARG;
    *out = h->counters;
    return CG_OK;
}

}  // extern "C"
