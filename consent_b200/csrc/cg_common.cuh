// cg_common.cuh — shared definitions of the B200 (sm_100a) CONSENT per-window correction path.
//
// Pipeline (one "chunk" = a few thousand windows whose workspaces fit the HBM budget):
//   k_plan    per-window sizes -> arena capacities          (host orchestration replaced: none)
//   k_scan    exclusive scans -> arena offsets
//   k_pack    ASCII -> 2-bit words + per-word tags          (reads are 2-bit stored upstream too: src/utils.cpp:21-54)
//   k_index   k-mer counts, solid list, template anchors    (BMEAN/bmean.cpp:43-114, 220-234)
//   k_chain   anchor pair scores + longest ordered chain    (bmean.cpp:161-216, 239-260)
//   k_split   distance statistics + region classification   (bmean.cpp:264-295, 324-418, 476-554)
//   k_poa     segmented partial-order alignment + vote      (bmean.cpp:585-698, spoa graph.cpp / *_alignment_engine*)
//   k_polish  stitch, weight, de Bruijn polish              (bmean.cpp:702-733, src/correctionMSA.cpp:6-49,
//                                                            src/correctionDBG.cpp, src/DBG.cpp)
//   k_gather  dense result arrays
//
// The same sources compile under tests/emu/simt_emu.h (-DCG_EMU, test infrastructure) so the kernels
// can be checked against the oracle on a CPU-only box.  The product build is nvcc only.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifndef CG_EMU
#include <cuda_runtime.h>
#define CG_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define CG_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define CG_NOINLINE __noinline__
// sm_100a-only pieces (the emulator header defines the same names for the CPU suite): SM id, L1 prefetch, TMA bulk copies.
__device__ __forceinline__ unsigned cg_smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ void cg_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// Two bulk copies HBM -> shared memory by the TMA engine (cp.async.bulk -> UBLKCP), `bytes` each (a multiple of 16; both sides 16-byte
// aligned), completing on one mbarrier.  cg_bulk_issue2: ONE thread; cg_bulk_wait: every thread, after a __syncthreads() behind the issue.
__device__ __forceinline__ void cg_bulk_issue2(void* dst0, const void* src0, void* dst1, const void* src1, unsigned bytes, uint64_t* bar) {
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(2u * bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst0)), "l"(src0), "r"(bytes), "r"(bar_a) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst1)), "l"(src1), "r"(bytes), "r"(bar_a) : "memory");
}
__device__ __forceinline__ void cg_bulk_wait(uint64_t* bar) {
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar_a) : "memory");
}
#else
#define CG_NOINLINE __attribute__((noinline))
#endif

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int16_t i16;
typedef int32_t i32;

#define CG_FULL 0xffffffffu
#define CG_NONE32 0xffffffffu
#define CG_NONE16 0xffffu

// ---- limits of this build (exceeding one -> CG_ERR_CAPACITY, never a silent wrong answer)
#define CG_KMAX 9u            // k_index counts k-mers of up to this length in a direct-addressed table (4^k <= 2^18); longer ones
                              // (k <= CG_KMAX_HASHED, the reference's own limit: 1 << 2k, BMEAN/bmean.cpp:46) in a hash table
#define CG_KMAX_HASHED 15u
#define CG_IDX_SOLID_CAP 32768u   // hashed path: solid k-mers of one window that can be sorted in shared memory
#define CG_TAB_BITS 15u       // 32768 u32 counters (128 KB of shared memory) per pass
#define CG_TK_MAX 2047u       // template k-mers per window (anchor hash has 4096 slots)
#define CG_N_MAX 4095u        // sequences per window
#define CG_LEN_MAX 6000u      // bases per sequence (int16 score matrix: 5*L < 32767)
#define CG_PW_CAP 8192u       // packed words (and tags) of one pile staged in shared memory

// ---- error / event flags (device -> host), OR-ed into CgChunk::flags[0]
enum {
    CG_FLAG_BAD_BASE = 1u,
    CG_FLAG_CAPACITY = 2u,
    CG_FLAG_INTERNAL = 4u,
    CG_FLAG_BAD_OFFSETS = 8u,
};

// ---- region kinds (easy_consensus cases, bmean.cpp:603-641)
enum { CG_REG_EMPTY = 0, CG_REG_COPY = 1, CG_REG_POA = 2 };

struct CgRegion {
    u32 kind;
    u32 n;            // sequences kept in the region
    u32 read, start, len;   // first kept segment (the consensus itself for CG_REG_COPY)
    u32 sum_len;      // total bases of the kept segments (bounds the POA graph and the consensus)
    u32 max_len;      // longest kept segment (routes the job to a scratch tier)
    u32 arena_off;    // CG_REG_POA: offset of the region consensus inside the window's arena slice
    u32 cons_len;     // CG_REG_POA: length of the region consensus (written by k_poa)
};

struct CgWin {
    u32 seq_begin;    // first sequence of the window (index into seq_off)
    u32 n_seqs;       // N
    u32 tlen;         // template length
    u32 tk;           // template k-mers = max(tlen-k+1, 0)
    u32 S;            // bmeanSup = min(commonKMers, N/2)            (src/correctionMSA.cpp:31)
    u32 n_cand;       // C: template positions with S <= count <= N  (stride of the position table)
    u32 n_alive;      // A: candidate anchors after fill+filter      (get_template)
    u32 n_chain;      // anchors of the longest ordered chain
    u32 n_regions;
    u32 n_solid;
    u32 stitched_len;
    u32 final_len;
    u32 final_beg;    // start of the final string inside its work slice
    u32 status;       // CG_WINDOW_*
    u32 bad;          // window skipped (capacity)
    u32 n_occ;        // k-mer occurrences of the pile
    u32 n_bases;
};

struct CgPoaScratch {     // one per resident POA warp (global memory, L1/L2 resident for small graphs)
    u32 vcap, ecap, scap, alncap, ncap;
    u64 hcap;             // score-matrix cells
    u8* letter; u8* in0; u8* nal; u8* leader; u8* marks; u8* check;
    u16* nseq; u16* aligned; u16* rank_of; u16* r2n;
    u32* in_head; u32* in_tail; u32* rdesc;
    u16* e_pred; u32* e_next;
    u32* stack;
    i32* aln_node; i32* aln_pos;
    u16* seg_read; u16* seg_start; u16* seg_len;
    i16* H;
};

struct CgCountersDev {
    u64 anchors, regions, poa_graphs, alignments, dp_cells, dp_pred_cells, solid_kmers, consensus_bytes, fallback_windows;
    u64 sequences, bases, windows;
    u64 tier_cells[4], tier_pred[4];      // k_poa2 tiers C1, G, W1, W2
    u64 error_windows;                    // windows over a limit of this build (status CG_WINDOW_ERROR)
};

struct CgChunk {
    // batch, resident in HBM (cg_upload)
    const char* bases;
    const u64* seq_off;
    const u32* win_seq_begin;
    u32 w0, nwin;                 // windows [w0, w0 + nwin) of the batch
    // parameters (src/correctionMSA.cpp:31-32,43-45)
    u32 k, solid, common, min_anchors;
    // 2-bit pile: word index of sequence s = seq_off[s]/16 + s - pword_base
    u32* pwords; u32* ptags; u64 pword_base;
    // per window
    CgWin* win;
    u64* off_solid; u64* off_slot; u64* off_pos; u64* off_reg; u64* off_arena;    // [nwin + 1]
    // arenas
    u32* solid_k; u32* solid_c;
    u16* slot_tpos; u32* slot_kmer; u16* anchors; u16* chain; u32* rel;
    u16* pos;
    CgRegion* regions;
    u8* arena;                    // region consensuses
    u8* fin;                      // per window: 3 slices of (2*n_bases+64) bytes: consensus, temp, path
    u32* visited;                 // per window ceil(solid_cap/32) words (bit per solid k-mer)
    // POA job queues (k_poa2.cuh tiers): 0 tier C1, 1 tier G, 2 tier W1 (filled by k_split), then the re-queue chain
    // (consent_b200.cu).  qctl[4*t + {0,1,2,3}] = jobs at the front of queue t, next job to take, jobs at
    // the back, capacity of the array.  A job that outgrows a tier is pushed to the front of that tier's overflow queue.
    u32* qctl;
    uint2* jobs_s; uint2* jobs_m; uint2* jobs_w;
    // k_index, k > CG_KMAX: open-addressing count tables in global memory, one per SM (slot = %smid), idx_cap entries each (power of two)
    u32* idx_keys; u32* idx_counts; u32 idx_cap, idx_slots;
    // status
    u32* flags;
    CgCountersDev* counters;
};

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ u32 cg_lane() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 cg_warp() { return threadIdx.x >> 5; }

__device__ __forceinline__ u64 cg_pword(const u64* seq_off, u32 s) { return (seq_off[s] >> 4) + (u64)s; }

// A0 C1 G2 T3 (BMEAN/utils.cpp:18-30); anything else -> 4
__device__ __forceinline__ u32 cg_base_code(u8 b) {
    u32 x = (b >> 1) & 3u;
    x ^= x >> 1;
    bool ok = (b == 'A') | (b == 'C') | (b == 'G') | (b == 'T');
    return ok ? x : 4u;
}

// k-mer starting at base `b` (0..15) of word w0, continuing into w1; bases are stored most significant first.
__device__ __forceinline__ u32 cg_kmer_at(u32 w0, u32 w1, u32 b, u32 k) {
    return __funnelshift_l(w1, w0, 2u * b) >> (32u - 2u * k);
}

__device__ __forceinline__ u32 cg_warp_sum(u32 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(CG_FULL, v, d);
    return v;
}
__device__ __forceinline__ u32 cg_warp_max(u32 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { u32 o = __shfl_xor_sync(CG_FULL, v, d); v = o > v ? o : v; }
    return v;
}
__device__ __forceinline__ u64 cg_warp_max64(u64 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { u64 o = __shfl_xor_sync(CG_FULL, v, d); v = o > v ? o : v; }
    return v;
}
// inclusive prefix sum over the warp
__device__ __forceinline__ u32 cg_warp_scan(u32 v) {
    u32 lane = cg_lane();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(CG_FULL, v, d); if (lane >= (u32)d) v += o; }
    return v;
}

// Block-wide exclusive scan of one value per thread; `scratch` holds >= 33 u32; returns the exclusive
// prefix and writes the block total to *total.  Contains two __syncthreads().
__device__ __forceinline__ u32 cg_block_scan(u32 v, u32* scratch, u32* total) {
    u32 lane = cg_lane(), w = cg_warp(), nw = (blockDim.x + 31u) >> 5;
    u32 inc = cg_warp_scan(v);
    if (lane == 31) scratch[w] = inc;
    __syncthreads();
    if (w == 0) {
        u32 x = lane < nw ? scratch[lane] : 0u;
        u32 xi = cg_warp_scan(x);
        scratch[lane] = xi - x;
        if (lane == 31) scratch[32] = xi;
    }
    __syncthreads();
    u32 r = inc - v + scratch[w];
    *total = scratch[32];
    __syncthreads();
    return r;
}

// comparable(x, mean)  BMEAN/bmean.cpp:286-295 (fp64 on purpose: mean may be 0 -> inf/NaN semantics)
__host__ __device__ __forceinline__ bool cg_comparable_mean(double x, double mean) {
    if (fabs(x - mean) < 5) return true;
    if (x / mean < 0.5 || x / mean > 2) return false;
    return true;
}
// comparable(x, deciles)  bmean.cpp:264-282
__device__ __forceinline__ bool cg_comparable_dec(double x, double lo, double hi) {
    if (x < hi + 5) {
        if (x > lo - 5) return true;
        if (x / lo < 0.5) return false;
        return true;
    } else {
        if (x / hi > 2) return false;
        return true;
    }
}
// The same two tests in integers, for the integer-valued arguments the path only ever has (lengths, positions, integer means; all below
// 2^31).  Exactly the fp64 results: |x - m| < 5 is exact; x / m < 0.5 <=> 2 x < m and x / m > 2 <=> x > 2 m, because a quotient of two
// such integers that is not exactly 0.5 (or 2) differs from it by at least 1 / (2 m) >= 2^-32, far more than the rounding of the
// division (2^-53 relative); m = 0: x / 0 = inf (> 2: not comparable) for x > 0 and |0 - 0| < 5 for x = 0, which the integer tests
// reproduce.  (One fp64 division per call was ~100 instructions of slow path on every read of every region.)
__host__ __device__ __forceinline__ bool cg_comparable_mean_u(u32 x, u32 m) {
    const u32 d = x > m ? x - m : m - x;
    if (d < 5u) return true;
    return !(2ull * x < (u64)m || (u64)x > 2ull * m);
}
__device__ __forceinline__ bool cg_comparable_dec_u(u32 x, u32 lo, u32 hi) {
    if ((u64)x < (u64)hi + 5u) {
        if ((u64)x + 5u > (u64)lo) return true;
        return !(2ull * x < (u64)lo);
    }
    return !((u64)x > 2ull * hi);
}

// Window view used by k_split / k_poa to evaluate split_reads (bmean.cpp:476-554) for one (region, read).
struct CgWinView {
    const u64* seq_off;   // &seq_off[seq_begin]
    const u16* pos;       // position table [N][C], value = position + 1, 0 = absent
    const u16* chain;     // slots of the chain anchors
    const u32* rel;       // integer means (average_distance_next_anchor)
    u32 N, C, nA;
};
__host__ __device__ __forceinline__ u32 cg_seq_len(const CgWinView& v, u32 r) { return (u32)(v.seq_off[r + 1] - v.seq_off[r]); }

// -> kept?  start/len of read r's piece in region g.
__host__ __device__ __forceinline__ bool cg_eval_segment(const CgWinView& v, u32 g, u32 r, u32* start, u32* len) {
    u32 len_r = cg_seq_len(v, r);
    if (v.nA == 0) {                                  // bmean.cpp:477-481 : [ [], Reads ]
        if (g != 1) return false;
        *start = 0; *len = len_r;
        return true;
    }
    const u16* prow = v.pos + (size_t)r * v.C;
    const u16* trow = v.pos;                          // read 0 = template
    if (g == 0) {                                     // :487-494
        u32 ap = prow[v.chain[0]];
        if (!ap) return false;
        u32 a = ap - 1;
        u32 l = a > len_r ? len_r : a;
        const i32 mi = (i32)trow[v.chain[0]] - 1;
        *start = 0; *len = l;
        if (mi < 0) return cg_comparable_mean((double)l, (double)mi) && l != 0;       // (the template holds every chain anchor: not reached)
        return cg_comparable_mean_u(l, (u32)mi) && l != 0;
    }
    if (g == v.nA) {                                  // :497-503
        u32 ap = prow[v.chain[v.nA - 1]];
        if (!ap) return false;
        u32 a = ap - 1;
        u32 l = len_r - a;
        u32 len0 = cg_seq_len(v, 0);
        const size_t ms = (size_t)len0 - (size_t)(int64_t)((i32)trow[v.chain[v.nA - 1]] - 1);
        *start = a; *len = l;
        if (ms >> 31) return cg_comparable_mean((double)l, (double)ms) && l != 0;      // (an anchor beyond the template's end: not reached)
        return cg_comparable_mean_u(l, (u32)ms) && l != 0;
    }
    u32 p1 = prow[v.chain[g - 1]], p2 = prow[v.chain[g]];   // :506-520
    if (!p1 || !p2) return false;
    u32 l = p2 - p1;
    *start = p1 - 1; *len = l;
    const u32 mr = v.rel[g - 1];
    if (mr >> 31) return cg_comparable_mean((double)l, (double)mr) && l != 0;
    return cg_comparable_mean_u(l, mr) && l != 0;
}
