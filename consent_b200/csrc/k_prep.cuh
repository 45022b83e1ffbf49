// k_prep.cuh — per-chunk planning, offset scans and ASCII -> 2-bit packing.
#pragma once
#include "cg_common.cuh"

__device__ __forceinline__ u64 cg_round_up(u64 v, u64 m) { return (v + m - 1) / m * m; }

// Every sequence of the batch: offsets non-decreasing, length within CG_LEN_MAX (cg_upload reads the flags back).
__global__ void k_validate(const u64* seq_off, u64 n_seqs, u32* flags) {
    u32 f = 0;
    for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < n_seqs; s += (u64)gridDim.x * blockDim.x) {
        const u64 a = seq_off[s], b = seq_off[s + 1];
        if (b < a) f |= CG_FLAG_BAD_OFFSETS;
        else if (b - a > CG_LEN_MAX) f |= CG_FLAG_CAPACITY;
    }
    if (f) atomicOr(flags, f);
}

// 2-bit input (cg_set_option "input_2bit"): base i of the batch is bits 2 (i & 3) .. of byte i >> 2, A 0 C 1 G 2 T 3 (cg_pack_bases_2bit,
// the way the reference's own read index holds reads: src/utils.cpp:21-54).  Expanded to the ASCII the kernels read; 4 bases per
// thread and step for the aligned body (one byte in, one 32-bit word out).
__global__ void k_unpack_2bit(const u8* packed, char* bases, u64 b0, u64 b1) {
    const u64 a0 = (b0 + 3) & ~(u64)3, a1 = b1 & ~(u64)3;              // aligned body [a0, a1)
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nth = (u64)gridDim.x * blockDim.x;
    if (a1 > a0)
        for (u64 q = (a0 >> 2) + tid; q < (a1 >> 2); q += nth) {
            const u32 v = packed[q];
            const u32 lut = 0x54474341u;                              // 'A' 'C' 'G' 'T' as bytes 0..3
            const u32 w = ((lut >> (8 * (v & 3u))) & 0xffu) | (((lut >> (8 * ((v >> 2) & 3u))) & 0xffu) << 8) |
                          (((lut >> (8 * ((v >> 4) & 3u))) & 0xffu) << 16) | (((lut >> (8 * ((v >> 6) & 3u))) & 0xffu) << 24);
            *(u32*)(bases + 4 * q) = w;
        }
    if (tid < 8) {                                                     // ragged head / tail (also the whole range when it is tiny)
        const u64 h1 = a1 > a0 ? a0 : b1, t0 = a1 > a0 ? a1 : b1;
        for (u64 i = b0 + tid; i < h1; i += 8) bases[i] = "ACGT"[(packed[i >> 2] >> (2 * (i & 3))) & 3u];
        for (u64 i = t0 + tid; i < b1; i += 8) bases[i] = "ACGT"[(packed[i >> 2] >> (2 * (i & 3))) & 3u];
    }
}

// One thread per window: sizes and arena capacities (written into the off_* arrays, scanned by k_scan).
__global__ void k_plan(CgChunk c) {
    u32 w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= c.nwin) return;
    u32 gw = c.w0 + w;
    u32 s0 = c.win_seq_begin[gw], s1 = c.win_seq_begin[gw + 1];
    u32 N = s1 - s0, k = c.k;
    u64 nocc = 0;
    bool bad = N > CG_N_MAX;
    for (u32 s = s0; s < s1; ++s) {
        u64 len = c.seq_off[s + 1] - c.seq_off[s];
        if (len > CG_LEN_MAX) bad = true;
        if (len >= k) nocc += len - k + 1;
    }
    u32 tlen = (u32)(c.seq_off[s0 + 1] - c.seq_off[s0]);
    u32 tk = tlen >= k ? tlen - k + 1 : 0;
    u64 nb = c.seq_off[s1] - c.seq_off[s0];
    if (tk > CG_TK_MAX) bad = true;
    // A window over a limit of this build is not processed: it comes back as its raw template with status CG_WINDOW_ERROR
    // (the reference has no such limits; the batch goes on).  Every later kernel sees a pile of one sequence without k-mers.
    if (bad) { N = 1; tk = 0; nocc = 0; nb = tlen; }
    CgWin W;
    W.seq_begin = s0; W.n_seqs = N; W.tlen = tlen; W.tk = tk;
    int S = (int)c.common < (int)N / 2 ? (int)c.common : (int)N / 2;     // src/correctionMSA.cpp:31
    W.S = (u32)S;
    W.n_cand = W.n_alive = W.n_chain = W.n_regions = W.n_solid = 0;
    W.stitched_len = W.final_len = W.final_beg = 0; W.status = 0; W.bad = bad ? 1u : 0u;
    W.n_occ = (u32)nocc; W.n_bases = (u32)nb;
    c.win[w] = W;
    c.off_solid[w] = cg_round_up(nocc / c.solid, 4);
    c.off_slot[w] = cg_round_up(tk, 8);
    c.off_pos[w] = cg_round_up((u64)tk * N, 8);
    c.off_reg[w] = (u64)tk + 2;
    c.off_arena[w] = cg_round_up(nb + N, 16);
}

// Exclusive in-place scan of up to 5 arrays of n entries (+ total at [n]); block b handles array b.
__global__ void k_scan(u64* a0, u64* a1, u64* a2, u64* a3, u64* a4, u32 n) {
    CG_DYN_SMEM(smem);
    u64* part = (u64*)smem;                       // blockDim.x entries
    u64* a = blockIdx.x == 0 ? a0 : blockIdx.x == 1 ? a1 : blockIdx.x == 2 ? a2 : blockIdx.x == 3 ? a3 : a4;
    u32 T = blockDim.x, t = threadIdx.x;
    u32 per = (n + T - 1) / T;
    u32 b = t * per, e = b + per < n ? b + per : n;
    u64 s = 0;
    if (a) for (u32 i = b; i < e; ++i) s += a[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        u64 run = 0;
        for (u32 i = 0; i < T; ++i) { u64 v = part[i]; part[i] = run; run += v; }
        if (a) a[n] = run;
    }
    __syncthreads();
    if (a) {
        u64 run = part[t];
        for (u32 i = b; i < e; ++i) { u64 v = a[i]; a[i] = run; run += v; }
    }
}

// One CTA per window, one thread per 16-base word of the pile (the words of a window are contiguous: sequence s starts at word
// (seq_off[s] >> 4) + s; a binary search over the window's sequence table, staged in shared memory, finds a word's sequence).
// A thread reads its 16 bases as five aligned 32-bit words, re-aligns them with funnel shifts and converts four bases at a
// time: code = ((b >> 1) & 3) ^ (its own high bit) maps A C G T to 0 1 2 3, anything else is caught by mapping the code back to
// its letter; a multiply gathers the four 2-bit codes into one byte.
// word: base i of the word at bits [31-2i, 30-2i] (so a k-mer read off the word is BMEAN's code, utils.cpp:18-30)
// tag : (read index in the window) << 16 | (word index in the read) << 4 | (k-mer starts in the word - 1); ~0 = none
#define CG_PACK_THREADS 256u
#define CG_PACK_SMEM_BYTES (4u * (CG_N_MAX + 2u))
__global__ void __launch_bounds__(CG_PACK_THREADS) k_pack(CgChunk c) {
    CG_DYN_SMEM(smem);
    u32* s_w0 = (u32*)smem;                          // first word of each sequence, relative to the window's first word
    const CgWin W = c.win[blockIdx.x];
    if (W.bad) return;                               // over a limit: untouched (k_plan)
    const u32 N = W.n_seqs, tid = threadIdx.x;
    const u64 gw0 = cg_pword(c.seq_off, W.seq_begin);
    for (u32 s = tid; s <= N; s += CG_PACK_THREADS) s_w0[s] = (u32)(cg_pword(c.seq_off, W.seq_begin + s) - gw0);
    __syncthreads();
    const u32 nw = s_w0[N];
    const u8* base = (const u8*)c.bases;
    u32 bad = 0;
    for (u32 g = tid; g < nw; g += CG_PACK_THREADS) {
        u32 lo = 0, hi = N;                          // last sequence whose first word is <= g
        while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (s_w0[mid] <= g) lo = mid; else hi = mid; }
        const u32 r = lo, wi = g - s_w0[r];
        const u64 off = c.seq_off[W.seq_begin + r];
        const u32 len = (u32)(c.seq_off[W.seq_begin + r + 1] - off);
        const u32 nkm = len >= c.k ? len - c.k + 1 : 0;
        const u32 left = len > 16 * wi ? len - 16 * wi : 0;        // bases of the sequence from this word on
        u32 word = 0;
        if (left) {
            const u64 a = off + 16ull * wi;
            const u32* p = (const u32*)(base + (a & ~3ull));
            const u32 sh = 8u * (u32)(a & 3ull);
            u32 raw[5];
#pragma unroll
            for (u32 i = 0; i < 5; ++i) raw[i] = (4 * i < left + (u32)(a & 3ull)) ? p[i] : 0x41414141u;     // words holding none of the bases are not read
#pragma unroll
            for (u32 i = 0; i < 4; ++i) {
                u32 x = __funnelshift_r(raw[i], raw[i + 1], sh);                               // bases 4i .. 4i + 3 of the word, first base lowest
                const u32 have = left > 4 * i ? (left - 4 * i < 4 ? left - 4 * i : 4u) : 0u;
                const u32 keep = have >= 4 ? 0xffffffffu : (1u << (8u * have)) - 1u;
                x = (x & keep) | (0x41414141u & ~keep);                                        // past the end: 'A' (code 0), never flagged
                u32 t = (x >> 1) & 0x03030303u;
                t ^= (t >> 1) & 0x01010101u;
                const u32 y = (t | (t >> 4)) & 0x00ff00ffu;
                const u32 sel = (y | (y >> 8)) & 0xffffu;
                bad |= __byte_perm(0x54474341u, 0u, sel) ^ x;                                  // 'A' 'C' 'G' 'T' by code against what was read
                word |= ((t * 0x40100401u) >> 24) << (24u - 8u * i);
            }
        }
        const u32 nv = nkm > 16 * wi ? (nkm - 16 * wi < 16 ? nkm - 16 * wi : 16) : 0;
        c.pwords[gw0 - c.pword_base + g] = word;
        c.ptags[gw0 - c.pword_base + g] = nv ? ((r << 16) | (wi << 4) | (nv - 1)) : CG_NONE32;
    }
    if (bad) atomicOr(c.flags, (u32)CG_FLAG_BAD_BASE);
}
