// k_reanchor.cuh — consensus re-anchoring on the device.
//
// Batched equivalent of alignConsensus (src/correctionAlignment.cpp:47-139): every window consensus of a read is
// located on the progressively corrected read by a local alignment (StripedSmithWaterman::Aligner defaults, match 2 /
// mismatch 2 / gap open 3 / gap extend 1: BMEAN/Complete-Striped-Smith-Waterman-Library/src/ssw_cpp.cpp:419-426,365-403,
// ssw.c:788-850), overlaps with the previous window are arbitrated by solid k-mers (correctionAlignment.cpp:93-118)
// and the aligned stretch of the read is replaced by the upper-cased consensus.
//
// Work layout: windows of one read depend on each other (each alignment sees the previous replacements), reads do
// not -> ONE WARP PER READ, persistent warps pulling reads (longest first) from a counter.
//
// The local alignment itself (ssw.c's striped SSE2 kernels sw_sse2_byte/word :158-575) is re-cut for a warp: every
// lane owns a block of up to CG_RA_RPL consecutive query rows whose H and E live in registers, reference columns
// flow through the lanes as a skewed wavefront (lane l works on column t-l at step t) and the only communication is
// one packed 32-bit shuffle per step carrying (H, F) of the block's last row and the column's base.  What the SSE2
// kernels return besides the score is reproduced exactly: ref_end = first column whose maximum reaches the final
// score, read_end = smallest query row holding it in that column (ssw.c:297-332), begin = the same scan over the
// reversed query prefix and the reference prefix read backwards, first column reaching the score (ssw.c:836-850,318).
// The CIGAR of the main alignment is never read by alignConsensus and is not computed; the one of the arbitration's
// sub-alignment is needed only through its I/D totals (getIndels, correctionAlignment.cpp:28-45) and comes from a
// faithful banded DP (banded_sw, ssw.c:577-758, band storage restated slot for slot because its edge handling is
// observable) run by lane 0 — it touches a few thousand cells on ~10 % of the windows.
#pragma once
#include "cg_common.cuh"

#define CG_RA_RPL 20u            // query rows per lane and band (640 rows per band)
#define CG_RA_WARPS 4u           // warps per CTA
#define CG_RA_PROF_BYTES (5u * (CG_RA_RPL / 2u) * 32u * 4u)      // query profile of one band: 5 reference letters x RPL / 2 row pairs x 32 lanes
#define CG_RA_QMAX 8000          // H must fit 14 bits of the wavefront message: 2 * rows <= 16383

enum { CG_RA_FLAG_DEGENERATE = 16u, CG_RA_FLAG_CAPACITY = 32u, CG_RA_FLAG_TRACEBACK = 64u };

struct CgReanchorArgs {
    // per-window results of the correction path
    const char* cons; const u64* cons_off; const u64* solid_off; const u32* solid_kmer;
    // templates: the resident batch (bases / seq_off / win_seq_begin) or, when that is null, compact copies (tpl / tpl_off)
    const char* bases; const u64* seq_off; const u32* win_seq_begin;
    const char* tpl; const u64* tpl_off;
    // reads
    u32 n_reads; const u32* order; const u32* read_win_begin; const u64* read_off; const char* read_bases; const u32* win_pos;
    u32 ws, ov, k;
    // outputs: per read a private slice of `head` that ends up holding the corrected read
    char* head; const u64* head_off; u32* out_len;
    // per resident warp scratch
    u8* scratch; u64 scratch_stride; u32 maxL, rmax; u64 dir_cap;
    u32 lines_in_smem;           // the three rolling lines of the banded DP live in shared memory (else in the scratch)
    u32* ctl;                    // [0] next read, [1] flags, [2..3] DP cells (u64)
};

__host__ __device__ inline u64 cg_ra_align16(u64 v) { return (v + 15) & ~(u64)15; }
__host__ __device__ inline u64 cg_ra_buf_bytes(u32 maxL) { return cg_ra_align16(2ull * maxL + 32); }
__host__ __device__ inline u64 cg_ra_bnd_bytes(u32 rmax) { return cg_ra_align16(8ull * rmax + 16); }
__host__ __device__ inline u64 cg_ra_line_bytes(u32 maxL) { return cg_ra_align16(4ull * (maxL + 8)); }
__host__ __device__ inline size_t cg_ra_smem_per_warp(u32 rmax, u32 maxL, bool lines_in_smem) {
    const size_t lines = lines_in_smem ? 3 * (size_t)cg_ra_line_bytes(maxL) : 0;
    return ((rmax + 15u) & ~15u) + (lines > CG_RA_PROF_BYTES ? lines : (size_t)CG_RA_PROF_BYTES);
}
__host__ __device__ inline u64 cg_ra_fixed_bytes(u32 maxL, u32 rmax) {
    return 3 * cg_ra_buf_bytes(maxL) + cg_ra_bnd_bytes(rmax) + 3 * cg_ra_line_bytes(maxL);
}

// ssw_cpp.cpp:11-28 kBaseTranslation: A/a 0, C/c 1, G/g 2, T/t 3 (U/u 0), everything else 4
__device__ __forceinline__ u32 cg_ra_code(char ch) {
    const u32 c = (u32)(u8)ch | 0x20u;
    return c == 'a' ? 0u : c == 'c' ? 1u : c == 'g' ? 2u : c == 't' ? 3u : c == 'u' ? 0u : 4u;
}
__device__ __forceinline__ char cg_ra_upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
__device__ __forceinline__ char cg_ra_lower(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; }

__device__ __forceinline__ int cg_ra_wmax(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const int o = __shfl_xor_sync(CG_FULL, v, d); v = o > v ? o : v; }
    return v;
}
__device__ __forceinline__ int cg_ra_wmin(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const int o = __shfl_xor_sync(CG_FULL, v, d); v = o < v ? o : v; }
    return v;
}
__device__ __forceinline__ int cg_ra_wsum(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(CG_FULL, v, d);
    return v;
}

struct CgRaEnd { int score, col, row; };

// One band of a scan: query rows [row0, row0 + 32 * RPL) (rows >= nq are padding), every reference column.  RPL rows per
// lane, all in registers.  The only loop-carried chain inside a lane is F: with hp = max(diag + s, E) and g = max(hp - 3, 0),
//   F' = max(F - 1, max(H - 3, 0)) = max(F - 1, g)          (H = max(hp, F), and F - 3 < F - 1)
// so a column costs RPL dependent VIADDMNMX and everything else (hp, g, H, E, the column maximum) is independent work.
//
// TWO CELLS PER INSTRUCTION (round 2; scores <= 2 * CG_RA_QMAX fit a signed halfword).  The rows of a lane cannot be paired in one
// register — the F chain runs down them — so a lane is split into two VIRTUAL lanes of RPL / 2 rows, A (low halves) and B (high
// halves), B one step behind A in the wavefront: at step t virtual lane v = 2 lane + half works on column t - v.  A's first row is
// fed by the previous lane's B (the shuffle, as before), B's first row by the same lane's A of the step before (registers).  H, E,
// F, the diagonal and the column maximum are packed pairs: VIADDMNMX.S16x2 / VIMNMX.S16x2 do both cells; the match scores come from a
// query profile in shared memory (below).
// Columns outside [0, nr) are computed too — before the first column the inputs are zeros and a code that matches nothing, which
// leaves the all-zero state untouched; after the last one nobody reads the state — only the maximum tracking and the band-boundary
// store look at the column index.  Column maximum and its first row: when every score of the scan stays below 2^11 (KEY16: alignments
// of up to 959 bases, i.e. every window of the usual sizes) the pair (h, row) is one halfword key h * 16 + (15 - i), maximised with the
// same packed instruction; otherwise the maximum alone is tracked and the row is found by compare on the steps that raise it (some lane
// of the warp does at nearly every step, so that path is only for the long alignments).
template <int RPL, bool KEY16>
__device__ __forceinline__ void cg_ra_band(const char* qsrc, int qfirst, int qstep, int nq, int row0, const u8* refc, int rfirst, int rstep,
                                           int nr, const u32* in_msgs, u32* out_msgs, bool first_band, bool last_band,
                                           int& best, int& bcol, int& brow, u32* prof, const u32 zero) {
    static_assert(RPL % 2 == 0, "two virtual lanes per lane");
    constexpr int R2 = RPL / 2;
    constexpr u32 NOMATCH = 6u;                                    // matches no query code (0..3, 5 = N, 7 = padding)
    const int lane = (int)(threadIdx.x & 31u);
    u32 H2[R2], E2[R2];
    const int my0 = row0 + lane * RPL;
    // Query profile of the band in shared memory: prof[(a * R2 + i) * 32 + lane] = score of reference letter a against row i of A (low
    // half) and against row i of B (high half); a = 4: a reference N / nothing, every row scores -2.  A cell pair's scores are then
    // two LDS and one byte permute instead of seven integer instructions (the ALU pipe is what bounds this kernel).
    __syncwarp();
#pragma unroll
    for (int i = 0; i < R2; ++i) {
        u32 ca = 7u, cb = 7u;                                      // padding row: matches nothing
        if (my0 + i < nq) { ca = cg_ra_code(qsrc[qfirst + qstep * (my0 + i)]); if (ca == 4u) ca = 5u; }      // N never matches, not even N (ssw_cpp.cpp:47-55)
        if (my0 + R2 + i < nq) { cb = cg_ra_code(qsrc[qfirst + qstep * (my0 + R2 + i)]); if (cb == 4u) cb = 5u; }
#pragma unroll
        for (u32 a = 0; a < 5; ++a) prof[(a * R2 + i) * 32 + lane] = (ca == a ? 2u : 0xfffeu) | ((cb == a ? 2u : 0xfffeu) << 16);
        H2[i] = 0; E2[i] = 0;
    }
    __syncwarp();
    const u32* pl = prof + lane;
    // `zero`: 0 in a register the compiler cannot fold (the kernel derives it from an argument): VIADDMNMX.S16x2 wants its third operand
    // in a register, and a literal 0 is re-materialised by a PRMT in front of every use (two more ALU-pipe instructions per cell pair)
    int bestA = best, bcolA = bcol, browA = brow, bestB = best, bcolB = bcol, browB = brow;
    u32 dA = 0, dB = 0, a_h = 0, a_f = 0, a_rc = NOMATCH;          // diagonals of the two first rows; A's last row of the step before
    u32 out_msg = NOMATCH;
    const int n_steps = nr + 63;
    for (int t = 0; t < n_steps; ++t) {
        // lane 0 is fed from shared memory (every lane reads the same word: a broadcast), the others from their neighbour's B
        u32 feed = NOMATCH;
        if (t < nr) feed = first_band ? (u32)refc[rfirst + rstep * t] : in_msgs[t];
        const u32 up_msg = __shfl_up_sync(CG_FULL, out_msg, 1);
        const u32 in_msg = lane == 0 ? feed : up_msg;
        const int cA = t - 2 * lane, cB = cA - 1;
        const u32 rcA = in_msg & 7u, hA_in = in_msg >> 18;
        const u32* pa = pl + (rcA < 4u ? rcA : 4u) * (R2 * 32);
        const u32* pb = pl + (a_rc < 4u ? a_rc : 4u) * (R2 * 32);
        u32 f2 = ((in_msg >> 4) & 0x3fffu) | (a_f << 16);
        u32 d2 = dA | (dB << 16);
        u32 h2 = 0, cm2 = 0;                                         // cm2: packed column maxima, or packed keys (KEY16)
#pragma unroll
        for (int i = 0; i < R2; ++i) {
            const u32 sc = __byte_perm(pa[32 * i], pb[32 * i], 0x7610);            // A's score for its column | B's for its own
            const u32 hp = cg_viaddmax2(d2, sc, E2[i]);                            // E, f >= 0: the floor at 0 is implied
            const u32 g = cg_viaddmax2(hp, 0xfffdfffdu, zero);
            h2 = cg_vmax2(hp, f2);
            f2 = cg_viaddmax2(f2, 0xffffffffu, g);
            d2 = H2[i]; H2[i] = h2;
            cm2 = cg_vmax2(cm2, KEY16 ? h2 * 16u + (u32)(15 - i) * 0x00010001u : h2);
            E2[i] = cg_viaddmax2(E2[i], 0xffffffffu, cg_viaddmax2(h2, 0xfffdfffdu, zero));
        }
        // hand-over: B's last row to the next lane, A's last row to this lane's B
        const u32 hB = h2 >> 16, fB = f2 >> 16;
        out_msg = (hB << 18) | (fB << 4) | a_rc;
        if (!last_band && lane == 31 && cB >= 0 && cB < nr) out_msgs[cB] = out_msg;
        dA = hA_in; dB = a_h;
        a_h = h2 & 0xffffu; a_f = f2 & 0xffffu; a_rc = rcA;
        if (KEY16) {
            const int kA = (int)(cm2 & 0xffffu), kB = (int)(cm2 >> 16);
            const int cmA = kA >> 4, cmB = kB >> 4;
            if (cA >= 0 && cA < nr && (cmA > bestA || (cmA == bestA && cA < bcolA && cmA > 0))) { bestA = cmA; bcolA = cA; browA = my0 + 15 - (kA & 15); }
            if (cB >= 0 && cB < nr && (cmB > bestB || (cmB == bestB && cB < bcolB && cmB > 0))) { bestB = cmB; bcolB = cB; browB = my0 + R2 + 15 - (kB & 15); }
        } else {
            const int cmA = (int)(cm2 & 0xffffu), cmB = (int)(cm2 >> 16);
            if (cA >= 0 && cA < nr && (cmA > bestA || (cmA == bestA && cA < bcolA && cmA > 0))) {
                int r = 0;
#pragma unroll
                for (int i = R2 - 1; i >= 0; --i) if ((int)(H2[i] & 0xffffu) == cmA) r = i;
                bestA = cmA; bcolA = cA; browA = my0 + r;
            }
            if (cB >= 0 && cB < nr && (cmB > bestB || (cmB == bestB && cB < bcolB && cmB > 0))) {
                int r = 0;
#pragma unroll
                for (int i = R2 - 1; i >= 0; --i) if ((int)(H2[i] >> 16) == cmB) r = i;
                bestB = cmB; bcolB = cB; browB = my0 + R2 + r;
            }
        }
    }
    // the lane's (score, first column, first row) = the better of its two virtual lanes (A's rows are the smaller ones)
    const bool takeB = bestB > bestA || (bestB == bestA && bcolB < bcolA);
    best = takeB ? bestB : bestA; bcol = takeB ? bcolB : bcolA; brow = takeB ? browB : browA;
}

// One scan of the local-alignment matrix by a warp.  Query row j = qsrc[qfirst + qstep * j], j in [0, nq); column c =
// refc[rfirst + rstep * c] (codes in shared memory), c in [0, nr).  Returns the score, the first column whose maximum
// reaches it and the smallest row holding it there.  bnd: 2 * rmax u32 of per-warp scratch (the wavefront messages of a
// band's last row, read back by lane 0 of the next band, when nq > 32 * CG_RA_RPL).
__device__ CG_NOINLINE CgRaEnd cg_ra_scan(const char* qsrc, int qfirst, int qstep, int nq, const u8* refc, int rfirst, int rstep,
                                          int nr, u32* bnd, u32 rmax, u32* prof, const u32 zero) {
    int best = 0, bcol = 0x7fffffff, brow = 0x7fffffff;
    const int band_rows = 32 * (int)CG_RA_RPL;
    const int n_bands = (nq + band_rows - 1) / band_rows;
    const bool key16 = 2 * min(nq, nr) + 128 < 2048;
    for (int b = 0; b < n_bands; ++b) {
        const int row0 = b * band_rows;
        const int rows = min(band_rows, nq - row0);
        const int rpl = (rows + 31) / 32;                          // warp-uniform
        const bool last_band = b + 1 == n_bands;
        const u32* in_msgs = bnd + (size_t)((b + 1) & 1) * rmax;    // written by band b - 1
        u32* out_msgs = bnd + (size_t)(b & 1) * rmax;
#define CG_RA_BAND(R, K) cg_ra_band<R, K>(qsrc, qfirst, qstep, nq, row0, refc, rfirst, rstep, nr, in_msgs, out_msgs, b == 0, last_band, best, bcol, brow, prof, zero)
        if (key16) {                                               // every score of the scan (and of its 64 run-out columns) below 2^11
            if (rpl <= 4) CG_RA_BAND(4, true);
            else if (rpl <= 8) CG_RA_BAND(8, true);
            else if (rpl <= 12) CG_RA_BAND(12, true);
            else if (rpl <= 16) CG_RA_BAND(16, true);
            else if (rpl <= 18) CG_RA_BAND(18, true);
            else CG_RA_BAND(20, true);
        } else if (rpl <= 10) CG_RA_BAND(10, false);
        else CG_RA_BAND(20, false);
#undef CG_RA_BAND
        __syncwarp();
    }
    CgRaEnd e;
    e.score = cg_ra_wmax(best);
    e.col = cg_ra_wmin(best == e.score ? bcol : 0x7fffffff);
    e.row = cg_ra_wmin((best == e.score && bcol == e.col) ? brow : 0x7fffffff);
    return e;
}

// merCounts[kmer] >= solidThresh on the window's sorted solid list
__device__ __forceinline__ bool cg_ra_solid_has(const u32* keys, u32 n, u32 key) {
    u32 lo = 0, hi = n;
    while (lo < hi) { const u32 m = (lo + hi) >> 1; if (keys[m] < key) lo = m + 1; else hi = m; }
    return lo < n && keys[lo] == key;
}
// nbSolidMers (correctionAlignment.cpp:6-15); str2num (BMEAN/utils.cpp:18-30): A0 C1 G2, anything else (lower case too) 3
__device__ int cg_ra_nb_solid(const char* s, u32 n, const u32* keys, u32 nkeys, u32 k) {
    const u32 lane = threadIdx.x & 31u;
    int nb = 0;
    for (u32 i = lane; i + k <= n; i += 32) {
        u32 v = 0;
        for (u32 t = 0; t < k; ++t) { const char c = s[i + t]; v = (v << 2) + (c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u); }
        nb += cg_ra_solid_has(keys, nkeys, v) ? 1 : 0;
    }
    return cg_ra_wsum(nb);
}
__device__ int cg_ra_nb_upper(const char* s, u32 n) {                       // nbUpperCase, correctionAlignment.cpp:17-26
    const u32 lane = threadIdx.x & 31u;
    int nb = 0;
    for (u32 i = lane; i < n; i += 32) nb += (s[i] >= 'A' && s[i] <= 'Z') ? 1 : 0;
    return cg_ra_wsum(nb);
}

// banded_sw (ssw.c:577-758) by one lane: band DP with the band doubled until the best cell reaches `score`, traceback
// from the last cell until read row 0; returns the numbers of I and D operations.  Storage follows the reference slot for
// slot (slot(i, j) = j - max(0, i - w) + 1; slot 0 and the slot right of the row's last column are zeroed before every
// row, ssw.c:624-627); the three direction codes of a cell are packed in one byte.
__device__ __forceinline__ int cg_ra_slot(int w, int i, int j) { const int x = i - w; return j - (x > 0 ? x : 0) + 1; }
__device__ CG_NOINLINE u32 cg_ra_banded(const u8* refc, const char* read, int refLen, int readLen, int score, int w, i32* h_prev,
                                        i32* e_line, i32* h_cur, u8* dir, u64 dir_cap, int* n_ins, int* n_del) {
    int best = 0, stride = 0;
    for (;;) {
        const int width = 2 * w + 3;
        stride = min(2 * w + 1, refLen);
        if ((u64)stride * (u64)readLen > dir_cap) return CG_RA_FLAG_CAPACITY;
        for (int s = 1; s < width - 1 && s < refLen + 4; ++s) h_prev[s] = 0;
        for (int i = 0; i < readLen; ++i) {
            const int beg = max(i - w, 0), end = min(i + w, refLen - 1), edge = min(end + 1, width - 1);
            int f = 0, last = 0;
            h_prev[0] = 0; e_line[0] = 0; h_prev[edge] = 0; e_line[edge] = 0; h_cur[0] = 0;
            u8* d = dir + (size_t)stride * (size_t)i;
            u32 rcode = cg_ra_code(read[i]);
            if (rcode == 4u) rcode = 5u;
            // slot(i, j - 1) = slot(i, j) - 1 and slot(i - 1, j - 1) = slot(i - 1, j) - 1: the left neighbour and the diagonal
            // of a cell are what the previous cell wrote / read, so they travel in registers (h_left, h_diag)
            int h_left = 0;                                              // h_cur[slot(i, beg - 1)] = h_cur[0]
            int h_diag = h_prev[cg_ra_slot(w, i - 1, beg - 1)];
            for (int j = beg; j <= end; ++j) {
                const int u = cg_ra_slot(w, i, j), up = cg_ra_slot(w, i - 1, j);
                const int h_up = h_prev[up];
                int a = i == 0 ? -3 : h_up - 3;
                int b = i == 0 ? -1 : e_line[up] - 1;
                const int e = a > b ? a : b;
                const u32 de = a > b ? 3u : 2u;
                e_line[u] = e;
                a = h_left - 3;
                b = f - 1;
                f = a > b ? a : b;
                const u32 df = a > b ? 5u : 4u;
                const int e1 = max(e, 0), f1 = max(f, 0);
                const int gap = max(e1, f1);
                const int diag = h_diag + ((u32)refc[j] == rcode ? 2 : -2);
                const int h = max(gap, diag);
                h_cur[u] = h;
                h_left = h; h_diag = h_up;
                best = max(best, h);
                const u32 dh = gap <= diag ? 1u : (e1 > f1 ? de : df);
                d[j - beg] = (u8)((de & 1u) | ((df & 1u) << 1) | (dh << 2));
                last = u;
            }
            for (int s = 1; s <= last; ++s) h_prev[s] = h_cur[s];
        }
        if (best >= score) break;
        w *= 2;
        if (w > 4 * (refLen + readLen) + 16) return CG_RA_FLAG_TRACEBACK;
    }
    int i = readLen - 1, j = refLen - 1, state = 2, ins = 0, del = 0;
    while (i > 0) {
        const int x = j - max(i - w, 0);
        if (j < 0 || x < 0 || x >= stride) return CG_RA_FLAG_TRACEBACK;       // the reference reads outside its matrix here
        const u32 cell = dir[(size_t)stride * (size_t)i + (size_t)x];
        const u32 code = state == 0 ? 2u + (cell & 1u) : state == 1 ? 4u + ((cell >> 1) & 1u) : (cell >> 2);
        if (code == 1u) { --i; --j; state = 2; }
        else if (code == 2u) { --i; state = 0; ++ins; }
        else if (code == 3u) { --i; state = 2; ++ins; }
        else if (code == 4u) { --j; state = 1; ++del; }
        else if (code == 5u) { --j; state = 2; ++del; }
        else return CG_RA_FLAG_TRACEBACK;
    }
    *n_ins = ins; *n_del = del;
    return 0;
}

// dst[0..n) <- src[0..n) inside one buffer (regions may overlap), by a warp
__device__ void cg_ra_move(char* base, u32 dst, u32 src, u32 n) {
    const u32 lane = threadIdx.x & 31u;
    if (dst == src || n == 0) return;
    if (dst < src) {
        for (u32 o = 0; o < n; o += 32) {
            const u32 i = o + lane;
            char c = 0;
            if (i < n) c = base[src + i];
            __syncwarp();
            if (i < n) base[dst + i] = c;
            __syncwarp();
        }
    } else {
        for (u32 done = 0; done < n; done += 32) {
            const u32 chunk = min(32u, n - done);
            const u32 o = n - done - chunk;
            char c = 0;
            if (lane < chunk) c = base[src + o + lane];
            __syncwarp();
            if (lane < chunk) base[dst + o + lane] = c;
            __syncwarp();
        }
    }
}

// 128 registers: 4 CTAs (16 warps) per SM.  No min-blocks clause: with it (4, 5 or 6 CTAs: 128 / 96 / 80 registers) ptxas schedules the
// unrolled row loop worse and the kernel loses 12 % (1 598 -> 1 400 GCUPS at 4; the extra warps of 5 and 6 do not win it back).
#define CG_RA_CTAS_PER_SM 4
__global__ void __launch_bounds__(CG_RA_WARPS * 32) k_reanchor(CgReanchorArgs P) {
    CG_DYN_SMEM(smem_raw);
    const u32 lane = threadIdx.x & 31u, wip = threadIdx.x >> 5;
    const u32 rpad = (P.rmax + 15u) & ~15u;
    // per warp: the reference codes, then ONE region shared by the scans' query profile and the three rolling lines of the banded
    // sub-alignment (lane 0 runs it between scans, every scan rebuilds its profile): a second region would cost a resident CTA
    const size_t smem_per_warp = cg_ra_smem_per_warp(P.rmax, P.maxL, P.lines_in_smem != 0);
    u8* refc = (u8*)smem_raw + (size_t)wip * smem_per_warp;
    u32* prof = (u32*)(refc + rpad);
    const u32 gw = blockIdx.x * CG_RA_WARPS + wip;
    u8* sc = P.scratch + (size_t)gw * P.scratch_stride;
    char* bufs[3];
    bufs[0] = (char*)sc; bufs[1] = bufs[0] + cg_ra_buf_bytes(P.maxL); bufs[2] = bufs[1] + cg_ra_buf_bytes(P.maxL);
    u32* bnd = (u32*)(bufs[2] + cg_ra_buf_bytes(P.maxL));
    i32* line0 = P.lines_in_smem ? (i32*)(refc + rpad) : (i32*)((u8*)bnd + cg_ra_bnd_bytes(P.rmax));
    i32* line1 = (i32*)((u8*)line0 + cg_ra_line_bytes(P.maxL));
    i32* line2 = (i32*)((u8*)line1 + cg_ra_line_bytes(P.maxL));
    u8* dir = (u8*)bnd + cg_ra_bnd_bytes(P.rmax) + 3 * cg_ra_line_bytes(P.maxL);
    u64 cells = 0;
    u32 flags = 0;
    const u32 zero = P.n_reads >> 31;                             // 0 (n_reads < 2^31), opaque to the compiler: cg_ra_band

    for (;;) {
        u32 slot = 0;
        if (lane == 0) slot = atomicAdd(&P.ctl[0], 1u);
        slot = __shfl_sync(CG_FULL, slot, 0);
        if (slot >= P.n_reads) break;
        const u32 r = P.order[slot];
        const u32 w0 = P.read_win_begin[r], w1 = P.read_win_begin[r + 1];
        const char* raw = P.read_bases + P.read_off[r];
        const u32 rawLen = (u32)(P.read_off[r + 1] - P.read_off[r]);
        char* head = P.head + P.head_off[r];
        if (w0 == w1) { if (lane == 0) P.out_len[r] = 0; continue; }          // CONSENT-correction.cpp:23-25
        u32 hlen = 0, t0 = 0;                                                 // outSequence = head[0..hlen) + lower(raw[t0..))
        int curPos = (int)P.win_pos[w0];                                      // startPos
        u32 oldEnd = 0, old_n = 0, oldW = w0;
        bool haveOld = false;
        int ic = 0, io = 1, it = 2;
        const char* oldp = bufs[io];

        for (u32 w = w0; w < w1; ++w) {
            const u64 c0 = P.cons_off[w], c1 = P.cons_off[w + 1];
            const bool isCons = (c1 - c0) >= (u64)P.k;                        // :72
            const char* src; u32 n;
            if (isCons) { src = P.cons + c0; n = (u32)(c1 - c0); }
            else if (P.win_seq_begin) { const u32 s0 = P.win_seq_begin[w]; src = P.bases + P.seq_off[s0]; n = (u32)(P.seq_off[s0 + 1] - P.seq_off[s0]); }
            else { src = P.tpl + P.tpl_off[w]; n = (u32)(P.tpl_off[w + 1] - P.tpl_off[w]); }
            const u32 outLen = hlen + (rawLen - t0);
            int alPos = curPos - (int)P.ov; if (alPos < 0) alPos = 0;          // :80
            int sizeAl = ((u64)alPos + P.ws + 2ull * P.ov >= (u64)outLen) ? (int)outLen - alPos : (int)(P.ws + 2 * P.ov);   // :81-85
            if (sizeAl <= 0 || n == 0) { flags |= CG_RA_FLAG_DEGENERATE; break; }
            if ((u32)sizeAl > P.rmax || n > P.maxL || n > CG_RA_QMAX) { flags |= CG_RA_FLAG_CAPACITY; break; }
            char* cur = bufs[ic];
            for (u32 i = lane; i < n; i += 32) cur[i] = src[i];
            // the read up to the end of the alignment region becomes part of the head
            {
                const u32 upto = (u32)alPos + (u32)sizeAl;
                if (upto > hlen) {
                    const u32 m = upto - hlen;
                    for (u32 i = lane; i < m; i += 32) head[hlen + i] = cg_ra_lower(raw[t0 + i]);
                    hlen += m; t0 += m;
                }
            }
            __syncwarp();
            for (u32 i = lane; i < (u32)sizeAl; i += 32) refc[i] = (u8)cg_ra_code(head[alPos + i]);
            __syncwarp();
            const CgRaEnd fw = cg_ra_scan(cur, 0, 1, (int)n, refc, 0, 1, sizeAl, bnd, P.rmax, prof, zero);                          // :87
            cells += (u64)n * (u64)sizeAl;
            if (fw.score <= 0) { flags |= CG_RA_FLAG_DEGENERATE; break; }
            const CgRaEnd bw = cg_ra_scan(cur, fw.row, -1, fw.row + 1, refc, fw.col, -1, fw.col + 1, bnd, P.rmax, prof, zero);
            cells += (u64)(fw.row + 1) * (u64)(bw.col + 1);
            const u32 beg = (u32)(fw.col - bw.col + alPos), end = (u32)(fw.col + alPos);                                  // :88-89
            const char* curp = cur + (fw.row - bw.row);                                                                   // :90
            u32 cn = (u32)(bw.row + 1);

            if (w != w0 && oldEnd >= beg) {                                                                               // :93
                const u32 overlap = oldEnd - beg + 1;
                if (isCons && old_n >= overlap && cn >= overlap) {                                                        // :95
                    const char* s1 = oldp + (old_n - overlap);
                    const char* s2 = curp;
                    bool differ = false;
                    for (u32 o = 0; o < overlap; o += 32) {
                        const u32 i = o + lane;
                        const bool ne = i < overlap && cg_ra_upper(s1[i]) != cg_ra_upper(s2[i]);
                        if (__ballot_sync(CG_FULL, ne)) { differ = true; break; }
                    }
                    if (differ) {
                        int n1, n2;
                        if (overlap >= P.k) {                                                                             // :99-101
                            const u64 a0 = haveOld ? P.solid_off[oldW] : 0, a1 = haveOld ? P.solid_off[oldW + 1] : 0;
                            n1 = cg_ra_nb_solid(s1, overlap, P.solid_kmer + a0, (u32)(a1 - a0), P.k);
                            n2 = cg_ra_nb_solid(s2, overlap, P.solid_kmer + P.solid_off[w], (u32)(P.solid_off[w + 1] - P.solid_off[w]), P.k);
                        } else { n1 = cg_ra_nb_upper(s1, overlap); n2 = cg_ra_nb_upper(s2, overlap); }                     // :103-104
                        if (n1 > n2) {                                                                                    // :106-117
                            __syncwarp();
                            for (u32 i = lane; i < overlap; i += 32) refc[i] = (u8)cg_ra_code(s2[i]);
                            __syncwarp();
                            const CgRaEnd f2 = cg_ra_scan(s1, 0, 1, (int)overlap, refc, 0, 1, (int)overlap, bnd, P.rmax, prof, zero);
                            cells += (u64)overlap * (u64)overlap;
                            int ins = 0, del = 0;
                            u32 bad = 0;
                            if (f2.score > 0) {
                                const CgRaEnd b2 = cg_ra_scan(s1, f2.row, -1, f2.row + 1, refc, f2.col, -1, f2.col + 1, bnd, P.rmax, prof, zero);
                                cells += (u64)(f2.row + 1) * (u64)(b2.col + 1);
                                const int rb = f2.col - b2.col, qb = f2.row - b2.row;
                                const int refLen = b2.col + 1, readLen = b2.row + 1;
                                if (lane == 0) {
                                    const int bw0 = (refLen > readLen ? refLen - readLen : readLen - refLen) + 1;
                                    bad = cg_ra_banded(refc + rb, s1 + qb, refLen, readLen, f2.score, bw0, line0, line1, line2, dir, P.dir_cap, &ins, &del);
                                }
                                bad = __shfl_sync(CG_FULL, bad, 0);
                                ins = __shfl_sync(CG_FULL, ins, 0);
                                del = __shfl_sync(CG_FULL, del, 0);
                            }   // score 0: the reference's CIGAR is "1M" plus clips: no I, no D
                            if (bad) { flags |= bad; break; }
                            const u32 cut = overlap - (u32)ins + (u32)del;
                            if (cut < cn) {
                                char* tmp = bufs[it];
                                for (u32 i = lane; i < overlap; i += 32) tmp[i] = s1[i];
                                for (u32 i = lane; i < cn - cut; i += 32) tmp[overlap + i] = curp[cut + i];
                                __syncwarp();
                                cn = overlap + cn - cut;
                                curp = tmp;
                                const int x = ic; ic = it; it = x;
                            } else cn = 0;
                        }
                    }
                }
            }

            if (cn != 0) {                                                                                                // :121
                if (isCons) {                                                                                             // :122-126
                    const u32 oldLen = end - beg + 1;
                    const u32 tail = hlen - (end + 1);
                    __syncwarp();
                    cg_ra_move(head, beg + cn, end + 1, tail);
                    for (u32 i = lane; i < cn; i += 32) head[beg + i] = cg_ra_upper(curp[i]);
                    hlen = hlen - oldLen + cn;
                    __syncwarp();
                }
                if (w + 1 < w1) {                                                                                         // :127-132
                    curPos = (int)((u32)curPos + P.win_pos[w + 1] - P.win_pos[w] - (end - beg + 1) + cn);
                    oldp = curp; old_n = cn; oldW = w; haveOld = true;
                    oldEnd = beg + cn - 1;
                    const int x = io; io = ic; ic = x;            // the buffer holding cur now holds old
                }
            }
        }
        // the untouched rest of the read
        {
            const u32 m = rawLen - t0;
            for (u32 i = lane; i < m; i += 32) head[hlen + i] = cg_ra_lower(raw[t0 + i]);
            hlen += m;
        }
        if (lane == 0) P.out_len[r] = hlen;
        __syncwarp();
    }
    if (lane == 0) {
        if (flags) atomicOr(&P.ctl[1], flags);
        atomicAdd((unsigned long long*)(P.ctl + 2), (unsigned long long)cells);
    }
}

// Post-filters: the tail of processRead (src/CONSENT-correction.cpp:49-59) on the re-anchored read in place.
//   trimRead(read, m)  src/utils.cpp:96-128   beg = start of the first run of m upper-case bases, end = last base of the last such
//                                             run; "" unless end > beg
//   dropRead(read)     src/utils.cpp:60-73    (float) upper-case bases / length < 0.1
// One CTA per read: len[r] becomes the length of the FASTA sequence line (0 = no record), skip[r] its first base.
// A read without any such run yields "" (the reference's descending `unsigned i >= 0` loop leaves the string there).
__global__ void __launch_bounds__(256) k_finish_reads(const char* head, const u64* head_off, u32* len, u32* skip, u32 n_reads, u32 m) {
    CG_DYN_SMEM(smem);
    u32* red = (u32*)smem;                                         // [0..8) minima, [8..16) maxima, [16..24) counts, [24..27) results
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (u32 r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const char* s = head + head_off[r];
        const u32 n = len[r];
        u32 lo = CG_NONE32, hi = 0;                                 // hi = last base of a run + 1 (0: none)
        for (u32 i = threadIdx.x; i + m <= n; i += blockDim.x) {
            bool run = true;
            for (u32 j = 0; j < m; ++j) { const char c = s[i + j]; run = run && c >= 'A' && c <= 'Z'; }
            if (run) { lo = lo < i ? lo : i; hi = i + m; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const u32 a = __shfl_xor_sync(CG_FULL, lo, d), b = __shfl_xor_sync(CG_FULL, hi, d);
            lo = a < lo ? a : lo; hi = b > hi ? b : hi;
        }
        if (lane == 0) { red[warp] = lo; red[8 + warp] = hi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 a = CG_NONE32, b = 0;
            for (u32 w = 0; w < nw; ++w) { a = red[w] < a ? red[w] : a; b = red[8 + w] > b ? red[8 + w] : b; }
            red[24] = a; red[25] = b;
        }
        __syncthreads();
        const u32 beg = red[24], endp1 = red[25];
        u32 cnt = 0;
        if (beg != CG_NONE32)
            for (u32 i = beg + threadIdx.x; i < endp1; i += blockDim.x) { const char c = s[i]; cnt += (c >= 'A' && c <= 'Z') ? 1u : 0u; }
        cnt = cg_warp_sum(cnt);
        if (lane == 0) red[16 + warp] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 out_len = 0, out_skip = 0;
            if (beg != CG_NONE32 && endp1 - 1u > beg) {            // end > beg (utils.cpp:123)
                u32 c = 0;
                for (u32 w = 0; w < nw; ++w) c += red[16 + w];
                const u32 L = endp1 - beg;
                const float frac = (float)(int)c / (float)(size_t)L;
                if (!((double)frac < 0.1)) { out_len = L; out_skip = beg; }
            }
            len[r] = out_len; skip[r] = out_skip;
        }
        __syncthreads();
    }
}

// cg_finish_resident: the reads of the piles, dense, from the device store (read r = store[pile_read[r]])
__global__ void k_reanchor_reads_from_store(const char* store, const u64* store_off, const u32* pile_read, const u64* read_off, char* out, u32 n_reads) {
    for (u32 r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const char* s = store + store_off[pile_read[r]];
        char* d = out + read_off[r];
        const u64 n = read_off[r + 1] - read_off[r];
        for (u64 i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    }
}

// corrected reads, dense: out[out_off[r] ..) <- head slice of read r (from base skip[r] on when the post-filters ran)
__global__ void k_reanchor_gather(const char* head, const u64* head_off, const u32* skip, const u64* out_off, char* out, u32 n_reads) {
    for (u32 r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const char* s = head + head_off[r] + (skip ? skip[r] : 0u);
        char* d = out + out_off[r];
        const u64 n = out_off[r + 1] - out_off[r];
        for (u64 i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    }
}
