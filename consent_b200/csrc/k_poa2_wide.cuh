// k_poa2_wide.cuh — the wide (all-global) tiers of k_poa2.cuh: score matrix, arg-max column and traceback for graphs of up to
// 1024 / 4096 nodes against segments of hundreds of bases — the whole-window alignments of shallow piles (SURVEY config 2, real
// coverage), where one matrix is ~1 MB and lives in HBM / L2.  Included by k_poa2.cuh; same reference semantics (spoa kSW linear,
// simd_alignment_engine_impl.hpp:712-1056).
//
// Wide layout of a matrix row (CH = ceil((L + 1) / 64) blocks, row stride 32 * CH words): lane l owns the 2 * CH CONSECUTIVE columns
// [2 CH l, 2 CH (l + 1)); word c of the lane (stored at row * 32 CH + 32 c + l: every store and load of a warp is one 128-byte line)
// holds column 2 CH l + c in its low half and column 2 CH l + CH + c in its high half.  With that
//   * the diagonal / left neighbour of word c is word c - 1 of the same lane (both halves at once); only word 0 needs the previous
//     lane: one shuffle + one byte permute per row instead of one per 64 columns;
//   * the in-row gap term H[j] = max(v[j], H[j-1] - 4) is a serial VIADDMNMX chain over the lane's words (two half-blocks at a time),
//     ONE max-plus scan over the 32 lane totals (5 shuffles per row, not 5 per 64 columns), and one VIADDMNMX per word to fold the
//     incoming carry back in;
//   * match / mismatch scores come from a per-alignment query profile in shared memory (one LDS per word).
// A 576-column row costs ~100 warp instructions (the 64-columns-per-chunk layout of the compact tiers: ~270).
// Columns past the end of the query are computed with "mismatch" scores: such a cell is always strictly below some real cell of its own
// or an earlier row, so neither the maximum, nor the first row reaching it, nor the count of rows reaching it can change.
//
// The traceback is done by the whole warp: lane t looks at the cell the path reaches after t diagonal steps along the
// first-predecessor chain (row descriptors in shared memory make the chain walk cheap), all lanes fetch their cell and its three
// neighbours at once (one memory round trip), a ballot finds how far the guess holds, those steps are committed together and the
// first lane that disagrees decides the next cell.  One round trip per run of matches instead of two or three per step.
#pragma once
#include <type_traits>

__device__ __forceinline__ u32 cg_viaddmax2(u32 a, u32 b, u32 c) { return __viaddmax_s16x2(a, b, c); }
__device__ __forceinline__ u32 cg_pcode(u32 ch) { return (ch >> 1) & 3u; }       // A 0, C 1, T 2, G 3: row of the query profile

// Row descriptor of the wide tiers (shared memory, one word per matrix row - 1): letter code | in-degree << 2 | first predecessor
// row << 8 | second predecessor row << (8 + RB).  Two predecessors cover ~95 % of the rows; the rest read the full vector (prow).
template <class T> struct CgWideRd {
    typedef typename std::conditional<(T::VCAP <= 1024), u32, u64>::type RdT;
    static constexpr u32 RB = T::VCAP <= 1024 ? 11 : 16;
    __device__ __forceinline__ static RdT* at(const CgPoa2G<T>& s) { return (RdT*)s.wrd(); }
    __device__ __forceinline__ static RdT pack(u32 letter, u32 deg, u32 p0, u32 p1) {
        return (RdT)(cg_pcode(letter) | (deg << 2)) | ((RdT)p0 << 8) | ((RdT)p1 << (8 + RB));
    }
    __device__ __forceinline__ static u32 code(RdT d) { return (u32)d & 3u; }
    __device__ __forceinline__ static u32 deg(RdT d) { return ((u32)d >> 2) & 31u; }
    __device__ __forceinline__ static u32 p0(RdT d) { return (u32)(d >> 8) & ((1u << RB) - 1u); }
    __device__ __forceinline__ static u32 p1(RdT d) { return (u32)(d >> (8 + RB)) & ((1u << RB) - 1u); }
};

// (word, half) of column j in the wide layout.  inv = ceil(2^20 / (2 CH)): exact for j < 2^20 / 66.
struct CgWideCol { u32 lane, c, half; };
__device__ __forceinline__ u32 cg_wide_inv(u32 CH) { return ((1u << 20) + 2u * CH - 1u) / (2u * CH); }
__device__ __forceinline__ CgWideCol cg_wide_col(u32 j, u32 CH, u32 inv) {
    CgWideCol w;
    w.lane = (j * inv) >> 20;
    const u32 t = j - w.lane * 2u * CH;
    w.half = t >= CH ? 1u : 0u;
    w.c = t - (w.half ? CH : 0u);
    return w;
}
__device__ __forceinline__ u32 cg_wide_cell(const u16* Hh, u32 row, u32 j, u32 CH, u32 inv) {
    const CgWideCol w = cg_wide_col(j, CH, inv);
    return Hh[(((size_t)row * CH + w.c) * 32u + w.lane) * 2u + w.half];
}

// Query profile: prof[(a * CH + c) * 32 + lane] = scores of the lane's word c against letter a (5 match, -10 otherwise / no column).
// Once per alignment: a plain loop (the code of these tiers has to stay small: their warps run in different phases and share one
// 32 KB instruction cache per SM).
__device__ CG_NOINLINE void cg_poa2w_profile(u32* prof, const u8* seq, u32 L, u32 CH) {
    const u32 lane = cg_lane();
#pragma unroll 1
    for (u32 c = 0; c < CH; ++c) {
        const u32 j0 = 2u * CH * lane + c, j1 = j0 + CH;
        const u32 q0 = (j0 >= 1 && j0 <= L) ? cg_pcode(seq[j0 - 1]) : 4u, q1 = (j1 <= L) ? cg_pcode(seq[j1 - 1]) : 4u;   // j1 >= 1 always
#pragma unroll
        for (u32 a = 0; a < 4; ++a) prof[(a * CH + c) * 32 + lane] = (q0 == a ? 5u : 0xfff6u) | (q1 == a ? 0x00050000u : 0xfff60000u);
    }
    __syncwarp();
}

// Lane totals -> the packed carry every word of the lane folds in.  lo_end / hi_end: the (unfixed) last cells of the lane's two
// half-blocks.  Returns (carry into the low half-block | carry into the high half-block << 16), as seen one column before the block.
__device__ __forceinline__ u32 cg_poa2w_carry(u32 hlast, u32 CH, u32 lane) {
    const i32 lo_end = (i32)(hlast & 0xffffu), hi_end = (i32)(hlast >> 16);          // scores are never negative
    const i32 B4 = 8 * (i32)CH;                                                     // 4 * (2 CH): decay across one lane
    const i32 t = max(hi_end, lo_end - 4 * (i32)CH);                                 // the lane's last column, nothing coming in
    i32 u = t + B4 * (i32)lane;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { const i32 o = __shfl_up_sync(CG_FULL, u, dd); u = o > u ? o : u; }   // lanes < dd get their own value back
    i32 e = __shfl_up_sync(CG_FULL, u, 1) - B4 * ((i32)lane - 1);                    // value one column before the lane's first
    if (lane == 0) e = 0;                                                            // nothing comes in (0 is below every H + 4)
    const i32 eh = max(e - 4 * (i32)CH, lo_end);
    return ((u32)e & 0xffffu) | ((u32)eh << 16);
}

// Score matrix, segments of up to 64 CH - 1 bases, the row in registers.  Rows with one or two predecessors (nearly all) take them
// from P0 / P1: the row just computed is copied there, any other row is LOADED ONE ROW AHEAD (right after the previous row's
// predecessors were consumed), so that the L2 round trip overlaps that row's scan instead of stalling this one.
template <int CH, class T>
__device__ CG_NOINLINE void cg_poa2w_dp(const CgPoa2G<T>& s, u32 V, CgPoa2Max& trk) {
    CG_P2_TYPES;
    typedef CgWideRd<T> Rd;
    const u32 lane = cg_lane();
    constexpr u32 RS = 32u * CH;
    u32* Hw = (u32*)s.H() + lane;
    const typename Rd::RdT* rd = Rd::at(s);
    const u32* prof = s.prof() + lane;
    const u32 keep0 = lane == 0 ? 0xffff0000u : 0xffffffffu;                         // column 0 stays 0
    u32 P0[CH], P1[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        P0[c] = 0; P1[c] = 0; Hw[32 * c] = 0;
    }
    __syncwarp();
    // a stored predecessor row.  (Keeping the last 2 / 4 rows in a shared-memory ring as well was measured: the 2.5 - 5 KB per warp cost
    // more in resident warps than the shorter loads gained: 165 k -> 157 k / 153 k windows/s at 20 sequences per window.)
#define CG_P2W_FETCH(P, prow_) { const u32* src_ = Hw + (size_t)(prow_) * RS; _Pragma("unroll") for (int c = 0; c < CH; ++c) P[c] = src_[32 * c]; }
    // candidates of one predecessor row S: max(diagonal + score, vertical - 4, 0)
#define CG_P2W_CAND(S, OUT, FIRST)                                                                                                  \
    {                                                                                                                               \
        const u32 top_ = __shfl_up_sync(CG_FULL, S[CH - 1], 1);                                                                      \
        const u32 c0_ = cg_viaddmax2_relu(__byte_perm(top_, S[CH - 1], 0x5432), pf[0], cg_vadd2(S[0], 0xfffcfffcu));                 \
        u32 t_[CH];                                                                                                                 \
        t_[0] = c0_;                                                                                                                \
        _Pragma("unroll") for (int c = 1; c < CH; ++c) t_[c] = cg_viaddmax2_relu(S[c - 1], pf[32 * c], cg_vadd2(S[c], 0xfffcfffcu)); \
        _Pragma("unroll") for (int c = 0; c < CH; ++c) OUT[c] = FIRST ? t_[c] : cg_vmax2(OUT[c], t_[c]);                             \
    }
    typename Rd::RdT dn = rd[0];
    {   // predecessors of the first row: matrix row 0 (zeros) is "the row just computed"
        // (its predecessors, if any, can only be row 0: P0 / P1 are zeros already)
    }
    u32* row = Hw + RS;
    for (u32 r = 0; r < V; ++r) {
        const typename Rd::RdT d = dn;
        if (r + 1 < V) dn = rd[r + 1];
        const u32 deg = Rd::deg(d);
        const u32* pf = prof + Rd::code(d) * RS;
        u32 val[CH];
        CG_P2W_CAND(P0, val, true)
        if (deg >= 2) {
            CG_P2W_CAND(P1, val, false)
            if (deg > 2) {                               // rare: the third and later predecessors (stored rows) go through P1 one at a time
                const VecT pr = s.prow(r);
#pragma unroll 1
                for (u32 e = 2; e < deg; ++e) {
                    const u32* src = Hw + (size_t)Pk::get(pr, e) * RS;
#pragma unroll
                    for (int c = 0; c < CH; ++c) P1[c] = src[32 * c];
                    CG_P2W_CAND(P1, val, false)
                }
            }
        }
        // the next row's stored predecessors: on their way while this row is scanned
        const bool more = r + 1 < V;
        const u32 ndeg = Rd::deg(dn), np0 = Rd::p0(dn), np1 = Rd::p1(dn);
        if (more && np0 != r + 1) { CG_P2W_FETCH(P0, np0) }
        if (more && ndeg >= 2 && np1 != r + 1) { CG_P2W_FETCH(P1, np1) }
        val[0] &= keep0;
        // in-row gap term: serial inside the lane, one scan across the lanes, carry folded back in
#pragma unroll
        for (int c = 1; c < CH; ++c) val[c] = cg_viaddmax2(val[c - 1], 0xfffcfffcu, val[c]);
        const u32 cw = cg_poa2w_carry(val[CH - 1], CH, lane);
        u32 m2 = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const u32 dec = (u32)((0u - 4u * (u32)(c + 1)) & 0xffffu) * 0x10001u;
            val[c] = cg_viaddmax2(cw, dec, val[c]);
            row[32 * c] = val[c];
            m2 = cg_vmax2(m2, val[c]);
        }
        // the row just computed, if the next row descends from it
        if (np0 == r + 1) { _Pragma("unroll") for (int c = 0; c < CH; ++c) P0[c] = val[c]; }
        if (ndeg >= 2 && np1 == r + 1) { _Pragma("unroll") for (int c = 0; c < CH; ++c) P1[c] = val[c]; }
        const u32 lo = m2 & 0xffffu, hi = m2 >> 16;
        const i32 m = cg_poa2_track(trk, (i32)(lo > hi ? lo : hi), r + 1);
        CG_P2_KEEP_ROWMAX(T, s, m, r + 1);
        row += RS;
        __syncwarp();
    }
#undef CG_P2W_CAND
#undef CG_P2W_FETCH
}

// Any length (CH blocks, known at run time): same layout and arithmetic, the row goes through memory, scores by comparison.
template <class T>
__device__ CG_NOINLINE void cg_poa2w_dp_any(const CgPoa2G<T>& s, u32 V, const u8* seq, u32 L, u32 CH, CgPoa2Max& trk) {
    CG_P2_TYPES;
    static_assert(5u * T::LCAP + 8u * T::CHMAX * 32u < 32767u, "packed 16-bit DP: score (<= 5 L) + the scan's lane offsets must fit a signed halfword");
    const u32 lane = cg_lane();
    const u32 RS = 32u * CH;
    typedef CgWideRd<T> Rd;
    u32* Hw = (u32*)s.H() + lane;
    const typename Rd::RdT* rd = Rd::at(s);
    const u32 keep0 = lane == 0 ? 0xffff0000u : 0xffffffffu;
    for (u32 c = 0; c < CH; ++c) Hw[32 * c] = 0;
    __syncwarp();
    for (u32 r = 0; r < V; ++r) {
        const typename Rd::RdT d = rd[r];
        const u32 ch = Rd::code(d), deg = Rd::deg(d), p0 = Rd::p0(d), np = deg ? deg : 1u;
        VecT pr = Pk::zero();
        if (deg > 1) pr = s.prow(r);
        u32* row = Hw + (size_t)(r + 1) * RS;
        u32 hprev = 0;
        for (u32 c = 0; c < CH; ++c) {
            const u32 j0 = 2u * CH * lane + c, j1 = j0 + CH;
            const u32 q0 = (j0 >= 1 && j0 <= L) ? cg_pcode(seq[j0 - 1]) : 4u, q1 = j1 <= L ? cg_pcode(seq[j1 - 1]) : 4u;
            const u32 sc = (q0 == ch ? 5u : 0xfff6u) | (q1 == ch ? 0x00050000u : 0xfff60000u);
            u32 val = 0;
#pragma unroll 1
            for (u32 e = 0; e < np; ++e) {
                const u32* src = Hw + (size_t)(deg > 1 ? Pk::get(pr, e) : p0) * RS;
                const u32 sv = src[32 * c];
                u32 dg;
                if (c == 0) { const u32 last = src[32 * (CH - 1)]; dg = __byte_perm(__shfl_up_sync(CG_FULL, last, 1), last, 0x5432); }
                else dg = src[32 * (c - 1)];
                val = cg_vmax2(val, cg_viaddmax2_relu(dg, sc, cg_vadd2(sv, 0xfffcfffcu)));
            }
            if (c == 0) val &= keep0;
            else val = cg_viaddmax2(hprev, 0xfffcfffcu, val);
            hprev = val;
            row[32 * c] = val;
        }
        const u32 cw = cg_poa2w_carry(hprev, CH, lane);
        u32 m2 = 0, dec = 0xfffcfffcu;
        for (u32 c = 0; c < CH; ++c) {
            const u32 h = cg_viaddmax2(cw, dec, row[32 * c]);
            row[32 * c] = h;
            m2 = cg_vmax2(m2, h);
            dec = cg_vadd2(dec, 0xfffcfffcu);
        }
        const u32 lo = m2 & 0xffffu, hi = m2 >> 16;
        const i32 m = cg_poa2_track(trk, (i32)(lo > hi ? lo : hi), r + 1);
        CG_P2_KEEP_ROWMAX(T, s, m, r + 1);
        __syncwarp();
    }
}

// First column (>= 1) of matrix row `row` equal to M (simd_alignment_engine_impl.hpp:860-862); 0 if none.
template <class T>
__device__ CG_NOINLINE u32 cg_poa2w_rowfirst(const CgPoa2G<T>& s, u32 row, u32 CH, u32 Wd, i32 M) {
    const u32 lane = cg_lane();
    const u32* hr = (const u32*)s.H() + (size_t)row * 32u * CH + lane;
    u32 best = 0xffffffffu;
#pragma unroll 1
    for (u32 c = 0; c < CH; ++c) {
        const u32 w = hr[32 * c];
        const u32 j0 = 2u * CH * lane + c, j1 = j0 + CH;
        if ((i32)(w >> 16) == M && j1 < Wd) best = j1 < best ? j1 : best;
        if ((i32)(w & 0xffffu) == M && j0 >= 1 && j0 < Wd) best = j0 < best ? j0 : best;
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) { const u32 o = __shfl_xor_sync(CG_FULL, best, dlt); best = o < best ? o : best; }
    return best == 0xffffffffu ? 0u : best;
}

// Traceback by the whole warp (priorities of simd_alignment_engine_impl.hpp:968-1004: diagonal over the predecessors in in-edge order,
// then vertical over them, then horizontal).  Pairs go to work[] in traceback order as node | qpos << W (all ones = none), like
// cg_poa2_traceback.  Returns their number.
template <class T>
__device__ CG_NOINLINE u32 cg_poa2w_traceback(const CgPoa2G<T>& s, const u8* seq, u32 CH, u32 bi, u32 bj, i32 M, bool* bad) {
    CG_P2_TYPES;
    typedef CgWideRd<T> Rd;
    enum { STOP = 0, DIAG = 1, DIAG1 = 2, VERT = 3, VERT1 = 4, HORIZ = 5, MULTI = 6 };
    const u32 lane = cg_lane();
    const u16* Hh = (const u16*)s.H();
    const typename Rd::RdT* rd = Rd::at(s);
    const u32 inv = cg_wide_inv(CH);
    u32 i = bi, j = bj, n = 0;
    i32 Hij = M;
    while (Hij != 0) {
        // ---- rows of the next 32 cells along the first-predecessor chain (lane t: t diagonal steps from (i, j))
        u32 row = i >= lane ? i - lane : 0u;
        typename Rd::RdT d = row ? rd[row - 1] : (typename Rd::RdT)0;
        u32 start = 0, nvalid = 32;
#pragma unroll 1
        for (u32 it = 0;; ++it) {
            const u32 brk = __ballot_sync(CG_FULL, row != 0 && Rd::p0(d) != row - 1) & (0xffffffffu << start);
            if (!brk) break;
            const u32 b = (u32)__ffs((int)brk) - 1u;
            if (it == 3 || b == 31) { nvalid = b + 1; break; }
            const u32 pb = __shfl_sync(CG_FULL, Rd::p0(d), (int)b);
            if (lane > b) {
                const u32 back = lane - b - 1;
                row = pb >= back ? pb - back : 0u;
                d = row ? rd[row - 1] : (typename Rd::RdT)0;
            }
            start = b + 1;
        }
        const u32 col = j >= lane ? j - lane : 0u;
        const bool live = row != 0 && col != 0;
        const u32 p0 = Rd::p0(d), p1 = Rd::p1(d), deg = Rd::deg(d);
        // ---- the cell, its left neighbour and the first two predecessors' cells: one round trip for the whole run
        i32 own = 0, hl = 0, hd = 0, hv = 0, hd1 = -1, hv1 = -1;
        u32 node = IDNONE;
        i32 sc = 0;
        if (live) {
            const CgWideCol w1 = cg_wide_col(col, CH, inv), w0 = cg_wide_col(col - 1, CH, inv);
            const size_t o1 = ((size_t)w1.c * 32u + w1.lane) * 2u + w1.half, o0 = ((size_t)w0.c * 32u + w0.lane) * 2u + w0.half;
            const u16* hr = Hh + (size_t)row * 64u * CH;
            const u16* hp = Hh + (size_t)p0 * 64u * CH;
            own = hr[o1]; hl = hr[o0]; hd = hp[o0]; hv = hp[o1];
            if (deg >= 2) { const u16* hq = Hh + (size_t)p1 * 64u * CH; hd1 = hq[o0]; hv1 = hq[o1]; }
            node = (u32)(s.rdesc(row - 1) >> (16 + W)) & IDNONE;
            sc = Rd::code(d) == cg_pcode(seq[col - 1]) ? 5 : -10;
        }
        u32 code = STOP;
        if (live && own != 0) {
            if (own == hd + sc) code = DIAG;
            else if (deg >= 2 && own == hd1 + sc) code = DIAG1;
            else if (deg > 2) code = MULTI;
            else if (own == hv - 4) code = VERT;
            else if (deg == 2 && own == hv1 - 4) code = VERT1;
            else code = HORIZ;
        }
        u32 stopm = __ballot_sync(CG_FULL, code != DIAG);
        if (nvalid < 32) stopm |= 0xffffffffu << nvalid;
        const u32 run = stopm ? (u32)__ffs((int)stopm) - 1u : 32u;              // lanes [0, run) step diagonally along the chain
        if (n + run + 1 > T::ALNCAP) { *bad = true; return 0; }                  // cannot happen: a path visits a cell once
        if (lane < run) s.work(n + lane) = (ItemT)(node | ((col - 1) << W));
        n += run;
        if (run == 32 || run == nvalid) {                                        // the guess held to its end: go on from where lane run - 1 leads
            const int src = (int)run - 1;
            i = __shfl_sync(CG_FULL, p0, src); j = __shfl_sync(CG_FULL, col, src) - 1u; Hij = __shfl_sync(CG_FULL, hd, src);
            continue;
        }
        // ---- lane `run` decides the next step
        const int src = (int)run;
        const u32 xcode = __shfl_sync(CG_FULL, code, src);
        if (xcode == STOP) break;
        const u32 xrow = __shfl_sync(CG_FULL, row, src), xcol = __shfl_sync(CG_FULL, col, src), xnode = __shfl_sync(CG_FULL, node, src);
        const i32 xown = __shfl_sync(CG_FULL, own, src), xhl = __shfl_sync(CG_FULL, hl, src);
        u32 ni = xrow, nj = xcol;
        i32 nH = 0;
        if (xcode == DIAG1) { ni = __shfl_sync(CG_FULL, p1, src); nj = xcol - 1; nH = __shfl_sync(CG_FULL, hd1, src); }
        else if (xcode == VERT) { ni = __shfl_sync(CG_FULL, p0, src); nH = __shfl_sync(CG_FULL, hv, src); }
        else if (xcode == VERT1) { ni = __shfl_sync(CG_FULL, p1, src); nH = __shfl_sync(CG_FULL, hv1, src); }
        else if (xcode == HORIZ) { nj = xcol - 1; nH = xhl; if (xown != xhl - 4) { *bad = true; return 0; } }
        else {                                                                   // three or more predecessors, no diagonal match among the first two
            const u32 xdeg = __shfl_sync(CG_FULL, deg, src);
            const i32 xsc = __shfl_sync(CG_FULL, sc, src);
            u32 pe = 0;
            i32 ed = -1, ev = -1;
            if (lane < xdeg) {
                pe = ((const IdT*)&s.prow(xrow - 1))[lane];
                const CgWideCol w1 = cg_wide_col(xcol, CH, inv), w0 = cg_wide_col(xcol - 1, CH, inv);
                const u16* hp = Hh + (size_t)pe * 64u * CH;
                ed = hp[((size_t)w0.c * 32u + w0.lane) * 2u + w0.half];
                ev = hp[((size_t)w1.c * 32u + w1.lane) * 2u + w1.half];
            }
            const u32 md = __ballot_sync(CG_FULL, lane < xdeg && xown == ed + xsc);
            const u32 mv = __ballot_sync(CG_FULL, lane < xdeg && xown == ev - 4);
            if (md) { const int e = __ffs((int)md) - 1; ni = __shfl_sync(CG_FULL, pe, e); nj = xcol - 1; nH = __shfl_sync(CG_FULL, ed, e); }
            else if (mv) { const int e = __ffs((int)mv) - 1; ni = __shfl_sync(CG_FULL, pe, e); nH = __shfl_sync(CG_FULL, ev, e); }
            else { nj = xcol - 1; nH = xhl; if (xown != xhl - 4) { *bad = true; return 0; } }
        }
        if (lane == 0) s.work(n) = (ItemT)((ni == xrow ? IDNONE : xnode) | ((nj == xcol ? IDNONE : (xcol - 1)) << W));
        ++n;
        i = ni; j = nj; Hij = nH;
    }
    __syncwarp();
    return n;
}
