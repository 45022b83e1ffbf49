// k_poa2.cuh — segmented partial-order alignment for small graphs (<= 254 nodes), one warp per region, everything in
// shared memory, and almost nothing left on a single lane.
//
// Same reference semantics as k_poa.cuh (consensus_SPOA BMEAN/bmean.cpp:585-599, vote :649-694, spoa 4.0.0 kSW linear
// m=5 n=-10 g=-4: simd_alignment_engine_impl.hpp:712-1056, Graph::add_alignment graph.cpp:155-272, add_sequence
// :274-292, add_edge :100-116, topological_sort :294-354, MSA columns :372-427); what changes is how the work is laid
// out for a warp:
//
//  * Lazy exact order.  spoa re-runs its DFS topological sort after every added sequence (graph.cpp:271).  The result
//    is read in exactly two places: (1) the order of the score-matrix rows, where it only decides WHICH cell wins when
//    several rows reach the maximum ("first row in rank order", simd...impl.hpp:828-833) — every cell value and the
//    traceback (in-edge order) are the same for any topological order; (2) the MSA column order at the end.
//    So this kernel maintains *a* valid topological order incrementally, in parallel (new nodes of an alignment are
//    spliced in front of the column of the next node they lead to; columns stay contiguous), and runs the reference's
//    DFS only when two rows tie for the maximum (~8 % of the alignments) and once before the vote.
//  * Packed node records: node ids are bytes; in-edges are 8 inline bytes per node (in-degree > 8 leaves the tier);
//    the aligned set (<= 3 others, one node per base in a column) shares a word with the in-degree.  Per row of the
//    matrix there is a descriptor (letter, in-degree, first predecessor row, node) and the 8 predecessor ROW indices,
//    rebuilt by all lanes after each change, so neither the DP nor the traceback chase pointers.
//  * Data-parallel graph update: every query position q maps to exactly one node (prefix chain, aligned part, suffix
//    chain — the id order of graph.cpp:195-201 comes from a warp scan), each node is touched once per alignment, and
//    there is an edge node(q-1) -> node(q) for every q: lanes over q, no conflicts.
//  * Traceback is the only per-alignment phase left on lane 0 (a dependent walk by nature), at two shared-memory
//    round trips per step.
//
// A job that outgrows the tier (nodes, cells, in-degree, segment length) is re-queued for the next tier before
// anything is committed.
#pragma once
#include "cg_common.cuh"
#include "k_poa.cuh"

template <u32 VCAP_, u32 HCELLS_, u32 LCAP_, u32 WARPS_, u32 CTAS_> struct CgPoa2Tier {
    static constexpr u32 VCAP = VCAP_, HCELLS = HCELLS_, LCAP = LCAP_, WARPS = WARPS_, CTAS_PER_SM = CTAS_;
    static constexpr u32 SEGCAP = 192, SCAP = 3 * VCAP_, ALNCAP = VCAP_ + LCAP_;
};
typedef CgPoa2Tier<128, 2048, 64, 4, 5> CgPoa2C1;     // 86 % of the regions of a 150-deep pile
typedef CgPoa2Tier<254, 8192, 120, 4, 2> CgPoa2C2;    // 99.4 %

#define CG_P2_NONE 0xffu

template <class T> struct CgPoa2Lay {
    static constexpr u32 r8(u32 v) { return (v + 7u) / 8u * 8u; }
    static constexpr u32 mx(u32 a, u32 b) { return a > b ? a : b; }
    static constexpr u32 WORK = r8(2 * mx(T::SCAP, T::ALNCAP));          // DFS stack (u16) | alignment pairs (u16)
    static constexpr u32 TMP = r8(mx(2 * T::VCAP, 4 * T::LCAP));         // DFS marks+check | update scratch (4 x LCAP)
    static constexpr u32 o_pred = 0, o_prow = o_pred + 8 * T::VCAP, o_H = o_prow + 8 * T::VCAP, o_rdesc = o_H + r8(2 * T::HCELLS),
                         o_meta = o_rdesc + 4 * T::VCAP, o_seg = o_meta + 4 * T::VCAP, o_nseq = o_seg + 4 * T::SEGCAP,
                         o_work = o_nseq + r8(2 * T::VCAP), o_tmp = o_work + WORK, o_letter = o_tmp + TMP,
                         o_r2n = o_letter + r8(T::VCAP), o_rank = o_r2n + 2 * r8(T::VCAP), o_xr2n = o_rank + r8(T::VCAP),
                         o_xlead = o_xr2n + r8(T::VCAP), o_seq = o_xlead + r8(T::VCAP), per_warp = r8(o_seq + r8(T::LCAP) + 8),
                         R2N_STRIDE = r8(T::VCAP);
    static constexpr u32 cta_bytes = per_warp * T::WARPS;
};

// meta word of a node: byte 0 = nal (bits 0-1) | "sequence 0 passes here" (bit 2) | in-degree (bits 4-7); bytes 1-3 = aligned ids
template <class T> struct CgPoa2G {
    typedef CgPoa2Lay<T> Lay;
    u32 wo;
    __device__ __forceinline__ u8* b() const { return cg_smem_base() + wo; }
    __device__ __forceinline__ u64& pred(u32 i) const { return ((u64*)(b() + Lay::o_pred))[i]; }
    __device__ __forceinline__ u64& prow(u32 i) const { return ((u64*)(b() + Lay::o_prow))[i]; }
    __device__ __forceinline__ i16* H() const { return (i16*)(b() + Lay::o_H); }
    __device__ __forceinline__ u32& rdesc(u32 i) const { return ((u32*)(b() + Lay::o_rdesc))[i]; }
    __device__ __forceinline__ u32& meta(u32 i) const { return ((u32*)(b() + Lay::o_meta))[i]; }
    __device__ __forceinline__ u32& seg(u32 i) const { return ((u32*)(b() + Lay::o_seg))[i]; }
    __device__ __forceinline__ u16& nseq(u32 i) const { return ((u16*)(b() + Lay::o_nseq))[i]; }
    __device__ __forceinline__ u16& work(u32 i) const { return ((u16*)(b() + Lay::o_work))[i]; }
    __device__ __forceinline__ u8& marks(u32 i) const { return b()[Lay::o_tmp + i]; }
    __device__ __forceinline__ u8& check(u32 i) const { return b()[Lay::o_tmp + T::VCAP + i]; }
    __device__ __forceinline__ u8& nodeq(u32 i) const { return b()[Lay::o_tmp + i]; }
    __device__ __forceinline__ u8& kindq(u32 i) const { return b()[Lay::o_tmp + T::LCAP + i]; }
    __device__ __forceinline__ u8& posq(u32 i) const { return b()[Lay::o_tmp + 2 * T::LCAP + i]; }
    __device__ __forceinline__ u8& anchq(u32 i) const { return b()[Lay::o_tmp + 3 * T::LCAP + i]; }
    __device__ __forceinline__ u8& letter(u32 i) const { return b()[Lay::o_letter + i]; }
    __device__ __forceinline__ u8& r2n(u32 which, u32 i) const { return b()[Lay::o_r2n + which * Lay::R2N_STRIDE + i]; }
    __device__ __forceinline__ u8& rank_of(u32 i) const { return b()[Lay::o_rank + i]; }
    __device__ __forceinline__ u8& xr2n(u32 i) const { return b()[Lay::o_xr2n + i]; }
    __device__ __forceinline__ u8& xlead(u32 i) const { return b()[Lay::o_xlead + i]; }
    __device__ __forceinline__ u8* seqbuf() const { return b() + Lay::o_seq; }
};

__device__ __forceinline__ u32 cg_byte64(u64 v, u32 i) { return (u32)(v >> (8u * i)) & 0xffu; }
__device__ __forceinline__ u32 cg_lt_mask() { return (1u << cg_lane()) - 1u; }

// ------------------------------------------------------------------ exact order: spoa's DFS (graph.cpp:294-354), lane 0
// Same walk as cg_poa_toposort (k_poa.cuh) on the packed records.  marks/check are pre-initialised (0 / 1) by the warp.
template <class T> __device__ __forceinline__ bool cg_poa2_dfs(const CgPoa2G<T>& s, u32 V) {
    u32 nrank = 0, sp = 0;
    for (u32 i = 0; i < V; ++i) {
        if (s.marks(i) != 0) continue;
        s.work(sp++) = (u16)i;
        while (sp != 0) {
            const u32 top = s.work(sp - 1);
            const u32 id = top & 0xffu;
            bool finish = (top & 0x100u) != 0;
            const u32 m = s.meta(id);
            const u32 nal = m & 3u;
            if (!finish) {
                if (s.marks(id) == 2) { --sp; continue; }
                const u32 sp0 = sp;
                const u32 deg = (m >> 4) & 15u;
                const u64 P = s.pred(id);
                if (sp + deg + 3 > T::SCAP) return false;
                for (u32 e = 0; e < deg; ++e) {
                    const u32 b = cg_byte64(P, e);
                    if (s.marks(b) != 2) s.work(sp++) = (u16)b;
                }
                if (s.check(id)) {
                    for (u32 a = 0; a < nal; ++a) {
                        const u32 aid = (m >> (8u * (a + 1))) & 0xffu;
                        if (s.marks(aid) != 2) { s.work(sp++) = (u16)aid; s.check(aid) = 0; }
                    }
                }
                if (sp == sp0) finish = true;
                else { s.marks(id) = 1; s.work(sp0 - 1) = (u16)(id | 0x100u); }
            }
            if (finish) {
                s.marks(id) = 2;
                if (s.check(id)) {
                    s.xr2n(nrank) = (u8)id; s.xlead(nrank) = 1; ++nrank;
                    for (u32 a = 0; a < nal; ++a) { s.xr2n(nrank) = (u8)((m >> (8u * (a + 1))) & 0xffu); s.xlead(nrank) = 0; ++nrank; }
                }
                --sp;
            }
        }
    }
    return true;
}

// ------------------------------------------------------------------ traceback (lane 0)
// simd_alignment_engine_impl.hpp:968-1004: diagonal over the predecessors in in-edge order, then vertical over them,
// then horizontal.  Pairs are written in traceback order (last pair first) as node | qpos << 8 (0xff = none).
template <class T> __device__ __forceinline__ u32 cg_poa2_traceback(const CgPoa2G<T>& s, const u8* seq, u32 Wd, u32 bi, u32 bj, bool* bad) {
    const i16* H = s.H();
    u32 i = bi, j = bj, n = 0;
    i32 Hij = H[i * Wd + j];
    while (Hij != 0) {
        const u32 d = s.rdesc(i - 1);
        const u64 pr = s.prow(i - 1);
        const u32 deg = (d >> 8) & 0xffu, np = deg ? deg : 1u;
        u32 pi_ = 0, pj_ = 0;
        i32 Hp = 0;
        bool found = false;
        if (j != 0) {
            const i32 sc = (d & 0xffu) == seq[j - 1] ? 5 : -10;
            for (u32 e = 0; e < np && !found; ++e) {
                const u32 p = deg ? cg_byte64(pr, e) : 0u;
                Hp = H[p * Wd + (j - 1)];
                if (Hij == Hp + sc) { pi_ = p; pj_ = j - 1; found = true; }
            }
        }
        if (!found) {
            for (u32 e = 0; e < np && !found; ++e) {
                const u32 p = deg ? cg_byte64(pr, e) : 0u;
                Hp = H[p * Wd + j];
                if (Hij == Hp - 4) { pi_ = p; pj_ = j; found = true; }
            }
        }
        if (!found && j != 0) {
            Hp = H[i * Wd + j - 1];
            if (Hij == Hp - 4) { pi_ = i; pj_ = j - 1; found = true; }
        }
        if (!found || n >= T::ALNCAP) { *bad = true; return 0; }      // inconsistent matrix: cannot happen
        s.work(n) = (u16)((i == pi_ ? CG_P2_NONE : (d >> 24)) | ((j == pj_ ? CG_P2_NONE : (j - 1)) << 8));
        ++n;
        i = pi_; j = pj_;
        Hij = Hp;
    }
    return n;
}

// ------------------------------------------------------------------ score matrix (all lanes), CH chunks of 32 columns
// tie: some other row reached this lane's maximum again (the caller then needs the exact row order).
template <int CH, class T>
__device__ __forceinline__ void cg_poa2_dp(const CgPoa2G<T>& s, u32 V, const u8* seq, u32 L, i32& bv, u32& bi, u32& bj, bool& tie) {
    const u32 lane = cg_lane(), Wd = L + 1;
    i16* H = s.H();
    u8 q[CH];
    bool act[CH];
    i32 prev[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const u32 j = 1 + 32 * c + lane;
        act[c] = j < Wd;
        q[c] = act[c] ? seq[j - 1] : (u8)0;
        prev[c] = 0;
    }
    for (u32 j = lane; j < Wd; j += 32) H[j] = 0;
    u32 desc_l = 0;
    __syncwarp();
    for (u32 r = 0; r < V; ++r) {
        if ((r & 31u) == 0) desc_l = r + lane < V ? s.rdesc(r + lane) : 0u;
        const u32 d = __shfl_sync(CG_FULL, desc_l, (int)(r & 31u));
        const u8 ch = (u8)(d & 0xffu);
        const u32 deg = (d >> 8) & 0xffu;
        const u32 p0 = (d >> 16) & 0xffu;
        i16* row = H + (r + 1) * Wd;
        i32 val[CH];
        if (deg <= 1 && p0 == r) {                       // predecessor = the row just computed (or the zero row for r = 0)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                i32 left = __shfl_up_sync(CG_FULL, prev[c], 1);
                const i32 l31 = c > 0 ? __shfl_sync(CG_FULL, prev[c > 0 ? c - 1 : 0], 31) : 0;
                if (lane == 0) left = l31;
                const i32 sc = q[c] == ch ? 5 : -10;
                const i32 a = left + sc, b = prev[c] - 4;
                val[c] = a > b ? a : b;
            }
        } else if (deg <= 1) {                           // one predecessor elsewhere, or none (virtual row 0)
            const i16* prow = H + p0 * Wd;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                val[c] = CG_POA_NEG;
                if (act[c]) {
                    const u32 j = 1 + 32 * c + lane;
                    const i32 sc = q[c] == ch ? 5 : -10;
                    const i32 a = (i32)prow[j - 1] + sc, b = (i32)prow[j] - 4;
                    val[c] = a > b ? a : b;
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) val[c] = CG_POA_NEG;
            const u64 pr = s.prow(r);
            for (u32 e = 0; e < deg; ++e) {
                const i16* prow = H + cg_byte64(pr, e) * Wd;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    if (act[c]) {
                        const u32 j = 1 + 32 * c + lane;
                        const i32 sc = q[c] == ch ? 5 : -10;
                        const i32 a = (i32)prow[j - 1] + sc, b = (i32)prow[j] - 4;
                        const i32 m = a > b ? a : b;
                        val[c] = m > val[c] ? m : val[c];
                    }
                }
            }
        }
        // clamp, then the in-row gap term as a max-plus prefix scan over u = H + 4j
        i32 u[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const i32 v0 = val[c] > 0 ? val[c] : 0;
            u[c] = act[c] ? v0 + 4 * (i32)(1 + 32 * c + lane) : CG_POA_NEG;
        }
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const i32 o = __shfl_up_sync(CG_FULL, u[c], dd);
                if (lane >= (u32)dd) u[c] = o > u[c] ? o : u[c];
            }
        }
        i32 carry = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            u[c] = u[c] > carry ? u[c] : carry;
            if (c + 1 < CH) carry = __shfl_sync(CG_FULL, u[c], 31);
            const i32 h = u[c] - 4 * (i32)(1 + 32 * c + lane);
            prev[c] = h;
            if (act[c]) {
                row[1 + 32 * c + lane] = (i16)h;
                if (h > bv) { bv = h; bi = r + 1; bj = 1 + 32 * c + lane; tie = false; }
                else if (h == bv && h > 0 && bi != r + 1) tie = true;
            }
        }
        if (lane == 0) row[0] = 0;
        __syncwarp();
    }
}

// ------------------------------------------------------------------ one job
// Returns the consensus length, or CG_NONE32 if the tier was outgrown (nothing is committed).
template <class T>
__device__ __forceinline__ u32 cg_poa2_job(const CgChunk& c, const CgPoa2G<T>& s, u32 w, u32 rg, u64* cnt_aln, u64* cnt_cells, u64* cnt_pred) {
    const u32 lane = cg_lane();
    const CgWin W = c.win[w];
    CgRegion* R = &c.regions[c.off_reg[w] + rg];
    if (R->n > T::SEGCAP || R->max_len > T::LCAP) return CG_NONE32;
    CgWinView v;
    v.seq_off = c.seq_off + W.seq_begin; v.pos = c.pos + c.off_pos[w]; v.chain = c.chain + c.off_slot[w];
    v.rel = c.rel + c.off_slot[w]; v.N = W.n_seqs; v.C = W.n_cand; v.nA = W.n_chain;
    const u8* bases = (const u8*)c.bases;

    // ---- the region's segments, in read order (split_reads): read | start << 12 | len << 25
    u32 nseg = 0;
    for (u32 rb = 0; rb < v.N; rb += 32) {
        const u32 r = rb + lane;
        u32 st = 0, ln = 0;
        const bool keep = r < v.N && cg_eval_segment(v, rg, r, &st, &ln);
        const u32 bal = __ballot_sync(CG_FULL, keep);
        if (keep) {
            const u32 idx = nseg + __popc(bal & cg_lt_mask());
            if (idx < T::SEGCAP) s.seg(idx) = r | (st << 12) | (ln << 25);
        }
        nseg += __popc(bal);
    }
    if (nseg > T::SEGCAP) return CG_NONE32;
    __syncwarp();

    u32 V = 0, nseqs = 0, cur = 0, sumdeg = 0;
    bool dfs_valid = false;
    u64 j_aln = 0, j_cells = 0, j_pred = 0;

    for (u32 si = 0; si < nseg; ++si) {
        const u32 sg = s.seg(si);
        const u32 L = sg >> 25;
        if (L == 0) continue;                                        // graph.cpp:160 — not a row of the MSA
        u8* seq = s.seqbuf();
        {
            const u8* src = bases + v.seq_off[sg & 0xfffu] + ((sg >> 12) & 0x1fffu);
            __syncwarp();
            for (u32 i = lane; i < L; i += 32) seq[i] = src[i];
            __syncwarp();
        }
        const u32 Wd = L + 1;
        u32 n_aln = 0;
        if (V != 0) {
            if ((V + 1) * Wd > T::HCELLS) return CG_NONE32;
            i32 bv = 0; u32 bi = 0, bj = 0;
            bool tie = false;
            if (L <= 32) cg_poa2_dp<1>(s, V, seq, L, bv, bi, bj, tie);
            else if (L <= 64) cg_poa2_dp<2>(s, V, seq, L, bv, bi, bj, tie);
            else cg_poa2_dp<4>(s, V, seq, L, bv, bi, bj, tie);
            j_aln += 1; j_cells += (u64)(V + 1) * L; j_pred += (u64)sumdeg * L;
            const u64 key = ((u64)(u32)bv << 32) | ((u64)(0xffffu - bi) << 16) | (u64)(0xffffu - bj);
            const u64 kb = cg_warp_max64(key);
            const i32 M = (i32)(kb >> 32);
            const u32 gbi = 0xffffu - (u32)((kb >> 16) & 0xffffu), gbj = 0xffffu - (u32)(kb & 0xffffu);
            const bool mytie = M > 0 && bv == M && (tie || bi != gbi);
            bi = gbi; bj = gbj;
            if (__any_sync(CG_FULL, mytie)) {
                // several rows reach the maximum: the winner is the first of them in spoa's own order (simd...impl.hpp:828-833)
                if (!dfs_valid) {
                    for (u32 i = lane; i < V; i += 32) { s.marks(i) = 0; s.check(i) = 1; }
                    __syncwarp();
                    bool ok = true;
                    if (lane == 0) ok = cg_poa2_dfs(s, V);
                    ok = __shfl_sync(CG_FULL, (u32)ok, 0) != 0;
                    if (!ok) return CG_NONE32;
                    dfs_valid = true;
                    __syncwarp();
                }
                const i16* H = s.H();
                u32 row = 0;
                for (u32 ib = 0; ib < V; ib += 32) {
                    const u32 i = ib + lane;
                    bool hit = false;
                    u32 myrow = 0;
                    if (i < V) {
                        myrow = (u32)s.rank_of(s.xr2n(i)) + 1;
                        const i16* hr = H + myrow * Wd;
                        for (u32 j = 1; j < Wd; ++j) hit = hit || (i32)hr[j] == M;
                    }
                    const u32 bal = __ballot_sync(CG_FULL, hit);
                    if (bal) { row = __shfl_sync(CG_FULL, myrow, __ffs((int)bal) - 1); break; }
                }
                bi = row; bj = 0;
                const i16* hr = H + row * Wd;
                for (u32 jb = 1; jb < Wd; jb += 32) {
                    const u32 j = jb + lane;
                    const u32 bal = __ballot_sync(CG_FULL, j < Wd && (i32)hr[j] == M);
                    if (bal) { bj = jb + (u32)__ffs((int)bal) - 1; break; }
                }
            }
            bool bad = false;
            if (lane == 0 && M > 0) n_aln = cg_poa2_traceback(s, seq, Wd, bi, bj, &bad);
            n_aln = __shfl_sync(CG_FULL, n_aln, 0);
            if (__shfl_sync(CG_FULL, (u32)bad, 0)) return CG_NONE32;
        }
        __syncwarp();

        // ---- graph update, all lanes (graph.cpp:155-272).  Pair t in path order = work[n_aln - 1 - t].
        u32 first_valid = L, last_valid = 0;
        {
            u32 mn = 0xffffu, mxq = 0;
            for (u32 t = lane; t < n_aln; t += 32) {
                const u32 qp = (u32)s.work(t) >> 8;
                if (qp != CG_P2_NONE) { mn = qp < mn ? qp : mn; mxq = qp > mxq ? qp : mxq; }
            }
#pragma unroll
            for (int dlt = 16; dlt > 0; dlt >>= 1) {
                const u32 o1 = __shfl_xor_sync(CG_FULL, mn, dlt), o2 = __shfl_xor_sync(CG_FULL, mxq, dlt);
                mn = o1 < mn ? o1 : mn; mxq = o2 > mxq ? o2 : mxq;
            }
            if (mn != 0xffffu) { first_valid = mn; last_valid = mxq; }
        }
        const bool has_aln = first_valid != L;
        const u32 V0 = V;
        const u32 n_prefix = first_valid, n_suffix = has_aln ? L - 1 - last_valid : 0u;
        const u32 mid_base = V0 + n_prefix + n_suffix;
        bool ovf = false;
        u32 n_mid_new = 0;
        // U1: the aligned part — which node does each consumed pair resolve to?
        for (u32 tb = 0; tb < n_aln; tb += 32) {
            const u32 t = tb + lane;
            u32 kind = 3, nn = 0, an = CG_P2_NONE, qp = CG_P2_NONE;
            u8 ch = 0;
            u32 m_an = 0;
            if (t < n_aln) {
                const u32 pr = s.work(n_aln - 1 - t);
                an = pr & 0xffu; qp = pr >> 8;
                if (qp != CG_P2_NONE) {
                    ch = seq[qp];
                    if (an == CG_P2_NONE) kind = 1;                                   // new node, not aligned to anything
                    else if (s.letter(an) == ch) { kind = 0; nn = an; }
                    else {
                        m_an = s.meta(an);
                        kind = 2;                                                     // new node in an's column ...
                        const u32 nal = m_an & 3u;
                        for (u32 a = 0; a < nal; ++a) {
                            const u32 aid = (m_an >> (8u * (a + 1))) & 0xffu;
                            if (s.letter(aid) == ch) { kind = 0; nn = aid; }          // ... unless the column already has the letter
                        }
                    }
                }
            }
            const bool isnew = kind == 1 || kind == 2;
            const u32 bal = __ballot_sync(CG_FULL, isnew);
            if (isnew) nn = mid_base + n_mid_new + __popc(bal & cg_lt_mask());
            n_mid_new += __popc(bal);
            if (isnew && nn >= T::VCAP) ovf = true;
            else if (kind == 2) {
                const u32 nal = m_an & 3u;
                if (nal >= 3) ovf = true;                                             // cannot happen with ACGT input
                else {
                    // aligned(nn) = aligned(an) + [an]; every member of the column appends nn (graph.cpp:232-243)
                    s.letter(nn) = ch; s.nseq(nn) = 0; s.pred(nn) = ~0ull;
                    s.meta(nn) = (nal + 1) | (m_an & 0xffffff00u & ~(0xffffffffu << (8u * (nal + 1)))) | (an << (8u * (nal + 1)));
                    for (u32 a = 0; a < nal; ++a) {
                        const u32 aid = (m_an >> (8u * (a + 1))) & 0xffu;
                        const u32 ma = s.meta(aid);
                        s.meta(aid) = (ma + 1) | (nn << (8u * ((ma & 3u) + 1)));
                    }
                    s.meta(an) = (m_an + 1) | (nn << (8u * (nal + 1)));
                }
            }
            if (qp != CG_P2_NONE && !(isnew && nn >= T::VCAP)) {
                s.nodeq(qp) = (u8)nn; s.kindq(qp) = (u8)kind; s.anchq(qp) = (u8)(kind == 1 ? CG_P2_NONE : an);
            }
        }
        V = mid_base + n_mid_new;
        if (V > T::VCAP) ovf = true;
        if (__any_sync(CG_FULL, ovf)) return CG_NONE32;
        __syncwarp();
        // U2a: every query position -> its node; new unaligned nodes are initialised here
        for (u32 qb = 0; qb < L; qb += 32) {
            const u32 q = qb + lane;
            if (q < L) {
                u32 node, kind;
                if (q < first_valid) { node = V0 + q; kind = 1; }
                else if (q > last_valid) { node = V0 + n_prefix + (q - last_valid - 1); kind = 1; }
                else { node = s.nodeq(q); kind = s.kindq(q); }
                if (q < first_valid || q > last_valid) { s.nodeq(q) = (u8)node; s.kindq(q) = 1; s.anchq(q) = CG_P2_NONE; }
                if (kind == 1) { s.letter(node) = seq[q]; s.nseq(node) = 0; s.meta(node) = 0; s.pred(node) = ~0ull; }
            }
        }
        __syncwarp();
        // U2b: one visit per position, one edge node(q-1) -> node(q) (an existing edge is reused, graph.cpp:105-110)
        bool changed = V != V0;
        for (u32 qb = 0; qb < L; qb += 32) {
            const u32 q = qb + lane;
            if (q < L) {
                const u32 node = s.nodeq(q);
                s.nseq(node) = (u16)(s.nseq(node) + 1);
                u32 m = s.meta(node);
                if (nseqs == 0) m |= 4u;
                if (q > 0) {
                    const u32 src = s.nodeq(q - 1);
                    const u32 deg = (m >> 4) & 15u;
                    u64 P = s.pred(node);
                    bool have = false;
                    for (u32 e = 0; e < deg; ++e) have = have || cg_byte64(P, e) == src;
                    if (!have) {
                        if (deg >= 8) ovf = true;
                        else {
                            P = (P & ~(0xffull << (8u * deg))) | ((u64)src << (8u * deg));
                            s.pred(node) = P;
                            m += 16u;
                            changed = true;
                        }
                    }
                }
                s.meta(node) = m;
            }
        }
        nseqs++;
        if (__any_sync(CG_FULL, ovf)) return CG_NONE32;
        if (!__any_sync(CG_FULL, changed)) { __syncwarp(); continue; }      // same nodes, same edges: same order, same rows
        dfs_valid = false;
        __syncwarp();

        // ---- the incremental order: splice the new nodes into the old order (columns stay contiguous)
        // U3a: position (in the old order) in front of which each new node goes
        u32 carry_min = CG_P2_NONE;                                   // smallest column start among the anchored positions behind
        for (int qb = (int)((L - 1) & ~31u); qb >= 0; qb -= 32) {
            const u32 q = (u32)qb + lane;
            u32 cs = CG_P2_NONE, ce = 0;
            u32 kind = 3;
            if (q < L) {
                kind = s.kindq(q);
                const u32 an = s.anchq(q);
                if (an != CG_P2_NONE) {                               // column of the anchor among the OLD nodes
                    const u32 m = s.meta(an);
                    const u32 nal = m & 3u;
                    const u32 r0 = s.rank_of(an);
                    cs = r0; ce = r0 + 1;
                    for (u32 a = 0; a < nal; ++a) {
                        const u32 aid = (m >> (8u * (a + 1))) & 0xffu;
                        if (aid < V0) { const u32 ra = s.rank_of(aid); cs = ra < cs ? ra : cs; ce = ra + 1 > ce ? ra + 1 : ce; }
                    }
                }
            }
            // exclusive suffix minimum of cs over q (positions never decrease along the path)
            u32 sm = cs;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                const u32 o = __shfl_down_sync(CG_FULL, sm, dlt);
                if (lane + (u32)dlt < 32u) sm = o < sm ? o : sm;
            }
            sm = sm < carry_min ? sm : carry_min;
            if (q < L) {
                u32 pos = CG_P2_NONE;
                if (kind == 2) pos = ce;
                else if (kind == 1) pos = sm == CG_P2_NONE ? V0 : sm;
                s.posq(q) = (u8)pos;
            }
            carry_min = __shfl_sync(CG_FULL, sm, 0);
        }
        __syncwarp();
        // U3b: the new nodes in q order -> work[k] = pos | node << 8
        u32 K = 0;
        for (u32 qb = 0; qb < L; qb += 32) {
            const u32 q = qb + lane;
            const bool isnew = q < L && s.kindq(q) != 0;
            const u32 bal = __ballot_sync(CG_FULL, isnew);
            if (isnew) s.work(K + __popc(bal & cg_lt_mask())) = (u16)((u32)s.posq(q) | ((u32)s.nodeq(q) << 8));
            K += __popc(bal);
        }
        __syncwarp();
        // U3c: old entry p moves up by the number of new nodes placed at or before it; new node k lands at pos_k + k
        const u32 nxt = cur ^ 1u;
        for (u32 pb = 0; pb < V0; pb += 32) {
            const u32 p = pb + lane;
            if (p < V0) {
                u32 lo = 0, hi = K;                                   // first k with pos_k > p
                while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (((u32)s.work(mid) & 0xffu) <= p) lo = mid + 1; else hi = mid; }
                const u32 node = s.r2n(cur, p);
                s.r2n(nxt, p + lo) = (u8)node;
                s.rank_of(node) = (u8)(p + lo);
            }
        }
        for (u32 kb2 = 0; kb2 < K; kb2 += 32) {
            const u32 k = kb2 + lane;
            if (k < K) {
                const u32 it = s.work(k);
                const u32 node = it >> 8, at = (it & 0xffu) + k;
                s.r2n(nxt, at) = (u8)node;
                s.rank_of(node) = (u8)at;
            }
        }
        cur = nxt;
        __syncwarp();
        // U4: row descriptors: letter | in-degree << 8 | first predecessor row << 16 | node << 24, and the 8 predecessor rows
        u32 sd = 0;
        for (u32 rb = 0; rb < V; rb += 32) {
            const u32 r = rb + lane;
            if (r < V) {
                const u32 node = s.r2n(cur, r);
                const u32 deg = (s.meta(node) >> 4) & 15u;
                const u64 P = s.pred(node);
                u64 rows = 0;
#pragma unroll
                for (u32 e = 0; e < 8; ++e)
                    if (e < deg) rows |= (u64)((u32)s.rank_of(cg_byte64(P, e)) + 1u) << (8u * e);
                s.prow(r) = rows;
                s.rdesc(r) = (u32)s.letter(node) | (deg << 8) | (((u32)rows & 0xffu) << 16) | (node << 24);
                sd += deg ? deg : 1u;
            }
        }
        sumdeg = cg_warp_sum(sd);
        __syncwarp();
    }

    // ---- exact column order for the vote
    if (!dfs_valid && V != 0) {
        for (u32 i = lane; i < V; i += 32) { s.marks(i) = 0; s.check(i) = 1; }
        __syncwarp();
        bool ok = true;
        if (lane == 0) ok = cg_poa2_dfs(s, V);
        ok = __shfl_sync(CG_FULL, (u32)ok, 0) != 0;
        if (!ok) return CG_NONE32;
        __syncwarp();
    }

    // ---- column vote (bmean.cpp:649-694) straight off the graph: a column = a leader and its aligned nodes
    u8* out = c.arena + c.off_arena[w] + R->arena_off;
    u32 outn = 0;
    for (u32 ib = 0; ib < V; ib += 32) {
        const u32 i = ib + lane;
        u8 emit = 0;
        if (i < V && s.xlead(i)) {
            u32 cnt[4] = {0, 0, 0, 0};
            u8 row0 = 0;
            const u32 node = s.xr2n(i);
            const u32 m = s.meta(node);
            const u32 na = m & 3u;
            for (u32 a = 0; a <= na; ++a) {
                const u32 x = a == 0 ? node : (m >> (8u * a)) & 0xffu;
                const u8 ch = s.letter(x);
                const u32 code = cg_base_code(ch) & 3u;
                cnt[code] = s.nseq(x);
                if (s.meta(x) & 4u) row0 = ch;
            }
            const u32 cA = cnt[0], cC = cnt[1], cG = cnt[2], cT = cnt[3];
            const u32 cM = nseqs - (cA + cC + cG + cT);
            if (cM > cA && cM > cC && cM > cT && cM > cG) emit = 0;
            else if (cA > cC && cA > cG && cA > cT) emit = 'A';
            else if (cC > cA && cC > cG && cC > cT) emit = 'C';
            else if (cG > cA && cG > cC && cG > cT) emit = 'G';
            else if (cT > cA && cT > cG && cT > cC) emit = 'T';
            else emit = row0;                                        // row 0's letter, if it has one here
        }
        const u32 bal = __ballot_sync(CG_FULL, emit != 0);
        if (emit) out[outn + __popc(bal & cg_lt_mask())] = emit;
        outn += __popc(bal);
    }
    *cnt_aln += j_aln; *cnt_cells += j_cells; *cnt_pred += j_pred;
    return outn;
}

// Persistent warps over the tier's queue (same protocol as cg_poa_drain).
template <class T>
__global__ void __launch_bounds__(T::WARPS * 32, T::CTAS_PER_SM) k_poa2(CgChunk c, u32 nwarps, const uint2* jobs, u32* qctl, uint2* jobs_next,
                                                                      u32* qnext) {
    const u32 gw = blockIdx.x * T::WARPS + cg_warp();
    if (gw >= nwarps) return;                       // warp-uniform; no block-wide barrier in this kernel
    CgPoa2G<T> s;
    s.wo = CgPoa2Lay<T>::per_warp * cg_warp();
    const u32 lane = cg_lane();
    const u32 nfront = qctl[0], nback = qctl[2], cap = qctl[3];
    u64 cnt_aln = 0, cnt_cells = 0, cnt_pred = 0;
    for (;;) {
        u32 j = 0;
        if (lane == 0) j = atomicAdd(&qctl[1], 1u);
        j = __shfl_sync(CG_FULL, j, 0);
        if (j >= nfront + nback) break;
        const uint2 job = j < nfront ? jobs[j] : jobs[cap - 1 - (j - nfront)];
        const u32 n = cg_poa2_job(c, s, job.x, job.y, &cnt_aln, &cnt_cells, &cnt_pred);
        if (lane == 0) {
            if (n == CG_NONE32) cg_queue_push_front(jobs_next, qnext, job);
            else c.regions[c.off_reg[job.x] + job.y].cons_len = n;
        }
        __syncwarp();
    }
    if (lane == 0 && cnt_aln) {
        atomicAdd((unsigned long long*)&c.counters->alignments, (unsigned long long)cnt_aln);
        atomicAdd((unsigned long long*)&c.counters->dp_cells, (unsigned long long)cnt_cells);
        atomicAdd((unsigned long long*)&c.counters->dp_pred_cells, (unsigned long long)cnt_pred);
    }
}
