// k_poa.cuh — the last-resort POA tier: one warp per region, everything in global memory, no limit on in-degree,
// graphs up to 65 535 nodes, segments up to 6000 bases.  Regions normally never get here: k_poa2.cuh's tiers take them
// (and document the algorithm); a region lands here only when a node collects more than 16 in-edges or the graph
// outgrows 4096 nodes / 2048-base segments / 4 M matrix cells.
//
// Replaces consensus_SPOA (BMEAN/bmean.cpp:585-599), the vote of easy_consensus (:649-694) and the spoa 4.0.0
// subset CONSENT uses: local (kSW) alignment with linear gaps m=5 n=-10 g=-4
// (spoa/src/simd_alignment_engine_impl.hpp:712-1056 == sisd_alignment_engine.cpp:260-435), Graph::add_alignment
// (graph.cpp:155-272), add_sequence (:274-292), add_edge (:100-116), topological_sort (:294-354) and
// generate_multiple_sequence_alignment (:372-427).
//
// What is kept of spoa's graph — exactly what its results depend on:
//   node  : letter, ordered in-edge list (predecessor order drives the traceback tie-breaks), aligned set
//           (<= 3 others: a column holds at most one node per base), #sequences through it, "sequence 0 is here"
//   order : rank <-> node from the explicit-stack DFS of topological_sort, re-run after every sequence as spoa does
// Out-edges and per-edge sequence labels are not stored: edge existence is tested on the in-list, and the MSA
// rows are never built — the vote only needs, per column, how many sequences pass through each of its nodes
// (gaps = the rest) and row 0's letter.  The score matrix is swept row by row in rank order with the 32 lanes across
// query columns (max-plus prefix scan for the in-row gap term); traceback, graph update and the DFS run on lane 0.
#pragma once
#include "cg_common.cuh"

#define CG_POA_WARPS_PER_CTA 4u
#define CG_POA_THREADS (CG_POA_WARPS_PER_CTA * 32u)
#define CG_POA_NEG (-(1 << 28))

// ------------------------------------------------------------------ storage policies
#ifdef CG_EMU
__device__ __forceinline__ u8* cg_smem_base() { return cg_emu::cur->blk->dyn_smem; }
#else
extern __shared__ __align__(128) unsigned char cg_dyn_smem_[];
__device__ __forceinline__ u8* cg_smem_base() { return cg_dyn_smem_; }
#endif

// Everything in global memory (per-warp scratch described by CgPoaScratch).
struct CgPoaGlobG {
    typedef u32 eidx;
    typedef u32 stk_t;
    static constexpr u32 ENONE = 0xffffffffu, STK_FLAG = 0x80000000u;
    CgPoaScratch gs;
    __device__ __forceinline__ u8& letter(u32 i) const { return gs.letter[i]; }
    __device__ __forceinline__ u8& in0(u32 i) const { return gs.in0[i]; }
    __device__ __forceinline__ u8& nal(u32 i) const { return gs.nal[i]; }
    __device__ __forceinline__ u8& leader(u32 i) const { return gs.leader[i]; }
    __device__ __forceinline__ u8& marks(u32 i) const { return gs.marks[i]; }
    __device__ __forceinline__ u8& check(u32 i) const { return gs.check[i]; }
    __device__ __forceinline__ u16& nseq(u32 i) const { return gs.nseq[i]; }
    __device__ __forceinline__ u16& aligned(u32 i) const { return gs.aligned[i]; }
    __device__ __forceinline__ u16& rank_of(u32 i) const { return gs.rank_of[i]; }
    __device__ __forceinline__ u16& r2n(u32 i) const { return gs.r2n[i]; }
    __device__ __forceinline__ u32& in_head(u32 i) const { return gs.in_head[i]; }
    __device__ __forceinline__ u32& in_tail(u32 i) const { return gs.in_tail[i]; }
    __device__ __forceinline__ u32& rdesc(u32 i) const { return gs.rdesc[i]; }
    __device__ __forceinline__ u16& e_pred(u32 i) const { return gs.e_pred[i]; }
    __device__ __forceinline__ u32& e_next(u32 i) const { return gs.e_next[i]; }
    __device__ __forceinline__ u32& stack(u32 i) const { return gs.stack[i]; }
    __device__ __forceinline__ void aln_set(u32 i, i32 node, i32 pos) const { gs.aln_node[i] = node; gs.aln_pos[i] = pos; }
    __device__ __forceinline__ i32 aln_node(u32 i) const { return gs.aln_node[i]; }
    __device__ __forceinline__ i32 aln_pos(u32 i) const { return gs.aln_pos[i]; }
    __device__ __forceinline__ i16* H() const { return gs.H; }
    __device__ __forceinline__ u8* seqbuf() const { return nullptr; }
    __device__ __forceinline__ u32 seqcap() const { return 0; }
    __device__ __forceinline__ u32 vcap() const { return gs.vcap; }
    __device__ __forceinline__ u32 ecap() const { return gs.ecap; }
    __device__ __forceinline__ u32 scap() const { return gs.scap; }
    __device__ __forceinline__ u32 alncap() const { return gs.alncap; }
    __device__ __forceinline__ u64 hcap() const { return gs.hcap; }
};

struct CgPoaState {
    u32 V, E, nseqs;
    bool ovf;
};

// ------------------------------------------------------------------ graph update (lane 0)
template <class G> __device__ __forceinline__ u32 cg_poa_add_node(const G& s, CgPoaState& g, u8 letter) {
    if (g.V >= s.vcap()) { g.ovf = true; return 0; }
    const u32 id = g.V++;
    s.letter(id) = letter; s.in0(id) = 0; s.nal(id) = 0; s.nseq(id) = 0;
    s.in_head(id) = (typename G::eidx)G::ENONE; s.in_tail(id) = (typename G::eidx)G::ENONE;
    return id;
}
// graph.cpp:100-116 — an existing edge is reused (only its label list would grow; labels are not needed here)
template <class G> __device__ __forceinline__ void cg_poa_add_edge(const G& s, CgPoaState& g, u32 b, u32 e) {
    for (u32 ee = s.in_head(e); ee != G::ENONE; ee = s.e_next(ee))
        if (s.e_pred(ee) == b) return;
    if (g.E >= s.ecap()) { g.ovf = true; return; }
    const u32 id = g.E++;
    s.e_pred(id) = (u16)b; s.e_next(id) = (typename G::eidx)G::ENONE;
    const u32 tl = s.in_tail(e);
    if (tl == G::ENONE) s.in_head(e) = (typename G::eidx)id; else s.e_next(tl) = (typename G::eidx)id;
    s.in_tail(e) = (typename G::eidx)id;
}
template <class G> __device__ __forceinline__ void cg_poa_visit(const G& s, const CgPoaState& g, u32 node) {
    s.nseq(node) = (u16)(s.nseq(node) + 1);
    if (g.nseqs == 0) s.in0(node) = 1;
}
// graph.cpp:274-292 — a chain of new nodes for seq[begin,end); -1 if empty
template <class G> __device__ __forceinline__ i32 cg_poa_add_sequence(const G& s, CgPoaState& g, const u8* seq, u32 begin, u32 end) {
    if (begin == end) return -1;
    const u32 first = cg_poa_add_node(s, g, seq[begin]);
    cg_poa_visit(s, g, first);
    u32 prev = first;
    for (u32 i = begin + 1; i < end && !g.ovf; ++i) {
        const u32 id = cg_poa_add_node(s, g, seq[i]);
        cg_poa_visit(s, g, id);
        cg_poa_add_edge(s, g, prev, id);
        prev = id;
    }
    return (i32)first;
}

// graph.cpp:155-272.  The alignment is stored in traceback order (last pair first): index n_aln-1 .. 0.
template <class G> __device__ __forceinline__ void cg_poa_add_alignment(const G& s, CgPoaState& g, u32 n_aln, const u8* seq, u32 L) {
    if (n_aln == 0) {
        cg_poa_add_sequence(s, g, seq, 0, L);
        g.nseqs++;
        return;
    }
    i32 first_valid = -1, last_valid = -1;
    for (i32 i = (i32)n_aln - 1; i >= 0; --i) {
        const i32 qp = s.aln_pos((u32)i);
        if (qp != -1) { if (first_valid == -1) first_valid = qp; last_valid = qp; }
    }
    const u32 tmp = g.V;
    cg_poa_add_sequence(s, g, seq, 0, (u32)first_valid);
    i32 head = tmp == g.V ? -1 : (i32)g.V - 1;
    const i32 tail = cg_poa_add_sequence(s, g, seq, (u32)last_valid + 1, L);
    for (i32 i = (i32)n_aln - 1; i >= 0 && !g.ovf; --i) {
        const i32 qp = s.aln_pos((u32)i);
        if (qp == -1) continue;
        const u8 letter = seq[qp];
        const i32 an = s.aln_node((u32)i);
        u32 nn;
        if (an == -1) {
            nn = cg_poa_add_node(s, g, letter);
        } else if (s.letter((u32)an) == letter) {
            nn = (u32)an;
        } else {
            i32 aligned_to = -1;
            const u32 na = s.nal((u32)an);
            for (u32 a = 0; a < na; ++a) {
                const u32 aid = s.aligned(3 * (u32)an + a);
                if (s.letter(aid) == letter) { aligned_to = (i32)aid; break; }
            }
            if (aligned_to == -1) {
                nn = cg_poa_add_node(s, g, letter);
                if (g.ovf) break;
                if (na >= 3) { g.ovf = true; break; }          // cannot happen with ACGT input
                for (u32 a = 0; a < na; ++a) {
                    const u32 aid = s.aligned(3 * (u32)an + a);
                    s.aligned(3 * nn + s.nal(nn)) = (u16)aid; s.nal(nn)++;
                    s.aligned(3 * aid + s.nal(aid)) = (u16)nn; s.nal(aid)++;
                }
                s.aligned(3 * nn + s.nal(nn)) = (u16)an; s.nal(nn)++;
                s.aligned(3 * (u32)an + s.nal((u32)an)) = (u16)nn; s.nal((u32)an)++;
            } else nn = (u32)aligned_to;
        }
        if (g.ovf) break;
        cg_poa_visit(s, g, nn);
        if (head != -1) cg_poa_add_edge(s, g, (u32)head, nn);
        head = (i32)nn;
    }
    if (tail != -1 && !g.ovf) cg_poa_add_edge(s, g, (u32)head, (u32)tail);
    g.nseqs++;
}

// graph.cpp:294-354 — the explicit-stack DFS over node ids, step for step, with one shortcut that cannot change
// its output: when a node comes back to the top of the stack after everything it pushed has been popped, all of
// its predecessors and aligned nodes are finished (each was pushed above it, or was finished already), so the
// reference's second scan of its lists always succeeds; a flag on the stack entry replaces that scan.
// marks/check are pre-initialised (0 / 1) by the warp.
template <class G> __device__ __forceinline__ void cg_poa_toposort(const G& s, CgPoaState& g) {
    u32 nrank = 0, sp = 0;
    const u32 V = g.V, scap = s.scap();
    for (u32 i = 0; i < V; ++i) {
        if (s.marks(i) != 0) continue;
        s.stack(sp++) = (typename G::stk_t)i;
        while (sp != 0) {
            const u32 top = s.stack(sp - 1);
            const u32 id = top & ~G::STK_FLAG;
            bool finish = (top & G::STK_FLAG) != 0;
            if (!finish) {
                if (s.marks(id) == 2) { --sp; continue; }
                const u32 sp0 = sp;
                for (u32 ee = s.in_head(id); ee != G::ENONE; ee = s.e_next(ee)) {
                    const u32 b = s.e_pred(ee);
                    if (s.marks(b) != 2) {
                        if (sp >= scap) { g.ovf = true; return; }
                        s.stack(sp++) = (typename G::stk_t)b;
                    }
                }
                if (s.check(id)) {
                    const u32 na = s.nal(id);
                    for (u32 a = 0; a < na; ++a) {
                        const u32 aid = s.aligned(3 * id + a);
                        if (s.marks(aid) != 2) {
                            if (sp >= scap) { g.ovf = true; return; }
                            s.stack(sp++) = (typename G::stk_t)aid; s.check(aid) = 0;
                        }
                    }
                }
                if (sp == sp0) finish = true;
                else { s.marks(id) = 1; s.stack(sp0 - 1) = (typename G::stk_t)(id | G::STK_FLAG); }
            }
            if (finish) {
                s.marks(id) = 2;
                if (s.check(id)) {
                    const u32 na = s.nal(id);
                    s.r2n(nrank) = (u16)id; s.leader(nrank) = 1; ++nrank;
                    for (u32 a = 0; a < na; ++a) { s.r2n(nrank) = s.aligned(3 * id + a); s.leader(nrank) = 0; ++nrank; }
                }
                --sp;
            }
        }
    }
}

// Traceback (simd_alignment_engine_impl.hpp:968-1004 == sisd_alignment_engine.cpp:392-431): diagonal over the
// predecessors in in-edge order, then vertical over them, then horizontal.  Returns the number of pairs.
template <class G> __device__ __forceinline__ u32 cg_poa_traceback(const G& s, CgPoaState& g, const u8* seq, u32 Wd, u32 bi, u32 bj) {
    const i16* H = s.H();
    const u32 alncap = s.alncap();
    u32 i = bi, j = bj, n = 0, pi_ = 0, pj_ = 0;
    i32 Hij = H[(size_t)i * Wd + j];
    while (Hij != 0) {
        const u32 node = s.r2n(i - 1);
        const u32 eh = s.in_head(node);
        bool found = false;
        i32 Hp = 0;
        if (j != 0) {
            const i32 sc = s.letter(node) == seq[j - 1] ? 5 : -10;
            if (eh == G::ENONE) {
                Hp = H[j - 1];
                if (Hij == Hp + sc) { pi_ = 0; pj_ = j - 1; found = true; }
            } else {
                for (u32 ee = eh; ee != G::ENONE && !found; ee = s.e_next(ee)) {
                    const u32 pr = (u32)s.rank_of(s.e_pred(ee)) + 1;
                    Hp = H[(size_t)pr * Wd + (j - 1)];
                    if (Hij == Hp + sc) { pi_ = pr; pj_ = j - 1; found = true; }
                }
            }
        }
        if (!found) {
            if (eh == G::ENONE) {
                Hp = H[j];
                if (Hij == Hp - 4) { pi_ = 0; pj_ = j; found = true; }
            } else {
                for (u32 ee = eh; ee != G::ENONE && !found; ee = s.e_next(ee)) {
                    const u32 pr = (u32)s.rank_of(s.e_pred(ee)) + 1;
                    Hp = H[(size_t)pr * Wd + j];
                    if (Hij == Hp - 4) { pi_ = pr; pj_ = j; found = true; }
                }
            }
        }
        if (!found && j != 0) {
            Hp = H[(size_t)i * Wd + j - 1];
            if (Hij == Hp - 4) { pi_ = i; pj_ = j - 1; found = true; }
        }
        if (!found || n >= alncap) { g.ovf = true; return 0; }       // inconsistent matrix: cannot happen
        s.aln_set(n, i == pi_ ? -1 : (i32)node, j == pj_ ? -1 : (i32)(j - 1));
        ++n;
        i = pi_; j = pj_;
        Hij = Hp;                                                    // the value just compared is the next cell
    }
    return n;
}

// Row descriptors: rank r -> letter | in-degree << 8 | (row index of the first predecessor) << 16  (0 = the virtual
// start row).  Built by all lanes after every topological sort, so that the row loop of the score matrix has no
// dependent pointer chasing: 32 descriptors are fetched at once and broadcast by shuffle.
template <class G> __device__ __forceinline__ void cg_poa_build_rdesc(const G& s, u32 V) {
    for (u32 r = cg_lane(); r < V; r += 32) {
        const u32 node = s.r2n(r);
        const u32 eh = s.in_head(node);
        u32 deg = 0;
        for (u32 ee = eh; ee != G::ENONE; ee = s.e_next(ee)) ++deg;
        const u32 p0 = eh == G::ENONE ? 0u : (u32)s.rank_of(s.e_pred(eh)) + 1u;
        s.rdesc(r) = (u32)s.letter(node) | ((deg > 255u ? 255u : deg) << 8) | (p0 << 16);
    }
}

// ------------------------------------------------------------------ score matrix (all lanes)
// CH chunks of 32 query columns per row held in registers.  The common row (one predecessor = the row just
// computed) needs no loads at all: the left neighbour comes from the adjacent lane.  Other rows read their
// predecessor rows back from the stored matrix.
template <int CH, class G>
__device__ __forceinline__ void cg_poa_dp(const G& s, u32 V, const u8* seq, u32 L, i32& bv, u32& bi, u32& bj, u64& pred_rows) {
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
        u32 deg = (d >> 8) & 0xffu;
        const u32 p0 = d >> 16;
        i16* row = H + (size_t)(r + 1) * Wd;
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
            const i16* prow = H + (size_t)p0 * Wd;
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
            u32 n = 0;
            for (u32 ee = s.in_head(s.r2n(r)); ee != G::ENONE; ee = s.e_next(ee)) {
                const i16* prow = H + (size_t)((u32)s.rank_of(s.e_pred(ee)) + 1) * Wd;
                ++n;
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
            deg = n;
        }
        pred_rows += deg ? deg : 1;
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
                if (h > bv) { bv = h; bi = r + 1; bj = 1 + 32 * c + lane; }
            }
        }
        if (lane == 0) row[0] = 0;
        __syncwarp();
    }
}

// Any length: chunks of 32 columns, every row read back from the stored matrix.
template <class G>
__device__ __forceinline__ void cg_poa_dp_any(const G& s, u32 V, const u8* seq, u32 L, i32& bv, u32& bi, u32& bj, u64& pred_rows) {
    const u32 lane = cg_lane(), Wd = L + 1;
    i16* H = s.H();
    for (u32 j = lane; j < Wd; j += 32) H[j] = 0;
    __syncwarp();
    for (u32 r = 0; r < V; ++r) {
        const u32 node = s.r2n(r);
        const u8 ch = s.letter(node);
        const u32 eh = s.in_head(node);
        i16* row = H + (size_t)(r + 1) * Wd;
        if (lane == 0) row[0] = 0;
        i32 carry = 0;
        for (u32 jb = 1; jb < Wd; jb += 32) {
            const u32 j = jb + lane;
            const bool act = j < Wd;
            i32 val = CG_POA_NEG;
            if (act) {
                const i32 sc = seq[j - 1] == ch ? 5 : -10;
                if (eh == G::ENONE) {
                    val = sc > -4 ? sc : -4;                  // virtual start row of zeros
                } else {
                    for (u32 ee = eh; ee != G::ENONE; ee = s.e_next(ee)) {
                        const i16* prow = H + (size_t)((u32)s.rank_of(s.e_pred(ee)) + 1) * Wd;
                        const i32 a = (i32)prow[j - 1] + sc, b = (i32)prow[j] - 4;
                        const i32 m = a > b ? a : b;
                        val = m > val ? m : val;
                    }
                }
                val = val > 0 ? val : 0;
            }
            i32 u = act ? val + 4 * (i32)j : CG_POA_NEG;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const i32 o = __shfl_up_sync(CG_FULL, u, d);
                if (lane >= (u32)d) u = o > u ? o : u;
            }
            u = u > carry ? u : carry;
            carry = __shfl_sync(CG_FULL, u, 31);
            if (act) {
                const i32 h = u - 4 * (i32)j;
                row[j] = (i16)h;
                if (h > bv) { bv = h; bi = r + 1; bj = j; }
            }
        }
        u32 deg = 0;
        for (u32 ee = eh; ee != G::ENONE; ee = s.e_next(ee)) ++deg;
        pred_rows += deg ? deg : 1;
        __syncwarp();
    }
}

// ------------------------------------------------------------------ one job
// Returns the consensus length, or CG_NONE32 if the tier's scratch was outgrown (nothing is committed).
template <class G>
__device__ __forceinline__ u32 cg_poa_job(const CgChunk& c, const G& s, u32 w, u32 rg, u64* cnt_aln, u64* cnt_cells, u64* cnt_pred) {
    const u32 lane = cg_lane();
    const CgWin W = c.win[w];
    CgRegion* R = &c.regions[c.off_reg[w] + rg];
    CgWinView v;
    v.seq_off = c.seq_off + W.seq_begin; v.pos = c.pos + c.off_pos[w]; v.chain = c.chain + c.off_slot[w];
    v.rel = c.rel + c.off_slot[w]; v.N = W.n_seqs; v.C = W.n_cand; v.nA = W.n_chain;
    const u8* bases = (const u8*)c.bases;
    const CgPoaScratch& gs = s.gs;

    // ---- the region's segments, in read order (split_reads)
    u32 nseg = 0;
    for (u32 rb = 0; rb < v.N; rb += 32) {
        const u32 r = rb + lane;
        u32 st = 0, ln = 0;
        const bool keep = r < v.N && cg_eval_segment(v, rg, r, &st, &ln);
        const u32 bal = __ballot_sync(CG_FULL, keep);
        if (keep) {
            const u32 idx = nseg + __popc(bal & ((1u << lane) - 1u));
            if (idx < gs.ncap) { gs.seg_read[idx] = (u16)r; gs.seg_start[idx] = (u16)st; gs.seg_len[idx] = (u16)ln; }
        }
        nseg += __popc(bal);
    }
    if (nseg > gs.ncap) return CG_NONE32;
    __syncwarp();

    CgPoaState g;
    g.V = 0; g.E = 0; g.nseqs = 0; g.ovf = false;
    u64 j_aln = 0, j_cells = 0, j_pred = 0;

    for (u32 si = 0; si < nseg; ++si) {
        const u32 L = gs.seg_len[si];
        if (L == 0) continue;                                        // graph.cpp:160 — not a row of the MSA
        const u8* seq = bases + v.seq_off[gs.seg_read[si]] + gs.seg_start[si];
        if (L <= s.seqcap()) {                                       // stage the segment next to the graph
            u8* sb = s.seqbuf();
            __syncwarp();
            for (u32 i = lane; i < L; i += 32) sb[i] = seq[i];
            __syncwarp();
            seq = sb;
        }
        const u32 Wd = L + 1;
        u32 n_aln = 0;
        if (g.V != 0) {
            if ((u64)(g.V + 1) * Wd > s.hcap()) return CG_NONE32;
            i32 bv = 0; u32 bi = 0, bj = 0;
            u64 pred_rows = 0;
            if (L <= 32) cg_poa_dp<1>(s, g.V, seq, L, bv, bi, bj, pred_rows);
            else if (L <= 64) cg_poa_dp<2>(s, g.V, seq, L, bv, bi, bj, pred_rows);
            else if (L <= 128) cg_poa_dp<4>(s, g.V, seq, L, bv, bi, bj, pred_rows);
            else if (L <= 256) cg_poa_dp<8>(s, g.V, seq, L, bv, bi, bj, pred_rows);
            else cg_poa_dp_any(s, g.V, seq, L, bv, bi, bj, pred_rows);
            j_aln += 1; j_cells += (u64)(g.V + 1) * L; j_pred += pred_rows * L;
            // first cell in row-major order holding the maximum (simd...impl.hpp:828-833, 860-862)
            const u64 key = ((u64)(u32)bv << 32) | ((u64)(0xffffu - bi) << 16) | (u64)(0xffffu - bj);
            const u64 kb = cg_warp_max64(key);
            bv = (i32)(kb >> 32); bi = 0xffffu - (u32)((kb >> 16) & 0xffffu); bj = 0xffffu - (u32)(kb & 0xffffu);
            if (lane == 0 && bv > 0) n_aln = cg_poa_traceback(s, g, seq, Wd, bi, bj);
        }
        // ---- graph update + topological order (sequential)
        const u32 V0 = g.V, E0 = g.E;
        if (lane == 0 && !g.ovf) cg_poa_add_alignment(s, g, n_aln, seq, L);
        g.V = __shfl_sync(CG_FULL, g.V, 0); g.E = __shfl_sync(CG_FULL, g.E, 0);
        g.nseqs = __shfl_sync(CG_FULL, g.nseqs, 0);
        g.ovf = __shfl_sync(CG_FULL, (u32)g.ovf, 0) != 0;
        if (g.ovf) return CG_NONE32;
        if (g.V == V0 && g.E == E0) { __syncwarp(); continue; }      // same nodes, same edges, same aligned sets: same order
        for (u32 i = lane; i < g.V; i += 32) { s.marks(i) = 0; s.check(i) = 1; }
        __syncwarp();
        if (lane == 0) cg_poa_toposort(s, g);
        g.ovf = __shfl_sync(CG_FULL, (u32)g.ovf, 0) != 0;
        if (g.ovf) return CG_NONE32;
        __syncwarp();
        for (u32 i = lane; i < g.V; i += 32) s.rank_of(s.r2n(i)) = (u16)i;
        __syncwarp();
        cg_poa_build_rdesc(s, g.V);
        __syncwarp();
    }

    // ---- column vote (bmean.cpp:649-694) straight off the graph: a column = a leader and its aligned nodes
    u8* out = c.arena + c.off_arena[w] + R->arena_off;
    u32 outn = 0;
    for (u32 ib = 0; ib < g.V; ib += 32) {
        const u32 i = ib + lane;
        u8 emit = 0;
        if (i < g.V && s.leader(i)) {
            u32 cnt[4] = {0, 0, 0, 0};
            u8 row0 = 0;
            const u32 node = s.r2n(i);
            const u32 na = s.nal(node);
            for (u32 a = 0; a <= na; ++a) {
                const u32 x = a == 0 ? node : (u32)s.aligned(3 * node + a - 1);
                const u8 ch = s.letter(x);
                const u32 code = cg_base_code(ch) & 3u;
                cnt[code] = s.nseq(x);
                if (s.in0(x)) row0 = ch;
            }
            const u32 cA = cnt[0], cC = cnt[1], cG = cnt[2], cT = cnt[3];
            const u32 cM = g.nseqs - (cA + cC + cG + cT);
            if (cM > cA && cM > cC && cM > cT && cM > cG) emit = 0;
            else if (cA > cC && cA > cG && cA > cT) emit = 'A';
            else if (cC > cA && cC > cG && cC > cT) emit = 'C';
            else if (cG > cA && cG > cC && cG > cT) emit = 'G';
            else if (cT > cA && cT > cG && cT > cC) emit = 'T';
            else emit = row0;                                        // row 0's letter, if it has one here
        }
        const u32 bal = __ballot_sync(CG_FULL, emit != 0);
        if (emit) out[outn + __popc(bal & ((1u << lane) - 1u))] = emit;
        outn += __popc(bal);
    }
    *cnt_aln += j_aln; *cnt_cells += j_cells; *cnt_pred += j_pred;
    return outn;
}

// ------------------------------------------------------------------ queues
// qctl[0] = jobs appended at the front (heavy, and jobs re-queued by the previous tier), qctl[1] = next job to take,
// qctl[2] = jobs appended at the back (light), qctl[3] = capacity of the array.  Front jobs are taken first, so the
// long jobs start early and the tail of the launch is made of short ones.
__device__ __forceinline__ void cg_queue_push_front(uint2* jobs, u32* qctl, uint2 job) { jobs[atomicAdd(&qctl[0], 1u)] = job; }

template <class G>
__device__ __forceinline__ void cg_poa_drain(const CgChunk& c, const G& s, const uint2* jobs, u32* qctl, uint2* jobs_next, u32* qnext) {
    const u32 lane = cg_lane();
    const u32 nfront = qctl[0], nback = qctl[2], cap = qctl[3];
    u64 cnt_aln = 0, cnt_cells = 0, cnt_pred = 0;
    for (;;) {
        u32 j = 0;
        if (lane == 0) j = atomicAdd(&qctl[1], 1u);
        j = __shfl_sync(CG_FULL, j, 0);
        if (j >= nfront + nback) break;
        const uint2 job = j < nfront ? jobs[j] : jobs[cap - 1 - (j - nfront)];
        const u32 n = cg_poa_job(c, s, job.x, job.y, &cnt_aln, &cnt_cells, &cnt_pred);
        if (lane == 0) {
            if (n == CG_NONE32) {
                if (jobs_next) cg_queue_push_front(jobs_next, qnext, job);
                else c.win[job.x].bad = 1;            // outgrew the last tier: raw template, status CG_WINDOW_ERROR
            } else {
                c.regions[c.off_reg[job.x] + job.y].cons_len = n;
            }
        }
        __syncwarp();
    }
    if (lane == 0 && cnt_aln) {
        atomicAdd((unsigned long long*)&c.counters->alignments, (unsigned long long)cnt_aln);
        atomicAdd((unsigned long long*)&c.counters->dp_cells, (unsigned long long)cnt_cells);
        atomicAdd((unsigned long long*)&c.counters->dp_pred_cells, (unsigned long long)cnt_pred);
    }
}

// Everything in global memory (any graph size up to the tier's capacities).
__global__ void __launch_bounds__(CG_POA_THREADS) k_poa(CgChunk c, const CgPoaScratch* scratch, u32 nwarps, const uint2* jobs, u32* qctl,
                                                      uint2* jobs_next, u32* qnext) {
    const u32 gw = blockIdx.x * CG_POA_WARPS_PER_CTA + cg_warp();
    if (gw >= nwarps) return;                       // warp-uniform; no block-wide barrier in this kernel
    CgPoaGlobG s;
    s.gs = scratch[gw];
    cg_poa_drain(c, s, jobs, qctl, jobs_next, qnext);
}
