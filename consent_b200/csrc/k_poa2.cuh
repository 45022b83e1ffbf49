// k_poa2.cuh — segmented partial-order alignment, one warp per region, with almost nothing left on a single lane.
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
//  * Packed node records: in-edges are 8 (byte ids) or 16 (16-bit ids) inline ids per node (more leaves these tiers); the aligned set
//    (<= 3 others, one node per base in a column) shares a word with the in-degree.  Per row of the matrix there is
//    a descriptor (letter, in-degree, first predecessor row, node) and the 8 predecessor ROW indices, rebuilt by all
//    lanes after each change, so neither the DP nor the traceback chase pointers.
//  * Data-parallel graph update: every query position q maps to exactly one node (prefix chain, aligned part, suffix
//    chain — the id order of graph.cpp:195-201 comes from a warp scan), each node is touched once per alignment, and
//    there is an edge node(q-1) -> node(q) for every q: lanes over q, no conflicts.
//  * Traceback is the only per-alignment phase left on lane 0 (a dependent walk by nature), at two memory round
//    trips per step.
//
// Tiers (one template, T = ids x storage x capacities).  What decides a tier's speed is how many warps an SM holds
// (every warp is a chain of dependent steps), i.e. shared memory per warp, so only small matrices live there:
//   C1 : byte ids (<= 128 nodes, segments <= 32: one column per lane), graph AND matrix (<= 2048 cells) in shared memory   20 warps / SM
//   G  : byte ids (<= 254 nodes), graph in shared memory, matrix in global memory (L2)            20 warps / SM
//   W1 : 16-bit ids (<= 1024 nodes, 1024-base segments), everything in global memory (L2)         24 warps / SM
//   W2 : 16-bit ids (<= 4096 nodes, 2048-base segments, 4 M cells), global memory                 2 warps / SM
// A job that outgrows its tier (nodes, cells, in-degree, segment length) is re-queued for the next one before anything
// is committed; in-degree > 16 and anything bigger end in k_poa (k_poa.cuh), which has no such limits.
#pragma once
#include "cg_common.cuh"
#include "k_poa.cuh"

// ------------------------------------------------------------------ id packing
struct __align__(16) CgVec16 { u64 w[4]; };             // 16 ids of 16 bits

template <class IdT> struct CgIdPack;
template <> struct CgIdPack<u8> {
    typedef u64 Vec;        // 8 ids
    typedef u32 Meta;       // flags | 3 aligned ids
    typedef u32 Rdesc;      // letter | in-degree << 8 | first predecessor row << 16 | node << 24
    typedef u16 Item;       // DFS stack entry (id | finish flag), alignment pair (node | qpos << 8), splice item (pos | node << 8)
    typedef u32 Seg;        // read | start << 12 | len << 25
    static constexpr u32 W = 8, NONE = 0xffu, PMAX = 8;      // PMAX: in-edges a node can hold in these tiers
    __device__ __forceinline__ static u32 get(Vec v, u32 e) { return (u32)(v >> (8u * e)) & 0xffu; }
    __device__ __forceinline__ static Vec set(Vec v, u32 e, u32 x) { return (v & ~(0xffull << (8u * e))) | ((u64)x << (8u * e)); }
    __device__ __forceinline__ static Vec none() { return ~0ull; }
    __device__ __forceinline__ static Vec zero() { return 0ull; }
    __device__ __forceinline__ static u32 first(Vec v) { return (u32)v & 0xffu; }
    __device__ __forceinline__ static Seg seg_pack(u32 read, u32 start, u32 len) { return read | (start << 12) | (len << 25); }
    __device__ __forceinline__ static u32 seg_read(Seg s) { return s & 0xfffu; }
    __device__ __forceinline__ static u32 seg_start(Seg s) { return (s >> 12) & 0x1fffu; }
    __device__ __forceinline__ static u32 seg_len(Seg s) { return s >> 25; }
};
template <> struct CgIdPack<u16> {
    typedef CgVec16 Vec;
    typedef u64 Meta;
    typedef u64 Rdesc;      // letter | in-degree << 8 | first predecessor row << 16 | node << 32
    typedef u32 Item;
    typedef u64 Seg;        // read | start << 16 | len << 32
    static constexpr u32 W = 16, NONE = 0xffffu, PMAX = 16;
    __device__ __forceinline__ static u32 get(const Vec& v, u32 e) {
        const u64 w = (e >> 2) == 0 ? v.w[0] : (e >> 2) == 1 ? v.w[1] : (e >> 2) == 2 ? v.w[2] : v.w[3];
        return (u32)(w >> (16u * (e & 3u))) & 0xffffu;
    }
    __device__ __forceinline__ static Vec set(Vec v, u32 e, u32 x) {
        const u64 m = ~(0xffffull << (16u * (e & 3u))), b = (u64)x << (16u * (e & 3u));
#pragma unroll
        for (u32 i = 0; i < 4; ++i) if ((e >> 2) == i) v.w[i] = (v.w[i] & m) | b;
        return v;
    }
    __device__ __forceinline__ static Vec none() { Vec v; v.w[0] = v.w[1] = v.w[2] = v.w[3] = ~0ull; return v; }
    __device__ __forceinline__ static Vec zero() { Vec v; v.w[0] = v.w[1] = v.w[2] = v.w[3] = 0; return v; }
    __device__ __forceinline__ static u32 first(const Vec& v) { return (u32)v.w[0] & 0xffffu; }
    __device__ __forceinline__ static Seg seg_pack(u32 read, u32 start, u32 len) { return (u64)read | ((u64)start << 16) | ((u64)len << 32); }
    __device__ __forceinline__ static u32 seg_read(Seg s) { return (u32)s & 0xffffu; }
    __device__ __forceinline__ static u32 seg_start(Seg s) { return (u32)(s >> 16) & 0xffffu; }
    __device__ __forceinline__ static u32 seg_len(Seg s) { return (u32)(s >> 32); }
};

// ------------------------------------------------------------------ tiers
enum { CG_P2_ALL_SMEM = 0, CG_P2_H_GLOBAL = 1, CG_P2_ALL_GLOBAL = 2 };
template <class IdT_, int STORE_, u32 VCAP_, u32 HCELLS_, u32 LCAP_, u32 SEGCAP_, u32 WARPS_, u32 CTAS_> struct CgPoa2Tier {
    typedef IdT_ IdT;
    static constexpr int STORE = STORE_;
    static constexpr bool SMEM = STORE_ != CG_P2_ALL_GLOBAL;            // the graph is in shared memory
    static constexpr bool H_SMEM = STORE_ == CG_P2_ALL_SMEM;            // ... and so is the score matrix
    static constexpr bool POSTPASS_MAX = H_SMEM;                         // then the maximum cell is found after the DP, else row by row
    static constexpr u32 VCAP = VCAP_, HCELLS = HCELLS_, LCAP = LCAP_, SEGCAP = SEGCAP_, WARPS = WARPS_, CTAS_PER_SM = CTAS_;
    static constexpr u32 SCAP = 3 * VCAP_ + 8, ALNCAP = VCAP_ + LCAP_;
    static constexpr u32 SEQCAP = LCAP_ < 512u ? LCAP_ : 512u;        // segments up to this long are staged next to the graph
    // all-global ("wide") tiers: score rows are stored lane-contiguous (k_poa2 wide layout below); segments of up to 64 * CHF - 1
    // bases run the register-resident DP with the query profile in shared memory, longer ones the any-length variant
    static constexpr u32 CHF = 10;
    static constexpr u32 CHMAX = (LCAP_ + 64u) / 64u;                  // 64-column blocks of the longest segment
};
typedef CgPoa2Tier<u8, CG_P2_ALL_SMEM, 128, 2048, 32, 192, 4, 5> CgPoa2C1;
typedef CgPoa2Tier<u8, CG_P2_H_GLOBAL, 254, 30976, 120, 192, 4, 5> CgPoa2GT;
#ifndef CG_W1_CTAS
#define CG_W1_CTAS 6
#endif
typedef CgPoa2Tier<u16, CG_P2_ALL_GLOBAL, 1024, 672u << 10, 1024, 1024, 4, CG_W1_CTAS> CgPoa2W1;    // 24 warps per SM (9 KB of shared memory per warp); 1025 rows x 640 columns
typedef CgPoa2Tier<u16, CG_P2_ALL_GLOBAL, 4096, 4u << 20, 2048, 4096, 4, 1> CgPoa2W2;    // 42 KB of shared memory per warp

template <class T> struct CgPoa2Lay {
    typedef CgIdPack<typename T::IdT> Pk;
    static constexpr size_t r16(size_t v) { return (v + 15u) / 16u * 16u; }
    static constexpr size_t mx(size_t a, size_t b) { return a > b ? a : b; }
    static constexpr size_t ID = sizeof(typename T::IdT);
    static constexpr size_t WORK = r16(sizeof(typename Pk::Item) * mx(T::SCAP, T::ALNCAP));   // DFS stack | alignment pairs | splice items
    static constexpr size_t TMP0 = r16(mx(2 * (size_t)T::VCAP, (3 * ID + 1) * T::LCAP));    // DFS marks+check | update scratch
    static constexpr size_t TMP = TMP0 + (T::STORE == CG_P2_ALL_GLOBAL ? r16(2 * ((size_t)T::VCAP + 1)) : 0);   // + the maximum of every matrix row (all-global tiers)
    static constexpr size_t o_pred = 0, o_prow = o_pred + sizeof(typename Pk::Vec) * T::VCAP, o_rdesc = o_prow + sizeof(typename Pk::Vec) * T::VCAP,
                            o_meta = o_rdesc + r16(sizeof(typename Pk::Rdesc) * T::VCAP), o_seg = o_meta + r16(sizeof(typename Pk::Meta) * T::VCAP),
                            o_nseq = o_seg + r16(sizeof(typename Pk::Seg) * T::SEGCAP), o_work = o_nseq + r16(2 * (size_t)T::VCAP),
                            o_tmp = o_work + WORK, o_letter = o_tmp + TMP, o_r2n = o_letter + r16(T::VCAP),
                            R2N_STRIDE = r16(ID * T::VCAP), o_rank = o_r2n + 2 * R2N_STRIDE, o_xr2n = o_rank + R2N_STRIDE,
                            o_xlead = o_xr2n + R2N_STRIDE, o_seq = o_xlead + r16(T::VCAP), o_H = o_seq + r16((size_t)T::SEQCAP + 16),
                            h_bytes = r16(2 * (size_t)T::HCELLS + 16),
                            per_warp = T::STORE == CG_P2_H_GLOBAL ? o_H : o_H + h_bytes,         // the graph slice (+ matrix unless it is apart)
                            scratch_per_warp = T::STORE == CG_P2_ALL_SMEM ? 0 : T::STORE == CG_P2_H_GLOBAL ? h_bytes : per_warp;
    // wide tiers, shared memory per warp: the row descriptors' low words (letter | in-degree << 8 | first predecessor row << 16) and
    // the query profile (4 letters x CHF blocks x 32 lanes of packed scores)
    // (k_poa2_wide.cuh: CgWideRd).  The profile's bytes double as the DFS's marks / check flags and the top of its stack.
    static constexpr size_t RDB = T::VCAP <= 1024 ? 4 : 8, DFS_STK = 512;
    static constexpr size_t o_wrd = 0, o_wprof = RDB * (size_t)T::VCAP,
                            wide_bytes = o_wprof + mx(512 * (size_t)T::CHF, 2 * (size_t)T::VCAP + sizeof(typename Pk::Item) * DFS_STK);
    static constexpr size_t cta_bytes = T::SMEM ? per_warp * T::WARPS : wide_bytes * T::WARPS;
};

// meta word of a node: low field = nal (bits 0-1) | "sequence 0 passes here" (bit 2) | in-degree (bits 3-7); then 3 aligned ids
template <class T> struct CgPoa2G {
    typedef CgPoa2Lay<T> Lay;
    typedef CgIdPack<typename T::IdT> Pk;
    typedef typename T::IdT IdT;
    u32 wo;                     // graph in shared memory: byte offset of this warp's slice (every access an LDS/STS with an immediate offset)
    u8* base;                   // this warp's global scratch slice: everything (all-global tiers) or the matrix alone
    u32 ws;                     // wide tiers: byte offset of this warp's shared-memory slice (row descriptors, query profile)
    __device__ __forceinline__ u8* wrd() const { return cg_smem_base() + ws + Lay::o_wrd; }
    __device__ __forceinline__ u32* prof() const { return (u32*)(cg_smem_base() + ws + Lay::o_wprof); }
    __device__ __forceinline__ u8* b() const { return T::SMEM ? cg_smem_base() + wo : base; }
    __device__ __forceinline__ typename Pk::Vec& pred(u32 i) const { return ((typename Pk::Vec*)(b() + Lay::o_pred))[i]; }
    __device__ __forceinline__ typename Pk::Vec& prow(u32 i) const { return ((typename Pk::Vec*)(b() + Lay::o_prow))[i]; }
    __device__ __forceinline__ i16* H() const { return T::STORE == CG_P2_H_GLOBAL ? (i16*)base : (i16*)(b() + Lay::o_H); }
    __device__ __forceinline__ typename Pk::Rdesc& rdesc(u32 i) const { return ((typename Pk::Rdesc*)(b() + Lay::o_rdesc))[i]; }
    __device__ __forceinline__ typename Pk::Meta& meta(u32 i) const { return ((typename Pk::Meta*)(b() + Lay::o_meta))[i]; }
    __device__ __forceinline__ typename Pk::Seg& seg(u32 i) const { return ((typename Pk::Seg*)(b() + Lay::o_seg))[i]; }
    __device__ __forceinline__ u16& nseq(u32 i) const { return ((u16*)(b() + Lay::o_nseq))[i]; }
    __device__ __forceinline__ typename Pk::Item& work(u32 i) const { return ((typename Pk::Item*)(b() + Lay::o_work))[i]; }
    __device__ __forceinline__ u8& marks(u32 i) const { return T::STORE == CG_P2_ALL_GLOBAL ? ((u8*)prof())[i] : b()[Lay::o_tmp + i]; }
    __device__ __forceinline__ u8& check(u32 i) const { return T::STORE == CG_P2_ALL_GLOBAL ? ((u8*)prof())[T::VCAP + i] : b()[Lay::o_tmp + T::VCAP + i]; }
    __device__ __forceinline__ typename Pk::Item& stk(u32 i) const {         // DFS stack: its top lives in shared memory on the wide tiers
        if (T::STORE == CG_P2_ALL_GLOBAL && i < Lay::DFS_STK) return ((typename Pk::Item*)((u8*)prof() + 2 * T::VCAP))[i];
        return work(i);
    }
    __device__ __forceinline__ IdT& nodeq(u32 i) const { return ((IdT*)(b() + Lay::o_tmp))[i]; }
    __device__ __forceinline__ IdT& posq(u32 i) const { return ((IdT*)(b() + Lay::o_tmp))[T::LCAP + i]; }
    __device__ __forceinline__ IdT& anchq(u32 i) const { return ((IdT*)(b() + Lay::o_tmp))[2 * T::LCAP + i]; }
    __device__ __forceinline__ u8& kindq(u32 i) const { return b()[Lay::o_tmp + 3 * Lay::ID * T::LCAP + i]; }
    __device__ __forceinline__ i16& rowmax(u32 i) const { return ((i16*)(b() + Lay::o_tmp + Lay::TMP0))[i]; }   // all-global tiers only
    __device__ __forceinline__ u8& letter(u32 i) const { return b()[Lay::o_letter + i]; }
    __device__ __forceinline__ IdT& r2n(u32 which, u32 i) const { return ((IdT*)(b() + Lay::o_r2n + which * Lay::R2N_STRIDE))[i]; }
    __device__ __forceinline__ IdT& rank_of(u32 i) const { return ((IdT*)(b() + Lay::o_rank))[i]; }
    __device__ __forceinline__ IdT& xr2n(u32 i) const { return ((IdT*)(b() + Lay::o_xr2n))[i]; }
    __device__ __forceinline__ u8& xlead(u32 i) const { return b()[Lay::o_xlead + i]; }
    __device__ __forceinline__ u8* seqbuf() const { return b() + Lay::o_seq; }
};

__device__ __forceinline__ u32 cg_lt_mask() { return (1u << cg_lane()) - 1u; }

#define CG_P2_TYPES                                   \
    typedef CgIdPack<typename T::IdT> Pk;             \
    typedef typename T::IdT IdT;                      \
    typedef typename Pk::Meta MetaT;                  \
    typedef typename Pk::Vec VecT;                    \
    typedef typename Pk::Item ItemT;                  \
    constexpr u32 W = Pk::W, IDNONE = Pk::NONE;       \
    constexpr MetaT IDMASK = (MetaT)Pk::NONE

// ------------------------------------------------------------------ exact order: spoa's DFS (graph.cpp:294-354), lane 0
// Same walk as cg_poa_toposort (k_poa.cuh) on the packed records.  marks/check are pre-initialised (0 / 1) by the warp.
template <class T> __device__ CG_NOINLINE bool cg_poa2_dfs(const CgPoa2G<T>& s, u32 V) {
    CG_P2_TYPES;
    constexpr u32 FIN = 1u << W;
    u32 nrank = 0, sp = 0;
    for (u32 i = 0; i < V; ++i) {
        if (s.marks(i) != 0) continue;
        s.stk(sp++) = (ItemT)i;
        while (sp != 0) {
            const u32 top = s.stk(sp - 1);
            const u32 id = top & IDNONE;
            bool finish = (top & FIN) != 0;
            const MetaT m = s.meta(id);
            VecT P = Pk::zero();
            if (T::STORE == CG_P2_ALL_GLOBAL && !finish) P = s.pred(id);     // global graph: both loads of the node leave together
            const u32 nal = (u32)m & 3u;
            if (!finish) {
                if (s.marks(id) == 2) { --sp; continue; }
                const u32 sp0 = sp;
                const u32 deg = ((u32)m >> 3) & 31u;
                if (T::STORE != CG_P2_ALL_GLOBAL) P = s.pred(id);
                if (sp + deg + 3 > T::SCAP) return false;
                for (u32 e = 0; e < deg; ++e) {
                    const u32 b = Pk::get(P, e);
                    if (s.marks(b) != 2) s.stk(sp++) = (ItemT)b;
                }
                if (s.check(id)) {
                    for (u32 a = 0; a < nal; ++a) {
                        const u32 aid = (u32)((m >> (W * (a + 1))) & IDMASK);
                        if (s.marks(aid) != 2) { s.stk(sp++) = (ItemT)aid; s.check(aid) = 0; }
                    }
                }
                if (sp == sp0) finish = true;
                else { s.marks(id) = 1; s.stk(sp0 - 1) = (ItemT)(id | FIN); }
            }
            if (finish) {
                s.marks(id) = 2;
                if (s.check(id)) {
                    s.xr2n(nrank) = (IdT)id; s.xlead(nrank) = 1; ++nrank;
                    for (u32 a = 0; a < nal; ++a) { s.xr2n(nrank) = (IdT)((m >> (W * (a + 1))) & IDMASK); s.xlead(nrank) = 0; ++nrank; }
                }
                --sp;
            }
        }
    }
    return true;
}

// The same walk for the first wide tier (ids of 10 bits, graph in global memory), where chasing every node's two records through L2
// on one lane cost an eighth of the tier's time: the warp first packs what the walk needs into ONE word per node in shared memory
// (first two in-edges | in-degree << 20 | aligned nodes << 25 | mark << 27 | check << 29: 4 KB over the idle query profile, plus a
// 512-entry stack of 16-bit items); only nodes with more than two in-edges or with aligned nodes (~5 %) still read their global records.
template <class T> struct CgDfsCompact {
    static constexpr bool USE = T::STORE == CG_P2_ALL_GLOBAL && T::VCAP <= 1024;
    static constexpr u32 STK = 512;
    __device__ __forceinline__ static u32* nt(const CgPoa2G<T>& s) { return s.prof(); }
    __device__ __forceinline__ static u16* stk(const CgPoa2G<T>& s) { return (u16*)(s.prof() + T::VCAP); }
};
template <class T> __device__ __forceinline__ void cg_poa2_dfs_prepare(const CgPoa2G<T>& s, u32 V) {
    CG_P2_TYPES;
    const u32 lane = cg_lane();
    if constexpr (CgDfsCompact<T>::USE) {
        u32* nt = CgDfsCompact<T>::nt(s);
        for (u32 i = lane; i < V; i += 32) {
            const MetaT m = s.meta(i);
            const VecT P = s.pred(i);
            const u32 deg = ((u32)m >> 3) & 31u;
            nt[i] = (deg > 0 ? Pk::get(P, 0) : 0u) | ((deg > 1 ? Pk::get(P, 1) : 0u) << 10) | (deg << 20) | (((u32)m & 3u) << 25) | (1u << 29);
        }
    } else if constexpr (T::SMEM) {
        static_assert(!T::SMEM || (sizeof(typename T::IdT) == 1 && 2 * T::VCAP <= CgPoa2Lay<T>::TMP0), "halfword node table over the marks / check bytes");
        u16* nt = (u16*)(s.b() + CgPoa2Lay<T>::o_tmp);
        for (u32 i = lane; i < V; i += 32) {
            const u32 m = (u32)s.meta(i);
            const u32 deg = (m >> 3) & 31u;
            nt[i] = (u16)((deg ? Pk::first(s.pred(i)) : 0u) | ((deg < 7u ? deg : 7u) << 8) | ((m & 3u) << 11) | 0x8000u);
        }
    } else {
        for (u32 i = lane; i < V; i += 32) { s.marks(i) = 0; s.check(i) = 1; }
    }
    __syncwarp();
}
template <class T> __device__ CG_NOINLINE bool cg_poa2_dfs_compact(const CgPoa2G<T>& s, u32 V) {
    CG_P2_TYPES;
    typedef CgDfsCompact<T> C;
    constexpr u32 FIN = 1u << 10, ID = 0x3ffu, MARK = 27, CHECK = 1u << 29;
    u32* nt = C::nt(s);
    u16* st16 = C::stk(s);
#define CG_DFS_PUSH(x) do { if (sp < C::STK) st16[sp] = (u16)(x); else s.work(sp) = (ItemT)(x); ++sp; } while (0)
#define CG_DFS_AT(i) ((i) < C::STK ? (u32)st16[i] : (u32)s.work(i))
    u32 nrank = 0, sp = 0;
    for (u32 i = 0; i < V; ++i) {
        if (((nt[i] >> MARK) & 3u) != 0) continue;
        CG_DFS_PUSH(i);
        while (sp != 0) {
            const u32 top = CG_DFS_AT(sp - 1);
            const u32 id = top & ID;
            bool finish = (top & FIN) != 0;
            const u32 e = nt[id];
            const u32 deg = (e >> 20) & 31u, nal = (e >> 25) & 3u;
            if (!finish) {
                if (((e >> MARK) & 3u) == 2) { --sp; continue; }
                const u32 sp0 = sp;
                if (sp + deg + 3 > T::SCAP) return false;
                if (deg <= 2) {
                    if (deg >= 1) { const u32 b = e & ID; if (((nt[b] >> MARK) & 3u) != 2) CG_DFS_PUSH(b); }
                    if (deg == 2) { const u32 b = (e >> 10) & ID; if (((nt[b] >> MARK) & 3u) != 2) CG_DFS_PUSH(b); }
                } else {
                    const VecT P = s.pred(id);
                    for (u32 q = 0; q < deg; ++q) { const u32 b = Pk::get(P, q); if (((nt[b] >> MARK) & 3u) != 2) CG_DFS_PUSH(b); }
                }
                if ((e & CHECK) && nal) {
                    const MetaT m = s.meta(id);
                    for (u32 a = 0; a < nal; ++a) {
                        const u32 aid = (u32)((m >> (W * (a + 1))) & IDMASK);
                        if (((nt[aid] >> MARK) & 3u) != 2) { CG_DFS_PUSH(aid); nt[aid] &= ~CHECK; }
                    }
                }
                if (sp == sp0) finish = true;
                else {
                    nt[id] = (nt[id] & ~(3u << MARK)) | (1u << MARK);
                    if (sp0 - 1 < C::STK) st16[sp0 - 1] = (u16)(id | FIN); else s.work(sp0 - 1) = (ItemT)(id | FIN);
                }
            }
            if (finish) {
                const u32 e2 = nt[id];
                nt[id] = (e2 & ~(3u << MARK)) | (2u << MARK);
                if (e2 & CHECK) {
                    s.xr2n(nrank) = (IdT)id; s.xlead(nrank) = 1; ++nrank;
                    if (nal) {
                        const MetaT m = s.meta(id);
                        for (u32 a = 0; a < nal; ++a) { s.xr2n(nrank) = (IdT)((m >> (W * (a + 1))) & IDMASK); s.xlead(nrank) = 0; ++nrank; }
                    }
                }
                --sp;
            }
        }
    }
#undef CG_DFS_PUSH
#undef CG_DFS_AT
    return true;
}
// The compact tiers (byte ids, graph in shared memory) keep one HALFWORD per node over the marks / check bytes: first in-edge |
// in-degree << 8 (7 = "7 or 8") | aligned nodes << 11 | mark << 13 | check << 15.  Nine nodes in ten have at most one in-edge and no
// aligned node: their visit reads that halfword and nothing else (the walk is one lane: every instruction saved is a whole issue slot).
template <class T> __device__ CG_NOINLINE bool cg_poa2_dfs_small(const CgPoa2G<T>& s, u32 V) {
    CG_P2_TYPES;
    constexpr u32 FIN = 1u << W, MK = 13, MKM = 3u << MK, CHECK = 0x8000u;
    u16* nt = (u16*)(s.b() + CgPoa2Lay<T>::o_tmp);
    u32 nrank = 0, sp = 0;
    for (u32 i = 0; i < V; ++i) {
        if ((nt[i] & MKM) != 0) continue;
        s.stk(sp++) = (ItemT)i;
        while (sp != 0) {
            const u32 top = s.stk(sp - 1);
            const u32 id = top & IDNONE;
            bool finish = (top & FIN) != 0;
            const u32 e = nt[id];
            const u32 degc = (e >> 8) & 7u, nal = (e >> 11) & 3u;
            if (!finish) {
                if ((e & MKM) == (2u << MK)) { --sp; continue; }
                const u32 sp0 = sp;
                if (degc <= 1) {
                    if (sp + degc + 3 > T::SCAP) return false;
                    if (degc) { const u32 b = e & 0xffu; if ((nt[b] & MKM) != (2u << MK)) s.stk(sp++) = (ItemT)b; }
                } else {
                    const u32 deg = degc < 7u ? degc : (((u32)s.meta(id) >> 3) & 31u);
                    if (sp + deg + 3 > T::SCAP) return false;
                    const VecT P = s.pred(id);
                    for (u32 q = 0; q < deg; ++q) { const u32 b = Pk::get(P, q); if ((nt[b] & MKM) != (2u << MK)) s.stk(sp++) = (ItemT)b; }
                }
                if ((e & CHECK) && nal) {
                    const MetaT m = s.meta(id);
                    for (u32 a = 0; a < nal; ++a) {
                        const u32 aid = (u32)((m >> (W * (a + 1))) & IDMASK);
                        if ((nt[aid] & MKM) != (2u << MK)) { s.stk(sp++) = (ItemT)aid; nt[aid] = (u16)(nt[aid] & ~CHECK); }
                    }
                }
                if (sp == sp0) finish = true;
                else { nt[id] = (u16)((nt[id] & ~MKM) | (1u << MK)); s.stk(sp0 - 1) = (ItemT)(id | FIN); }
            }
            if (finish) {
                const u32 e2 = nt[id];
                nt[id] = (u16)((e2 & ~MKM) | (2u << MK));
                if (e2 & CHECK) {
                    s.xr2n(nrank) = (IdT)id; s.xlead(nrank) = 1; ++nrank;
                    if (nal) {
                        const MetaT m = s.meta(id);
                        for (u32 a = 0; a < nal; ++a) { s.xr2n(nrank) = (IdT)((m >> (W * (a + 1))) & IDMASK); s.xlead(nrank) = 0; ++nrank; }
                    }
                }
                --sp;
            }
        }
    }
    return true;
}

// marks / check (or the packed table) by the warp, the walk by lane 0, the verdict to every lane
template <class T> __device__ __forceinline__ bool cg_poa2_dfs_run(const CgPoa2G<T>& s, u32 V) {
    cg_poa2_dfs_prepare(s, V);
    bool ok = true;
    if (cg_lane() == 0) {
        if constexpr (CgDfsCompact<T>::USE) ok = cg_poa2_dfs_compact(s, V);
        else if constexpr (T::SMEM) ok = cg_poa2_dfs_small(s, V);
        else ok = cg_poa2_dfs(s, V);
    }
    ok = __shfl_sync(CG_FULL, (u32)ok, 0) != 0;
    __syncwarp();
    return ok;
}

// A line into L1 ahead of its use (no registers held): the next row's predecessor vector during the DP.  (Prefetching the matrix
// rows a traceback is about to reach — four rows ~7 steps ahead, or one row 8 steps ahead along the first-predecessor chain —
// measured no gain on the wide tiers and was dropped.)
#define CG_P2_PREFETCH(p) cg_prefetch_l1(p)

// ------------------------------------------------------------------ traceback (compact tiers), by the whole warp
// simd_alignment_engine_impl.hpp:968-1004: diagonal over the predecessors in in-edge order, then vertical over them,
// then horizontal.  Pairs are written in traceback order (last pair first) as node | qpos << W (all ones = none).
// Lane t looks at the cell the path reaches after t diagonal steps along the first-predecessor chain (row descriptors are in
// shared memory), all lanes fetch their cell and its neighbours at once, a ballot finds how far the guess holds: those steps are
// committed together, the first lane that disagrees decides the next cell.  A run of matches costs one round (~50 warp
// instructions) instead of ~25 single-lane instructions per step (the wide tiers' version: k_poa2_wide.cuh).
template <class T> __device__ CG_NOINLINE u32 cg_poa2_traceback(const CgPoa2G<T>& s, const u8* seq, u32 Ws, u32 bi, u32 bj, i32 M, bool* bad) {
    CG_P2_TYPES;
    enum { STOP = 0, DIAG = 1, DIAG1 = 2, VERT = 3, VERT1 = 4, HORIZ = 5, MULTI = 6 };
    const u32 lane = cg_lane();
    const i16* H = s.H();
    u32 i = bi, j = bj, n = 0;
    i32 Hij = M;
    while (Hij != 0) {
        u32 row = i >= lane ? i - lane : 0u;
        u32 d = row ? (u32)s.rdesc(row - 1) : 0u;
        u32 start = 0, nvalid = 32;
#pragma unroll 1
        for (u32 it = 0;; ++it) {
            const u32 brk = __ballot_sync(CG_FULL, row != 0 && ((d >> 16) & IDNONE) != row - 1) & (0xffffffffu << start);
            if (!brk) break;
            const u32 b = (u32)__ffs((int)brk) - 1u;
            if (it == 3 || b == 31) { nvalid = b + 1; break; }
            const u32 pb = __shfl_sync(CG_FULL, (d >> 16) & IDNONE, (int)b);
            if (lane > b) {
                const u32 back = lane - b - 1;
                row = pb >= back ? pb - back : 0u;
                d = row ? (u32)s.rdesc(row - 1) : 0u;
            }
            start = b + 1;
        }
        const u32 col = j >= lane ? j - lane : 0u;
        const bool live = row != 0 && col != 0;
        const u32 p0 = (d >> 16) & IDNONE, deg = (d >> 8) & 0xffu;
        u32 p1 = 0;
        i32 own = 0, hl = 0, hd = 0, hv = 0, hd1 = -1, hv1 = -1, sc = 0;
        if (live) {
            const i16* hr = H + (size_t)row * Ws + col;
            const i16* hp = H + (size_t)p0 * Ws + col;
            own = hr[0]; hl = hr[-1]; hd = hp[-1]; hv = hp[0];
            if (deg >= 2) { p1 = Pk::get(s.prow(row - 1), 1); const i16* hq = H + (size_t)p1 * Ws + col; hd1 = hq[-1]; hv1 = hq[0]; }
            sc = (d & 0xffu) == seq[col - 1] ? 5 : -10;
        }
        const u32 node = (d >> (16 + W)) & IDNONE;
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
        const u32 run = stopm ? (u32)__ffs((int)stopm) - 1u : 32u;
        if (n + run + 1 > T::ALNCAP) { *bad = true; return 0; }          // cannot happen: a path visits a cell once
        if (lane < run) s.work(n + lane) = (ItemT)(node | ((col - 1) << W));
        n += run;
        if (run == 32 || run == nvalid) {
            const int src = (int)run - 1;
            i = __shfl_sync(CG_FULL, p0, src); j = __shfl_sync(CG_FULL, col, src) - 1u; Hij = __shfl_sync(CG_FULL, hd, src);
            continue;
        }
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
        else {                                                            // three or more predecessors, no diagonal match among the first two
            const u32 xdeg = __shfl_sync(CG_FULL, deg, src);
            const i32 xsc = __shfl_sync(CG_FULL, sc, src);
            const VecT pr = s.prow(xrow - 1);
            u32 pe = 0;
            i32 ed = -1, ev = -1;
            if (lane < xdeg) {
                pe = Pk::get(pr, lane);
                const i16* hp = H + (size_t)pe * Ws + xcol;
                ed = hp[-1]; ev = hp[0];
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

// ------------------------------------------------------------------ score matrix (all lanes), CH chunks of 32 columns
// The common row (one predecessor = the row just computed) needs no loads: the left neighbour comes from the adjacent lane.
// The in-row gap term H[i][j] = max(v[j], H[i][j-1] - 4) is the max-plus prefix scan H[i][j] = max_{t<=j}(v[t] + 4t) - 4j.
// Lanes past the end of the query compute harmless values (a scan only moves data towards higher lanes) and store nothing,
// so the loop carries no activity masks; shfl_up hands a lane its own value back when the source is out of range, so the
// scan needs no lane tests either.  The loop keeps one running maximum per lane; the winning cell is found afterwards
// (cg_poa2_find_max), which costs cells/32 loads instead of a dozen instructions per row.
// Column 0 of every row is zeroed beforehand.  Returns the lane's maximum over its columns.
// Where the matrix lives in global memory the post-pass over it is expensive, so those tiers find the maximum as they go:
// one warp reduction per row (redux.sync) and warp-uniform bookkeeping: the maximum, the first row holding it (in this
// kernel's row order) and how many rows hold it.
struct CgPoa2Max { i32 M; u32 row, nrows; };
__device__ __forceinline__ i32 cg_poa2_track(CgPoa2Max& t, i32 lane_max, u32 row) {
    const i32 m = __reduce_max_sync(CG_FULL, lane_max);
    if (m > t.M) { t.M = m; t.row = row; t.nrows = 1; }
    else if (m == t.M) t.nrows++;
    return m;
}
// The wide tiers keep every row's maximum: when several rows tie for the matrix maximum, the first of them in spoa's order is
// found from this array instead of scanning the rows of a matrix that lives in HBM.
#define CG_P2_KEEP_ROWMAX(T, s, m, row) do { if (T::STORE == CG_P2_ALL_GLOBAL && cg_lane() == 0) (s).rowmax(row) = (i16)(m); } while (0)


template <int CH, class T>
__device__ __forceinline__ i32 cg_poa2_dp(const CgPoa2G<T>& s, u32 V, const u8* seq, u32 L, u32 Ws) {
    CG_P2_TYPES;
    const u32 lane = cg_lane(), Wd = L + 1;
    i16* H = s.H();
    u8 q[CH];
    i32 prev[CH], j4[CH];
    bool act[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const u32 j = 1 + 32 * c + lane;
        act[c] = j < Wd;
        q[c] = act[c] ? seq[j - 1] : (u8)0;
        prev[c] = 0;
        j4[c] = 4 * (i32)j;
    }
    for (u32 j = lane; j < Wd; j += 32) H[j] = 0;
    for (u32 r = lane; r < V; r += 32) H[(size_t)(r + 1) * Ws] = 0;
    __syncwarp();
    i32 bv = 0;
    u32 dnext = (u32)s.rdesc(0);
    i16* row = H + Ws + 1 + lane;                        // cell (r + 1, 1 + lane)
    for (u32 r = 0; r < V; ++r) {
        const u32 d = dnext;
        if (r + 1 < V) dnext = (u32)s.rdesc(r + 1);      // one row ahead: off the dependent chain
        const u8 ch = (u8)(d & 0xffu);
        const u32 deg = (d >> 8) & 0xffu;
        const u32 p0 = (d >> 16) & IDNONE;
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
            const i16* prow = H + (size_t)p0 * Ws + lane;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                val[c] = 0;
                if (act[c]) {
                    const i32 sc = q[c] == ch ? 5 : -10;
                    const i32 a = (i32)prow[32 * c] + sc, b = (i32)prow[32 * c + 1] - 4;
                    val[c] = a > b ? a : b;
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) val[c] = 0;
            const VecT pr = s.prow(r);
#pragma unroll 1
            for (u32 e = 0; e < deg; ++e) {
                const i16* prow = H + (size_t)Pk::get(pr, e) * Ws + lane;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    if (act[c]) {
                        const i32 sc = q[c] == ch ? 5 : -10;
                        const i32 a = (i32)prow[32 * c] + sc, b = (i32)prow[32 * c + 1] - 4;
                        const i32 m = a > b ? a : b;
                        val[c] = m > val[c] ? m : val[c];
                    }
                }
            }
        }
        // clamp at 0 (a zero start for val above is the same clamp), then the in-row gap term as a max-plus prefix scan
        i32 u[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) u[c] = (val[c] > 0 ? val[c] : 0) + j4[c];
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const i32 o = __shfl_up_sync(CG_FULL, u[c], dd);       // lanes < dd get their own value back
                u[c] = o > u[c] ? o : u[c];
            }
        }
        i32 carry = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            u[c] = u[c] > carry ? u[c] : carry;
            if (c + 1 < CH) carry = __shfl_sync(CG_FULL, u[c], 31);
            const i32 h = u[c] - j4[c];
            prev[c] = h;
            if (act[c]) {
                row[32 * c] = (i16)h;
                bv = h > bv ? h : bv;
            }
        }
        row += Ws;
        __syncwarp();
    }
    return bv;
}

// The same for longer segments with TWO columns per lane, as packed 16-bit halves of one register (VIADD.16x2 /
// VIMNMX.S16x2 / VIADDMNMX.S16x2 are native on sm_100a): lane l of chunk c owns columns 64c + 2l (low half) and
// 64c + 2l + 1 (high half), so a row of up to 64 columns costs one pass of the scan instead of two, and one 32-bit store.
// Column 0 is lane 0's low half: it is forced to 0 in every row.  Rows have an even stride, so the pairs are aligned.
__device__ __forceinline__ u32 cg_vadd2(u32 a, u32 b) { return __vadd2(a, b); }
__device__ __forceinline__ u32 cg_vmax2(u32 a, u32 b) { return __vmaxs2(a, b); }
__device__ __forceinline__ u32 cg_viaddmax2_relu(u32 a, u32 b, u32 c) { return __viaddmax_s16x2_relu(a, b, c); }
template <int CH, class T>
__device__ __forceinline__ i32 cg_poa2_dp2(const CgPoa2G<T>& s, u32 V, const u8* seq, u32 L, u32 Ws, CgPoa2Max& trk) {
    CG_P2_TYPES;
    const u32 lane = cg_lane(), Wd = L + 1;
    i16* H = s.H();
    u32 q0[CH], q1[CH];                                  // query letters of the two columns (0 = none)
    u32 prev[CH], j4[CH], nj4[CH], keep[CH], actm[CH];
    bool act[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const u32 j0 = 64 * c + 2 * lane, j1 = j0 + 1;
        act[c] = j0 < Wd;
        q0[c] = (j0 >= 1 && j0 < Wd) ? seq[j0 - 1] : 0u;
        q1[c] = j1 < Wd ? seq[j1 - 1] : 0u;
        prev[c] = 0;
        j4[c] = (4 * j0) | ((4 * j1) << 16);
        nj4[c] = ((0u - 4 * j0) & 0xffffu) | ((0u - 4 * j1) << 16);
        keep[c] = j0 == 0 ? 0xffff0000u : 0xffffffffu;   // column 0 stays 0
        actm[c] = (j0 < Wd ? 0xffffu : 0u) | (j1 < Wd ? 0xffff0000u : 0u);
    }
    for (u32 j = lane; j < Ws; j += 32) H[j] = 0;
    __syncwarp();
    u32 bv2 = 0;
    u32 dnext = (u32)s.rdesc(0);
    u32* row = (u32*)(H + Ws) + lane;                    // cells (r + 1, 2 lane) and (r + 1, 2 lane + 1)
    for (u32 r = 0; r < V; ++r) {
        const u32 d = dnext;
        if (r + 1 < V) {
            dnext = (u32)s.rdesc(r + 1);
            if (!T::SMEM) CG_P2_PREFETCH(&s.prow(r + 1));
            // (Measured and dropped: an L1 prefetch of the next row's first stored predecessor row on the G tier: 50.4 -> 52.0 ms per 16 384 windows.)
        }
        const u32 ch = d & 0xffu;
        const u32 deg = (d >> 8) & 0xffu;
        const u32 p0 = (d >> 16) & IDNONE;
        u32 val[CH];
        if (deg <= 1 && p0 == r) {                       // predecessor = the row just computed (or the zero row for r = 0)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                u32 up = __shfl_up_sync(CG_FULL, prev[c], 1);
                const u32 l31 = c > 0 ? __shfl_sync(CG_FULL, prev[c > 0 ? c - 1 : 0], 31) : 0u;
                if (lane == 0) up = l31;
                const u32 left = __funnelshift_l(up, prev[c], 16);            // (column 2l - 1, column 2l) of the row above
                const u32 sc = (q0[c] == ch ? 5u : 0xfff6u) | (q1[c] == ch ? 0x00050000u : 0xfff60000u);
                val[c] = cg_viaddmax2_relu(left, sc, cg_vadd2(prev[c], 0xfffcfffcu));
            }
        } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) val[c] = 0;
            const VecT pr = s.prow(r);
            const u32 np = deg ? deg : 1u;
#pragma unroll 1
            for (u32 e = 0; e < np; ++e) {
                const i16* prow = H + (size_t)(deg ? Pk::get(pr, e) : 0u) * Ws + 2 * lane;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    if (act[c]) {
                        const u32 sc = (q0[c] == ch ? 5u : 0xfff6u) | (q1[c] == ch ? 0x00050000u : 0xfff60000u);
                        const u32 h01 = *(const u32*)(prow + 64 * c);                 // columns 2l, 2l + 1
                        const u32 hm1 = (c == 0 && lane == 0) ? 0u : (u32)(u16)prow[64 * c - 1];
                        const u32 diag = hm1 | (h01 << 16);                           // columns 2l - 1, 2l
                        val[c] = cg_vmax2(val[c], cg_viaddmax2_relu(diag, sc, cg_vadd2(h01, 0xfffcfffcu)));
                    }
                }
            }
        }
        // in-row gap term: max-plus prefix scan over u = H + 4j, first inside the lane, then across lanes
        u32 u[CH], x[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            u[c] = cg_vadd2(val[c] & keep[c], j4[c]);
            u[c] = cg_vmax2(u[c], u[c] << 16);                              // high column also sees the low one
            x[c] = __byte_perm(u[c], 0, 0x3232);                            // the lane's maximum in both halves
        }
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = cg_vmax2(x[c], __shfl_up_sync(CG_FULL, x[c], dd));
        }
        u32 carry = 0, rowmax = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            u32 e = __shfl_up_sync(CG_FULL, x[c], 1);                       // maximum over the lanes before this one
            if (lane == 0) e = 0;
            e = cg_vmax2(e, carry);
            if (c + 1 < CH) carry = cg_vmax2(carry, __shfl_sync(CG_FULL, x[c], 31));
            const u32 h = cg_vadd2(cg_vmax2(u[c], e), nj4[c]) & keep[c];
            prev[c] = h;
            if (act[c]) row[32 * c] = h;
            if (T::H_SMEM) bv2 = cg_vmax2(bv2, h & actm[c]);
            else rowmax = c == 0 ? (h & actm[c]) : cg_vmax2(rowmax, h & actm[c]);
        }
        if (!T::H_SMEM) {
            const u32 lo = rowmax & 0xffffu, hi = rowmax >> 16;          // scores are never negative
            const i32 m = cg_poa2_track(trk, (i32)(lo > hi ? lo : hi), r + 1);
            CG_P2_KEEP_ROWMAX(T, s, m, r + 1);
        }
        row += Ws / 2;
        __syncwarp();
    }
    const i32 lo = (i32)(i16)(bv2 & 0xffffu), hi = (i32)(i16)(bv2 >> 16);
    return lo > hi ? lo : hi;
}

#include "k_poa2_wide.cuh"

// Where is the maximum M (> 0)?  Lanes over rows, each scanning its row.  Returns the number of rows that hold M; if it is
// exactly one, (*bi, *bj) is the first cell of that row equal to M (simd_alignment_engine_impl.hpp:860-862).
template <class T>
__device__ __forceinline__ u32 cg_poa2_find_max(const CgPoa2G<T>& s, u32 V, u32 Wd, u32 Ws, i32 M, u32* bi, u32* bj) {
    const u32 lane = cg_lane();
    const i16* H = s.H();
    u32 nrows = 0, frow = 0, fcol = 0;
#pragma unroll 1
    for (u32 rb = 0; rb < V; rb += 32) {
        const u32 r = rb + lane;
        u32 col = 0;
        if (r < V) {
            const i16* hr = H + (size_t)(r + 1) * Ws;
#pragma unroll 4
            for (u32 j = Wd - 1; j >= 1; --j) col = (i32)hr[j] == M ? j : col;
        }
        const u32 bal = __ballot_sync(CG_FULL, col != 0);
        if (bal && nrows == 0) {
            const int src = __ffs((int)bal) - 1;
            frow = rb + (u32)src + 1;
            fcol = __shfl_sync(CG_FULL, col, src);
        }
        nrows += __popc(bal);
    }
    *bi = frow; *bj = fcol;
    return nrows;
}

// ------------------------------------------------------------------ debug builds (-DCG_POA_TIMING): per-job phase clocks
#ifdef CG_POA_TIMING
struct CgJobTiming { u32 tier, w, rg, nseg, V, maxL; long long t0, t1, ph[8]; };   // ph: dp, max/tie, traceback, update, splice+rows, dfs, setup, vote
__device__ CgJobTiming cg_dbg_jobs[1 << 16];
__device__ u32 cg_dbg_njobs;
#define CG_T_DECL long long tph_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tlast_ = clock64(); const long long tjob0_ = tlast_
#define CG_T_MARK(i) do { const long long now_ = clock64(); tph_[i] += now_ - tlast_; tlast_ = now_; } while (0)
#define CG_T_DONE(a_, b_, c_, d_, e_, f_) do { if (cg_lane() == 0) { const u32 k_ = atomicAdd(&cg_dbg_njobs, 1u); if (k_ < (1u << 16)) { \
        CgJobTiming& J = cg_dbg_jobs[k_]; J.tier = a_; J.w = b_; J.rg = c_; J.nseg = d_; J.V = e_; J.maxL = f_; J.t0 = tjob0_; J.t1 = clock64(); \
        for (int q_ = 0; q_ < 8; ++q_) J.ph[q_] = tph_[q_]; } } } while (0)
#else
#define CG_T_DECL
#define CG_T_MARK(i)
#define CG_T_DONE(a_, b_, c_, d_, e_, f_)
#endif

// ------------------------------------------------------------------ one job
// Returns the consensus length, or CG_NONE32 if the tier was outgrown (nothing is committed).
template <class T>
__device__ __forceinline__ u32 cg_poa2_job(const CgChunk& c, const CgPoa2G<T>& s, u32 w, u32 rg, u64* cnt_aln, u64* cnt_cells, u64* cnt_pred) {
    CG_P2_TYPES;
    const u32 lane = cg_lane();
    const CgWin W_ = c.win[w];
    CgRegion* R = &c.regions[c.off_reg[w] + rg];
    if (R->n > T::SEGCAP || R->max_len > T::LCAP) return CG_NONE32;
    CgWinView v;
    v.seq_off = c.seq_off + W_.seq_begin; v.pos = c.pos + c.off_pos[w]; v.chain = c.chain + c.off_slot[w];
    v.rel = c.rel + c.off_slot[w]; v.N = W_.n_seqs; v.C = W_.n_cand; v.nA = W_.n_chain;
    const u8* bases = (const u8*)c.bases;

    // ---- the region's segments, in read order (split_reads)
    u32 nseg = 0;
#pragma unroll 1
    for (u32 rb = 0; rb < v.N; rb += 32) {
        const u32 r = rb + lane;
        u32 st = 0, ln = 0;
        const bool keep = r < v.N && cg_eval_segment(v, rg, r, &st, &ln);
        const u32 bal = __ballot_sync(CG_FULL, keep);
        if (keep) {
            const u32 idx = nseg + __popc(bal & cg_lt_mask());
            if (idx < T::SEGCAP) s.seg(idx) = Pk::seg_pack(r, st, ln);
        }
        nseg += __popc(bal);
    }
    if (nseg > T::SEGCAP) return CG_NONE32;
    __syncwarp();

    u32 V = 0, nseqs = 0, cur = 0, sumdeg = 0;
    bool dfs_valid = false;
    bool cache_ok = false;                                           // the previous sequence's alignment can be applied again (below)
    u32 cache_len = 0, cache_naln = 0;
    CG_T_DECL;
    CG_T_MARK(6);
    u64 j_aln = 0, j_cells = 0, j_pred = 0;

    for (u32 si = 0; si < nseg; ++si) {
        const typename Pk::Seg sg = s.seg(si);
        const u32 L = Pk::seg_len(sg);
        if (L == 0) continue;                                        // graph.cpp:160 — not a row of the MSA
        const u8* seq = bases + v.seq_off[Pk::seg_read(sg)] + Pk::seg_start(sg);
        // The same string as the sequence before it, whose alignment left the graph as it was (same nodes, same edges, hence the same row
        // order)?  Then the score matrix, the winning cell, the tie order and the traceback would all come out the same: the alignment
        // pairs of that sequence are still in `work` and are applied again.  In regions of up to 16 bases (half the jobs of the config-3
        // shape) two thirds of the sequences repeat an earlier one, and 45 % of their alignments are such repeats.
        bool reuse = false;
        if (L <= T::SEQCAP) {                                        // stage the segment next to the graph (where the previous one still is)
            u8* sb = s.seqbuf();
            const bool cand = T::SMEM && cache_ok && L == cache_len;
            bool same = true;
            __syncwarp();
            for (u32 i = lane; i < L; i += 32) { const u8 ch = seq[i]; if (T::SMEM) same = same && ch == sb[i]; sb[i] = ch; }
            __syncwarp();
            reuse = cand && __all_sync(CG_FULL, same);
            seq = sb;
        }
        constexpr bool WIDE = T::STORE == CG_P2_ALL_GLOBAL;    // lane-contiguous rows of 64 CH columns (k_poa2_wide.cuh)
        u32 CHw = (L + 64u) / 64u;
#ifdef CG_W1_ONE_CH
        if (WIDE) CHw = CHw <= 10 ? 10u : CHw;
#else
        if (WIDE) CHw = CHw <= 4 ? 4u : CHw <= 7 ? 7u : CHw <= 10 ? 10u : CHw;
#endif
        const u32 Wd = L + 1, Ws = WIDE ? 64u * CHw : (L + 2) & ~1u;   // columns 0..L; rows are stored with an even stride
        u32 n_aln = 0;
        if (reuse) {
            n_aln = cache_naln;
            j_aln += 1; j_cells += (u64)(V + 1) * L; j_pred += (u64)sumdeg * L;      // the work counters describe the algorithm, not this shortcut
        } else if (V != 0) {
            if ((u64)(V + 1) * Ws > (u64)T::HCELLS) return CG_NONE32;
            i32 bv = 0;
            CgPoa2Max trk;
            trk.M = 0; trk.row = 0; trk.nrows = 0;
            if constexpr (WIDE) {
                // three block counts only (rows of 256 / 448 / 640 columns): every variant is ~3 KB of code that 24 warps share
#ifndef CG_W1_ONE_CH
                if (CHw <= 4) { cg_poa2w_profile(s.prof(), seq, L, 4); cg_poa2w_dp<4>(s, V, trk); }
                else if (CHw <= 7) { cg_poa2w_profile(s.prof(), seq, L, 7); cg_poa2w_dp<7>(s, V, trk); }
                else
#endif
                if (CHw <= 10) { cg_poa2w_profile(s.prof(), seq, L, 10); cg_poa2w_dp<10>(s, V, trk); }
                else cg_poa2w_dp_any(s, V, seq, L, CHw, trk);
            }
            else if (L <= 32 && T::H_SMEM) bv = cg_poa2_dp<1>(s, V, seq, L, Ws);
            else if (L <= 63) bv = cg_poa2_dp2<1>(s, V, seq, L, Ws, trk);
            else if (T::LCAP > 63 && L <= 127) bv = cg_poa2_dp2<(T::LCAP > 63 ? 2 : 1)>(s, V, seq, L, Ws, trk);
            else if (T::LCAP > 127 && L <= 255) bv = cg_poa2_dp2<(T::LCAP > 127 ? 4 : 1)>(s, V, seq, L, Ws, trk);
            else if (T::LCAP > 255 && L <= 511) bv = cg_poa2_dp2<(T::LCAP > 255 ? 8 : 1)>(s, V, seq, L, Ws, trk);
            else bv = 0;
            CG_T_MARK(0);
            j_aln += 1; j_cells += (u64)(V + 1) * L; j_pred += (u64)sumdeg * L;
            i32 M;
            u32 bi = 0, bj = 0, nrows;
            if (T::H_SMEM) {
                M = (i32)cg_warp_max((u32)bv);
                nrows = M > 0 ? cg_poa2_find_max(s, V, Wd, Ws, M, &bi, &bj) : 0u;
            } else {
                M = trk.M; nrows = trk.nrows; bi = trk.row;
                if constexpr (WIDE) { if (M > 0 && nrows == 1) bj = cg_poa2w_rowfirst(s, bi, CHw, Wd, M); }
                else if (M > 0 && nrows == 1) {              // first cell of that row equal to M
                    const i16* hr = s.H() + (size_t)bi * Ws;
#pragma unroll 1
                    for (u32 jb = 1; jb < Wd; jb += 32) {
                        const u32 j = jb + lane;
                        const u32 bal = __ballot_sync(CG_FULL, j < Wd && (i32)hr[j] == M);
                        if (bal) { bj = jb + (u32)__ffs((int)bal) - 1; break; }
                    }
                }
            }
            if (M > 0 && nrows > 1) {
                // several rows reach the maximum: the winner is the first of them in spoa's own order (simd...impl.hpp:828-833)
                if (!dfs_valid) {
                    CG_T_MARK(1);
                    if (!cg_poa2_dfs_run(s, V)) return CG_NONE32;
                    dfs_valid = true;
                    CG_T_MARK(5);
                }
                const i16* H = s.H();
                u32 row = 0;
#pragma unroll 1
                for (u32 ib = 0; ib < V; ib += 32) {
                    const u32 i = ib + lane;
                    bool hit = false;
                    u32 myrow = 0;
                    if (i < V) {
                        myrow = (u32)s.rank_of(s.xr2n(i)) + 1;
                        if (T::STORE == CG_P2_ALL_GLOBAL) hit = (i32)s.rowmax(myrow) == M;
                        else {
                            const i16* hr = H + (size_t)myrow * Ws;
                            for (u32 j = 1; j < Wd; ++j) hit = hit || (i32)hr[j] == M;
                        }
                    }
                    const u32 bal = __ballot_sync(CG_FULL, hit);
                    if (bal) { row = __shfl_sync(CG_FULL, myrow, __ffs((int)bal) - 1); break; }
                }
                bi = row; bj = 0;
                if constexpr (WIDE) bj = cg_poa2w_rowfirst(s, row, CHw, Wd, M);
                else {
                    const i16* hr = H + (size_t)row * Ws;
#pragma unroll 1
                    for (u32 jb = 1; jb < Wd; jb += 32) {
                        const u32 j = jb + lane;
                        const u32 bal = __ballot_sync(CG_FULL, j < Wd && (i32)hr[j] == M);
                        if (bal) { bj = jb + (u32)__ffs((int)bal) - 1; break; }
                    }
                }
            }
            bool bad = false;
            CG_T_MARK(1);
            if constexpr (WIDE) {
                if (M > 0) n_aln = cg_poa2w_traceback(s, seq, CHw, bi, bj, M, &bad);     // the whole warp
                if (bad) return CG_NONE32;
            } else {
                if (M > 0) n_aln = cg_poa2_traceback(s, seq, Ws, bi, bj, M, &bad);           // the whole warp
                if (bad) return CG_NONE32;
            }
        }
        __syncwarp();
        CG_T_MARK(2);

        // ---- graph update, all lanes (graph.cpp:155-272).  Pair t in path order = work[n_aln - 1 - t].
        u32 first_valid = L, last_valid = 0;
        {
            u32 mn = 0xffffffffu, mxq = 0;
#pragma unroll 1
            for (u32 t = lane; t < n_aln; t += 32) {
                const u32 qp = (u32)s.work(t) >> W;
                if (qp != IDNONE) { mn = qp < mn ? qp : mn; mxq = qp > mxq ? qp : mxq; }
            }
#pragma unroll
            for (int dlt = 16; dlt > 0; dlt >>= 1) {
                const u32 o1 = __shfl_xor_sync(CG_FULL, mn, dlt), o2 = __shfl_xor_sync(CG_FULL, mxq, dlt);
                mn = o1 < mn ? o1 : mn; mxq = o2 > mxq ? o2 : mxq;
            }
            if (mn != 0xffffffffu) { first_valid = mn; last_valid = mxq; }
        }
        const bool has_aln = first_valid != L;
        const u32 V0 = V;
        const u32 n_prefix = first_valid, n_suffix = has_aln ? L - 1 - last_valid : 0u;
        const u32 mid_base = V0 + n_prefix + n_suffix;
        bool ovf = false;
        u32 n_mid_new = 0;
        // U1: the aligned part — which node does each consumed pair resolve to?
#pragma unroll 1
        for (u32 tb = 0; tb < n_aln; tb += 32) {
            const u32 t = tb + lane;
            u32 kind = 3, nn = 0, an = IDNONE, qp = IDNONE;
            u8 ch = 0;
            MetaT m_an = 0;
            if (t < n_aln) {
                const u32 pr = s.work(n_aln - 1 - t);
                an = pr & IDNONE; qp = pr >> W;
                if (qp != IDNONE) {
                    ch = seq[qp];
                    if (an == IDNONE) kind = 1;                                       // new node, not aligned to anything
                    else if (s.letter(an) == ch) { kind = 0; nn = an; }
                    else {
                        m_an = s.meta(an);
                        kind = 2;                                                     // new node in an's column ...
                        const u32 nal = (u32)m_an & 3u;
#pragma unroll 1
                        for (u32 a = 0; a < nal; ++a) {
                            const u32 aid = (u32)((m_an >> (W * (a + 1))) & IDMASK);
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
                const u32 nal = (u32)m_an & 3u;
                if (nal >= 3) ovf = true;                                             // cannot happen with ACGT input
                else {
                    // aligned(nn) = aligned(an) + [an]; every member of the column appends nn (graph.cpp:232-243)
                    const u32 sh = W * (nal + 1);
                    s.letter(nn) = ch; s.nseq(nn) = 0; s.pred(nn) = Pk::none();
                    s.meta(nn) = (MetaT)(nal + 1) | (m_an & ~IDMASK & (((MetaT)1 << sh) - 1)) | ((MetaT)an << sh);
#pragma unroll 1
                    for (u32 a = 0; a < nal; ++a) {
                        const u32 aid = (u32)((m_an >> (W * (a + 1))) & IDMASK);
                        const MetaT ma = s.meta(aid);
                        s.meta(aid) = (ma + 1) | ((MetaT)nn << (W * (((u32)ma & 3u) + 1)));
                    }
                    s.meta(an) = (m_an + 1) | ((MetaT)nn << sh);
                }
            }
            if (qp != IDNONE && !(isnew && nn >= T::VCAP)) {
                s.nodeq(qp) = (IdT)nn; s.kindq(qp) = (u8)kind; s.anchq(qp) = (IdT)(kind == 1 ? IDNONE : an);
            }
        }
        V = mid_base + n_mid_new;
        if (V > T::VCAP) ovf = true;
        if (__any_sync(CG_FULL, ovf)) return CG_NONE32;
        __syncwarp();
        // U2a: every query position -> its node; new unaligned nodes are initialised here
#pragma unroll 1
        for (u32 qb = 0; qb < L; qb += 32) {
            const u32 q = qb + lane;
            if (q < L) {
                u32 node, kind;
                if (q < first_valid) { node = V0 + q; kind = 1; }
                else if (q > last_valid) { node = V0 + n_prefix + (q - last_valid - 1); kind = 1; }
                else { node = s.nodeq(q); kind = s.kindq(q); }
                if (q < first_valid || q > last_valid) { s.nodeq(q) = (IdT)node; s.kindq(q) = 1; s.anchq(q) = (IdT)IDNONE; }
                if (kind == 1) { s.letter(node) = seq[q]; s.nseq(node) = 0; s.meta(node) = 0; s.pred(node) = Pk::none(); }
            }
        }
        __syncwarp();
        // U2b: one visit per position, one edge node(q-1) -> node(q) (an existing edge is reused, graph.cpp:105-110)
        bool changed = V != V0;
#pragma unroll 1
        for (u32 qb = 0; qb < L; qb += 32) {
            const u32 q = qb + lane;
            if (q < L) {
                const u32 node = s.nodeq(q);
                s.nseq(node) = (u16)(s.nseq(node) + 1);
                MetaT m = s.meta(node);
                if (nseqs == 0) m |= 4u;
                if (q > 0) {
                    const u32 src = s.nodeq(q - 1);
                    const u32 deg = ((u32)m >> 3) & 31u;
                    const VecT P = s.pred(node);
                    bool have = false;
#pragma unroll 1
                    for (u32 e = 0; e < deg; ++e) have = have || Pk::get(P, e) == src;
                    if (!have) {
                        if (deg >= Pk::PMAX) ovf = true;
                        else {
                            s.pred(node) = Pk::set(P, deg, src);
                            m += 8u;
                            changed = true;
                        }
                    }
                }
                s.meta(node) = m;
            }
        }
        nseqs++;
        if (__any_sync(CG_FULL, ovf)) return CG_NONE32;
        if (!__any_sync(CG_FULL, changed)) {                         // same nodes, same edges: same order, same rows
            cache_ok = V0 != 0 && L <= T::SEQCAP; cache_len = L; cache_naln = n_aln;
            __syncwarp(); CG_T_MARK(3); continue;
        }
        cache_ok = false;
        dfs_valid = false;
        __syncwarp();
        CG_T_MARK(3);

        // ---- the incremental order: splice the new nodes into the old order (columns stay contiguous)
        // U3a: position (in the old order) in front of which each new node goes
        u32 carry_min = IDNONE;                                       // smallest column start among the anchored positions behind
#pragma unroll 1
        for (int qb = (int)((L - 1) & ~31u); qb >= 0; qb -= 32) {
            const u32 q = (u32)qb + lane;
            u32 cs = IDNONE, ce = 0;
            u32 kind = 3;
            if (q < L) {
                kind = s.kindq(q);
                const u32 an = s.anchq(q);
                if (an != IDNONE) {                                   // column of the anchor among the OLD nodes
                    const MetaT m = s.meta(an);
                    const u32 nal = (u32)m & 3u;
                    const u32 r0 = s.rank_of(an);
                    cs = r0; ce = r0 + 1;
#pragma unroll 1
                    for (u32 a = 0; a < nal; ++a) {
                        const u32 aid = (u32)((m >> (W * (a + 1))) & IDMASK);
                        if (aid < V0) { const u32 ra = s.rank_of(aid); cs = ra < cs ? ra : cs; ce = ra + 1 > ce ? ra + 1 : ce; }
                    }
                }
            }
            // suffix minimum of cs over q (positions never decrease along the path; unanchored positions hold "none")
            u32 sm = cs;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                const u32 o = __shfl_down_sync(CG_FULL, sm, dlt);
                if (lane + (u32)dlt < 32u) sm = o < sm ? o : sm;
            }
            sm = sm < carry_min ? sm : carry_min;
            if (q < L) {
                u32 pos = IDNONE;
                if (kind == 2) pos = ce;
                else if (kind == 1) pos = sm == IDNONE ? V0 : sm;
                s.posq(q) = (IdT)pos;
            }
            carry_min = __shfl_sync(CG_FULL, sm, 0);
        }
        __syncwarp();
        // U3b: the new nodes in q order -> work[k] = pos | node << W
        u32 K = 0;
#pragma unroll 1
        for (u32 qb = 0; qb < L; qb += 32) {
            const u32 q = qb + lane;
            const bool isnew = q < L && s.kindq(q) != 0;
            const u32 bal = __ballot_sync(CG_FULL, isnew);
            if (isnew) {
                const u32 at = K + __popc(bal & cg_lt_mask()), ps = s.posq(q);
                s.work(at) = (ItemT)(ps | ((u32)s.nodeq(q) << W));
                if constexpr (T::STORE == CG_P2_ALL_GLOBAL) ((u16*)s.prof())[at] = (u16)ps;    // wide tiers: the search keys of U3c in shared memory
            }
            K += __popc(bal);
        }
        __syncwarp();
        // U3c: old entry p moves up by the number of new nodes placed at or before it; new node k lands at pos_k + k
        const u32 nxt = cur ^ 1u;
#pragma unroll 1
        for (u32 pb = 0; pb < V0; pb += 32) {
            const u32 p = pb + lane;
            if (p < V0) {
                u32 lo = 0, hi = K;                                   // first k with pos_k > p
#pragma unroll 1
                while (lo < hi) {
                    const u32 mid = (lo + hi) >> 1;
                    const u32 key = T::STORE == CG_P2_ALL_GLOBAL ? (u32)((const u16*)s.prof())[mid] : ((u32)s.work(mid) & IDNONE);
                    if (key <= p) lo = mid + 1; else hi = mid;
                }
                const u32 node = s.r2n(cur, p);
                s.r2n(nxt, p + lo) = (IdT)node;
                s.rank_of(node) = (IdT)(p + lo);
            }
        }
#pragma unroll 1
        for (u32 kb2 = 0; kb2 < K; kb2 += 32) {
            const u32 k = kb2 + lane;
            if (k < K) {
                const u32 it = s.work(k);
                const u32 node = it >> W, at = (it & IDNONE) + k;
                s.r2n(nxt, at) = (IdT)node;
                s.rank_of(node) = (IdT)at;
            }
        }
        cur = nxt;
        __syncwarp();
        // U4: row descriptors and the predecessor rows
        u32 sd = 0;
#pragma unroll 1
        for (u32 rb = 0; rb < V; rb += 32) {
            const u32 r = rb + lane;
            if (r < V) {
                const u32 node = s.r2n(cur, r);
                const u32 deg = ((u32)s.meta(node) >> 3) & 31u;
                const VecT P = s.pred(node);
                VecT rows = Pk::zero();
#pragma unroll 1
                for (u32 e = 0; e < deg; ++e) rows = Pk::set(rows, e, (u32)s.rank_of(Pk::get(P, e)) + 1u);
                s.prow(r) = rows;
                const u32 rd_lo = (u32)s.letter(node) | (deg << 8) | (Pk::first(rows) << 16);
                s.rdesc(r) = (typename Pk::Rdesc)rd_lo | ((typename Pk::Rdesc)node << (16 + W));
                if constexpr (T::STORE == CG_P2_ALL_GLOBAL)
                    CgWideRd<T>::at(s)[r] = CgWideRd<T>::pack((u32)s.letter(node), deg, Pk::first(rows), deg > 1 ? Pk::get(rows, 1) : 0u);
                sd += deg ? deg : 1u;
            }
        }
        sumdeg = cg_warp_sum(sd);
        __syncwarp();
        CG_T_MARK(4);
    }

    // ---- exact column order for the vote
    if (!dfs_valid && V != 0) {
        if (!cg_poa2_dfs_run(s, V)) return CG_NONE32;
    }

    CG_T_MARK(5);
    // ---- column vote (bmean.cpp:649-694) straight off the graph: a column = a leader and its aligned nodes
    u8* out = c.arena + c.off_arena[w] + R->arena_off;
    u32 outn = 0;
#pragma unroll 1
    for (u32 ib = 0; ib < V; ib += 32) {
        const u32 i = ib + lane;
        u8 emit = 0;
        if (i < V && s.xlead(i)) {
            u32 cnt[4] = {0, 0, 0, 0};
            u8 row0 = 0;
            const u32 node = s.xr2n(i);
            const MetaT m = s.meta(node);
            const u32 na = (u32)m & 3u;
            for (u32 a = 0; a <= na; ++a) {
                const u32 x = a == 0 ? node : (u32)((m >> (W * a)) & IDMASK);
                const u8 ch = s.letter(x);
                const u32 code = cg_base_code(ch) & 3u;
                cnt[code] = s.nseq(x);
                if ((u32)s.meta(x) & 4u) row0 = ch;
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
    CG_T_MARK(7);
    CG_T_DONE((u32)T::VCAP, w, rg, nseg, V, R->max_len);
    return outn;
}

// Persistent warps over the tier's queue (same protocol as cg_poa_drain).  scratch: per-warp slices of
// CgPoa2Lay<T>::scratch_per_warp bytes (nothing for the all-shared-memory tier).
template <class T>
__global__ void __launch_bounds__(T::WARPS * 32, T::CTAS_PER_SM) k_poa2(CgChunk c, u8* scratch, u32 nwarps, const uint2* jobs, u32* qctl,
                                                                      uint2* jobs_next, u32* qnext) {
    const u32 gw = blockIdx.x * T::WARPS + cg_warp();
    if (gw >= nwarps) return;                       // warp-uniform; no block-wide barrier in this kernel
    CgPoa2G<T> s;
    s.wo = (u32)CgPoa2Lay<T>::per_warp * cg_warp();
    s.ws = (u32)CgPoa2Lay<T>::wide_bytes * cg_warp();
    s.base = T::STORE == CG_P2_ALL_SMEM ? nullptr : scratch + CgPoa2Lay<T>::scratch_per_warp * (size_t)gw;
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
        constexpr u32 tier = T::VCAP <= 128 ? 0u : T::VCAP <= 254 ? 1u : T::VCAP <= 1024 ? 2u : 3u;
        atomicAdd((unsigned long long*)&c.counters->tier_cells[tier], (unsigned long long)cnt_cells);
        atomicAdd((unsigned long long*)&c.counters->tier_pred[tier], (unsigned long long)cnt_pred);
    }
}
