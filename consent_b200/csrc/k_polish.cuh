// k_polish.cuh — stitch the region consensuses, weight by k-mer solidity, polish with the local de Bruijn graph.
//
// Replaces global_consensus (BMEAN/bmean.cpp:702-733), the tail of computeConsensusReadCorrection
// (src/correctionMSA.cpp:33-48), weightConsensus (:6-27), polishCorrection / getNextSrc / getNextDst / getAnchors
// (src/correctionDBG.cpp:13-204) and getNeighbours / extendLeft / extendRight / link (src/DBG.cpp:18-169).
//
// The "graph" is the window's solid k-mer list (sorted by k-mer, produced by k_index): membership and counts by
// binary search, `visited` (a std::set<string> that lives for the whole window in the reference,
// correctionDBG.cpp:94) as one bit per list entry.  The reference's string surgery is done in place:
//   head : the weak prefix is cut by moving the start of the string, extendLeft writes into the freed bytes,
//          what was not covered is still there                        (correctionDBG.cpp:116-130)
//   tail : symmetric with extendRight                                 (:189-202)
//   link : the recursion of DBG.cpp:99-169 becomes an explicit stack of frames over ONE path buffer (every
//          recursive call extends its caller's string by one base); depth <= maxBranches + 2.
// One warp per window: all lanes stitch and weight, lane 0 walks the graph (inherently sequential, and a few
// percent of the path).
#pragma once
#include "cg_common.cuh"

#define CG_POLISH_WARPS_PER_CTA 4u
#define CG_POLISH_THREADS (CG_POLISH_WARPS_PER_CTA * 32u)
#define CG_MAX_BRANCHES 50u

__device__ __forceinline__ bool cg_is_upper(u8 ch) { return ch >= 'A' && ch <= 'Z'; }
__device__ __forceinline__ u8 cg_to_upper(u8 ch) { return (ch >= 'a' && ch <= 'z') ? (u8)(ch - 32) : ch; }
__device__ __forceinline__ u8 cg_to_lower(u8 ch) { return (ch >= 'A' && ch <= 'Z') ? (u8)(ch + 32) : ch; }
// str2num of an upper-cased character (BMEAN/utils.cpp:18-30): A0 C1 G2, anything else 3.  Every byte this kernel looks at is one of
// ACGTacgt (the input is validated, k_prep.cuh: k_pack; the consensus is built from input bases; weightConsensus only changes case), and
// for those eight ((ch >> 1) & 3) ^ its own high bit is exactly that code.
__device__ __forceinline__ u32 cg_char_code(u8 ch) { const u32 x = ((u32)ch >> 1) & 3u; return x ^ (x >> 1); }
__device__ __forceinline__ u32 cg_code_of(const u8* p, u32 k) {
    u32 r = 0;
#pragma unroll 1
    for (u32 i = 0; i < k; ++i) r = (r << 2) | cg_char_code(p[i]);
    return r;
}

struct CgDbg {
    const u32* sk; const u32* sc; u32 ns;
    u32 k, mask;
    u32* visited;
    const u32* bstart; u32 shift;        // bstart[b] = first list entry whose k-mer's top 10 bits are >= b (1025 entries, shared memory)
};
#define CG_DBG_BUCKETS 1024u
#define CG_DBG_BUCKET_BITS 10u
// The bucket index of a window's solid list, by the whole warp: one coalesced pass over the list marks where the top bits change, a
// suffix minimum fills the empty buckets.  A lookup then bisects ~ns / 1024 entries instead of ns (the walk is a chain of dependent lookups).
__device__ __forceinline__ void cg_dbg_build_index(const u32* sk, u32 ns, u32 shift, u32* bstart) {
    const u32 lane = cg_lane();
    for (u32 b = lane; b <= CG_DBG_BUCKETS; b += 32) bstart[b] = ns;
    __syncwarp();
    for (u32 ib = 0; ib < ns; ib += 32) {
        const u32 i = ib + lane;
        if (i < ns) {
            const u32 b = sk[i] >> shift;
            if (i == 0 || (sk[i - 1] >> shift) != b) bstart[b] = i;
        }
    }
    __syncwarp();
    // suffix minimum: lane l owns the buckets [PER l, PER (l + 1))
    constexpr u32 PER = CG_DBG_BUCKETS / 32u;
    u32 m = ns;
    for (int q = (int)PER - 1; q >= 0; --q) { const u32 x = bstart[PER * lane + q]; m = x < m ? x : m; }
    u32 sm = m;                                           // minimum over this lane's buckets and every later lane's
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) { const u32 o = __shfl_down_sync(CG_FULL, sm, dlt); if (lane + (u32)dlt < 32u) sm = o < sm ? o : sm; }
    u32 run = __shfl_down_sync(CG_FULL, sm, 1);          // ... over the later lanes only
    if (lane == 31) run = ns;
    for (int q = (int)PER - 1; q >= 0; --q) { const u32 x = bstart[PER * lane + q]; run = x < run ? x : run; bstart[PER * lane + q] = run; }
    __syncwarp();
}
// first entry >= code
__device__ __forceinline__ u32 cg_dbg_lower(const CgDbg& d, u32 code) {
    const u32 b = code >> d.shift;
    u32 lo = d.bstart[b], hi = d.bstart[b + 1];
    while (lo < hi) { const u32 m = (lo + hi) >> 1; if (d.sk[m] < code) lo = m + 1; else hi = m; }
    return lo;
}
__device__ __forceinline__ i32 cg_dbg_find(const CgDbg& d, u32 code) {
    const u32 lo = cg_dbg_lower(d, code);
    return (lo < d.ns && d.sk[lo] == code) ? (i32)lo : -1;
}
struct CgNb { u32 code[4]; u32 cnt[4]; u32 idx[4]; u32 n; };
// getNeighbours (DBG.cpp:18-54): successors generated in A,C,G,T order; predecessors (left) through the reverse
// complement trick, i.e. with first letter T,G,C,A; then sorted by count, descending — std::sort on <= 4 elements
// is an insertion sort, i.e. stable.  The four successors are consecutive keys: one lookup, then the entries that follow it.
__device__ CG_NOINLINE void cg_dbg_neighbours(const CgDbg& d, u32 code, bool left, CgNb& nb) {
    nb.n = 0;
    if (!left) {
        const u32 base4 = (code << 2) & d.mask;
        const u32 lo = cg_dbg_lower(d, base4);
        u32 key[4];
#pragma unroll
        for (u32 j = 0; j < 4; ++j) key[j] = lo + j < d.ns ? d.sk[lo + j] : 0xffffffffu;
#pragma unroll
        for (u32 j = 0; j < 4; ++j)
            if (key[j] - base4 < 4u) { nb.code[nb.n] = key[j]; nb.cnt[nb.n] = d.sc[lo + j]; nb.idx[nb.n] = lo + j; nb.n++; }
    } else {
        for (u32 i = 0; i < 4; ++i) {
            const u32 cand = ((3u - i) << (2 * (d.k - 1))) | (code >> 2);
            const i32 ix = cg_dbg_find(d, cand);
            if (ix >= 0) { nb.code[nb.n] = cand; nb.cnt[nb.n] = d.sc[ix]; nb.idx[nb.n] = (u32)ix; nb.n++; }
        }
    }
    for (u32 i = 1; i < nb.n; ++i) {
        const u32 cc = nb.code[i], ct = nb.cnt[i], ci = nb.idx[i];
        u32 j = i;
        while (j > 0 && ct > nb.cnt[j - 1]) { nb.code[j] = nb.code[j - 1]; nb.cnt[j] = nb.cnt[j - 1]; nb.idx[j] = nb.idx[j - 1]; --j; }
        nb.code[j] = cc; nb.cnt[j] = ct; nb.idx[j] = ci;
    }
}
__device__ __forceinline__ u8 cg_base_char(u32 code2) { return code2 == 0 ? 'A' : code2 == 1 ? 'C' : code2 == 2 ? 'G' : 'T'; }

struct CgLinkFrame { u32 len, dist; u32 code[4]; u32 idx[4]; u32 n, it; };

// link (DBG.cpp:99-169).  P[0..*plen) holds the source k-mer on entry and the path on success.
__device__ CG_NOINLINE bool cg_dbg_link(const CgDbg& d, u32 tgt, u32* curBranches, u8* P, u32* plen, u32 pcap, u32 LRLen, bool* ovf) {
    CgLinkFrame fr[CG_MAX_BRANCHES + 6];
    u32 depth = 0;
    u32 len = *plen, dist = 0;
    for (;;) {
        // ---- function entry
        bool failed = false, found = false;
        CgNb nb; nb.n = 0;
        u32 it = 0;
        if (*curBranches > CG_MAX_BRANCHES || dist > LRLen) failed = true;
        if (!failed) {
            const u32 src = cg_code_of(P + len - d.k, d.k);
            found = src == tgt;
            cg_dbg_neighbours(d, src, false, nb);
            while (!found && nb.n == 1 && it != nb.n && dist <= LRLen) {          // DBG.cpp:119-138
                const u32 cur = nb.code[it], ix = nb.idx[it];
                const bool seen = (d.visited[ix >> 5] >> (ix & 31u)) & 1u;
                found = cur == tgt;
                if (!found && !seen) {
                    d.visited[ix >> 5] |= 1u << (ix & 31u);
                    if (len + 1 > pcap) { *ovf = true; return false; }
                    P[len++] = cg_base_char(cur & 3u);
                    dist += 1;
                    cg_dbg_neighbours(d, cur, false, nb);
                    it = 0;
                } else if (found) {
                    if (len + 1 > pcap) { *ovf = true; return false; }
                    P[len++] = cg_base_char(cur & 3u);
                } else ++it;
            }
        }
        // ---- branching loop (DBG.cpp:141-160), re-entered after a failed recursive call
        for (;;) {
            bool descend = false;
            if (!failed) {
                while (!found && nb.n > 1 && it != nb.n && dist <= LRLen) {
                    const u32 cur = nb.code[it], ix = nb.idx[it];
                    const bool seen = (d.visited[ix >> 5] >> (ix & 31u)) & 1u;
                    found = cur == tgt;
                    if (!found && !seen) {
                        d.visited[ix >> 5] |= 1u << (ix & 31u);
                        (*curBranches)++;
                        if (depth + 1 >= CG_MAX_BRANCHES + 6 || len + 1 > pcap) { *ovf = true; return false; }
                        CgLinkFrame& f = fr[depth++];
                        f.len = len; f.dist = dist; f.n = nb.n; f.it = it;
                        for (u32 q = 0; q < 4; ++q) { f.code[q] = nb.code[q]; f.idx[q] = nb.idx[q]; }
                        P[len++] = cg_base_char(cur & 3u);
                        dist += 1;
                        descend = true;
                        found = false;
                        break;
                    } else if (found) {
                        if (len + 1 > pcap) { *ovf = true; return false; }
                        P[len++] = cg_base_char(cur & 3u);
                    } else ++it;
                }
            }
            if (descend) break;                       // recursive call: back to "function entry"
            if (!failed && found) { *plen = len; return true; }      // every caller returns 1 at once
            // this call returns 0
            if (depth == 0) return false;
            const CgLinkFrame& f = fr[--depth];
            len = f.len; dist = f.dist; nb.n = f.n; it = f.it + 1;
            for (u32 q = 0; q < 4; ++q) { nb.code[q] = f.code[q]; nb.idx[q] = f.idx[q]; }
            failed = false; found = false;
        }
    }
}

// getNextSrc / getNextDst (correctionDBG.cpp:13-43)
__device__ __forceinline__ int cg_next_src(const u8* r, u32 n, u32 beg, u32 m) {
    u32 nb = 0, i = beg;
    while (i < n && (cg_is_upper(r[i]) || nb < m)) { if (cg_is_upper(r[i])) nb++; else nb = 0; i++; }
    return nb >= m ? (int)i - 1 : -1;
}
__device__ __forceinline__ int cg_next_dst(const u8* r, u32 n, u32 beg, u32 m) {
    u32 nb = 0, i = beg;
    while (i < n && nb < m) { if (cg_is_upper(r[i])) nb++; else nb = 0; i++; }
    return nb >= m ? (int)i - 1 : -1;
}

struct CgAnchorPair { u32 spos, dpos; int occ; };
// getAnchors (correctionDBG.cpp:47-91): pairs of k-mers unique inside their zone, stable-sorted by
// count(src)+count(dst) descending, first `nbmax` kept.
__device__ CG_NOINLINE u32 cg_get_anchors(const CgDbg& d, const u8* srcZone, const u8* dstZone, u32 zlen, u32 nbmax, CgAnchorPair* out) {
    const u32 k = d.k, nk = zlen - k + 1;
    CgAnchorPair res[16];
    u32 scode[4], dcode[4];
    for (u32 i = 0; i < nk; ++i) { scode[i] = cg_code_of(srcZone + i, k); dcode[i] = cg_code_of(dstZone + i, k); }
    u32 n = 0;
    for (u32 i = 0; i < nk; ++i) {
        u32 os = 0;
        for (u32 j = 0; j < nk; ++j) os += scode[i] == scode[j];
        if (os != 1) continue;
        for (u32 q = 0; q < nk; ++q) {
            u32 od = 0;
            for (u32 j = 0; j < nk; ++j) od += dcode[q] == dcode[j];
            if (od != 1) continue;
            const i32 is = cg_dbg_find(d, scode[i]), id = cg_dbg_find(d, dcode[q]);
            res[n].spos = i; res[n].dpos = q;
            res[n].occ = (int)((is >= 0 ? d.sc[is] : 0u) + (id >= 0 ? d.sc[id] : 0u));
            ++n;
        }
    }
    for (u32 i = 1; i < n; ++i) {
        const CgAnchorPair t = res[i];
        u32 j = i;
        while (j > 0 && t.occ > res[j - 1].occ) { res[j] = res[j - 1]; --j; }
        res[j] = t;
    }
    const u32 m = n < nbmax ? n : nbmax;
    for (u32 i = 0; i < m; ++i) out[i] = res[i];
    return m;
}

// polishCorrection (correctionDBG.cpp:93-204).  buf[0..cap) holds the weighted consensus at [0,n); returns the
// start offset and length of the polished string inside buf.
__device__ CG_NOINLINE void cg_polish(const CgDbg& d, u8* buf, u32 cap, u32 n_in, u8* P, u32 pcap, u32* out_beg, u32* out_n, bool* ovf) {
    const u32 k = d.k, zone = 3, zlen = k + zone;
    u32 beg = 0, n = n_in;                     // the string is buf[beg, beg+n)
    u8* cr = buf;
    u32 tmpSrcBeg = 0, tmpSrcEnd = 0, tmpDstBeg = 0, tmpDstEnd = 0;

    u32 i = 0;
    while (i < n && !cg_is_upper(cr[i])) i++;
    if (i > 0 && i < n && n - i >= k) {                                           // :116-130
        const u32 extLen = i;
        beg = i; n -= i;
        CgNb nb;
        u32 dist = 0;
        u32 code = cg_code_of(buf + beg, k);
        cg_dbg_neighbours(d, code, true, nb);
        while (nb.n == 1 && dist < extLen) {                                      // extendLeft, DBG.cpp:56-75
            code = nb.code[0];
            buf[--beg] = cg_base_char(code >> (2 * (k - 1)));
            ++n; ++dist;
            cg_dbg_neighbours(d, code, true, nb);
        }
        // what the extension did not cover is still in front of it
        n += beg; beg = 0;
        i = dist;
    }
    while (i < n) {                                                               // :133-187
        const int srcEnd = cg_next_src(cr, n, i, zlen);
        const int dstEnd = cg_next_dst(cr, n, (u32)(srcEnd + 1), zlen);
        const int srcBeg = srcEnd - (int)k - (int)zone + 1;
        const int dstBeg = dstEnd - (int)k - (int)zone + 1;
        if (srcEnd != -1 && dstEnd != -1) {
            u32 plen = 0;
            CgAnchorPair anchors[5];
            const u32 nAnch = cg_get_anchors(d, cr + srcBeg, cr + dstBeg, zlen, 5, anchors);
            u32 anchorNb = 0;
            while (anchorNb < nAnch && plen == 0) {
                const u32 sp = anchors[anchorNb].spos, dp = anchors[anchorNb].dpos;
                tmpSrcBeg = (u32)srcBeg + sp; tmpSrcEnd = tmpSrcBeg + k - 1;
                tmpDstBeg = (u32)dstBeg + dp; tmpDstEnd = tmpDstBeg + k - 1;
                const u32 scode = cg_code_of(cr + tmpSrcBeg, k), dcode = cg_code_of(cr + tmpDstBeg, k);
                if (scode != dcode) {
                    u32 curBranches = 0;
                    const u32 gap = tmpDstBeg - tmpSrcEnd - 1;
                    // 15.0 / 100.0 * 2.0 * gap + gap + merSize, evaluated left to right in fp64, no FMA  (:163)
                    const double t2 = __dadd_rn(__dadd_rn(__dmul_rn(15.0 / 100.0 * 2.0, (double)gap), (double)gap), (double)k);
                    const u32 maxSize = (u32)t2;
                    for (u32 q = 0; q < k; ++q) P[q] = cg_to_upper(cr[tmpSrcBeg + q]);
                    u32 pl = k;
                    const bool ok = cg_dbg_link(d, dcode, &curBranches, P, &pl, pcap, maxSize, ovf);
                    if (*ovf) { *out_beg = 0; *out_n = n; return; }
                    plen = ok ? pl : 0;
                }
                anchorNb++;
            }
            if (plen != 0) {                                                      // :169-183
                const u32 l = tmpDstEnd - tmpSrcBeg + 1;
                long b = -1;
                for (u32 x = 0; x + l <= n; ++x) {                                // std::string::find of the old stretch
                    u32 t = 0;
                    while (t < l && cr[x + t] == cr[tmpSrcBeg + t]) ++t;
                    if (t == l) { b = (long)x; break; }
                }
                if (b != -1) {
                    const u32 tail = n - (u32)b - l;
                    if ((u64)b + plen + tail > cap) { *ovf = true; *out_beg = 0; *out_n = n; return; }
                    if (plen > l) for (u32 t = tail; t-- > 0;) cr[(u32)b + plen + t] = cr[(u32)b + l + t];
                    else if (plen < l) for (u32 t = 0; t < tail; ++t) cr[(u32)b + plen + t] = cr[(u32)b + l + t];
                    for (u32 t = 0; t < plen; ++t) cr[(u32)b + t] = P[t];
                    n = (u32)b + plen + tail;
                    i = (u32)b;
                } else i = tmpDstBeg > i ? tmpDstBeg : (u32)dstBeg;
            } else i = tmpDstBeg > i ? tmpDstBeg : (u32)dstBeg;
        } else i = n;
    }
    i = n - 1;                                                                    // :189-202
    while (i > 0 && !cg_is_upper(cr[i])) i--;
    if (i > 0 && i < n - 1 && i + 1 >= k) {
        const u32 extLen = n - 1 - i, oldn = n;
        n = i + 1;
        CgNb nb;
        u32 dist = 0;
        u32 code = cg_code_of(cr + n - k, k);
        cg_dbg_neighbours(d, code, false, nb);
        while (nb.n != 0 && dist < extLen) {                                      // extendRight, DBG.cpp:77-96
            code = nb.code[0];
            cr[n++] = cg_base_char(code & 3u);
            ++dist;
            cg_dbg_neighbours(d, code, false, nb);
        }
        if (dist < extLen) n = oldn;           // the uncovered rest of the old tail is still behind it
    }
    *out_beg = beg; *out_n = n;
}

// One thread per window: stitched length (sum of the region consensuses) and the size of its work slices.
__global__ void k_stitch_len(CgChunk c, u64* off_fin) {
    const u32 w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= c.nwin) return;
    const CgWin W = c.win[w];
    const CgRegion* regs = c.regions + c.off_reg[w];
    u32 len = 0;
    if (!W.bad)
        for (u32 g = 0; g < W.n_regions; ++g) {
            const CgRegion R = regs[g];
            len += R.kind == CG_REG_COPY ? R.len : R.kind == CG_REG_POA ? R.cons_len : 0u;
        }
    c.win[w].stitched_len = len;
    const u32 m = len > W.tlen ? len : W.tlen;
    off_fin[w] = ((u64)2 * m + 64 + 15) / 16 * 16 * 2;          // two slices: consensus, path
}

#define CG_POLISH_SMEM_BYTES (CG_POLISH_WARPS_PER_CTA * (CG_DBG_BUCKETS + 8u) * 4u)
__global__ void __launch_bounds__(CG_POLISH_THREADS) k_polish(CgChunk c, const u64* off_fin) {
    CG_DYN_SMEM(smem);
    const u32 lane = cg_lane();
    const u32 w = blockIdx.x * CG_POLISH_WARPS_PER_CTA + cg_warp();
    if (w >= c.nwin) return;                        // warp-uniform
    const CgWin W = c.win[w];
    const u8* bases = (const u8*)c.bases;
    const u64* soff = c.seq_off + W.seq_begin;
    const u32 cap = (u32)((off_fin[w + 1] - off_fin[w]) / 2);
    u8* buf = c.fin + off_fin[w];
    u8* P = buf + cap;
    const u32 n = W.stitched_len, k = c.k;
    if (n == 0) {                                   // "could not build a consensus": the raw template (correctionMSA.cpp:34-36)
        const u8* t = bases + soff[0];              // (also what a window over a limit of this build comes back as: status 2)
        for (u32 i = lane; i < W.tlen; i += 32) buf[i] = t[i];
        if (lane == 0) {
            c.win[w].final_len = W.tlen; c.win[w].final_beg = 0; c.win[w].status = W.bad ? 2u : 1u;
            if (W.bad) atomicAdd((unsigned long long*)&c.counters->error_windows, 1ull);
            else atomicAdd((unsigned long long*)&c.counters->fallback_windows, 1ull);
            atomicAdd((unsigned long long*)&c.counters->consensus_bytes, (unsigned long long)W.tlen);
        }
        return;
    }
    // ---- stitch (global_consensus)
    const CgRegion* regs = c.regions + c.off_reg[w];
    const u8* arena = c.arena + c.off_arena[w];
    u32 at = 0;
    for (u32 g = 0; g < W.n_regions; ++g) {
        const CgRegion R = regs[g];
        const u8* src; u32 l;
        if (R.kind == CG_REG_COPY) { src = bases + soff[R.read] + R.start; l = R.len; }
        else if (R.kind == CG_REG_POA) { src = arena + R.arena_off; l = R.cons_len; }
        else continue;
        for (u32 i = lane; i < l; i += 32) buf[at + i] = src[i];
        at += l;
    }
    __syncwarp();
    u32 fbeg = 0, fn = n;
    if (n >= k) {
        // ---- weightConsensus: case of base p = solidity of the k-mer starting at min(p, n-k)
        CgDbg d;
        d.sk = c.solid_k + c.off_solid[w]; d.sc = c.solid_c + c.off_solid[w]; d.ns = W.n_solid;
        d.k = k; d.mask = (k >= 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
        d.visited = c.visited + (c.off_solid[w] / 32 + w);
        u32* bstart = (u32*)smem + cg_warp() * (CG_DBG_BUCKETS + 8u);
        d.shift = 2 * k > CG_DBG_BUCKET_BITS ? 2 * k - CG_DBG_BUCKET_BITS : 0u;
        cg_dbg_build_index(d.sk, d.ns, d.shift, bstart);
        d.bstart = bstart;
        const u32 last = n - k;                     // positions beyond the last k-mer follow it
        const u32 solid_last = cg_dbg_find(d, cg_code_of(buf + last, k)) >= 0 ? 1u : 0u;
        __syncwarp();
        for (u32 p = lane; p < n; p += 32) {
            const u32 sdl = p <= last ? (cg_dbg_find(d, cg_code_of(buf + p, k)) >= 0 ? 1u : 0u) : solid_last;
            buf[p] = sdl ? cg_to_upper(buf[p]) : cg_to_lower(buf[p]);
        }
        __syncwarp();
        // ---- polishCorrection (lane 0)
        const u32 vwords = (W.n_solid + 31) / 32;
        for (u32 i = lane; i < vwords; i += 32) d.visited[i] = 0;
        __syncwarp();
        if (lane == 0) {
            bool ovf = false;
            cg_polish(d, buf, cap, n, P, cap, &fbeg, &fn, &ovf);
            if (ovf) c.win[w].bad = 1;
        }
        fbeg = __shfl_sync(CG_FULL, fbeg, 0);
        fn = __shfl_sync(CG_FULL, fn, 0);
        if (__shfl_sync(CG_FULL, (u32)c.win[w].bad, 0)) {       // the polish outgrew its work slice: raw template, status CG_WINDOW_ERROR
            const u8* t = bases + soff[0];
            __syncwarp();
            for (u32 i = lane; i < W.tlen; i += 32) buf[i] = t[i];
            if (lane == 0) {
                c.win[w].final_len = W.tlen; c.win[w].final_beg = 0; c.win[w].status = 2;
                atomicAdd((unsigned long long*)&c.counters->error_windows, 1ull);
                atomicAdd((unsigned long long*)&c.counters->consensus_bytes, (unsigned long long)W.tlen);
            }
            return;
        }
    }
    if (lane == 0) {
        c.win[w].final_len = fn; c.win[w].final_beg = fbeg; c.win[w].status = 0;
        atomicAdd((unsigned long long*)&c.counters->consensus_bytes, (unsigned long long)fn);
    }
}

// Dense outputs.  cons_off / solid_off hold the (already scanned) chunk-local offsets.
__global__ void __launch_bounds__(256) k_gather(CgChunk c, const u64* off_fin, const u64* cons_off, const u64* solid_off,
                                              u8* out_cons, u32* out_sk, u32* out_sc, u8* out_status,
                                              u64 cons_base, u64 solid_base, u64* g_cons_off, u64* g_solid_off) {
    const u32 w = blockIdx.x;
    const CgWin W = c.win[w];
    const u8* src = c.fin + off_fin[w] + W.final_beg;
    u8* dst = out_cons + cons_off[w];
    for (u32 i = threadIdx.x; i < W.final_len; i += blockDim.x) dst[i] = src[i];
    const u32* sk = c.solid_k + c.off_solid[w];
    const u32* sc = c.solid_c + c.off_solid[w];
    for (u32 i = threadIdx.x; i < W.n_solid; i += blockDim.x) { out_sk[solid_off[w] + i] = sk[i]; out_sc[solid_off[w] + i] = sc[i]; }
    if (threadIdx.x == 0) {
        out_status[w] = (u8)W.status;
        g_cons_off[w] = cons_base + cons_off[w];
        g_solid_off[w] = solid_base + solid_off[w];
        atomicAdd((unsigned long long*)&c.counters->solid_kmers, (unsigned long long)W.n_solid);
    }
}

// final_len / n_solid -> u64 arrays for k_scan
__global__ void k_out_sizes(CgChunk c, u64* cons_off, u64* solid_off) {
    const u32 w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= c.nwin) return;
    cons_off[w] = c.win[w].final_len;
    solid_off[w] = c.win[w].n_solid;
}
