// k_extract.cuh — window extraction on the device.
//
// Phase A of processRead (src/CONSENT-correction.cpp:21-35) for every read pile of a batch:
//   k_ex_positions   getCoverages + getAlignmentWindowsPositions      (src/alignmentWindows.cpp:5-85)
//   k_ex_sizes       the per-overlap clipping arithmetic of getAlignmentWindowsSequences (src/alignmentWindows.cpp:87-149):
//                    which overlaps contribute to a window, from where in the target read, how many bases
//   k_ex_copy        the pile itself: bases cut from the read store (reverse-complemented for '-' overlaps,
//                    src/reverseComplement.cpp:6-24) straight into the resident window batch, plus seq_off
// The read store is shipped once (1 byte per base); a window's N sequences are never assembled on the host.  k_ex_copy is the
// one kernel of this repository that is bound by HBM bandwidth in the plain sense: it reads ~1 byte and writes 1 byte per pile
// base (7.8 GB of piles at the config-3 shape).
#pragma once
#include "cg_common.cuh"

enum { CG_EX_FLAG_BAD_OVERLAP = 256u, CG_EX_FLAG_EMPTY_PILE = 512u, CG_EX_FLAG_CAPACITY = 1024u, CG_EX_FLAG_SUBSTR = 2048u };

__device__ __forceinline__ u32 cg_ex_warp_or(u32 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v |= __shfl_xor_sync(CG_FULL, v, d);
    return v;
}

struct CgOverlapDev { u32 t_read, strand, q_start, q_end, t_start, t_end, t_length; };

struct CgExtractArgs {
    // store + piles
    const u64* store_off; const char* store; u32 n_store;
    u32 n_piles; const u32* pile_read; const u32* pile_qlen; const u32* pile_ov_begin; const CgOverlapDev* ov;
    u32 min_support, ws, ovl, k;
    // k_ex_positions: coverage scratch (pile p owns cov[cov_off[p] .. +qlen+1)), windows of pile p at slots win_cap_off[p] ..
    u32* cov; const u64* cov_off; const u64* win_cap_off; u32* cap_beg; u32* cap_end; u32* n_win;
    // dense windows (after the host's prefix over n_win)
    u32 n_windows; const u32* win_pile; const u32* win_beg; const u32* win_end;
    // per KEPT piece (sequence s of the batch: the slots of window w start at win_seq_begin[w], its first one is the template):
    // piece length, first source base (bit 63: walk backwards + complement), byte offset of the piece inside its window.
    // k_ex_sizes runs twice: mode 0 only counts (win_nseq, win_nbytes), mode 1 writes the pieces once win_seq_begin is known —
    // memory is per kept piece, not per (window, overlap of its pile): a contig's pile may hold tens of thousands of overlaps
    // of which a window sees its local coverage (CONSENT-polish: maxSupport 20000, CONSENT-polish:43).
    u32 mode;
    u32* slot_len; u64* slot_src; u32* slot_loc;
    u32* win_nseq; u32* win_nbytes;
    // outputs
    const u32* win_seq_begin; const u64* win_base; u64* seq_off; char* bases;
    u32* flags;
};

// Coverage (getCoverages :5-25) by the whole warp, then the two scans of getAlignmentWindowsPositions (:27-85) by lane 0 —
// a strictly sequential automaton (the cursor steps back by the window overlap after every window), ~tplLen steps per read.
__global__ void k_ex_positions(CgExtractArgs A) {
    const u32 lane = threadIdx.x & 31u;
    const u32 p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= A.n_piles) return;
    const u32 tplLen = A.pile_qlen[p];
    u32* cov = A.cov + A.cov_off[p];
    const u32 o0 = A.pile_ov_begin[p], o1 = A.pile_ov_begin[p + 1];
    u32 bad = 0;
    if (tplLen == 0) bad = CG_EX_FLAG_BAD_OVERLAP;
    for (u32 i = lane; i <= tplLen; i += 32) cov[i] = 0;
    __syncwarp();
    // difference array: +1 at qStart, -1 after qEnd
    for (u32 o = o0 + lane; o < o1; o += 32) {
        const CgOverlapDev a = A.ov[o];
        if (a.q_start > a.q_end) continue;
        if (a.q_end >= tplLen) { bad = CG_EX_FLAG_BAD_OVERLAP; continue; }             // the reference writes past its array
        atomicAdd(&cov[a.q_start], 1u);
        atomicSub(&cov[a.q_end + 1], 1u);
    }
    __syncwarp();
    bad = __ballot_sync(CG_FULL, bad != 0) ? CG_EX_FLAG_BAD_OVERLAP : 0u;
    // inclusive prefix sum, 32 positions per round
    u32 carry = 0;
    for (u32 b = 0; b <= tplLen && !bad; b += 32) {
        const u32 i = b + lane;
        u32 v = i <= tplLen ? cov[i] : 0u;
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(CG_FULL, v, d); if (lane >= d) v += o; }
        v += carry;
        if (i <= tplLen) cov[i] = v;
        carry = __shfl_sync(CG_FULL, v, 31);
    }
    __syncwarp();
    if (lane != 0) return;
    u32 n = 0;
    if (!bad) {
        u32* wb = A.cap_beg + A.win_cap_off[p];
        u32* we = A.cap_end + A.win_cap_off[p];
        const u32 cap = (u32)(A.win_cap_off[p + 1] - A.win_cap_off[p]);
        const u32 ws = A.ws, minSup = A.min_support;
        u32 curLen = 0, beg = 0, i = 0;
        while (i < tplLen) {                                                            // :38-55
            if (curLen >= ws) {
                if (n < cap) { wb[n] = beg; we[n] = beg + curLen - 1; }
                ++n;
                if (A.ovl) i = i - A.ovl;
                beg = i; curLen = 0;
            }
            if (cov[i] < minSup) { curLen = 0; i++; beg = i; } else { curLen++; i++; }
        }
        u32 pushed = 0, end = tplLen - 1;                                               // :57-80: the last window
        curLen = 0; i = tplLen - 1;
        while (i > 0 && !pushed) {
            if (curLen >= ws) {
                if (n < cap) { wb[n] = end - curLen + 1; we[n] = end; }
                ++n; pushed = 1; end = i; curLen = 0;
            }
            if (cov[i] < minSup) { curLen = 0; i--; end = i; } else { curLen++; i--; }
        }
        if (n > cap) { bad = CG_EX_FLAG_CAPACITY; n = 0; }
    }
    A.n_win[p] = n;
    if (bad) atomicOr(A.flags, bad);
}

// One warp per window: lanes over the slots (template + every overlap of the pile, in order): piece length and source, then
// the kept pieces' ranks and byte offsets inside the window (warp scans) and the window totals.
__global__ void k_ex_sizes(CgExtractArgs A) {
    const u32 lane = threadIdx.x & 31u;
    const u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= A.n_windows) return;
    const u32 p = A.win_pile[w];
    const u32 qBeg = A.win_beg[w], end = A.win_end[w];
    const u32 o0 = A.pile_ov_begin[p], nal = A.pile_ov_begin[p + 1] - o0;
    const u64 sb = A.mode ? (u64)A.win_seq_begin[w] : 0ull;
    const u32 q = A.pile_read[p];
    u32 bad = 0, rank_base = 0, byte_base = 0;
    for (u32 s0 = 0; s0 < nal + 1; s0 += 32) {
        const u32 s = s0 + lane;
        u32 len = 0;
        u64 src = 0;                                            // index of the first output base in the store; bit 63: walk backwards + complement
        if (s == 0) {                                           // the template, :95-100
            const u64 qlen = A.store_off[q + 1] - A.store_off[q];
            len = end - qBeg + 1;
            if ((u64)qBeg + len - 1 >= qlen) { bad |= CG_EX_FLAG_EMPTY_PILE; len = 0; }
            src = A.store_off[q] + qBeg;
        } else if (s <= nal) {
            const CgOverlapDev o = A.ov[o0 + s - 1];
            u32 tBeg = o.t_start, tEnd = o.t_end, length = end - qBeg + 1;
            u32 shift = qBeg > o.q_start ? qBeg - o.q_start : 0u;
            if (((o.q_start <= qBeg && o.q_end > qBeg) || (end <= o.q_end && o.q_start < end)) && o.t_start + shift <= o.t_end) {   // :114
                if (qBeg < o.q_start && o.q_end < end) {                                                                               // :116-120
                    shift = 0;
                    tBeg = (u32)max(0, (int)o.t_start - ((int)o.q_start - (int)qBeg));
                    tEnd = (u32)min((int)o.t_length - 1, (int)o.t_end + ((int)end - (int)o.q_end));
                    length = tEnd - tBeg + 1;
                } else if (qBeg < o.q_start) {                                                                                         // :121-124
                    shift = 0;
                    tBeg = (u32)max(0, (int)o.t_start - ((int)o.q_start - (int)qBeg));
                    length = (u32)min((int)length, min((int)o.t_length - 1, (int)tBeg + (int)length - 1) - (int)tBeg + 1);
                } else if (o.q_end < end) {                                                                                            // :125-128
                    tEnd = (u32)min((int)o.t_length - 1, (int)o.t_end + ((int)end - (int)o.q_end));
                    length = (u32)min((int)length, (int)tEnd - max(0, (int)tEnd - (int)length + 1) + 1);
                }
                if (o.t_read >= A.n_store) { bad |= CG_EX_FLAG_BAD_OVERLAP; }
                else {
                    const u64 tlen = A.store_off[o.t_read + 1] - A.store_off[o.t_read];
                    if ((u64)tBeg > tlen) bad |= CG_EX_FLAG_SUBSTR;                      // substr(tBeg, ..) throws
                    else {
                        u64 n1 = (u64)(u32)(tEnd - tBeg + 1);                           // substr(tBeg, tEnd - tBeg + 1), :130
                        if (n1 > tlen - tBeg) n1 = tlen - tBeg;
                        if ((u64)shift > n1) bad |= CG_EX_FLAG_SUBSTR;                  // .substr(shift, length) throws, :135
                        else {
                            u64 n2 = length;
                            if (n2 > n1 - shift) n2 = n1 - shift;
                            if (n2 >= A.k) {                                            // :138-140
                                len = (u32)n2;
                                if (n2 > CG_LEN_MAX) { bad |= CG_EX_FLAG_CAPACITY; len = 0; }
                                src = o.strand ? ((A.store_off[o.t_read] + tBeg + n1 - 1 - shift) | (1ull << 63))
                                               : (A.store_off[o.t_read] + tBeg + shift);
                            }
                        }
                    }
                }
            }
        }
        // exclusive scans of (kept, len) over the 32 slots of this round
        u32 kept = len ? 1u : 0u, r = kept, b = len;
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) {
            const u32 ro = __shfl_up_sync(CG_FULL, r, d), bo = __shfl_up_sync(CG_FULL, b, d);
            if (lane >= d) { r += ro; b += bo; }
        }
        if (A.mode && len) {
            const u64 at = sb + rank_base + r - 1;
            A.slot_len[at] = len;
            A.slot_src[at] = src;
            A.slot_loc[at] = byte_base + b - len;
        }
        rank_base += __shfl_sync(CG_FULL, r, 31);
        byte_base += __shfl_sync(CG_FULL, b, 31);
    }
    bad = cg_ex_warp_or(bad);
    if (lane == 0) {
        if (!A.mode) { A.win_nseq[w] = rank_base; A.win_nbytes[w] = byte_base; }
        if (bad) atomicOr(A.flags, bad);
    }
}

__device__ __forceinline__ char cg_ex_comp(char c) {            // reverseComplement.cpp:34-43 (the store only holds A, C, G, T)
    return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : (char)0;
}
// four bases at once: A (0x41) <-> T (0x54) is x ^ 0x15, C (0x43) <-> G (0x47) is x ^ 0x04; bit 1 tells the pairs apart
__device__ __forceinline__ u32 cg_ex_comp4(u32 x) { return x ^ 0x15151515u ^ (((x >> 1) & 0x01010101u) * 0x11u); }
// the 4 bytes at byte address `a` of a buffer of aligned 32-bit words (little endian): two aligned loads + a funnel shift
__device__ __forceinline__ u32 cg_ex_load4(const char* base, u64 a) {
    const u32* w = (const u32*)(base + (a & ~(u64)3));
    const u32 sh = (u32)(a & 3) * 8;
    const u32 lo = w[0];
    return sh ? __funnelshift_r(lo, w[1], sh) : lo;
}

// One piece by one warp: dst[0..len) <- store[src..] (forward) or the complement of store[src], store[src-1], ... (reverse).
// Head bytes up to the first 16-byte boundary of dst and the tail go byte by byte; the body moves 16 bytes per lane and step
// (one 128-bit store, aligned 32-bit loads re-aligned by funnel shifts), 512 bytes per warp and step.
__device__ __forceinline__ void cg_ex_copy_piece(char* dst, const char* store, u64 src, bool reverse, u32 len, u32 lane) {
    const u32 head = min(len, (u32)((16 - ((u64)dst & 15)) & 15));
    const u32 body = (len - head) & ~15u;
    if (!reverse) {
        if (lane < head) dst[lane] = store[src + lane];
        for (u32 o = head + lane * 16; o < head + body; o += 512) {
            uint4 v;
            v.x = cg_ex_load4(store, src + o); v.y = cg_ex_load4(store, src + o + 4);
            v.z = cg_ex_load4(store, src + o + 8); v.w = cg_ex_load4(store, src + o + 12);
            *(uint4*)(dst + o) = v;
        }
        for (u32 o = head + body + lane; o < len; o += 32) dst[o] = store[src + o];
    } else {
        if (lane < head) dst[lane] = cg_ex_comp(store[src - lane]);
        for (u32 o = head + lane * 16; o < head + body; o += 512) {
            uint4 v;                                            // output bytes o .. o+3 = complement of store[src-o], store[src-o-1], ...
            v.x = cg_ex_comp4(__byte_perm(cg_ex_load4(store, src - o - 3), 0, 0x0123));
            v.y = cg_ex_comp4(__byte_perm(cg_ex_load4(store, src - o - 7), 0, 0x0123));
            v.z = cg_ex_comp4(__byte_perm(cg_ex_load4(store, src - o - 11), 0, 0x0123));
            v.w = cg_ex_comp4(__byte_perm(cg_ex_load4(store, src - o - 15), 0, 0x0123));
            *(uint4*)(dst + o) = v;
        }
        for (u32 o = head + body + lane; o < len; o += 32) dst[o] = cg_ex_comp(store[src - o]);
    }
}

// One CTA per window: every kept piece copied from the store into the window's slice of the batch, seq_off filled.
__global__ void __launch_bounds__(256) k_ex_copy(CgExtractArgs A) {
    for (u32 w = blockIdx.x; w < A.n_windows; w += gridDim.x) {
        const u32 seq0 = A.win_seq_begin[w], nslot = A.win_seq_begin[w + 1] - seq0;
        const u64 wbase = A.win_base[w];
        const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, nwarps = blockDim.x >> 5;
        for (u32 s = threadIdx.x; s < nslot; s += blockDim.x) A.seq_off[seq0 + s] = wbase + A.slot_loc[seq0 + s];
        // pieces: one warp per piece
        for (u32 s = warp; s < nslot; s += nwarps) {
            const u32 len = A.slot_len[seq0 + s];
            const u64 src = A.slot_src[seq0 + s];
            cg_ex_copy_piece(A.bases + wbase + A.slot_loc[seq0 + s], A.store, src & ~(1ull << 63), (src >> 63) != 0, len, lane);
        }
    }
}

// the read store as the reference stores reads: upper-cased, anything but A, C, G becomes T (src/utils.cpp:21-32,189)
__global__ void k_ex_normalise(char* store, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const char c = store[i] & ~0x20;
        store[i] = (c == 'A' || c == 'C' || c == 'G') ? c : 'T';
    }
}
