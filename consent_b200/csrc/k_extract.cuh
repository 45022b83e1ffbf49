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
    u32 n_windows; const u32* win_pile; const u32* win_beg; const u32* win_end; const u64* slot_base;   // slot_base[w]: first slot of window w
    // per slot (slot 0 of a window = the template): piece length (0: not in the pile), first source base, step (+1 / -1 = revcomp)
    u32* slot_len; u64* slot_src; u32* slot_loc;       // slot_loc: byte offset of the piece inside its window
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
    const u64 sb = A.slot_base[w];
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
        if (s <= nal) {
            A.slot_len[sb + s] = len;
            A.slot_src[sb + s] = src;
            A.slot_loc[sb + s] = byte_base + b - len;
        }
        rank_base += __shfl_sync(CG_FULL, r, 31);
        byte_base += __shfl_sync(CG_FULL, b, 31);
    }
    bad = cg_ex_warp_or(bad);
    if (lane == 0) {
        A.win_nseq[w] = rank_base;
        A.win_nbytes[w] = byte_base;
        if (bad) atomicOr(A.flags, bad);
    }
}

__device__ __forceinline__ char cg_ex_comp(char c) {            // reverseComplement.cpp:34-43 (the store only holds A, C, G, T)
    return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : (char)0;
}

// One CTA per window: every kept piece copied from the store into the window's slice of the batch, seq_off filled.
__global__ void __launch_bounds__(256) k_ex_copy(CgExtractArgs A) {
    for (u32 w = blockIdx.x; w < A.n_windows; w += gridDim.x) {
        const u32 p = A.win_pile[w];
        const u32 nslot = A.pile_ov_begin[p + 1] - A.pile_ov_begin[p] + 1;
        const u64 sb = A.slot_base[w];
        const u64 wbase = A.win_base[w];
        const u32 seq0 = A.win_seq_begin[w];
        const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, nwarps = blockDim.x >> 5;
        // seq_off: the rank of a kept slot = number of kept slots before it (recomputed per 32 slots by warp 0)
        if (warp == 0) {
            u32 rank = 0;
            for (u32 s0 = 0; s0 < nslot; s0 += 32) {
                const u32 s = s0 + lane;
                const u32 len = s < nslot ? A.slot_len[sb + s] : 0u;
                const u32 m = __ballot_sync(CG_FULL, len != 0);
                if (len) A.seq_off[seq0 + rank + __popc(m & ((1u << lane) - 1u))] = wbase + A.slot_loc[sb + s];
                rank += __popc(m);
            }
        }
        // pieces: one warp per piece, 32 consecutive bases per step (coalesced on both sides)
        for (u32 s = warp; s < nslot; s += nwarps) {
            const u32 len = A.slot_len[sb + s];
            if (!len) continue;
            const u64 src = A.slot_src[sb + s];
            char* dst = A.bases + wbase + A.slot_loc[sb + s];
            if (src >> 63) {
                const char* from = A.store + (src & ~(1ull << 63));
                for (u32 i = lane; i < len; i += 32) dst[i] = cg_ex_comp(*(from - i));
            } else {
                const char* from = A.store + src;
                for (u32 i = lane; i < len; i += 32) dst[i] = from[i];
            }
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
