// k_chain.cuh — longest ordered chain of template anchors.
//
// Replaces anchors_ordered_according2reads (BMEAN/bmean.cpp:161-186), longest_ordered_chain_from_anchors
// (:190-216) and longest_ordered_chain (:239-260).
//
// score(a,b), a before b on the template = -1 if some read holds both with pos(a) > pos(b), else the number
// of reads holding both.  The reference merges two location lists per pair (A^2/2 * N work).  Here:
//   - "reads holding both" = popcount(present[a] & present[b]) over N-bit masks;
//   - "-1" pairs are found per read: the anchors present in a read (a sparse subset) are compacted by ballot
//     and only those are compared pairwise, setting bits of an A x A inversion matrix.
// The chain DP itself is a backward sweep (best[start] depends on best[i], i > start), one warp, lanes over i,
// arg-max by a packed 64-bit key that encodes the reference's tie rules:
//   inner scan : larger length, then larger total score, then SMALLER i   (:198-199, strict '>')
//   top level  : larger length, then larger score, then LARGER start      (:244-254, scan from the end)
#pragma once
#include "cg_common.cuh"

#define CG_CHAIN_THREADS 128u
#define CG_CHAIN_WARPS (CG_CHAIN_THREADS / 32u)

__host__ __device__ inline u32 cg_chain_nwp(u32 N) { return ((N + 31u) / 32u) | 1u; }   // odd stride: no bank conflicts
__host__ __device__ inline size_t cg_chain_smem(u32 A, u32 N) {
    size_t aw = (A + 31u) / 32u;
    return 4 * ((size_t)A * cg_chain_nwp(N) + (size_t)A * aw + (size_t)CG_CHAIN_WARPS * A + 2 * (size_t)A) + 64;
}

// Two launches: pass 0 with a shared-memory size that fits ordinary windows (many CTAs per SM) defers the windows that
// need more; pass 1, with the device's maximum, takes those (and flags what does not fit even then).
#define CG_CHAIN_DEFERRED 0xffffffffu
__global__ void __launch_bounds__(CG_CHAIN_THREADS) k_chain(CgChunk c, u32 smem_bytes, u32 pass) {
    CG_DYN_SMEM(smem);
    const u32 w = blockIdx.x, tid = threadIdx.x, lane = cg_lane(), warp = cg_warp();
    const CgWin W = c.win[w];
    const u32 A = W.n_alive, N = W.n_seqs, C = W.n_cand, S = W.S;
    if (W.bad) return;
    if (pass == 1 && W.n_chain != CG_CHAIN_DEFERRED) return;
    if (A == 0) {
        if (tid == 0) c.win[w].n_chain = 0;
        return;
    }
    if (cg_chain_smem(A, N) > smem_bytes) {
        if (tid == 0) {
            if (pass == 0) c.win[w].n_chain = CG_CHAIN_DEFERRED;
            else { c.win[w].n_chain = 0; c.win[w].bad = 1; }      // -> raw template, status CG_WINDOW_ERROR
        }
        return;
    }
    const u32 NWp = cg_chain_nwp(N), AW = (A + 31u) / 32u;
    u32* pres = (u32*)smem;                       // [A][NWp]  bit r of pres[a]: read r holds anchor a
    u32* inv = pres + (size_t)A * NWp;            // [A][AW]   bit b of inv[a]: some read has pos(a) > pos(b)
    u32* clist = inv + (size_t)A * AW;            // [warps][A] (anchor << 16 | pos+1) of the anchors present in one read
    u32* bscore = clist + (size_t)CG_CHAIN_WARPS * A;
    u16* blen = (u16*)(bscore + A);
    u16* bnext = blen + A;
    const u64 slot_base = c.off_slot[w];
    const u16* anchors = c.anchors + slot_base;
    const u16* pos = c.pos + c.off_pos[w];

    for (u32 i = tid; i < A * NWp + A * AW; i += CG_CHAIN_THREADS) pres[i] = 0;
    __syncthreads();

    u32* cl = clist + (size_t)warp * A;
    for (u32 r = warp; r < N; r += CG_CHAIN_WARPS) {
        const u16* prow = pos + (size_t)r * C;
        u32 m = 0;
        for (u32 ab = 0; ab < A; ab += 128) {           // four independent loads in flight per lane
            u32 pv[4];
#pragma unroll
            for (u32 t = 0; t < 4; ++t) {
                const u32 a = ab + 32 * t + lane;
                pv[t] = a < A ? prow[anchors[a]] : 0u;
            }
#pragma unroll
            for (u32 t = 0; t < 4; ++t) {
                const u32 a = ab + 32 * t + lane;
                const u32 p = pv[t];
                const u32 bal = __ballot_sync(CG_FULL, p != 0);
                if (p) {
                    cl[m + __popc(bal & ((1u << lane) - 1u))] = (a << 16) | p;
                    atomicOr(&pres[a * NWp + (r >> 5)], 1u << (r & 31u));
                }
                m += __popc(bal);
            }
        }
        __syncwarp();
        if (r != 0) {                              // the template holds its anchors in order by construction
            // an anchor is the left end of an inversion iff a smaller position follows it: suffix minima, back to front
            u32 carry = 0xffffu;
            for (int jb = (int)((m ? m - 1 : 0) & ~31u); jb >= 0; jb -= 32) {
                const u32 j = (u32)jb + lane;
                const u32 ej = j < m ? cl[j] : 0xffffffffu;
                const u32 pj = ej & 0xffffu;
                u32 sm = pj;                          // inclusive suffix minimum inside the chunk
#pragma unroll
                for (int dlt = 1; dlt < 32; dlt <<= 1) {
                    const u32 o = __shfl_down_sync(CG_FULL, sm, dlt);
                    if (lane + (u32)dlt < 32u) sm = o < sm ? o : sm;
                }
                u32 after = __shfl_down_sync(CG_FULL, sm, 1);     // minimum over the lanes behind this one ...
                if (lane == 31) after = 0xffffu;
                after = after < carry ? after : carry;            // ... and over the chunks behind
                if (j < m && after < pj) {
                    const u32 aj = ej >> 16;
                    for (u32 l = j + 1; l < m; ++l) {
                        const u32 el = cl[l];
                        if ((el & 0xffffu) < pj) atomicOr(&inv[aj * AW + ((el >> 16) >> 5)], 1u << ((el >> 16) & 31u));
                    }
                }
                const u32 first = __shfl_sync(CG_FULL, sm, 0);
                carry = first < carry ? first : carry;
            }
        }
        __syncwarp();
    }
    __syncthreads();

    if (warp != 0) return;
    const u64 VALID = 1ull << 63;
    for (int start = (int)A - 1; start >= 0; --start) {
        u64 best = 0;
        const u32* ps = pres + (size_t)start * NWp;
        const u32* is = inv + (size_t)start * AW;
        for (u32 ib = (u32)start + 1; ib < A; ib += 32) {
            const u32 i = ib + lane;
            if (i < A && !((is[i >> 5] >> (i & 31u)) & 1u)) {
                const u32* pi = pres + (size_t)i * NWp;
                u32 score = 0;
                for (u32 q = 0; q < NWp; ++q) score += __popc(ps[q] & pi[q]);
                if (score >= S) {
                    const u64 key = VALID | ((u64)blen[i] << 40) | ((u64)(bscore[i] + score) << 16) | (u64)(0xffffu - i);
                    if (key > best) best = key;
                }
            }
        }
        best = cg_warp_max64(best);
        if (lane == 0) {
            if (best == 0) { blen[start] = 0; bscore[start] = 0; bnext[start] = CG_NONE16; }
            else {
                blen[start] = (u16)(((best >> 40) & 0x7fffffu) + 1);
                bscore[start] = (u32)((best >> 16) & 0xffffffu);
                bnext[start] = (u16)(0xffffu - (u32)(best & 0xffffu));
            }
        }
        __syncwarp();
    }
    u64 top = 0;
    for (u32 ib = 0; ib < A; ib += 32) {
        const u32 i = ib + lane;
        if (i < A && blen[i] > 0) {
            const u64 key = ((u64)blen[i] << 40) | ((u64)bscore[i] << 16) | (u64)i;
            if (key > top) top = key;
        }
    }
    top = cg_warp_max64(top);
    if (lane == 0) {
        u32 n = 0;
        if (top != 0) {
            u32 i = (u32)(top & 0xffffu);
            u16* chain = c.chain + slot_base;
            while (i != CG_NONE16) { chain[n++] = anchors[i]; i = bnext[i]; }
        }
        c.win[w].n_chain = n;
    }
}
