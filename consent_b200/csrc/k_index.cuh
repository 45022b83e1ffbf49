// k_index.cuh — k-mer index of one pile: solid k-mer list, template anchor candidates, position table.
//
// Replaces fill_index_kmers (BMEAN/bmean.cpp:43-84), filter_index_kmers (:88-114), get_template (:220-234).
//
// The reference keeps a hash map kmer -> vector<(read,pos)> of all ~N*(L-k+1) occurrences.  Only two things
// are ever read from it: (1) merCounts = total occurrences of k-mers seen >= solid times (:72-76) and
// (2) the location lists of k-mers that occur in the template, are never repeated inside one read (:58-60,
// 66-68, 78-82) and are present in >= S reads (:101).  So this kernel
//   - counts all k-mers of the pile in a direct-addressed shared-memory table; the solid entries are compacted
//     *in key order* (the sorted list the ABI returns) and the counts of the template's k-mers are picked up;
//   - keeps template positions with S <= count <= N (necessary for "alive"), gives each a slot, and in one
//     more sweep over the pile records pos[read][slot] = position + 1;
//   - a slot is alive iff no read holds its k-mer twice (every occurrence sets one bit per (read, slot) in shared memory;
//     finding it set is a second occurrence).
//
// One CTA (1024 threads) per window.  The 2-bit pile (words + tags, <= 64 KB) is bulk-copied HBM -> shared memory by the
// TMA engine (cp.async.bulk + mbarrier) while the counters are zeroed; every sweep then reads it from there.  Bigger piles
// (polishing depth) are read through L1/L2.
//
// Counting, k <= 9, pile staged (every window of the BASELINE configs): BYTE counters, 2^17 keys per pass in the 128 KB table
// (two passes for k = 9, one below), one shared-memory atomic per occurrence (the add lands in the key's byte of its 32-bit
// word and returns the old value).  The occurrence that lifts a counter to `solid` sets the key's bit in a 16 KB bitmap, so
// the solid list of a pass is read off 4096 words instead of 131072 counters.  A counter that reaches 255 raises a flag and
// the window is counted again in 32 bits (8 passes of 2^15 keys, the path of the deeper piles), so the result never depends
// on the counter width; a k-mer seen 255 times in one window means a low-complexity pile.
// k = 10 .. 15: an open-addressing table per SM in global memory (below).
#pragma once
#include "cg_common.cuh"

#ifndef CG_IDX_THREADS
#define CG_IDX_THREADS 1024u
#endif
#define CG_IDX_SMEM_BYTES (131072u + 8u * CG_PW_CAP + 2u * 2048u * 4u + 16384u + 64u * 4u + 16u)     // table | pile words, tags | template k-mers, counts | solid bitmap | misc, barrier
__global__ void __launch_bounds__(CG_IDX_THREADS, 1) k_index(CgChunk c) {
    CG_DYN_SMEM(smem);
    u32* tab = (u32*)smem;
    u32* pile_s = tab + 32768;
    u32* tags_s = pile_s + CG_PW_CAP;
    u32* tkmer = tags_s + CG_PW_CAP;
    u32* tcount = tkmer + 2048;
    u32* sbm = tcount + 2048;                        // byte-counter path: one bit per key of the pass, set when the key turns solid
    u32* misc = sbm + 4096;

    const u32 w = blockIdx.x, tid = threadIdx.x, lane = cg_lane(), warp = cg_warp();
    const u32 T = CG_IDX_THREADS, NWARPS = CG_IDX_THREADS / 32;
    const CgWin W = c.win[w];
    if (W.bad) return;                               // k_plan left n_cand = n_alive = n_solid = 0
    const u32 k = c.k, N = W.n_seqs, tk = W.tk, S = W.S;

    const u64 g0 = cg_pword(c.seq_off, W.seq_begin) - c.pword_base;
    const u32 nw = (u32)(cg_pword(c.seq_off, W.seq_begin + N) - c.pword_base - g0);

    const u32 solid_thr = c.solid;
    const u64 solid_base = c.off_solid[w];
    u32 nsolid = 0;
    const u32* pw;
    const u32* pt;
    const bool hashed = k > CG_KMAX;
    // ---- stage the 2-bit pile in shared memory (TMA bulk copy), if it fits
    bool staged = false;
    if (!hashed) {
        const u64 ga = g0 & ~3ull;                    // 16-byte aligned source
        const u32 shift = (u32)(g0 - ga);
        const u32 ncopy = (shift + nw + 1 + 3) & ~3u; // + the look-ahead word; multiple of 16 bytes
        if (ncopy <= CG_PW_CAP) {
            u64* bar = (u64*)(misc + 64);
            if (tid == 0) cg_bulk_issue2(pile_s, c.pwords + ga, tags_s, c.ptags + ga, ncopy * 4u, bar);
            for (u32 i = tid; i < 32768; i += T) tab[i] = 0;         // the first pass's counters, while the copy is in flight
            if (tid == 0) misc[0] = 0;
            __syncthreads();                                          // the barrier's init is visible to every waiter
            cg_bulk_wait(bar);
            pw = pile_s + shift;
            pt = tags_s + shift;
            staged = true;
        }
    }
    bool counted = false;
    if (staged) {
        // ---- byte counters: 2^17 keys per pass
        for (u32 p = tid; p < tk; p += T) {
            tkmer[p] = cg_kmer_at(pw[p >> 4], pw[(p >> 4) + 1], p & 15u, k);
            tcount[p] = 0;
        }
        const u32 key_bits = 2 * k < 17u ? 2 * k : 17u;
        const u32 key_mask = (1u << key_bits) - 1u;
        const u32 npass = 1u << (2 * k - key_bits);
        const u32 nbw = ((1u << key_bits) + 31u) / 32u;        // words of the solid bitmap
        const u32 per_warp = (nbw + NWARPS - 1) / NWARPS;
        const u32 vb = warp * per_warp < nbw ? warp * per_warp : nbw;
        const u32 ve = vb + per_warp < nbw ? vb + per_warp : nbw;
        const u32 thrm1 = solid_thr - 1u;                      // a key turns solid when its counter leaves this value (solid >= 1)
        const u8* tab8 = (const u8*)tab;
        bool ok = true;
        for (u32 pass = 0; pass < npass && ok; ++pass) {
            if (pass)
                for (u32 i = tid; i < 32768; i += T) tab[i] = 0;
            for (u32 i = tid; i < nbw; i += T) sbm[i] = 0;
            __syncthreads();
            u32 top = 0;
            for (u32 g = tid; g < nw; g += T) {
                const u32 tag = pt[g];
                if (tag == CG_NONE32) continue;
                const u32 nv = (tag & 15u) + 1;
                const u32 w0 = pw[g], w1 = pw[g + 1];
#pragma unroll
                for (u32 b = 0; b < 16; ++b) {
                    if (b < nv) {
                        const u32 km = cg_kmer_at(w0, w1, b, k);
                        if ((km >> key_bits) == pass) {
                            const u32 key = km & key_mask;
                            const u32 old = atomicAdd(&tab[key >> 2], 1u << (8u * (key & 3u)));
                            const u32 ob = __byte_perm(old, 0, 0x4440u | (key & 3u));     // the key's byte before the add
                            top = ob > top ? ob : top;
                            if (ob == thrm1) atomicOr(&sbm[key >> 5], 1u << (key & 31u));
                        }
                    }
                }
            }
            if (top == 255u) misc[0] = 1;                      // a counter wrapped: recount in 32 bits
            __syncthreads();
            if (misc[0]) { ok = false; break; }
            for (u32 p = tid; p < tk; p += T) {
                const u32 km = tkmer[p];
                if ((km >> key_bits) == pass) tcount[p] = tab8[km & key_mask];
            }
            // solid entries of this pass, in key order: count per warp range, scan, write
            u32 wc = 0;
            for (u32 v = vb + lane; v < ve; v += 32) wc += (u32)__popc(sbm[v]);
            wc = cg_warp_sum(wc);
            if (lane == 0) misc[1 + warp] = wc;
            __syncthreads();
            u32 woff = 0, total = 0;
            for (u32 i = 0; i < NWARPS; ++i) { const u32 v = misc[1 + i]; if (i < warp) woff += v; total += v; }
            if (wc) {
                u64 run = solid_base + nsolid + woff;
                for (u32 v0 = vb; v0 < ve; v0 += 32) {
                    const u32 v = v0 + lane;
                    u32 mm = v < ve ? sbm[v] : 0u;
                    const u32 n = (u32)__popc(mm);
                    const u32 inc = cg_warp_scan(n);
                    u64 idx = run + inc - n;
                    while (mm) {
                        const u32 key = 32u * v + (u32)__ffs((int)mm) - 1u;
                        mm &= mm - 1u;
                        c.solid_k[idx] = (pass << key_bits) | key;
                        c.solid_c[idx] = tab8[key];
                        ++idx;
                    }
                    run += __shfl_sync(CG_FULL, inc, 31);
                }
            }
            nsolid += total;
            __syncthreads();
        }
        counted = ok;
        if (!ok) nsolid = 0;
    }
    if (counted) {
    } else
    if (hashed) {
        // ---- k = 10 .. 15: 4^k keys do not fit a direct table.  Every occurrence goes into an open-addressing table in global
        // memory (one table per SM, L2-resident: atomicCAS claims the key, atomicAdd counts), the template's k-mers are looked up,
        // the solid keys are collected, sorted in shared memory (bitonic, <= 32768 keys) and written in key order with their counts.
        pw = c.pwords + g0;
        pt = c.ptags + g0;
        const u32 slot = cg_smid();                  // one table per SM (the emulator: per CTA)
        const u32 cap = c.idx_cap, hmask = cap - 1u;
        bool ovf = slot >= c.idx_slots || (u64)2 * W.n_occ > (u64)cap;
        u32* hk = c.idx_keys + (size_t)(ovf ? 0u : slot) * cap;
        u32* hc = c.idx_counts + (size_t)(ovf ? 0u : slot) * cap;
        if (!ovf) {
            for (u32 i = tid; i < cap; i += T) { hk[i] = CG_NONE32; hc[i] = 0; }
            for (u32 p = tid; p < tk; p += T) { tkmer[p] = cg_kmer_at(pw[p >> 4], pw[(p >> 4) + 1], p & 15u, k); tcount[p] = 0; }
            if (tid == 0) misc[0] = 0;
            __syncthreads();
            for (u32 g = tid; g < nw; g += T) {
                const u32 tag = pt[g];
                if (tag == CG_NONE32) continue;
                const u32 nv = (tag & 15u) + 1;
                const u32 w0 = pw[g], w1 = pw[g + 1];
                for (u32 b = 0; b < nv; ++b) {
                    const u32 km = cg_kmer_at(w0, w1, b, k);
                    u32 hh = (km * 2654435761u) & hmask;
                    for (;;) {
                        const u32 old = atomicCAS(&hk[hh], CG_NONE32, km);
                        if (old == CG_NONE32 || old == km) { atomicAdd(&hc[hh], 1u); break; }
                        hh = (hh + 1) & hmask;
                    }
                }
            }
            __syncthreads();
            for (u32 p = tid; p < tk; p += T) {
                const u32 km = tkmer[p];
                u32 hh = (km * 2654435761u) & hmask;
                while (hk[hh] != km) hh = (hh + 1) & hmask;               // the template's own k-mers are all in the table
                tcount[p] = hc[hh];
            }
            // solid keys -> tab[] (any order), then sorted
            for (u32 i = tid; i < cap; i += T) {
                if (hk[i] != CG_NONE32 && hc[i] >= solid_thr) {
                    const u32 at = atomicAdd(&misc[0], 1u);
                    if (at < CG_IDX_SOLID_CAP) tab[at] = hk[i];
                }
            }
            __syncthreads();
            nsolid = misc[0];
            if (nsolid > CG_IDX_SOLID_CAP) ovf = true;
        }
        if (ovf) {                                   // more k-mers than this path holds: the window is left uncorrected (CG_WINDOW_ERROR)
            if (tid == 0) { CgWin* Wg = &c.win[w]; Wg->bad = 1; Wg->n_solid = 0; Wg->n_cand = 0; Wg->n_alive = 0; }
            return;
        }
        u32 npow = 1;
        while (npow < nsolid) npow <<= 1;
        for (u32 i = nsolid + tid; i < npow; i += T) tab[i] = CG_NONE32;
        __syncthreads();
        for (u32 kk = 2; kk <= npow; kk <<= 1)
            for (u32 j = kk >> 1; j > 0; j >>= 1) {
                for (u32 i = tid; i < npow; i += T) {
                    const u32 l = i ^ j;
                    if (l > i) {
                        const u32 a = tab[i], b = tab[l];
                        if (((i & kk) == 0) == (a > b)) { tab[i] = b; tab[l] = a; }
                    }
                }
                __syncthreads();
            }
        for (u32 i = tid; i < nsolid; i += T) {
            const u32 km = tab[i];
            u32 hh = (km * 2654435761u) & hmask;
            while (hk[hh] != km) hh = (hh + 1) & hmask;
            c.solid_k[solid_base + i] = km;
            c.solid_c[solid_base + i] = hc[hh];
        }
        __syncthreads();
    } else {
    if (!staged) {
        pw = c.pwords + g0;
        pt = c.ptags + g0;
    }

    // ---- template k-mers (read 0 starts at word 0 of the pile)
    for (u32 p = tid; p < tk; p += T) {
        tkmer[p] = cg_kmer_at(pw[p >> 4], pw[(p >> 4) + 1], p & 15u, k);
        tcount[p] = 0;
    }

    // ---- counting passes
    const u32 tab_bits = 2 * k < CG_TAB_BITS ? 2 * k : CG_TAB_BITS;
    const u32 tab_n = 1u << tab_bits, tab_mask = tab_n - 1;
    const u32 npass = 1u << (2 * k - tab_bits);
    u32 per_warp = ((tab_n + NWARPS - 1) / NWARPS + 31u) & ~31u;
    const u32 wb = warp * per_warp < tab_n ? warp * per_warp : tab_n;
    const u32 we = wb + per_warp < tab_n ? wb + per_warp : tab_n;

    for (u32 pass = 0; pass < npass; ++pass) {
        for (u32 i = tid; i < tab_n; i += T) tab[i] = 0;
        __syncthreads();
        for (u32 g = tid; g < nw; g += T) {
            const u32 tag = pt[g];
            if (tag == CG_NONE32) continue;
            const u32 nv = (tag & 15u) + 1;
            const u32 w0 = pw[g], w1 = pw[g + 1];
#pragma unroll
            for (u32 b = 0; b < 16; ++b) {
                if (b < nv) {
                    const u32 km = cg_kmer_at(w0, w1, b, k);
                    if ((km >> tab_bits) == pass) atomicAdd(&tab[km & tab_mask], 1u);
                }
            }
        }
        __syncthreads();
        for (u32 p = tid; p < tk; p += T) {
            const u32 km = tkmer[p];
            if ((km >> tab_bits) == pass) tcount[p] = tab[km & tab_mask];
        }
        // solid entries of this pass, in key order: count per warp range, scan, write
        u32 wc = 0;
        for (u32 b = wb; b < we; b += 32) {
            const u32 i = b + lane;
            const bool f = i < we && tab[i] >= solid_thr;
            wc += __popc(__ballot_sync(CG_FULL, f));
        }
        if (lane == 0) misc[warp] = wc;
        __syncthreads();
        u32 woff = 0, total = 0;
        for (u32 i = 0; i < NWARPS; ++i) { const u32 v = misc[i]; if (i < warp) woff += v; total += v; }
        u64 run = solid_base + nsolid + woff;
        for (u32 b = wb; b < we; b += 32) {
            const u32 i = b + lane;
            const u32 cnt = i < we ? tab[i] : 0;
            const bool f = i < we && cnt >= solid_thr;
            const u32 m = __ballot_sync(CG_FULL, f);
            if (f) {
                const u64 idx = run + __popc(m & ((1u << lane) - 1u));
                c.solid_k[idx] = (pass << tab_bits) | i;
                c.solid_c[idx] = cnt;
            }
            run += __popc(m);
        }
        nsolid += total;
        __syncthreads();
    }
    }

    // ---- candidate anchors: template positions with S <= count <= N, slots in template order
    u32* bitmap = tab;                      // 4^k bits
    u32* hkey = tab + 8192;                 // 4096 keys
    u16* hval = (u16*)(tab + 12288);        // 4096 slots
    u32* scnt = tab + 14336;                // count by slot
    u32* filled = tab + 16384;              // reads holding the slot | "some read holds it twice"
    u32* pairs = tab + 18432;               // one bit per (read, slot): 14336 words
    const u32 bm_words = hashed ? 0u : ((1u << (2 * k)) >= 32 ? (1u << (2 * k)) / 32 : 1);      // pre-filter of the position sweep (4^k bits)
    for (u32 i = tid; i < bm_words; i += T) bitmap[i] = 0;
    for (u32 i = tid; i < 4096; i += T) hkey[i] = CG_NONE32;
    for (u32 i = tid; i < 2048; i += T) filled[i] = 0;
    const u64 slot_base = c.off_slot[w];
    u32 flags4 = 0, nloc = 0;
#pragma unroll
    for (u32 q = 0; q < 4; ++q) {
        const u32 p = tid * 4 + q;
        if (p < tk) {
            const u32 cnt = tcount[p];
            if (cnt >= S && cnt <= N) { flags4 |= 1u << q; ++nloc; }
        }
    }
    u32 C = 0;
    u32 slot = cg_block_scan(nloc, misc, &C);
    // A slot is alive iff no read holds its k-mer twice.  Small tables: one bit per (read, slot) in shared memory tells a second
    // occurrence apart as it is recorded; big ones: the reads holding each slot are counted from the position table afterwards.
    const bool pair_bits = (u64)C * N <= 14336ull * 32ull;
    if (pair_bits) for (u32 i = tid; i < (C * N + 31u) / 32u; i += T) pairs[i] = 0;
#pragma unroll
    for (u32 q = 0; q < 4; ++q) {
        if (flags4 & (1u << q)) {
            const u32 p = tid * 4 + q, km = tkmer[p];
            scnt[slot] = tcount[p];
            c.slot_tpos[slot_base + slot] = (u16)p;
            c.slot_kmer[slot_base + slot] = km;
            if (!hashed) atomicOr(&bitmap[km >> 5], 1u << (km & 31u));
            u32 h = (km * 2654435761u) >> 20;
            for (;;) {
                const u32 old = atomicCAS(&hkey[h], CG_NONE32, km);
                if (old == CG_NONE32) { hval[h] = (u16)slot; break; }    // one writer per entry.  A k-mer that occurs twice in the template
                if (old == km) { if (pair_bits) filled[slot] = 1; break; }   // keeps its first claimant's slot: neither slot can be alive
                h = (h + 1) & 4095u;                                      // (count > reads holding it), whichever it is
            }
            ++slot;
        }
    }
    // zero the position table [N][C]
    u16* pos = c.pos + c.off_pos[w];
    {
        u32* p32 = (u32*)pos;
        const u32 n32 = (u32)(((u64)C * N + 1) / 2);
        for (u32 i = tid; i < n32; i += T) p32[i] = 0;
    }
    __syncthreads();

    // ---- positions of the candidate k-mers in every read
    if (C) {
        for (u32 g = tid; g < nw; g += T) {
            const u32 tag = pt[g];
            if (tag == CG_NONE32) continue;
            const u32 nv = (tag & 15u) + 1, r = tag >> 16, wi = (tag >> 4) & 0xfffu;
            const u32 w0 = pw[g], w1 = pw[g + 1];
#pragma unroll
            for (u32 b = 0; b < 16; ++b) {
                if (b < nv) {
                    const u32 km = cg_kmer_at(w0, w1, b, k);
                    if (hashed || ((bitmap[km >> 5] >> (km & 31u)) & 1u)) {
                        u32 h = (km * 2654435761u) >> 20;
                        while (hkey[h] != km && hkey[h] != CG_NONE32) h = (h + 1) & 4095u;
                        if (hkey[h] == km) {
                            const u32 sl = hval[h], cell = r * C + sl;
                            pos[cell] = (u16)(16 * wi + b + 1);
                            if (pair_bits && ((atomicOr(&pairs[cell >> 5], 1u << (cell & 31u)) >> (cell & 31u)) & 1u)) filled[sl] = 1;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- alive <=> every occurrence is in a different read; anchors = alive slots in template order
    if (!pair_bits) {
        for (u32 r = warp; r < N; r += NWARPS) {
            const u16* prow = pos + (size_t)r * C;
            for (u32 sb = 0; sb < C; sb += 32) {
                const u32 s = sb + lane;
                if (s < C && prow[s] != 0) atomicAdd(&filled[s], 1u);
            }
        }
        __syncthreads();
    }
    flags4 = 0; nloc = 0;
#pragma unroll
    for (u32 q = 0; q < 4; ++q) {
        const u32 s = tid * 4 + q;
        if (s < C && (pair_bits ? filled[s] == 0 : filled[s] == scnt[s])) { flags4 |= 1u << q; ++nloc; }
    }
    u32 A = 0;
    u32 a = cg_block_scan(nloc, misc, &A);
#pragma unroll
    for (u32 q = 0; q < 4; ++q)
        if (flags4 & (1u << q)) c.anchors[slot_base + a++] = (u16)(tid * 4 + q);
    if (tid == 0) {
        CgWin* Wg = &c.win[w];
        Wg->n_solid = nsolid; Wg->n_cand = C; Wg->n_alive = A;
    }
}
