// k_split.cuh — distance statistics between consecutive chain anchors, region classification, POA job list.
//
// Replaces deciles / comparable / average_distance_next_anchor (BMEAN/bmean.cpp:264-295, 324-418),
// get_position + split_reads (:422-430, 476-554) and the trivial cases of easy_consensus (:603-641).
//
// The reference materialises every region as a vector of strings.  Here a region is never copied: a kept
// segment is (read, start, len) recomputed from the position table by cg_eval_segment, so this kernel only
// *classifies* regions (empty / one distinct string -> copy / needs POA), sizes their consensus slot and
// appends POA jobs to the chunk-wide queue that k_poa's persistent warps drain.
//
// The two order statistics of the reference's sorted distance vector (V[floor((n-1)*0.2)], V[ceil((n-1)*0.8)])
// are found by a warp radix select over the bits of the largest distance; nothing is sorted.
#pragma once
#include "cg_common.cuh"

#define CG_SPLIT_THREADS 256u
#define CG_SPLIT_WARPS (CG_SPLIT_THREADS / 32u)
// POA job routing.  The final graph size of a region is predicted from its longest segment L and its depth n
// (fit on PacBio-profile piles: V ~ 1.52 L + 0.022 L n - 7, +12 % margin); a wrong guess only costs a re-queue.
// class 0/1: tier C1 (heavy / light), 2/3: tier G (heavy / light), 4/5: wide tiers (heavy / light).  Capacities mirror k_poa2.cuh.
#define CG_POA_C1_LCAP 32u
#define CG_POA_C1_VCAP 128u
#define CG_POA_C1_CELLS 2048u
#define CG_POA_G_LCAP 120u
#define CG_POA_G_VCAP 254u
#define CG_POA_C_SEGCAP 192u
#define CG_POA_C1_HEAVY 24000u     // sequences x predicted cells: front of the queue (drained first)
#define CG_POA_G_HEAVY 400000u
#define CG_POA_W_HEAVY 3000000u    // wide tiers: the long jobs first, so that the short ones fill the tail of the launch
#define CG_POA_NCLASS 6u
__device__ __forceinline__ u32 cg_poa_class(u32 n, u32 L) {
    u32 vhat = (L * (1557u + 23u * n)) >> 10;
    vhat = vhat > 8u ? vhat - 7u : 1u;
    vhat += vhat >> 3;
    const u32 cells = (vhat + 1u) * (L + 1u);
    const u32 cost = n * cells;
    if (n > CG_POA_C_SEGCAP || L > CG_POA_G_LCAP || vhat > CG_POA_G_VCAP) return (u64)n * cells >= CG_POA_W_HEAVY ? 4u : 5u;
    if (L <= CG_POA_C1_LCAP && vhat <= CG_POA_C1_VCAP && cells <= CG_POA_C1_CELLS) return cost >= CG_POA_C1_HEAVY ? 0u : 1u;
    return cost >= CG_POA_G_HEAVY ? 2u : 3u;
}

// idx-th smallest (0-based) distance of the pair (s1,s2) over reads holding both.
__device__ __forceinline__ u32 cg_select_distance(const u16* pos, u32 C, u32 N, u32 s1, u32 s2, u32 nbits, u32 idx) {
    const u32 lane = cg_lane();
    u64 prefix = 0;
    u32 rem = idx;
    for (int b = (int)nbits - 1; b >= 0; --b) {
        u32 cnt0 = 0;
        for (u32 rb = 0; rb < N; rb += 32) {
            const u32 r = rb + lane;
            if (r < N) {
                const u32 p1 = pos[(size_t)r * C + s1], p2 = pos[(size_t)r * C + s2];
                if (p1 && p2) {
                    const u64 d = (u32)(p2 - p1);
                    if ((d >> (b + 1)) == (prefix >> (b + 1)) && !((d >> b) & 1ull)) ++cnt0;
                }
            }
        }
        cnt0 = cg_warp_sum(cnt0);
        if (rem >= cnt0) { rem -= cnt0; prefix |= 1ull << b; }
    }
    return (u32)prefix;
}

// The same over distances held in registers: lane l keeps reads l, l + 32, ... (0xffffffff = the read lacks an anchor).
#define CG_SPLIT_REG_READS 8u       // piles of up to 256 sequences take this path
__device__ __forceinline__ u32 cg_select_distance_reg(const u32 (&d)[CG_SPLIT_REG_READS], u32 nbits, u32 idx) {
    u32 prefix = 0, rem = idx;
    for (int b = (int)nbits - 1; b >= 0; --b) {
        u32 cnt0 = 0;
#pragma unroll
        for (u32 t = 0; t < CG_SPLIT_REG_READS; ++t)
            cnt0 += ((d[t] >> (b + 1)) == (prefix >> (b + 1)) && !((d[t] >> b) & 1u)) ? 1u : 0u;
        cnt0 = cg_warp_sum(cnt0);
        if (rem >= cnt0) { rem -= cnt0; prefix |= 1u << b; }
    }
    return prefix;
}

__global__ void __launch_bounds__(CG_SPLIT_THREADS) k_split(CgChunk c) {
    const u32 w = blockIdx.x, lane = cg_lane(), warp = cg_warp();
    const CgWin W = c.win[w];
    if (W.bad) return;                               // n_regions stays 0: no POA jobs, nothing to stitch
    const u32 N = W.n_seqs, C = W.n_cand, nA = W.n_chain;
    const u64 slot_base = c.off_slot[w];
    const u16* chain = c.chain + slot_base;
    const u16* pos = c.pos + c.off_pos[w];
    u32* rel = c.rel + slot_base;
    CgRegion* regs = c.regions + c.off_reg[w];
    const u8* wbases = (const u8*)c.bases;

    // ---- average_distance_next_anchor: one warp per consecutive anchor pair
    for (u32 i = warp; i + 1 < nA; i += CG_SPLIT_WARPS) {
        const u32 s1 = chain[i], s2 = chain[i + 1];
        if (N <= 32u * CG_SPLIT_REG_READS) {            // distances once into registers
            u32 d[CG_SPLIT_REG_READS];
            u32 n = 0, mx = 0;
#pragma unroll
            for (u32 t = 0; t < CG_SPLIT_REG_READS; ++t) {
                const u32 r = 32 * t + lane;
                d[t] = 0xffffffffu;
                if (r < N) {
                    const u32 p1 = pos[(size_t)r * C + s1], p2 = pos[(size_t)r * C + s2];
                    if (p1 && p2) { d[t] = p2 - p1; ++n; mx = d[t] > mx ? d[t] : mx; }
                }
            }
            n = cg_warp_sum(n);
            mx = cg_warp_max(mx);
            if (n == 0) {                                   // unreachable: chain edges share >= max(S,1) reads
                if (lane == 0) { rel[i] = 0; atomicOr(c.flags, (u32)CG_FLAG_INTERNAL); }
                continue;
            }
            const u32 nbits = 32u - (u32)__clz((int)mx);
            const u32 ilo = (u32)floor((double)(n - 1) * 0.2);     // deciles, bmean.cpp:324-327
            const u32 ihi = (u32)ceil((double)(n - 1) * 0.8);
            const u32 lo = cg_select_distance_reg(d, nbits, ilo);
            const u32 hi = cg_select_distance_reg(d, nbits, ihi);
            u32 sum = 0, cnt = 0;
#pragma unroll
            for (u32 t = 0; t < CG_SPLIT_REG_READS; ++t)
                if (d[t] != 0xffffffffu && cg_comparable_dec_u(d[t], lo, hi)) { sum += d[t]; ++cnt; }      // distances are < 2^16
            sum = cg_warp_sum(sum);
            cnt = cg_warp_sum(cnt);
            if (lane == 0) rel[i] = cnt ? sum / cnt : 0u;           // integer division, bmean.cpp:401
            continue;
        }
        u32 n = 0, mx = 0;
        for (u32 rb = 0; rb < N; rb += 32) {
            const u32 r = rb + lane;
            bool has = false;
            if (r < N) {
                const u32 p1 = pos[(size_t)r * C + s1], p2 = pos[(size_t)r * C + s2];
                if (p1 && p2) { has = true; const u32 d = p2 - p1; mx = d > mx ? d : mx; }
            }
            n += __popc(__ballot_sync(CG_FULL, has));
        }
        mx = cg_warp_max(mx);
        if (n == 0) {                                   // unreachable: chain edges share >= max(S,1) reads
            if (lane == 0) { rel[i] = 0; atomicOr(c.flags, (u32)CG_FLAG_INTERNAL); }
            continue;
        }
        const u32 nbits = 32u - (u32)__clz((int)mx);
        const u32 ilo = (u32)floor((double)(n - 1) * 0.2);     // deciles, bmean.cpp:324-327
        const u32 ihi = (u32)ceil((double)(n - 1) * 0.8);
        const double lo = (double)cg_select_distance(pos, C, N, s1, s2, nbits, ilo);
        const double hi = (double)cg_select_distance(pos, C, N, s1, s2, nbits, ihi);
        u32 sum = 0, cnt = 0;
        for (u32 rb = 0; rb < N; rb += 32) {
            const u32 r = rb + lane;
            if (r < N) {
                const u32 p1 = pos[(size_t)r * C + s1], p2 = pos[(size_t)r * C + s2];
                if (p1 && p2) {
                    const u32 d = p2 - p1;
                    if (cg_comparable_dec((double)d, lo, hi)) { sum += d; ++cnt; }
                }
            }
        }
        sum = cg_warp_sum(sum);
        cnt = cg_warp_sum(cnt);
        if (lane == 0) rel[i] = cnt ? sum / cnt : 0u;           // integer division, bmean.cpp:401
    }
    __syncthreads();

    // ---- split_reads + trivial cases of easy_consensus: one warp per region
    u32 nreg = nA ? nA + 1 : 2;
    if (nreg < c.min_anchors) nreg = 0;                         // MSABMAAC bails out, bmean.cpp:796-805
    CgWinView v;
    v.seq_off = c.seq_off + W.seq_begin; v.pos = pos; v.chain = chain; v.rel = rel; v.N = N; v.C = C; v.nA = nA;
    for (u32 g = warp; g < nreg; g += CG_SPLIT_WARPS) {
        u32 n = 0, r0 = 0, st0 = 0, ln0 = 0, sum = 0, mxl = 0;
        bool same = true;
        for (u32 rb = 0; rb < N; rb += 32) {
            const u32 r = rb + lane;
            u32 st = 0, ln = 0;
            const bool keep = r < N && cg_eval_segment(v, g, r, &st, &ln);
            const u32 bal = __ballot_sync(CG_FULL, keep);
            if (bal && n == 0) {
                const int src = __ffs((int)bal) - 1;
                r0 = __shfl_sync(CG_FULL, r, src); st0 = __shfl_sync(CG_FULL, st, src); ln0 = __shfl_sync(CG_FULL, ln, src);
            }
            bool eq = true;
            if (keep) {
                sum += ln;
                mxl = ln > mxl ? ln : mxl;
                eq = ln == ln0;
                if (eq) {
                    const u8* a = wbases + v.seq_off[r] + st;
                    const u8* b = wbases + v.seq_off[r0] + st0;
                    for (u32 t = 0; t < ln; ++t) if (a[t] != b[t]) { eq = false; break; }
                }
            }
            const bool alleq = __all_sync(CG_FULL, eq);
            same = same && alleq;
            n += __popc(bal);
        }
        sum = cg_warp_sum(sum);
        mxl = cg_warp_max(mxl);
        if (lane == 0) {
            CgRegion R;
            R.kind = n == 0 ? CG_REG_EMPTY : (n == 1 || same) ? CG_REG_COPY : CG_REG_POA;
            R.n = n; R.read = r0; R.start = st0; R.len = ln0; R.sum_len = sum; R.max_len = mxl; R.arena_off = 0; R.cons_len = 0;
            regs[g] = R;
        }
    }
    __syncthreads();

    // ---- consensus slots and POA jobs (warp 0).  Jobs are routed by their longest segment: short ones to the
    // shared-memory tier, the rest to the medium tier (either re-queues what it cannot hold); inside a queue the
    // heavy jobs (sequences x longest segment) go to the front, which is drained first.
    if (warp != 0) return;
    u32 cnt[CG_POA_NCLASS];
#pragma unroll
    for (u32 q = 0; q < CG_POA_NCLASS; ++q) cnt[q] = 0;
    for (u32 gb = 0; gb < nreg; gb += 32) {
        const u32 g = gb + lane;
        u32 cls = CG_POA_NCLASS;
        if (g < nreg && regs[g].kind == CG_REG_POA) cls = cg_poa_class(regs[g].n, regs[g].max_len);
#pragma unroll
        for (u32 q = 0; q < CG_POA_NCLASS; ++q) cnt[q] += __popc(__ballot_sync(CG_FULL, cls == q));
    }
    u32 njobs = 0;
#pragma unroll
    for (u32 q = 0; q < CG_POA_NCLASS; ++q) njobs += cnt[q];
    // class -> (queue, end): 0 q0 front, 1 q0 back, 2 q1 front, 3 q1 back, 4 q2 front, 5 q2 back
    u32 base[CG_POA_NCLASS];
#pragma unroll
    for (u32 q = 0; q < CG_POA_NCLASS; ++q) {
        base[q] = 0;
        const u32 ctl = 4u * (q >> 1) + 2u * (q & 1u);
        if (lane == 0 && cnt[q]) base[q] = atomicAdd(&c.qctl[ctl], cnt[q]);
        base[q] = __shfl_sync(CG_FULL, base[q], 0);
    }
    const u32 cap_s = c.qctl[3], cap_m = c.qctl[7], cap_w = c.qctl[11];
    u32 run_off = 0;
    for (u32 gb = 0; gb < nreg; gb += 32) {
        const u32 g = gb + lane;
        const bool isp = g < nreg && regs[g].kind == CG_REG_POA;
        u32 cls = CG_POA_NCLASS;
        if (isp) cls = cg_poa_class(regs[g].n, regs[g].max_len);
        const u32 sz = isp ? regs[g].sum_len : 0u;
        const u32 inc = cg_warp_scan(sz);
        u32 my = 0;
#pragma unroll
        for (u32 q = 0; q < CG_POA_NCLASS; ++q) {
            const u32 bal = __ballot_sync(CG_FULL, cls == q);
            if (cls == q) my = base[q] + __popc(bal & ((1u << lane) - 1u));
            base[q] += __popc(bal);
        }
        if (isp) {
            regs[g].arena_off = run_off + inc - sz;
            const uint2 job = make_uint2(w, g);
            if (cls == 0) c.jobs_s[my] = job;
            else if (cls == 1) c.jobs_s[cap_s - 1 - my] = job;
            else if (cls == 2) c.jobs_m[my] = job;
            else if (cls == 3) c.jobs_m[cap_m - 1 - my] = job;
            else if (cls == 4) c.jobs_w[my] = job;
            else c.jobs_w[cap_w - 1 - my] = job;
        }
        run_off += __shfl_sync(CG_FULL, inc, 31);
    }
    if (lane == 0) {
        c.win[w].n_regions = nreg;
        atomicAdd((unsigned long long*)&c.counters->anchors, (unsigned long long)nA);
        atomicAdd((unsigned long long*)&c.counters->regions, (unsigned long long)(nA ? nA + 1 : 2));
        atomicAdd((unsigned long long*)&c.counters->poa_graphs, (unsigned long long)njobs);
    }
}

// The wide tier's queue in descending order of predicted cost.  Its jobs are long (a warp spends milliseconds on a whole-window
// graph) and few per warp (config-2 shape: 16 700 jobs on 3 552 resident warps), so the launch ends when the unluckiest warp does:
// two classes (heavy first, light last) left a fifth of the launch as tail; longest-first over 512 cost buckets (16 per octave)
// brings it within a few percent of the mean (the order has no influence on any result: jobs are independent).
// One CTA: histogram of the buckets, exclusive scan, scatter.  ctl = the queue's four control words (front count, cursor, back count, capacity).
#define CG_QSORT_THREADS 1024u
#define CG_QSORT_BUCKETS 512u
__global__ void __launch_bounds__(CG_QSORT_THREADS) k_poa_sort_queue(CgChunk c, const uint2* jobs, u32* ctl, uint2* sorted) {
    CG_DYN_SMEM(smem);
    u32* hist = (u32*)smem;                          // CG_QSORT_BUCKETS + 64 words
    u32* scratch = hist + CG_QSORT_BUCKETS;
    const u32 tid = threadIdx.x;
    const u32 nfront = ctl[0], nback = ctl[2], cap = ctl[3], n = nfront + nback;
    if (tid < CG_QSORT_BUCKETS) hist[tid] = 0;
    __syncthreads();
    for (u32 pass = 0; pass < 2; ++pass) {
        for (u32 j = tid; j < n; j += CG_QSORT_THREADS) {
            const uint2 job = j < nfront ? jobs[j] : jobs[cap - 1 - (j - nfront)];
            const CgRegion& R = c.regions[c.off_reg[job.x] + job.y];
            const u32 L = R.max_len, ns = R.n;
            u32 vhat = (L * (1557u + 23u * ns)) >> 10;            // the predicted graph size of cg_poa_class
            vhat = vhat > 8u ? vhat - 7u : 1u;
            vhat += vhat >> 3;
            const float cost = (float)ns * (float)(vhat + 1u) * (float)(L + 1u);
            const u32 key = __float_as_uint(cost) >> 19;          // exponent and four mantissa bits: 16 buckets per octave
            const u32 lo = 127u << 4;                              // cost >= 1
            u32 b = key > lo ? key - lo : 0u;
            b = b < CG_QSORT_BUCKETS ? CG_QSORT_BUCKETS - 1u - b : 0u;                     // heaviest first
            const u32 at = atomicAdd(&hist[b], 1u);
            if (pass) sorted[at] = job;
        }
        __syncthreads();
        if (pass == 0) {
            u32 total = 0;
            const u32 v = tid < CG_QSORT_BUCKETS ? hist[tid] : 0u;
            const u32 ex = cg_block_scan(v, scratch, &total);
            if (tid < CG_QSORT_BUCKETS) hist[tid] = ex;
            __syncthreads();
        }
    }
    if (tid == 0) { ctl[0] = n; ctl[2] = 0; }
}
