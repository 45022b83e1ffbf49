// k_ingest.cuh — PAF ingest on the device.
//
// Every getNextReadPile (src/alignmentPiles.cpp:22-58) of one PAF text at once:
//   k_paf_count / k_paf_lines   where the lines are (newline positions; 16 bytes per thread, the text is read twice)
//   k_names_build               read names -> store index: open-addressing table keyed by a position-weighted byte sum (a sum, so that
//                               the parser's warp can build it 32 bytes per step), verified by byte compare; a name listed twice
//                               resolves to its last entry (`index[header] =`, src/utils.cpp:186)
//   k_paf_parse                 Overlap(line) (src/Overlap.h:26-60): one warp per line: tabs ranked by ballot + popcount, the two names hashed /
//                               looked up / compared by the whole warp, one lane per remaining column
//   k_paf_heads / k_paf_piles   consecutive lines with the same qName form a pile; an empty line ends one (:29-37)
//   k_paf_select                std::sort(rbegin, rend) by resMatches + the cut to maxSupport (:39-42).  std::sort is not stable and
//                               the reference is a libstdc++ program: the order of overlaps with equal resMatches — hence which of
//                               them survive the cut and in which order they enter a window's pile — is whatever that library's
//                               introsort leaves.  Lane 0 of the pile's warp replays that algorithm move for move on the pile's keys
//                               in shared memory (bits/stl_algo.h: __introsort_loop, threshold 16, depth limit 2·floor(log2 n),
//                               __move_median_to_first, __unguarded_partition, heapsort via __partial_sort, __final_insertion_sort);
//                               the other lanes load the keys and write the kept overlaps out.
//   k_in_scan                   the small exclusive scans in between (tiles, 256-line blocks, piles)
// Byte / integer work: the line kernels run at 2 TB/s, the parser is bound by instruction issue (645 warp instructions per line), the
// per-pile replay by the latency of one lane (profiles/r01_ingest_summary.md).
#pragma once
#include "cg_common.cuh"
#include "k_extract.cuh"      // CgOverlapDev

enum { CG_IN_FLAG_COLUMNS = 1u, CG_IN_FLAG_NUMBER = 2u, CG_IN_FLAG_NAME = 4u };
#define CG_IN_SAME_Q 0xfffffffeu   // CgPafRec::q of a line whose qName is the previous line's (same pile): not looked up again
#define CG_IN_TILE 4096u      // bytes of text per CTA of k_paf_count / k_paf_lines: 256 threads x 16 bytes

struct CgPafRec { u32 q, qlen, res; CgOverlapDev o; };            // 10 x u32; q == CG_NONE32: an empty line

struct CgIngestArgs {
    const char* text; u64 nbytes;                                   // padded with zeros to a multiple of CG_IN_TILE
    u64* tile_cnt; u32 n_tiles;                                     // newlines per tile -> exclusive scan
    u64* nl_pos; u64 n_lines;                                       // position of the i-th newline
    const char* names; const u64* name_off; u32 n_names; u32* slots; u32 slot_mask;
    CgPafRec* rec;                                                  // [n_lines]
    u32* head;                                                      // [n_lines] (heads before the line inside its 256-line block) << 1 | line is the first of a pile
    u64* head_tile;                                                 // [n_lines / 256 + 1] heads per block -> exclusive scan
    u32 n_piles; u32* pile_first; u32* pile_last;
    u64* keep;                                                      // [n_piles + 1] overlaps kept -> exclusive scan = pile_ov_begin
    u32 max_support;
    u64* sort_scratch; u32 smem_cap;                                // piles over smem_cap lines sort in sort_scratch[first ..]
    u32* pile_read; u32* pile_qlen; CgOverlapDev* ov; u32* res;
    u32* ctl;                                                       // [0] flags, [1] longest pile, [2] non-empty lines
};

// 0x80 in every byte of w that equals '\n' (exact: no borrow between bytes)
__device__ __forceinline__ u32 cg_in_nl_mask(u32 w) {
    const u32 x = w ^ 0x0a0a0a0au;
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
}

__global__ void __launch_bounds__(256) k_paf_count(CgIngestArgs A) {
    CG_DYN_SMEM(smem);
    u32* scratch = (u32*)smem;
    const u64 at = (u64)blockIdx.x * CG_IN_TILE + (u64)threadIdx.x * 16u;
    const uint4 v = *(const uint4*)(A.text + at);
    const u32 n = (u32)(__popc(cg_in_nl_mask(v.x)) + __popc(cg_in_nl_mask(v.y)) + __popc(cg_in_nl_mask(v.z)) + __popc(cg_in_nl_mask(v.w)));
    u32 total;
    cg_block_scan(n, scratch, &total);
    if (threadIdx.x == 0) A.tile_cnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k_paf_lines(CgIngestArgs A) {
    CG_DYN_SMEM(smem);
    u32* scratch = (u32*)smem;
    const u64 at = (u64)blockIdx.x * CG_IN_TILE + (u64)threadIdx.x * 16u;
    const uint4 v = *(const uint4*)(A.text + at);
    const u32 m[4] = {cg_in_nl_mask(v.x), cg_in_nl_mask(v.y), cg_in_nl_mask(v.z), cg_in_nl_mask(v.w)};
    const u32 n = (u32)(__popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]));
    u32 total;
    u64 rank = A.tile_cnt[blockIdx.x] + cg_block_scan(n, scratch, &total);
#pragma unroll
    for (u32 j = 0; j < 4; ++j) {
        u32 x = m[j];
        while (x) {
            const u32 bit = (u32)__ffs((int)x) - 1u;                 // bit 7 of byte b -> 8b + 7 (little endian: byte b at address +b)
            A.nl_pos[rank++] = at + 4u * j + (bit >> 3);
            x &= x - 1u;
        }
    }
}

// Name hash: a position-weighted byte sum with a final mix — a sum, so that a warp can build it 32 bytes per step
// (cg_in_hash_warp) and one thread byte by byte (cg_in_hash, table build) with the same result.
__device__ __forceinline__ u32 cg_in_weight(u32 i) { return (i * 0x9E3779B1u + 0x85EBCA6Bu) | 1u; }
__device__ __forceinline__ u32 cg_in_mix(u32 h, u32 n) {
    h += n * 0x27D4EB2Fu;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}
__device__ __forceinline__ u32 cg_in_hash(const char* s, u32 n) {
    u32 h = 0;
    for (u32 i = 0; i < n; ++i) h += ((u32)(u8)s[i] + 1u) * cg_in_weight(i);
    return cg_in_mix(h, n);
}
__device__ __forceinline__ u32 cg_in_hash_warp(const char* s, u32 n, u32 lane) {
    u32 h = 0;
    for (u32 base = 0; base < n; base += 32u) {
        const u32 i = base + lane;
        if (i < n) h += ((u32)(u8)s[i] + 1u) * cg_in_weight(i);
    }
    return cg_in_mix(cg_warp_sum(h), n);
}
__device__ __forceinline__ bool cg_in_same(const char* a, const char* b, u32 n) {
    for (u32 i = 0; i < n; ++i) if (a[i] != b[i]) return false;
    return true;
}

// Exclusive in-place scan of n u64 entries (+ total at [n]) by one CTA of 1024 threads: contiguous chunk per thread, block scan
// of the chunk sums by warp shuffles.  For the few-thousand-entry arrays of this file (tiles, 256-line blocks, piles).
__global__ void __launch_bounds__(1024) k_in_scan(u64* a, u32 n) {
    CG_DYN_SMEM(smem);
    u64* part = (u64*)smem;                                         // 33 entries
    const u32 t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const u32 per = (n + 1023u) / 1024u;
    const u32 b = t * per < n ? t * per : n, e = b + per < n ? b + per : n;
    u64 s = 0;
    for (u32 i = b; i < e; ++i) s += a[i];
    u64 inc = s;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) { const u64 o = __shfl_up_sync(CG_FULL, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) part[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const u64 x = part[lane];
        u64 xi = x;
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) { const u64 o = __shfl_up_sync(CG_FULL, xi, d); if (lane >= d) xi += o; }
        part[lane] = xi - x;
        if (lane == 31) part[32] = xi;
    }
    __syncthreads();
    u64 run = part[warp] + inc - s;
    for (u32 i = b; i < e; ++i) { const u64 v = a[i]; a[i] = run; run += v; }
    if (t == 0) a[n] = part[32];
}

// slots[] = store index + 1 (0 = free).  Two entries with the same name share a slot, which keeps the larger index.
__global__ void k_names_build(CgIngestArgs A) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_names) return;
    const char* s = A.names + A.name_off[i];
    const u32 n = (u32)(A.name_off[i + 1] - A.name_off[i]);
    u32 slot = cg_in_hash(s, n) & A.slot_mask;
    for (;;) {
        const u32 old = atomicCAS(&A.slots[slot], 0u, i + 1u);
        if (old == 0u) return;
        const u32 j = old - 1u;
        if ((u32)(A.name_off[j + 1] - A.name_off[j]) == n && cg_in_same(A.names + A.name_off[j], s, n)) { atomicMax(&A.slots[slot], i + 1u); return; }
        slot = (slot + 1u) & A.slot_mask;
    }
}

// The whole warp looks one name up: hash and byte compare 32 bytes per step, the probe sequence is uniform.
__device__ __forceinline__ u32 cg_in_lookup_warp(const CgIngestArgs& A, const char* s, u32 n, u32 lane) {
    u32 slot = cg_in_hash_warp(s, n, lane) & A.slot_mask;
    for (;;) {
        const u32 v = A.slots[slot];
        if (v == 0u) return CG_NONE32;
        const u32 j = v - 1u;
        const u64 o = A.name_off[j];
        if ((u32)(A.name_off[j + 1] - o) == n) {
            const char* c = A.names + o;
            u32 diff = 0;
            for (u32 base = 0; base < n && !diff; base += 32u) {
                const u32 i = base + lane;
                diff = __ballot_sync(CG_FULL, i < n && c[i] != s[i]);
            }
            if (!diff) return j;
        }
        slot = (slot + 1u) & A.slot_mask;
    }
}

// One warp per line.  Columns (src/Overlap.h:30-58): 0 qName, 1 qLength, 2 qStart, 3 qEnd (+1), 4 strand, 5 tName, 6 tLength,
// 7 tStart, 8 tEnd (+1), 9 resMatches, 10 alBlockLen, 11 mapQual; anything after the 12th column is ignored.  Numbers are what
// stoi() takes from a PAF: digits first, whatever follows them ignored, at most INT_MAX.
__global__ void __launch_bounds__(256) k_paf_parse(CgIngestArgs A) {
    const u32 lane = threadIdx.x & 31u;
    const u64 i = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= A.n_lines) return;
    const u64 b = i ? A.nl_pos[i - 1] + 1u : 0u;
    const u32 len = (u32)(A.nl_pos[i] - b);
    const char* line = A.text + b;
    u32* out = (u32*)(A.rec + i);
    if (len == 0u) { if (lane == 0) out[0] = CG_NONE32; return; }
    // the first 12 tabs: every lane holding one knows its rank (ballot + popc) and posts its position for the two columns it bounds
    CG_DYN_SMEM(smem);
    u32* tab = (u32*)smem + (threadIdx.x >> 5) * 12u;
    u32 ntabs = 0;
    for (u32 base = 0; base < len && ntabs < 12u; base += 32u) {
        const u32 p = base + lane;
        const bool is_tab = p < len && line[p] == '\t';
        const u32 m = __ballot_sync(CG_FULL, is_tab);
        const u32 t = ntabs + (u32)__popc(m & ((1u << lane) - 1u));
        if (is_tab && t < 12u) tab[t] = p;
        ntabs += (u32)__popc(m);
    }
    __syncwarp();
    if (ntabs > 12u) ntabs = 12u;
    if (ntabs < 11u) { if (lane == 0) { atomicOr(A.ctl, (u32)CG_IN_FLAG_COLUMNS); out[0] = CG_NONE32; } return; }
    // the two names, by the whole warp
    const u32 t4 = tab[4] + 1u;
    // Consecutive lines of a pile share their qName (a 150-deep pile: 150 lines, one name): a line that repeats the previous line's name
    // — the same bytes up to and including the tab — skips the hash, the probe and the compare against the table and is marked instead;
    // the pile takes its read from its first line.  (An empty line in between ends the pile, src/alignmentPiles.cpp:29-37: it has no name.)
    bool same_q = false;
    if (i > 0) {
        const u64 pb = i > 1 ? A.nl_pos[i - 2] + 1u : 0u;
        const u32 plen = (u32)(A.nl_pos[i - 1] - pb), n0 = tab[0];
        if (plen > n0) {
            const char* pl = A.text + pb;
            u32 diff = 0;
            for (u32 base = 0; base <= n0 && !diff; base += 32u) {
                const u32 x = base + lane;
                diff = __ballot_sync(CG_FULL, x <= n0 && pl[x] != line[x]);
            }
            same_q = diff == 0u;
        }
    }
    const u32 qid = same_q ? CG_IN_SAME_Q : cg_in_lookup_warp(A, line, tab[0], lane);
    const u32 tid = cg_in_lookup_warp(A, line + t4, tab[5] - t4, lane);
    // the other columns: lane f owns column f = [st, en)
    u32 st = 0, en = len;
    if (lane >= 1u && lane <= ntabs) st = tab[lane - 1u] + 1u;
    if (lane < ntabs) en = tab[lane];
    const char* s = line + st;
    const u32 n = en - st;
    u32 v = 0, bad = 0;
    if (lane == 0 || lane == 5) {
        v = lane == 0 ? qid : tid;
        if (v == CG_NONE32) bad = CG_IN_FLAG_NAME;
    } else if (lane == 4) {
        v = (n == 1u && s[0] == '+') ? 0u : 1u;
    } else if (lane < 12u) {
        u32 nd = 0;
        for (; nd < n && nd < 11u && s[nd] >= '0' && s[nd] <= '9'; ++nd) v = v * 10u + (u32)(s[nd] - '0');
        if (nd == 0u || nd > 10u) bad = CG_IN_FLAG_NUMBER;
        else if (nd == 10u) {                                       // may exceed INT_MAX (or 32 bits): decide in 64
            u64 x = 0;
            for (u32 j = 0; j < 10u; ++j) x = x * 10u + (u64)(s[j] - '0');
            if (x > 0x7fffffffull) bad = CG_IN_FLAG_NUMBER;
        }
        if (lane == 3 || lane == 8) v -= 1u;                       // "Has to be -1" (Overlap.h:37,48); 0 wraps like the reference's unsigned
    }
    // CgPafRec slots: q, qlen, res, t_read, strand, q_start, q_end, t_start, t_end, t_length
    // (one nibble per lane: 0 1 5 6 4 3 9 7 8 2, then none)
    const u32 nib = lane < 16u ? (u32)(0xFFFFFF2879346510ull >> (4u * lane)) & 15u : 15u;
    const u32 slot = nib == 15u ? CG_NONE32 : nib;
    const u32 anybad = __ballot_sync(CG_FULL, bad != 0u);
    if (anybad) {
        if (bad) atomicOr(A.ctl, bad);
        if (lane == 0) out[0] = CG_NONE32;
        return;
    }
    if (slot != CG_NONE32) out[slot] = v;
}

// Pile index of a line = heads before it: a block scan here, one small scan over the block totals, summed up by the readers.
__global__ void __launch_bounds__(256) k_paf_heads(CgIngestArgs A) {
    CG_DYN_SMEM(smem);
    u32* scratch = (u32*)smem;
    const u64 i = (u64)blockIdx.x * 256u + threadIdx.x;
    u32 hd = 0;
    if (i < A.n_lines) {
        const u32 q = A.rec[i].q;
        if (q != CG_NONE32 && q != CG_IN_SAME_Q) { hd = 1; if (i) { const u32 pq = A.rec[i - 1].q; if (pq == q) hd = 0; } }
    }
    u32 total;
    const u32 before = cg_block_scan(hd, scratch, &total);
    if (i < A.n_lines) A.head[i] = (before << 1) | hd;
    if (threadIdx.x == 0) A.head_tile[blockIdx.x] = total;
}

// after the scan of head_tile[]: first / last line of every pile
__global__ void k_paf_piles(CgIngestArgs A) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_lines) return;
    if (A.rec[i].q == CG_NONE32) return;
    const u32 h = A.head[i];
    const bool is_head = (h & 1u) != 0u;
    const u32 before = (u32)A.head_tile[i >> 8] + (h >> 1);
    const u32 pid = is_head ? before : before - 1u;
    if (is_head) A.pile_first[pid] = (u32)i;
    const bool is_last = i + 1 == A.n_lines || A.rec[i + 1].q == CG_NONE32 || (A.head[i + 1] & 1u) != 0u;
    if (is_last) A.pile_last[pid] = (u32)i;
}

__global__ void k_paf_sizes(CgIngestArgs A) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.n_piles) return;
    const u32 n = A.pile_last[p] - A.pile_first[p] + 1u;
    A.keep[p] = n < A.max_support ? n : A.max_support;
    atomicMax(A.ctl + 1, n);
    atomicAdd(A.ctl + 2, n);                                        // non-empty lines
}

// ---- libstdc++'s std::sort on a[0..n), elements = key << 32 | payload, compared by key only --------------------------------
#define CG_IN_LT(x, y) ((u32)((x) >> 32) < (u32)((y) >> 32))
template <class P> __device__ __forceinline__ void cg_in_swap(P a, u32 i, u32 j) { const u64 t = a[i]; a[i] = a[j]; a[j] = t; }

template <class P> __device__ void cg_in_adjust_heap(P a, u32 first, i32 hole, i32 len, u64 value) {
    const i32 top = hole;
    i32 child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (CG_IN_LT(a[first + child], a[first + child - 1])) child--;
        a[first + hole] = a[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[first + hole] = a[first + child - 1];
        hole = child - 1;
    }
    i32 parent = (hole - 1) / 2;                                   // __push_heap
    while (hole > top && CG_IN_LT(a[first + parent], value)) { a[first + hole] = a[first + parent]; hole = parent; parent = (hole - 1) / 2; }
    a[first + hole] = value;
}
template <class P> __device__ void cg_in_heap_sort(P a, u32 first, i32 len) {
    if (len >= 2) {
        i32 parent = (len - 2) / 2;
        for (;;) { const u64 v = a[first + parent]; cg_in_adjust_heap(a, first, parent, len, v); if (parent == 0) break; parent--; }
    }
    while (len > 1) { --len; const u64 v = a[first + len]; a[first + len] = a[first]; cg_in_adjust_heap(a, first, 0, len, v); }
}
template <class P> __device__ __forceinline__ void cg_in_linear_insert(P a, u32 last) {
    const u64 v = a[last];
    u32 next = last - 1u;
    while (CG_IN_LT(v, a[next])) { a[last] = a[next]; last = next; --next; }
    a[last] = v;
}
template <class P> __device__ void cg_in_insertion_sort(P a, u32 first, u32 last) {
    if (first == last) return;
    for (u32 i = first + 1u; i != last; ++i) {
        if (CG_IN_LT(a[i], a[first])) { const u64 v = a[i]; for (u32 j = i; j > first; --j) a[j] = a[j - 1u]; a[first] = v; }
        else cg_in_linear_insert(a, i);
    }
}
template <class P> __device__ void cg_in_std_sort(P a, u32 n) {
    if (n == 0u) return;
    u32 sf[68], sl[68]; i32 sd[68];                                // the recursion of __introsort_loop: right part first
    i32 sp = 0;
    sf[0] = 0; sl[0] = n; sd[0] = 2 * (31 - __clz((int)n));
    ++sp;
    while (sp > 0) {
        --sp;
        u32 first = sf[sp], last = sl[sp];
        i32 depth = sd[sp];
        while (last - first > 16u) {
            if (depth == 0) { cg_in_heap_sort(a, first, (i32)(last - first)); break; }
            --depth;
            const u32 mid = first + (last - first) / 2u, x = first + 1u, c = last - 1u;
            if (CG_IN_LT(a[x], a[mid])) {                          // __move_median_to_first(first, first + 1, mid, last - 1)
                if (CG_IN_LT(a[mid], a[c])) cg_in_swap(a, first, mid); else if (CG_IN_LT(a[x], a[c])) cg_in_swap(a, first, c); else cg_in_swap(a, first, x);
            } else if (CG_IN_LT(a[x], a[c])) cg_in_swap(a, first, x);
            else if (CG_IN_LT(a[mid], a[c])) cg_in_swap(a, first, c);
            else cg_in_swap(a, first, mid);
            u32 lo = first + 1u, hi = last;                         // __unguarded_partition(first + 1, last, first)
            const u64 pivot = a[first];
            for (;;) {
                while (CG_IN_LT(a[lo], pivot)) ++lo;
                --hi;
                while (CG_IN_LT(pivot, a[hi])) --hi;
                if (!(lo < hi)) break;
                cg_in_swap(a, lo, hi);
                ++lo;
            }
            sf[sp] = first; sl[sp] = lo; sd[sp] = depth; ++sp;      // the caller continues with [first, cut) ...
            first = lo;                                             // ... after the call on [cut, last)
        }
    }
    if (n > 16u) {
        cg_in_insertion_sort(a, 0u, 16u);
        for (u32 i = 16u; i != n; ++i) cg_in_linear_insert(a, i);
    } else cg_in_insertion_sort(a, 0u, n);
}

// One warp (= one CTA) per pile.
__global__ void __launch_bounds__(32) k_paf_select(CgIngestArgs A) {
    CG_DYN_SMEM(smem);
    const u32 lane = threadIdx.x;
    for (u32 p = blockIdx.x; p < A.n_piles; p += gridDim.x) {
        const u32 first = A.pile_first[p], n = A.pile_last[p] - first + 1u;
        u64* a = n <= A.smem_cap ? (u64*)smem : A.sort_scratch + first;
        for (u32 j = lane; j < n; j += 32u) a[j] = ((u64)A.rec[first + (n - 1u - j)].res << 32) | (n - 1u - j);   // the reversed range
        __syncwarp();
        if (lane == 0) cg_in_std_sort(a, n);
        __syncwarp();
        const u64 o0 = A.keep[p];
        const u32 keep = (u32)(A.keep[p + 1] - o0);
        for (u32 j = lane; j < keep; j += 32u) {
            const CgPafRec r = A.rec[first + (u32)a[n - 1u - j]];
            A.ov[o0 + j] = r.o;
            A.res[o0 + j] = r.res;
            if (j == 0) { A.pile_read[p] = A.rec[first].q; A.pile_qlen[p] = r.qlen; }      // alignments.begin(); the pile's read: its first line's (k_paf_parse)
        }
        __syncwarp();
    }
}
