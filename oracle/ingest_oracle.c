/*
 * ingest_oracle.c — TEST INFRASTRUCTURE.  CPU oracle for PAF ingest and the post-filters (rank 4).
 *
 * A plain-C restatement of
 *     Overlap(std::string)              src/Overlap.h:26-60        one PAF line -> one record
 *     getNextReadPile                   src/alignmentPiles.cpp:22-58  grouping, std::sort(rbegin, rend), cut to maxSupport
 *     Overlap::operator<                src/Overlap.h:90-96        resMatches only
 *     trimRead / dropRead / nbCorBases  src/utils.cpp:60-73,96-128
 * written from the reference's behaviour (no reference source is copied).  It only CHECKS the CUDA path
 * (consent_b200/csrc/k_ingest.cuh).
 *
 * std::sort is not stable, so which of several overlaps with equal resMatches survive the cut — and in which order they
 * enter the piles — depends on the sort algorithm.  The reference is built with libstdc++; `lsort` below restates that
 * library's std::sort (GCC bits/stl_algo.h + bits/stl_heap.h: __introsort_loop with threshold 16 and depth limit
 * 2·floor(log2 n), __move_median_to_first, __unguarded_partition, heapsort through __partial_sort, __final_insertion_sort)
 * as a sequence of comparisons and moves on an array; the reverse iterators become a reversed copy.
 *
 * PARITY PINNING: pinned against the UNMODIFIED reference (oracle/_ref: ref_ingest_paf drives the reference's own
 * getNextReadPile over the same text; ref_sort_desc calls std::sort(rbegin, rend) on Overlap records; ref_finish_reads
 * calls trimRead / dropRead) on seeded PAF texts, adversarial key sequences that reach the heapsort branch, and the PAF
 * of the shipped example (tests/test_ingest.py).
 *
 * What the reference does not survive is rejected instead of imitated: a text that does not end with a newline (its
 * getline loop re-reads the last line forever when that line continues a pile, alignmentPiles.cpp:29-35), lines with
 * fewer than 12 columns or a numeric column stoi() throws on, names outside the read index (operator[] then yields an empty
 * read), and — for trimRead — a read without any upper-case base (`unsigned i; while (i >= 0 ...)` walks off the string).
 */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "consent_oracle.h"

/* ---- libstdc++ std::sort, restated.  Elements are u64 = key << 32 | payload; only the key is compared. ---------------- */
typedef uint64_t el;
static int adv_on;                                               /* oracle_sort_adversary: comparisons answered by an adversary */
static int adv_lt(el a, el b);
static long heap_sorts;                                          /* times the depth limit was reached (tests want to see it) */
#define LT(a, b) (adv_on ? adv_lt((a), (b)) : (((a) >> 32) < ((b) >> 32)))

static void swp(el* a, el* b) { el t = *a; *a = *b; *b = t; }

static void push_heap_(el* first, long hole, long top, el value) {
    long parent = (hole - 1) / 2;
    while (hole > top && LT(first[parent], value)) { first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2; }
    first[hole] = value;
}
static void adjust_heap_(el* first, long hole, long len, el value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (LT(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_(first, hole, top, value);
}
static void heap_sort_(el* first, long len) {                    /* __partial_sort(first, last, last) */
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) { el v = first[parent]; adjust_heap_(first, parent, len, v); if (parent == 0) break; parent--; }
    }
    while (len > 1) { --len; el v = first[len]; first[len] = first[0]; adjust_heap_(first, 0, len, v); }
}
static void median_to_first_(el* result, el* a, el* b, el* c) {
    if (LT(*a, *b)) {
        if (LT(*b, *c)) swp(result, b); else if (LT(*a, *c)) swp(result, c); else swp(result, a);
    } else if (LT(*a, *c)) swp(result, a);
    else if (LT(*b, *c)) swp(result, c);
    else swp(result, b);
}
static el* partition_(el* first, el* last, el* pivot) {
    for (;;) {
        while (LT(*first, *pivot)) ++first;
        --last;
        while (LT(*pivot, *last)) --last;
        if (!(first < last)) return first;
        swp(first, last);
        ++first;
    }
}
static void introsort_loop_(el* first, el* last, long depth) {
    while (last - first > 16) {
        if (depth == 0) { ++heap_sorts; heap_sort_(first, last - first); return; }
        --depth;
        el* mid = first + (last - first) / 2;
        median_to_first_(first, first + 1, mid, last - 1);
        el* cut = partition_(first + 1, last, first);
        introsort_loop_(cut, last, depth);
        last = cut;
    }
}
static void linear_insert_(el* last) {
    el v = *last;
    el* next = last - 1;
    while (LT(v, *next)) { *last = *next; last = next; --next; }
    *last = v;
}
static void insertion_sort_(el* first, el* last) {
    if (first == last) return;
    for (el* i = first + 1; i != last; ++i) {
        if (LT(*i, *first)) { el v = *i; memmove(first + 1, first, (size_t)(i - first) * sizeof(el)); *first = v; }
        else linear_insert_(i);
    }
}
static void lsort(el* a, long n);
static void lsort(el* a, long n) {
    if (n == 0) return;
    long lg = 0;
    for (long m = n; m > 1; m >>= 1) ++lg;
    introsort_loop_(a, a + n, 2 * lg);
    if (n > 16) {
        insertion_sort_(a, a + 16);
        for (el* i = a + 16; i != a + n; ++i) linear_insert_(i);
    } else insertion_sort_(a, a + n);
}

/* std::sort(v.rbegin(), v.rend()) on records whose operator< compares `keys`: order[i] = index of the record that ends up
 * at position i. */
void oracle_sort_desc(const uint32_t* keys, uint32_t n, uint32_t* order) {
    el* a = (el*)malloc(((size_t)n + 1) * sizeof(el));
    for (uint32_t j = 0; j < n; ++j) a[j] = ((el)keys[n - 1 - j] << 32) | (n - 1 - j);      /* the reversed range */
    lsort(a, (long)n);
    for (uint32_t i = 0; i < n; ++i) order[i] = (uint32_t)a[n - 1 - i];
    free(a);
}

/* An input on which the sort above exhausts its depth limit and falls back to heapsort (McIlroy's adversary, "A killer
 * adversary for quicksort", 1999, played against lsort): keys[0..n) for oracle_sort_desc / ref_sort_desc.  Returns how many
 * times heapsort was entered while the adversary played. */
static uint32_t* adv_val; static uint32_t adv_gas, adv_solid, adv_cand;
static int adv_lt(el a, el b) {
    uint32_t x = (uint32_t)a, y = (uint32_t)b;
    if (adv_val[x] == adv_gas && adv_val[y] == adv_gas) { if (x == adv_cand) adv_val[x] = adv_solid++; else adv_val[y] = adv_solid++; }
    if (adv_val[x] == adv_gas) adv_cand = x; else if (adv_val[y] == adv_gas) adv_cand = y;
    return adv_val[x] < adv_val[y];
}
long oracle_sort_adversary(uint32_t n, uint32_t* keys) {
    el* a = (el*)malloc(((size_t)n + 1) * sizeof(el));
    adv_val = (uint32_t*)malloc(((size_t)n + 1) * 4);
    adv_gas = n ? n - 1 : 0; adv_solid = 0; adv_cand = 0;
    for (uint32_t j = 0; j < n; ++j) { a[j] = j; adv_val[j] = adv_gas; }
    heap_sorts = 0;
    adv_on = 1; lsort(a, (long)n); adv_on = 0;
    const long hs = heap_sorts;
    for (uint32_t j = 0; j < n; ++j) keys[n - 1 - j] = adv_val[j];              /* position j of the reversed range */
    free(a); free(adv_val);
    return hs;
}
long oracle_sort_heap_sorts(void) { return heap_sorts; }

/* ---- names -> store index --------------------------------------------------------------------------------------------- */
typedef struct { const cg_read_names* t; } name_ctx;
static const cg_read_names* g_names;
static int name_cmp_ids(const void* x, const void* y) {
    uint32_t a = *(const uint32_t*)x, b = *(const uint32_t*)y;
    size_t la = (size_t)(g_names->name_off[a + 1] - g_names->name_off[a]), lb = (size_t)(g_names->name_off[b + 1] - g_names->name_off[b]);
    int c = memcmp(g_names->names + g_names->name_off[a], g_names->names + g_names->name_off[b], la < lb ? la : lb);
    if (c) return c;
    if (la != lb) return la < lb ? -1 : 1;
    return a < b ? -1 : a > b;
}
static long name_find(const cg_read_names* t, const uint32_t* sorted, const char* s, size_t n) {
    long lo = 0, hi = (long)t->n_reads - 1, hit = -1;
    while (lo <= hi) {                                            /* last entry among equal names: `index[header] =` overwrites */
        long mid = (lo + hi) / 2;
        uint32_t id = sorted[mid];
        size_t l = (size_t)(t->name_off[id + 1] - t->name_off[id]);
        int c = memcmp(t->names + t->name_off[id], s, l < n ? l : n);
        if (c == 0 && l != n) c = l < n ? -1 : 1;
        if (c == 0) { hit = mid; lo = mid + 1; }
        else if (c < 0) lo = mid + 1;
        else hi = mid - 1;
    }
    return hit < 0 ? -1 : (long)sorted[hit];
}

/* stoi() restricted to what a PAF holds: digits first, anything after them ignored, value <= INT_MAX */
static int parse_int(const char* s, size_t n, uint32_t* out) {
    if (n == 0 || s[0] < '0' || s[0] > '9') return 0;
    uint64_t v = 0;
    for (size_t i = 0; i < n && s[i] >= '0' && s[i] <= '9'; ++i) { v = v * 10 + (uint64_t)(s[i] - '0'); if (v > (uint64_t)INT_MAX) return 0; }
    *out = (uint32_t)v;
    return 1;
}

typedef struct { uint32_t q, qlen, res; cg_overlap o; } rec;
typedef struct { uint32_t *pile_read, *pile_qlen, *ovb, *res; cg_overlap* ov; } pile_owner;

int oracle_ingest_paf(const char* paf, uint64_t nbytes, const cg_read_names* names, uint32_t max_support, cg_pile_set* out) {
    if (!out || !names || (nbytes && !paf) || max_support == 0) return CG_ERR_INVALID_ARG;
    if (nbytes && paf[nbytes - 1] != '\n') return CG_ERR_INVALID_ARG;
    uint32_t* sorted = (uint32_t*)malloc(((size_t)names->n_reads + 1) * 4);
    for (uint32_t i = 0; i < names->n_reads; ++i) sorted[i] = i;
    g_names = names;
    qsort(sorted, names->n_reads, 4, name_cmp_ids);
    uint64_t n_lines = 0;
    for (uint64_t i = 0; i < nbytes; ++i) n_lines += paf[i] == '\n';
    rec* R = (rec*)malloc((size_t)(n_lines + 1) * sizeof(rec));
    uint8_t* brk = (uint8_t*)calloc((size_t)n_lines + 1, 1);     /* an empty line precedes record i */
    uint64_t nr = 0;
    int rc = CG_OK, pending_break = 1;
    for (uint64_t p = 0; p < nbytes && rc == CG_OK;) {
        uint64_t e = p;
        while (paf[e] != '\n') ++e;
        if (e == p) { pending_break = 1; p = e + 1; continue; }
        uint64_t fs[13];
        int nf = 0;
        fs[0] = p;
        for (uint64_t i = p; i < e && nf < 12; ++i) if (paf[i] == '\t') fs[++nf] = i + 1;
        if (nf < 12) { if (nf < 11) { rc = CG_ERR_INVALID_ARG; break; } fs[12] = e + 1; }   /* column 12 may end the line */
#define FLD(k) (paf + fs[k]), (size_t)(fs[(k) + 1] - 1 - fs[k])
        rec r;
        uint32_t v[8], mapq;
        long q = name_find(names, sorted, FLD(0)), t = name_find(names, sorted, FLD(5));
        int ok = q >= 0 && t >= 0 && parse_int(FLD(1), &v[0]) && parse_int(FLD(2), &v[1]) && parse_int(FLD(3), &v[2]) && parse_int(FLD(6), &v[3]) &&
                 parse_int(FLD(7), &v[4]) && parse_int(FLD(8), &v[5]) && parse_int(FLD(9), &v[6]) && parse_int(FLD(10), &v[7]) && parse_int(FLD(11), &mapq);
        if (!ok) { rc = CG_ERR_INVALID_ARG; break; }
        r.q = (uint32_t)q; r.qlen = v[0]; r.res = v[6];
        r.o.t_read = (uint32_t)t;
        r.o.strand = !(fs[5] - 1 - fs[4] == 1 && paf[fs[4]] == '+');
        r.o.q_start = v[1]; r.o.q_end = v[2] - 1u; r.o.t_length = v[3]; r.o.t_start = v[4]; r.o.t_end = v[5] - 1u;
#undef FLD
        brk[nr] = (uint8_t)pending_break; pending_break = 0;
        R[nr++] = r;
        p = e + 1;
    }
    free(sorted);
    pile_owner* ow = (pile_owner*)calloc(1, sizeof *ow);
    ow->pile_read = (uint32_t*)malloc((size_t)(nr + 1) * 4); ow->pile_qlen = (uint32_t*)malloc((size_t)(nr + 1) * 4);
    ow->ovb = (uint32_t*)malloc((size_t)(nr + 2) * 4); ow->res = (uint32_t*)malloc((size_t)(nr + 1) * 4);
    ow->ov = (cg_overlap*)malloc((size_t)(nr + 1) * sizeof(cg_overlap));
    uint32_t np = 0, no = 0;
    ow->ovb[0] = 0;
    uint32_t* keys = (uint32_t*)malloc((size_t)(nr + 1) * 4);
    uint32_t* order = (uint32_t*)malloc((size_t)(nr + 1) * 4);
    for (uint64_t a = 0; a < nr && rc == CG_OK;) {
        uint64_t b = a + 1;
        while (b < nr && !brk[b] && R[b].q == R[a].q) ++b;
        const uint32_t n = (uint32_t)(b - a);
        for (uint32_t i = 0; i < n; ++i) keys[i] = R[a + i].res;
        oracle_sort_desc(keys, n, order);
        const uint32_t keep = n > max_support ? max_support : n;
        ow->pile_read[np] = R[a].q;
        ow->pile_qlen[np] = R[a + order[0]].qlen;                 /* alignments.begin()->qLength */
        for (uint32_t i = 0; i < keep; ++i) { ow->ov[no] = R[a + order[i]].o; ow->res[no] = R[a + order[i]].res; ++no; }
        ow->ovb[++np] = no;
        a = b;
    }
    free(keys); free(order); free(R); free(brk);
    out->n_piles = np; out->pile_read = ow->pile_read; out->pile_qlen = ow->pile_qlen; out->pile_ov_begin = ow->ovb;
    out->overlaps = ow->ov; out->res_matches = ow->res; out->n_lines = nr; out->owner_ = ow;
    if (rc != CG_OK) { oracle_free_pile_set(out); memset(out, 0, sizeof *out); }
    return rc;
}

void oracle_free_pile_set(cg_pile_set* s) {
    if (!s || !s->owner_) return;
    pile_owner* ow = (pile_owner*)s->owner_;
    free(ow->pile_read); free(ow->pile_qlen); free(ow->ovb); free(ow->res); free(ow->ov); free(ow);
    s->owner_ = NULL;
}

/* ---- trimRead (utils.cpp:96-128) + dropRead (:60-73) on the output of alignConsensus --------------------------------- */
static int up(char c) { return 'A' <= c && c <= 'Z'; }

typedef struct { uint64_t* off; char* bases; } cor_owner;

int oracle_finish_reads(const cg_corrected* in, uint32_t trim_mer, cg_corrected* out) {
    if (!in || !out) return CG_ERR_INVALID_ARG;
    const uint32_t R = in->n_reads;
    cor_owner* ow = (cor_owner*)calloc(1, sizeof *ow);
    ow->off = (uint64_t*)calloc((size_t)R + 1, 8);
    ow->bases = (char*)malloc((size_t)(R ? in->read_off[R] : 0) + 1);
    uint64_t w = 0;
    for (uint32_t r = 0; r < R; ++r) {
        const char* s = in->bases + in->read_off[r];
        const uint64_t len = in->read_off[r + 1] - in->read_off[r];
        uint64_t b0 = 0, n_out = len;
        if (trim_mer && len) {
            uint64_t i = 0, n = 0;
            while (i < len && n < trim_mer) { if (up(s[i])) n++; else n = 0; i++; }
            const uint64_t beg = i - trim_mer;                    /* wraps like the reference's unsigned when i < merSize */
            int found = n >= trim_mer;
            uint64_t j = len, m = 0;                               /* j = i + 1 of the reference's descending scan */
            while (j > 0 && m < trim_mer) { if (up(s[j - 1])) m++; else m = 0; j--; }
            if (!found || m < trim_mer) { b0 = 0; n_out = 0; }    /* no run at all: the reference indexes s[-1] */
            else {
                const uint64_t end = (j - 1) + trim_mer;          /* i + merSize with i = j - 1 */
                if ((uint32_t)end > (uint32_t)beg) { b0 = beg; n_out = end - beg + 1; } else n_out = 0;
            }
            if (n_out) {                                           /* dropRead on the trimmed string */
                int nb = 0;
                for (uint64_t x = 0; x < n_out; ++x) nb += up(s[b0 + x]);
                if ((float)nb / (float)n_out < 0.1) n_out = 0;
            }
        }
        if (n_out) memcpy(ow->bases + w, s + b0, n_out);
        w += n_out;
        ow->off[r + 1] = w;
    }
    ow->bases[w] = 0;
    out->n_reads = R; out->read_off = ow->off; out->bases = ow->bases; out->owner_ = ow;
    return CG_OK;
}

void oracle_free_finished(cg_corrected* c) {
    if (!c || !c->owner_) return;
    cor_owner* ow = (cor_owner*)c->owner_;
    free(ow->off); free(ow->bases); free(ow);
    c->owner_ = NULL;
}
