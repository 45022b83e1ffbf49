/*
 * reanchor_oracle.c — TEST INFRASTRUCTURE.  CPU oracle for consensus re-anchoring.
 *
 * A plain-C restatement of
 *     alignConsensus                          src/correctionAlignment.cpp:47-139
 *     nbSolidMers / nbUpperCase / getIndels   src/correctionAlignment.cpp:6-45
 * and of what it needs from the vendored SSW library
 *     StripedSmithWaterman::Aligner::Align    BMEAN/Complete-Striped-Smith-Waterman-Library/src/ssw_cpp.cpp:365-403
 *     ssw_align                               .../src/ssw.c:788-891
 *     sw_sse2_byte / sw_sse2_word             .../src/ssw.c:158-366 / 392-575
 *     banded_sw                               .../src/ssw.c:577-758
 * written from the reference's behaviour (no reference source is copied).  It only CHECKS the CUDA path
 * (consent_b200/csrc/k_reanchor.cuh): tests/, smoke() and bench.py's CPU legs may load it, the product never does.
 *
 * PARITY PINNING: the reference has no golden vectors for this path; the oracle is pinned against the UNMODIFIED
 * reference (oracle/_ref: ref_reanchor_reads calls the reference's own alignConsensus, which links the vendored
 * ssw.c / ssw_cpp.cpp) in tests/test_reanchor.py, and tests/golden/reanchor_golden.json holds reference outputs
 * for where /root/reference is absent.
 *
 * The striped SIMD kernels are restated by their results, not their lanes: both sw_sse2_byte and sw_sse2_word
 * compute the exact affine-gap local alignment matrix H (E and F floored at 0; the "lazy F" loops only complete F
 * inside a column, ssw.c:266-295,489-501), so what has to be reproduced is how the end points are picked:
 *   score    = max H
 *   ref_end  = first reference column (in scan order) whose column maximum reaches the final score  (ssw.c:297-313: the
 *              saved column is replaced only on a strictly larger maximum)
 *   read_end = smallest query index holding `score` in that column                                  (ssw.c:323-332)
 *   begin    = the same scan on the reversed query prefix / reference prefix, stopped at the first column whose
 *              maximum equals `score`                                                               (ssw.c:836-850,318)
 * The byte kernel is abandoned for the word kernel when a score reaches 255 - bias (ssw.c:305,806-810); both give the
 * same H, so only the degenerate score-0 case can tell them apart (end_ref starts at -1 in the byte kernel, which is
 * the one whose result is kept when nothing overflows).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "consent_oracle.h"

/* ssw_cpp.cpp:11-28 kBaseTranslation: A/a 0, C/c 1, G/g 2, T/t 3, (U/u 0), everything else 4 */
static int8_t base_code(char c) {
    switch (c) {
        case 'A': case 'a': case 'U': case 'u': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}
/* ssw_cpp.cpp:30-57 BuildSwScoreMatrix with the Aligner defaults (ssw_cpp.cpp:419-426): match 2, mismatch -2, N -2 */
static int sub_score(int8_t a, int8_t b) { return (a < 4 && a == b) ? 2 : -2; }
enum { GAP_O = 3, GAP_E = 1 };

/* how often the rarely taken branches ran since load: [0] overlapping windows, [1] differing overlaps,
 * [2] arbitration won by the previous window (sub-alignment + banded traceback), [3] band doublings,
 * [4] windows emptied by the arbitration */
static uint64_t g_branch[8];
void oracle_reanchor_branches(uint64_t out[8]) { memcpy(out, g_branch, sizeof g_branch); }
#define BRANCH(i) __atomic_fetch_add(&g_branch[i], 1, __ATOMIC_RELAXED)

typedef struct { int score, ref, read; } sw_end;

/* One scan of the local-alignment matrix, reference columns in the order first, first+step, ... (n_ref of them),
 * query rows 0..n_q-1 (already reversed by the caller for the backward scan).  terminate > 0 stops after the first
 * column whose maximum equals it.  ssw.c:158-366 / 392-575 by result (see the header). */
static sw_end sw_scan(const int8_t* q, int n_q, const int8_t* ref, int first, int step, int n_ref, int terminate,
                      uint64_t* cells) {
    int* H = (int*)calloc((size_t)n_q + 1, sizeof(int));
    int* E = (int*)calloc((size_t)n_q + 1, sizeof(int));
    int* best_col = (int*)calloc((size_t)n_q + 1, sizeof(int));
    sw_end r; r.score = 0; r.ref = -1; r.read = n_q - 1;
    for (int c = 0, i = first; c < n_ref; ++c, i += step) {
        int diag = 0, F = 0, colmax = 0;
        for (int j = 0; j < n_q; ++j) {
            int h = diag + sub_score(q[j], ref[i]);
            if (h < E[j]) h = E[j];
            if (h < F) h = F;                      /* E, F >= 0: the floor at 0 is implied */
            diag = H[j]; H[j] = h;
            if (h > colmax) colmax = h;
            int open = h - GAP_O; if (open < 0) open = 0;
            int e = E[j] - GAP_E; if (e < 0) e = 0;
            E[j] = e > open ? e : open;
            int f = F - GAP_E; if (f < 0) f = 0;
            F = f > open ? f : open;
        }
        if (cells) *cells += (uint64_t)n_q;
        if (colmax > r.score) { r.score = colmax; r.ref = i; memcpy(best_col, H, sizeof(int) * (size_t)n_q); }
        if (terminate > 0 && colmax == terminate) break;
    }
    for (int j = 0; j < n_q; ++j) if (best_col[j] == r.score && j < r.read) { r.read = j; break; }
    free(H); free(E); free(best_col);
    return r;
}

typedef struct { int score, ref_begin, ref_end, query_begin, query_end; } sw_aln;

/* ssw_align (ssw.c:788-850) without the CIGAR */
static sw_aln sw_locate(const int8_t* q, int n_q, const int8_t* ref, int n_ref, uint64_t* cells) {
    sw_aln a;
    sw_end f = sw_scan(q, n_q, ref, 0, 1, n_ref, 0, cells);
    a.score = f.score; a.ref_end = f.ref; a.query_end = f.read;
    int n_rq = f.read + 1, n_rr = f.ref + 1;
    int8_t* rq = (int8_t*)malloc((size_t)(n_rq > 0 ? n_rq : 1));
    for (int j = 0; j < n_rq; ++j) rq[j] = q[f.read - j];
    sw_end b = sw_scan(rq, n_rq, ref, n_rr - 1, -1, n_rr, f.score, cells);
    free(rq);
    a.ref_begin = b.ref; a.query_begin = f.read - b.read;
    return a;
}

/* banded_sw (ssw.c:577-758): global-ish banded affine DP over read rows x ref columns with the band doubled until the
 * best cell reaches `score`, then a traceback from the last cell that only stops at read row 0.  Only the totals of
 * I and D operations are needed (getIndels, correctionAlignment.cpp:28-45).
 *
 * Storage is restated as the reference lays it out because the layout is observable: three rolling lines indexed by
 * band slot (slot(i,j) = j - max(0, i-w) + 1), where slot 0 and the slot just right of the row's last column are
 * zeroed before every row (ssw.c:624-627) — on rows i <= w+1 that "edge" slot can be the previous row's last column.
 * Returns 0 on success, -1 if the traceback would leave the matrix (undefined behaviour in the reference). */
static int band_slot(int w, int i, int j) { int x = i - w; if (x < 0) x = 0; return j - x + 1; }
static int banded_indels(const int8_t* ref, const int8_t* read, int refLen, int readLen, int score, int w0,
                         int* n_ins, int* n_del) {
    int w = w0, best = 0;
    int n_line = refLen + 4;
    int* h_prev = (int*)calloc((size_t)n_line, sizeof(int));
    int* e_line = (int*)calloc((size_t)n_line, sizeof(int));
    int* h_cur = (int*)calloc((size_t)n_line, sizeof(int));
    uint8_t* dir = NULL;            /* per row: 3 codes per band slot, [E-choice, F-choice, H-choice] */
    int stride = 0;
    for (;;) {
        const int width = 2 * w + 3;
        stride = 3 * (2 * w + 1);
        free(dir);
        dir = (uint8_t*)calloc((size_t)stride * (size_t)readLen + 8, 1);
        /* lines are as long as the widest band slot that can be touched: min(width-1, refLen) + 1 */
        for (int s = 1; s < width - 1 && s < n_line; ++s) h_prev[s] = 0;
        for (int i = 0; i < readLen; ++i) {
            int beg = i - w > 0 ? i - w : 0;
            int end = i + w < refLen - 1 ? i + w : refLen - 1;
            int edge = end + 1 < width - 1 ? end + 1 : width - 1;
            int f = 0, last = 0;
            h_prev[0] = e_line[0] = h_prev[edge] = e_line[edge] = h_cur[0] = 0;
            uint8_t* d = dir + (size_t)stride * (size_t)i;
            for (int j = beg; j <= end; ++j) {
                const int u = band_slot(w, i, j), up = band_slot(w, i - 1, j), left = band_slot(w, i, j - 1),
                          dg = band_slot(w, i - 1, j - 1);
                const int x = (j - beg) * 3;
                int a = i == 0 ? -GAP_O : h_prev[up] - GAP_O;
                int b = i == 0 ? -GAP_E : e_line[up] - GAP_E;
                e_line[u] = a > b ? a : b;
                d[x] = a > b ? 3 : 2;
                a = h_cur[left] - GAP_O;
                b = f - GAP_E;
                f = a > b ? a : b;
                d[x + 1] = a > b ? 5 : 4;
                const int e1 = e_line[u] > 0 ? e_line[u] : 0, f1 = f > 0 ? f : 0;
                const int gap = e1 > f1 ? e1 : f1;
                const int diag = h_prev[dg] + sub_score(ref[j], read[i]);
                h_cur[u] = gap > diag ? gap : diag;
                if (h_cur[u] > best) best = h_cur[u];
                d[x + 2] = gap <= diag ? 1 : (e1 > f1 ? d[x] : d[x + 1]);
                last = u;
            }
            for (int s = 1; s <= last; ++s) h_prev[s] = h_cur[s];
        }
        if (best >= score) break;
        w *= 2; BRANCH(3);
        if (w > 4 * (refLen + readLen) + 16) { free(h_prev); free(e_line); free(h_cur); free(dir); return -1; }
    }
    /* traceback (ssw.c:672-719) */
    int i = readLen - 1, j = refLen - 1, state = 2, ins = 0, del = 0, rc = 0;
    while (i > 0) {
        int beg = i - w > 0 ? i - w : 0;
        int x = j - beg;
        if (j < 0 || x < 0 || x > 2 * w) { rc = -1; break; }
        switch (dir[(size_t)stride * (size_t)i + (size_t)x * 3 + (size_t)state]) {
            case 1: --i; --j; state = 2; break;
            case 2: --i; state = 0; ++ins; break;
            case 3: --i; state = 2; ++ins; break;
            case 4: --j; state = 1; ++del; break;
            case 5: --j; state = 2; ++del; break;
            default: rc = -1; i = 0; break;
        }
    }
    *n_ins = ins; *n_del = del;
    free(h_prev); free(e_line); free(h_cur); free(dir);
    return rc;
}

/* Aligner::Align + getIndels for the overlap arbitration (correctionAlignment.cpp:107-110): query seq1, reference seq2. */
static int sub_align_indels(const char* s1, const char* s2, int n, int* ins, int* del, uint64_t* cells) {
    int8_t* q = (int8_t*)malloc((size_t)n + 1);
    int8_t* r = (int8_t*)malloc((size_t)n + 1);
    for (int i = 0; i < n; ++i) { q[i] = base_code(s1[i]); r[i] = base_code(s2[i]); }
    sw_aln a = sw_locate(q, n, r, n, cells);
    int rc = 0;
    *ins = *del = 0;
    if (a.score > 0) {
        int refLen = a.ref_end - a.ref_begin + 1, readLen = a.query_end - a.query_begin + 1;
        int bw = refLen > readLen ? refLen - readLen : readLen - refLen;
        rc = banded_indels(r + a.ref_begin, q + a.query_begin, refLen, readLen, a.score, bw + 1, ins, del);
    }   /* score 0: the reference's CIGAR is "1M" plus clips (ssw.c:858-861 on a 1x1 problem): no I, no D */
    free(q); free(r);
    return rc;
}

static int is_upper(char c) { return 'A' <= c && c <= 'Z'; }
static char up(char c) { return ('a' <= c && c <= 'z') ? (char)(c - 32) : c; }
static char low(char c) { return ('A' <= c && c <= 'Z') ? (char)(c + 32) : c; }

/* merCounts[kmer] >= solidThresh with merCounts restricted to its solid entries (sorted list) */
static int solid_has(const uint32_t* keys, uint64_t n, uint32_t key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t m = (lo + hi) / 2; if (keys[m] < key) lo = m + 1; else hi = m; }
    return lo < n && keys[lo] == key;
}
/* nbSolidMers, correctionAlignment.cpp:6-15; str2num (BMEAN/utils.cpp:18-30) maps A0 C1 G2, anything else — lower
 * case included — to 3 */
static int nb_solid_mers(const char* s, unsigned n, const uint32_t* keys, uint64_t nkeys, unsigned k) {
    int nb = 0;
    for (unsigned i = 0; i + k <= n; ++i) {
        uint32_t v = 0;
        for (unsigned t = 0; t < k; ++t) { char c = s[i + t]; v = (v << 2) + (c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u); }
        nb += solid_has(keys, nkeys, v);
    }
    return nb;
}
static int nb_upper(const char* s, unsigned n) { int nb = 0; for (unsigned i = 0; i < n; ++i) nb += is_upper(s[i]); return nb; }

typedef struct { char* p; size_t n, cap; } buf_t;
static void buf_reserve(buf_t* b, size_t n) { if (n + 1 > b->cap) { b->cap = (n + 1) * 2; b->p = (char*)realloc(b->p, b->cap); } }
static void buf_set(buf_t* b, const char* s, size_t n) { buf_reserve(b, n); if (n) memmove(b->p, s, n); b->n = n; b->p[n] = 0; }
/* std::string::replace(pos, len, s) */
static void buf_replace(buf_t* b, size_t pos, size_t len, const char* s, size_t n) {
    if (pos > b->n) return;
    if (len > b->n - pos) len = b->n - pos;
    size_t nn = b->n - len + n;
    buf_reserve(b, nn);
    memmove(b->p + pos + n, b->p + pos + len, b->n - pos - len);
    memcpy(b->p + pos, s, n);
    b->n = nn; b->p[nn] = 0;
}

/* alignConsensus, correctionAlignment.cpp:47-139, for one read.  out receives the corrected read. */
static int reanchor_read(const cg_batch* win, const cg_results* cons, const cg_reads* rd, const cg_params* p, uint32_t r,
                         buf_t* out, uint64_t* cells) {
    const uint32_t w0 = rd->read_win_begin[r], w1 = rd->read_win_begin[r + 1];
    const unsigned k = p->mer_size, ws = rd->window_size, ov = rd->window_overlap;
    const char* raw = rd->read_bases + rd->read_off[r];
    const size_t rawLen = (size_t)(rd->read_off[r + 1] - rd->read_off[r]);
    buf_set(out, raw, rawLen);
    for (size_t i = 0; i < rawLen; ++i) out->p[i] = low(out->p[i]);                       /* :57-58 */
    if (w0 == w1) { out->n = 0; out->p[0] = 0; return 0; }                                /* CONSENT-correction.cpp:23-25 */
    int curPos = (int)rd->win_pos[w0];                                                    /* startPos */
    unsigned oldEnd = 0;
    buf_t cur = {0, 0, 0}, old = {0, 0, 0}, tmp = {0, 0, 0};
    buf_set(&cur, "", 0); buf_set(&old, "", 0);
    uint32_t oldW = w0;   /* window whose solid list is oldMers; empty map before the first assignment */
    int haveOld = 0, rc = 0;
    for (uint32_t w = w0; w < w1; ++w) {
        const uint64_t c0 = cons->cons_off[w], c1 = cons->cons_off[w + 1];
        const int isCons = (c1 - c0) >= k;                                                /* :72 */
        if (isCons) buf_set(&cur, cons->cons + c0, (size_t)(c1 - c0));
        else { uint32_t s0 = win->win_seq_begin[w]; buf_set(&cur, win->bases + win->seq_off[s0], (size_t)(win->seq_off[s0 + 1] - win->seq_off[s0])); }
        int alPos = curPos - (int)ov; if (alPos < 0) alPos = 0;                            /* :80 */
        int sizeAl;
        if ((size_t)alPos + ws + 2 * ov >= out->n) sizeAl = (int)out->n - alPos; else sizeAl = (int)(ws + 2 * ov);   /* :81-85 */
        if (sizeAl <= 0 || cur.n == 0) { rc = -1; break; }     /* the reference would allocate a negative length here */
        int8_t* q = (int8_t*)malloc(cur.n);
        int8_t* t = (int8_t*)malloc((size_t)sizeAl);
        for (size_t i = 0; i < cur.n; ++i) q[i] = base_code(cur.p[i]);
        for (int i = 0; i < sizeAl; ++i) t[i] = base_code(out->p[alPos + i]);
        sw_aln a = sw_locate(q, (int)cur.n, t, sizeAl, cells);                              /* :87 */
        free(q); free(t);
        if (a.score <= 0) { rc = -1; break; }                  /* begin/end are -1 in the reference: nothing sensible follows */
        unsigned beg = (unsigned)(a.ref_begin + alPos), end = (unsigned)(a.ref_end + alPos); /* :88-89 */
        buf_set(&tmp, cur.p + a.query_begin, (size_t)(a.query_end - a.query_begin + 1));    /* :90 */
        buf_set(&cur, tmp.p, tmp.n);
        if (w != w0 && oldEnd >= beg) {                                                     /* :93 */
            unsigned overlap = oldEnd - beg + 1;
            BRANCH(0);
            if (isCons && old.n >= overlap && cur.n >= overlap) {                           /* :95 */
                const char* s1 = old.p + (old.n - overlap);
                const char* s2 = cur.p;
                int same = 1;
                for (unsigned i = 0; i < overlap; ++i) if (up(s1[i]) != up(s2[i])) { same = 0; break; }
                if (!same) {
                    int n1, n2;
                    BRANCH(1);
                    if (overlap >= k) {                                                     /* :99-101 */
                        const uint64_t a0 = haveOld ? cons->solid_off[oldW] : 0, a1 = haveOld ? cons->solid_off[oldW + 1] : 0;
                        n1 = nb_solid_mers(s1, overlap, cons->solid_kmer + a0, a1 - a0, k);
                        n2 = nb_solid_mers(s2, overlap, cons->solid_kmer + cons->solid_off[w], cons->solid_off[w + 1] - cons->solid_off[w], k);
                    } else { n1 = nb_upper(s1, overlap); n2 = nb_upper(s2, overlap); }      /* :103-104 */
                    if (n1 > n2) {                                                          /* :106-117 */
                        int ins = 0, del = 0;
                        BRANCH(2);
                        if (sub_align_indels(s1, s2, (int)overlap, &ins, &del, cells) != 0) { rc = -1; break; }
                        unsigned cut = overlap - (unsigned)ins + (unsigned)del;
                        if (cut < cur.n) {
                            buf_set(&tmp, s1, overlap);
                            buf_reserve(&tmp, overlap + cur.n);
                            memcpy(tmp.p + overlap, cur.p + cut, cur.n - cut);
                            tmp.n = overlap + cur.n - cut; tmp.p[tmp.n] = 0;
                            buf_set(&cur, tmp.p, tmp.n);
                        } else { buf_set(&cur, "", 0); BRANCH(4); }
                    }
                }
            }
        }
        if (cur.n != 0) {                                                                   /* :121 */
            if (isCons) {                                                                   /* :122-126 */
                buf_set(&tmp, cur.p, cur.n);
                for (size_t i = 0; i < tmp.n; ++i) tmp.p[i] = up(tmp.p[i]);
                buf_replace(out, beg, end - beg + 1, tmp.p, tmp.n);
            }
            if (w + 1 < w1) {                                                               /* :127-132 */
                curPos = curPos + (int)rd->win_pos[w + 1] - (int)rd->win_pos[w] - (int)(end - beg + 1) + (int)cur.n;
                buf_set(&old, cur.p, cur.n);
                oldW = w; haveOld = 1;
                oldEnd = beg + (unsigned)cur.n - 1;
            }
        }
    }
    free(cur.p); free(old.p); free(tmp.p);
    return rc;
}

typedef struct { uint64_t* off; char* bases; } corrected_owner;
typedef struct {
    const cg_batch* win; const cg_results* cons; const cg_reads* rd; const cg_params* p;
    buf_t* res; uint32_t* next; int* rc; uint64_t cells;
} job_t;
static void* reanchor_worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (;;) {
        uint32_t r = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (r >= j->rd->n_reads) break;
        if (reanchor_read(j->win, j->cons, j->rd, j->p, r, &j->res[r], &j->cells) != 0) __atomic_store_n(j->rc, -1, __ATOMIC_RELAXED);
    }
    return NULL;
}

static uint64_t g_reanchor_cells;
uint64_t oracle_reanchor_cells(void) { return g_reanchor_cells; }


int oracle_reanchor_reads(const cg_batch* win, const cg_results* cons, const cg_reads* reads, const cg_params* p,
                          int threads, cg_corrected* out, double* seconds) {
    if (!win || !cons || !reads || !p || !out) return CG_ERR_INVALID_ARG;
    const uint32_t R = reads->n_reads;
    buf_t* res = (buf_t*)calloc(R ? R : 1, sizeof(buf_t));
    uint32_t next = 0; int rc = 0;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    job_t jobs[256]; pthread_t th[256];
    for (int t = 0; t < threads; ++t) { job_t j = {win, cons, reads, p, res, &next, &rc, 0}; jobs[t] = j; }
    for (int t = 1; t < threads; ++t) pthread_create(&th[t], NULL, reanchor_worker, &jobs[t]);
    reanchor_worker(&jobs[0]);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    g_reanchor_cells = 0;
    for (int t = 0; t < threads; ++t) g_reanchor_cells += jobs[t].cells;
    corrected_owner* ow = (corrected_owner*)calloc(1, sizeof *ow);
    ow->off = (uint64_t*)calloc((size_t)R + 1, sizeof(uint64_t));
    uint64_t tot = 0;
    for (uint32_t r = 0; r < R; ++r) { tot += res[r].n; ow->off[r + 1] = tot; }
    ow->bases = (char*)malloc(tot + 1);
    for (uint32_t r = 0; r < R; ++r) { if (res[r].n) memcpy(ow->bases + ow->off[r], res[r].p, res[r].n); free(res[r].p); }
    free(res);
    out->n_reads = R; out->read_off = ow->off; out->bases = ow->bases; out->owner_ = ow;
    return rc == 0 ? CG_OK : CG_ERR_INVALID_ARG;
}

void oracle_free_corrected(cg_corrected* c) {
    if (c && c->owner_) { corrected_owner* ow = (corrected_owner*)c->owner_; free(ow->off); free(ow->bases); free(ow); c->owner_ = NULL; }
}
