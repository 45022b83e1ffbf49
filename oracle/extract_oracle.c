/*
 * extract_oracle.c — TEST INFRASTRUCTURE.  CPU oracle for window extraction.
 *
 * A plain-C restatement of phase A of processRead (src/CONSENT-correction.cpp:21-35):
 *     getCoverages                      src/alignmentWindows.cpp:5-25
 *     getAlignmentWindowsPositions      src/alignmentWindows.cpp:27-85
 *     getAlignmentWindowsSequences      src/alignmentWindows.cpp:87-149
 *     rev_comp::run                     src/reverseComplement.cpp:6-24
 * written from the reference's behaviour (no reference source is copied).  It only CHECKS the CUDA path
 * (consent_b200/csrc/k_extract.cuh).
 *
 * PARITY PINNING: pinned against the UNMODIFIED reference functions (oracle/_ref: ref_extract_windows) on seeded piles
 * and on the overlaps of the shipped example (tests/test_extract.py).  The example fixture tests/golden/example_windows.txt.gz
 * — piles cut by the reference itself — is reproduced from tests/golden/example_small.paf.gz where the two overlap.
 *
 * Inputs on which the reference reads outside its buffers (an overlap ending beyond qLength, a window beyond the stored
 * read: empty pile dereferenced at CONSENT-correction.cpp:36) are rejected instead of imitated.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "consent_oracle.h"

typedef struct { uint32_t* v; size_t n, cap; } v32;
typedef struct { uint64_t* v; size_t n, cap; } v64;
typedef struct { char* v; size_t n, cap; } vch;
static void p32(v32* a, uint32_t x) { if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 64; a->v = (uint32_t*)realloc(a->v, a->cap * 4); } a->v[a->n++] = x; }
static void p64(v64* a, uint64_t x) { if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 64; a->v = (uint64_t*)realloc(a->v, a->cap * 8); } a->v[a->n++] = x; }
static void pch(vch* a, const char* s, size_t n) {
    if (a->n + n + 1 > a->cap) { a->cap = 2 * (a->n + n + 1); a->v = (char*)realloc(a->v, a->cap); }
    if (n) memcpy(a->v + a->n, s, n);
    a->n += n;
}

typedef struct { v32 wsb, rwb, wpos, wend; v64 soff, roff; vch bases, rbases; } ws_owner;

static char comp(char c) {                                       /* reverseComplement.cpp:34-43; anything else maps to 0 */
    switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
                 case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c'; default: return 0; }
}
static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

int oracle_extract_windows(const cg_piles* p, unsigned merSize, cg_window_set* out) {
    if (!p || !out) return CG_ERR_INVALID_ARG;
    ws_owner* ow = (ws_owner*)calloc(1, sizeof *ow);
    int rc = CG_OK;
    p32(&ow->wsb, 0); p32(&ow->rwb, 0); p64(&ow->soff, 0); p64(&ow->roff, 0);
    const unsigned ws = p->window_size, ovl = p->window_overlap, minSup = p->min_support;
    for (uint32_t pi = 0; pi < p->n_piles && rc == CG_OK; ++pi) {
        const cg_overlap* al = p->overlaps + p->pile_ov_begin[pi];
        const uint32_t nal = p->pile_ov_begin[pi + 1] - p->pile_ov_begin[pi];
        const uint32_t q = p->pile_read[pi];
        const char* qseq = p->store_bases + p->store_off[q];
        const uint64_t qlenStored = p->store_off[q + 1] - p->store_off[q];
        const unsigned tplLen = p->pile_qlen[pi];
        const size_t firstWin = ow->wpos.n;
        if (tplLen == 0) { rc = CG_ERR_INVALID_ARG; break; }
        if (nal) {                                               /* processRead is only called on non-empty piles */
            /* getCoverages :5-25 */
            unsigned* cov = (unsigned*)calloc((size_t)tplLen + 1, sizeof(unsigned));
            for (uint32_t a = 0; a < nal && rc == CG_OK; ++a) {
                if (al[a].q_start > al[a].q_end) continue;
                if (al[a].q_end >= tplLen) { rc = CG_ERR_INVALID_ARG; break; }       /* the reference writes past its array */
                for (unsigned i = al[a].q_start; i <= al[a].q_end; ++i) cov[i]++;
            }
            if (rc == CG_OK) {
                /* getAlignmentWindowsPositions :27-85 */
                unsigned curLen = 0, beg = 0, i = 0;
                while (i < tplLen) {
                    if (curLen >= ws) {
                        p32(&ow->wpos, beg); p32(&ow->wend, beg + curLen - 1);
                        if (ovl) i = i - ovl;
                        beg = i; curLen = 0;
                    }
                    if (cov[i] < minSup) { curLen = 0; i++; beg = i; } else { curLen++; i++; }
                }
                int pushed = 0;
                unsigned end = tplLen - 1;
                curLen = 0; i = tplLen - 1;
                while (i > 0 && !pushed) {
                    if (curLen >= ws) { p32(&ow->wpos, end - curLen + 1); p32(&ow->wend, end); pushed = 1; end = i; curLen = 0; }
                    if (cov[i] < minSup) { curLen = 0; i--; end = i; } else { curLen++; i--; }
                }
            }
            free(cov);
        }
        /* getAlignmentWindowsSequences :87-149 for every window */
        for (size_t w = firstWin; w < ow->wpos.n && rc == CG_OK; ++w) {
            const unsigned qBeg = ow->wpos.v[w], end = ow->wend.v[w];
            unsigned length = end - qBeg + 1, shift;
            if ((uint64_t)qBeg + length - 1 >= qlenStored) { rc = CG_ERR_INVALID_ARG; break; }        /* :95-97: empty pile */
            pch(&ow->bases, qseq + qBeg, length); p64(&ow->soff, ow->bases.n);
            for (uint32_t a = 0; a < nal; ++a) {
                const cg_overlap* o = &al[a];
                unsigned tBeg = o->t_start, tEnd = o->t_end;
                length = end - qBeg + 1;
                shift = qBeg > o->q_start ? qBeg - o->q_start : 0;
                if (!(((o->q_start <= qBeg && o->q_end > qBeg) || (end <= o->q_end && o->q_start < end)) && o->t_start + shift <= o->t_end)) continue;
                if (qBeg < o->q_start && o->q_end < end) {
                    shift = 0;
                    tBeg = (unsigned)imax(0, (int)o->t_start - ((int)o->q_start - (int)qBeg));
                    tEnd = (unsigned)imin((int)o->t_length - 1, (int)o->t_end + ((int)end - (int)o->q_end));
                    length = tEnd - tBeg + 1;
                } else if (qBeg < o->q_start) {
                    shift = 0;
                    tBeg = (unsigned)imax(0, (int)o->t_start - ((int)o->q_start - (int)qBeg));
                    length = (unsigned)imin((int)length, imin((int)o->t_length - 1, (int)tBeg + (int)length - 1) - (int)tBeg + 1);
                } else if (o->q_end < end) {
                    tEnd = (unsigned)imin((int)o->t_length - 1, (int)o->t_end + ((int)end - (int)o->q_end));
                    length = (unsigned)imin((int)length, (int)tEnd - imax(0, (int)tEnd - (int)length + 1) + 1);
                }
                const char* tseq = p->store_bases + p->store_off[o->t_read];
                const uint64_t tlen = p->store_off[o->t_read + 1] - p->store_off[o->t_read];
                if ((uint64_t)tBeg > tlen) { rc = CG_ERR_INVALID_ARG; break; }                         /* substr throws */
                uint64_t n1 = (uint64_t)(unsigned)(tEnd - tBeg + 1);                                  /* substr(tBeg, tEnd - tBeg + 1) */
                if (n1 > tlen - tBeg) n1 = tlen - tBeg;
                if ((uint64_t)shift > n1) { rc = CG_ERR_INVALID_ARG; break; }                          /* substr throws */
                uint64_t n2 = length;                                                                 /* .substr(shift, length) */
                if (n2 > n1 - shift) n2 = n1 - shift;
                if (n2 < merSize) continue;                                                           /* :142-144 */
                char* buf = (char*)malloc(n2 + 1);
                for (uint64_t x = 0; x < n2; ++x)
                    buf[x] = o->strand ? comp(tseq[tBeg + n1 - 1 - (shift + x)]) : tseq[tBeg + shift + x];
                pch(&ow->bases, buf, n2); p64(&ow->soff, ow->bases.n);
                free(buf);
            }
            p32(&ow->wsb, (uint32_t)(ow->soff.n - 1));
        }
        pch(&ow->rbases, qseq, qlenStored); p64(&ow->roff, ow->rbases.n);
        p32(&ow->rwb, (uint32_t)ow->wpos.n);
    }
    if (rc != CG_OK) {
        free(ow->wsb.v); free(ow->rwb.v); free(ow->wpos.v); free(ow->wend.v); free(ow->soff.v); free(ow->roff.v); free(ow->bases.v); free(ow->rbases.v); free(ow);
        return rc;
    }
    if (ow->wpos.n == 0) { p32(&ow->wpos, 0); p32(&ow->wend, 0); ow->wpos.n = ow->wend.n = 0; }
    pch(&ow->bases, "", 0); pch(&ow->rbases, "", 0);
    out->batch.n_windows = (uint32_t)(ow->wsb.n - 1);
    out->batch.win_seq_begin = ow->wsb.v; out->batch.seq_off = ow->soff.v; out->batch.bases = ow->bases.v;
    out->reads.n_reads = p->n_piles; out->reads.read_win_begin = ow->rwb.v; out->reads.read_off = ow->roff.v;
    out->reads.read_bases = ow->rbases.v; out->reads.win_pos = ow->wpos.v;
    out->reads.window_size = ws; out->reads.window_overlap = ovl;
    out->win_end = ow->wend.v;
    out->owner_ = ow;
    return CG_OK;
}

void oracle_free_window_set(cg_window_set* s) {
    if (!s || !s->owner_) return;
    ws_owner* ow = (ws_owner*)s->owner_;
    free(ow->wsb.v); free(ow->rwb.v); free(ow->wpos.v); free(ow->wend.v); free(ow->soff.v); free(ow->roff.v); free(ow->bases.v); free(ow->rbases.v); free(ow);
    s->owner_ = NULL;
}
