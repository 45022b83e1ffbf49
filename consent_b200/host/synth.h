/* synth.h — deterministic synthetic alignment-pile windows.
 *
 * Shared by the GPU path, the oracle and the reference harness so that all three
 * consume byte-identical inputs.  Not part of the reference: CONSENT ships no
 * generator; the window shape follows the reference's own windowing (500-base
 * windows, src/alignmentWindows.cpp:27-85) and its PacBio / ONT error profiles
 * (CONSENT-correct:185,187).
 *
 * Window w of a run with seed s draws from std::mt19937_64(s * 1000003 + w):
 *   truth[i]  = rng() & 3                          i in [0, len)
 *   for every sequence (the template, index 0, included) and every i:
 *       u = (rng() >> 11) * 2^-53
 *       u >= e                 -> emit truth[i]
 *       u/e <  p_sub           -> emit (truth[i] + 1 + rng() % 3) & 3      (substitution)
 *       u/e <  p_sub + p_ins   -> emit rng() & 3, then truth[i]            (insertion)
 *       else                   -> emit nothing                             (deletion)
 * Bases are written as ASCII ACGT (A0 C1 G2 T3).
 */
#ifndef CONSENT_B200_SYNTH_H
#define CONSENT_B200_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cg_synth_spec {
    uint64_t seed;
    uint32_t first_window;   /* index of the first window to generate                  */
    uint32_t n_windows;
    uint32_t n_seqs;         /* sequences per window, template included                */
    uint32_t truth_len;      /* 500                                                    */
    double   err;            /* total error rate e: PB 0.15, ONT 0.10                  */
    double   p_sub, p_ins;   /* shares of e: PB 0.10/0.60 (del 0.30), ONT 0.40/0.20    */
} cg_synth_spec;

/* Upper bound of the bytes one call writes into `bases`. */
uint64_t cg_synth_max_bases(const cg_synth_spec* s);
/* Fills win_seq_begin[n_windows+1], seq_off[n_windows*n_seqs+1] and bases; returns bytes written.
 * threads <= 1 is serial; the output does not depend on the thread count. */
uint64_t cg_synth_windows(const cg_synth_spec* s, uint32_t* win_seq_begin, uint64_t* seq_off,
                          char* bases, int threads);

/* Synthetic reads with their windows, for the re-anchoring path.
 *
 * Read r of a run with seed s draws from std::mt19937_64(s * 7000003 + r):
 *   truth[i] = rng() & 3, i in [0, truth_len); the read = truth through the error channel above,
 *   remembering for every read base the truth index it was emitted at;
 *   windows as getAlignmentWindowsPositions cuts them on a fully covered read
 *   (src/alignmentWindows.cpp:27-85): starts 0, ws-ov, 2(ws-ov), ... while a full window fits,
 *   then one last window [len-ws, len-1];
 *   pile of window [a, b] = the read's own bases [a, b] (the template) followed by n_seqs-1 fresh
 *   channel copies of truth[t(a) .. t(b)].  Every `thin_every`-th window (if non-zero) only gets
 *   `thin_seqs` sequences (poorly covered windows, template fall-backs). */
typedef struct cg_synth_read_spec {
    uint64_t seed;
    uint32_t first_read, n_reads;
    uint32_t n_seqs;         /* sequences per window, template included */
    uint32_t truth_len;      /* bases of truth per read                 */
    uint32_t window_size, window_overlap;
    uint32_t thin_every, thin_seqs;
    double   err, p_sub, p_ins;
} cg_synth_read_spec;

/* Upper bounds for the buffers of one cg_synth_reads call. */
void cg_synth_reads_bounds(const cg_synth_read_spec* s, uint64_t* max_windows, uint64_t* max_seqs,
                           uint64_t* max_bases, uint64_t* max_read_bases);
/* Fills the cg_batch arrays (win_seq_begin, seq_off, bases) and the cg_reads arrays (read_win_begin,
 * read_off, read_bases, win_pos); returns the number of windows. */
uint64_t cg_synth_reads(const cg_synth_read_spec* s, uint32_t* win_seq_begin, uint64_t* seq_off, char* bases,
                        uint32_t* read_win_begin, uint64_t* read_off, char* read_bases, uint32_t* win_pos);

/* Synthetic read piles over a random genome, for the window-extraction path (BASELINE config 4 shape
 * without the overlapper): genome[i] = rng() & 3 from std::mt19937_64(seed * 9000011); read r = `read_len` genome bases
 * from a uniform start, on a random strand, through the error channel; the overlaps of a pile are derived from the true
 * genome coordinates (what minimap2 would report up to its end-point jitter), sorted by overlap length (descending, ties by
 * read index) and cut at max_support, as getNextReadPile leaves them (src/alignmentPiles.cpp:22-58).
 * Piles are built for reads [0, n_piles). */
typedef struct cg_synth_pile_spec {
    uint64_t seed;
    uint32_t genome_len, n_reads, read_len, n_piles, max_support, min_overlap;
    double   err, p_sub, p_ins;
} cg_synth_pile_spec;
/* Sizes for the caller's buffers: total store bases and total overlaps (exact). Returns an opaque handle. */
void* cg_synth_piles_build(const cg_synth_pile_spec* s, uint64_t* store_bases, uint64_t* n_overlaps);
/* Copies the built set out (overlaps as 7 uint32 per record in cg_overlap order) and frees the handle. */
void  cg_synth_piles_fetch(void* handle, uint64_t* store_off, char* store, uint32_t* pile_read, uint32_t* pile_qlen,
                           uint32_t* pile_ov_begin, uint32_t* overlaps7);

#ifdef __cplusplus
}
#endif
#endif
