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

#ifdef __cplusplus
}
#endif
#endif
