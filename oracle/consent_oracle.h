/* consent_oracle.h — TEST INFRASTRUCTURE (CPU oracle), see consent_oracle.c. */
#ifndef CONSENT_ORACLE_H
#define CONSENT_ORACLE_H
#include "consent_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Same contract as the reference harness (oracle/ref_harness.cpp: ref_correct_windows). */
int   oracle_correct_windows(const cg_batch* in, const cg_params* p, int threads, int with_status,
                             cg_results* out, double* seconds);
void  oracle_free_results(cg_results* r);
/* Per-stage text dump of window w, same format as ref_dump_window. malloc'ed. */
char* oracle_dump_window(const cg_batch* in, uint32_t w, const cg_params* p);
/* MSA rows (newline separated) of one POA run over seqs[0..n). malloc'ed. */
char* oracle_spoa_msa(const char* const* seqs, uint32_t n);
/* Heaviest-bundle consensus of one POA run with scores m / x / gap — only to replay spoa's own known-answer tests
 * (BMEAN/spoa/test/spoa_test.cpp) against the on-path DP / traceback / graph code. malloc'ed. */
char* oracle_spoa_consensus(const char* const* seqs, uint32_t n, int m, int x, int gap);
void  oracle_free_text(char* p);
/* Work counters accumulated by oracle_correct_windows since the last reset. */
void  oracle_reset_counters(void);
void  oracle_get_counters(cg_counters* out);

/* Re-anchoring oracle (oracle/reanchor_oracle.c): alignConsensus once per read; same contract as
 * ref_reanchor_reads of the reference harness. */
int   oracle_reanchor_reads(const cg_batch* win, const cg_results* cons, const cg_reads* reads, const cg_params* p,
                            int threads, cg_corrected* out, double* seconds);
void  oracle_free_corrected(cg_corrected* c);
/* DP cells (forward + backward scans of every alignment) swept by the last oracle_reanchor_reads. */
uint64_t oracle_reanchor_cells(void);

/* Window-extraction oracle (oracle/extract_oracle.c): phase A of processRead; same contract as ref_extract_windows. */
int   oracle_extract_windows(const cg_piles* p, unsigned merSize, cg_window_set* out);
void  oracle_free_window_set(cg_window_set* s);

/* PAF-ingest and post-filter oracle (oracle/ingest_oracle.c); same contracts as ref_ingest_paf / ref_sort_desc /
 * ref_finish_reads of the reference harness. */
int   oracle_ingest_paf(const char* paf, uint64_t nbytes, const cg_read_names* names, uint32_t max_support, cg_pile_set* out);
void  oracle_free_pile_set(cg_pile_set* s);
void  oracle_sort_desc(const uint32_t* keys, uint32_t n, uint32_t* order);
int   oracle_finish_reads(const cg_corrected* in, uint32_t trim_mer, cg_corrected* out);
void  oracle_free_finished(cg_corrected* c);

#ifdef __cplusplus
}
#endif
#endif
