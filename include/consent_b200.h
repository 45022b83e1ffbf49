/*
 * consent_b200.h — C ABI of the B200-native CONSENT per-window correction path.
 *
 * This is the drop-in boundary.  The reference has no FFI layer: its seam is the
 * in-process C++ call
 *
 *   std::pair<std::string, robin_hood::unordered_map<kmer,unsigned>>
 *   computeConsensusReadCorrection(readId, piles, pilesPos, minSupport, merSize,
 *        commonKMers, minAnchors, solidThresh, windowSize, maxMSA, path)
 *                                        (reference src/correctionMSA.h:8, body
 *                                         src/correctionMSA.cpp:29-49)
 *   computeConsensusAssemblyPolishing(id, ...same..., nbThreads)
 *                                        (src/correctionMSA.h:10, body :51-70)
 *
 * called once per window from processRead (src/CONSENT-correction.cpp:34-44) and
 * from the CTPL jobs of processContig (src/CONSENT-polishing.cpp:49,66).  One
 * window per call cannot feed a GPU, so the replacement is the same function
 * *batched over windows*: every entry point below takes W windows and returns,
 * per window, exactly what the reference call returns for that window:
 *
 *   - the consensus string, mixed case (upper = solid k-mer support), byte for
 *     byte what `.first` of the reference pair holds,
 *   - the solid k-mers of the pile with their occurrence counts, i.e. every
 *     (kmer, count) of `.second` with count >= solidThresh, sorted by k-mer
 *     value (the reference map additionally holds zero-count keys inserted by
 *     operator[] look-ups; its consumers only look keys up and absent == 0:
 *     src/correctionAlignment.cpp:6-15,103-104),
 *   - a status byte telling whether MSABMAAC produced a consensus or the raw
 *     template was returned (src/correctionMSA.cpp:34-36).
 *
 * Unused reference arguments (readId, pilesPos, minSupport, windowSize, maxMSA,
 * path, id, nbThreads — never read on this path, BMEAN/bmean.cpp:585-599,738)
 * are not part of the ABI.
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 * All functions return CG_OK (0) or a negative cg_status; no exceptions.
 */
#ifndef CONSENT_B200_H
#define CONSENT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CG_ABI_VERSION 5

typedef enum cg_status {
    CG_OK                 =  0,
    CG_ERR_INVALID_ARG    = -1,  /* null pointer, empty pile, k out of range ...     */
    CG_ERR_NO_DEVICE      = -2,  /* no CUDA device / extension not usable            */
    CG_ERR_CUDA           = -3,  /* a CUDA runtime call failed, see cg_last_error    */
    CG_ERR_OUT_OF_MEMORY  = -4,
    CG_ERR_BAD_BASE       = -5,  /* a base outside {A,C,G,T} (reads are 2-bit stored
                                    upstream, src/utils.cpp:21-54,189, so the
                                    reference never sees one)                       */
    CG_ERR_CAPACITY       = -6,  /* a per-window limit of this build was exceeded    */
    CG_ERR_STATE          = -7   /* call sequence error (run before upload ...)      */
} cg_status;

/* Parameters actually read by the path (src/correctionMSA.cpp:31-32,43-45).
 * Defaults of the CONSENT-correct wrapper: k=9, solid=4, commonKMers=8,
 * minAnchors=2 (reference CONSENT-correct:42-50). */
typedef struct cg_params {
    uint32_t mer_size;      /* -k  merSize      : k-mer length, 2..15 (reference: 1<<(2k) overflows at 16, bmean.cpp:46);
                               up to 9 the k-mers are counted in a direct-addressed table, 10..15 in a hash table */
    uint32_t solid_thresh;  /* -f  solidThresh  : min occurrences for a solid k-mer     */
    uint32_t common_kmers;  /* -c  commonKMers  : anchor support cap; S = min(c, N/2)   */
    uint32_t min_anchors;   /* -A  minAnchors   : MSABMAAC bails out if regions < this  */
} cg_params;

/* W windows.  Window w owns sequences [win_seq_begin[w], win_seq_begin[w+1]);
 * its first sequence is the template (piles[0]).  Sequence s owns bytes
 * [seq_off[s], seq_off[s+1]) of `bases` (ASCII, upper-case ACGT, no terminator).
 * Caller owns all input buffers. */
typedef struct cg_batch {
    uint32_t        n_windows;
    const uint32_t* win_seq_begin;   /* [n_windows + 1] */
    const uint64_t* seq_off;         /* [n_seqs + 1], n_seqs = win_seq_begin[n_windows] */
    const char*     bases;
} cg_batch;

#define CG_WINDOW_CONSENSUS   0   /* MSABMAAC produced a consensus                      */
#define CG_WINDOW_TEMPLATE    1   /* fell back to the raw template (correctionMSA.cpp:34-36) */
#define CG_WINDOW_ERROR       2   /* the window is over a limit of this build (more than 4095 sequences, a sequence over 6000
                                     bases, a template over 2047 k-mers, an anchor table or POA graph beyond the largest
                                     workspace): it comes back as its raw template (with the solid k-mers counted before the
                                     limit was met, none if it was met at once), the rest of the batch is unaffected (the reference has no such limits and never fails a window)      */

/* Results for W windows, owned by the library until cg_free_results(). */
typedef struct cg_results {
    uint32_t  n_windows;
    uint64_t* cons_off;      /* [n_windows + 1] byte offsets into cons                 */
    char*     cons;          /* consensus strings, mixed case, concatenated            */
    uint8_t*  status;        /* [n_windows] CG_WINDOW_*                                */
    uint64_t* solid_off;     /* [n_windows + 1] offsets into solid_kmer/solid_count    */
    uint32_t* solid_kmer;    /* 2-bit packed k-mers (A0 C1 G2 T3, first base most
                                significant: BMEAN/utils.cpp:18-30), ascending         */
    uint32_t* solid_count;   /* total occurrences in the pile (>= solid_thresh)        */
    void*     owner_;        /* private                                                */
} cg_results;

typedef struct cg_handle cg_handle;

/* Device / lifetime ------------------------------------------------------ */
int         cg_abi_version(void);
int         cg_device_count(void);
/* Create a context bound to CUDA device `device` (own stream, own workspaces). */
int         cg_create(int device, const cg_params* params, cg_handle** out);
void        cg_destroy(cg_handle* h);
const char* cg_last_error(const cg_handle* h);   /* h may be NULL: last create error */

/* Tuning knobs that never change results: workspace budget per chunk of windows ("chunk_budget_bytes",
 * "chunk_max_windows", "lanes"), resident warps of the POA tiers ("poa_c1_warps", "poa_g_warps", "poa_wide{1,2}_warps"),
 * last-resort POA scratch ("poa_tier{1,2}_{warps,nodes,cells}"). */
int         cg_set_option(cg_handle* h, const char* key, long long value);
/* Two more options change what crosses the bus, not what is computed:
 *   "input_2bit" 1          cg_upload / cg_correct_windows read `bases` as 2 bits per base (base i of the batch at bits 2 (i & 3) of byte
 *                           i >> 2; A 0 C 1 G 2 T 3 — cg_pack_bases_2bit), the form the reference's own read index keeps reads in
 *                           (src/utils.cpp:21-54): a quarter of the bytes over PCIe, expanded on the device.  seq_off stays in bases.
 *   "results_with_solid" 0  the solid k-mer lists are not downloaded (solid_off reads all zeros); they stay in HBM, where
 *                           cg_reanchor_reads / cg_finish_reads given the live results read them. */
void        cg_pack_bases_2bit(const char* ascii, uint64_t n, uint8_t* out, int threads);

/* One-shot call: host buffers in, host buffers out (H2D, kernels, D2H inside).
 * This is the batched equivalent of calling computeConsensusReadCorrection on
 * every window of `in`.  Blocking; one call at a time per handle. */
int  cg_correct_windows(cg_handle* h, const cg_batch* in, cg_results* out);
void cg_free_results(cg_results* r);

/* Staged calls (what cg_correct_windows does, split so a caller — and bench.py —
 * can keep a batch resident in HBM and time the kernels alone). */
int  cg_upload(cg_handle* h, const cg_batch* in);   /* H2D + 2-bit packing on device  */
int  cg_run(cg_handle* h);                          /* every kernel of the path       */
int  cg_download(cg_handle* h, cg_results* out);    /* D2H of the results of cg_run   */

/* Consensus re-anchoring ---------------------------------
 * Batched equivalent of
 *
 *   std::string alignConsensus(rawRead, sequence, consensuses, merCounts, pilesPos,
 *                              templates, startPos, windowSize, windowOverlap,
 *                              solidThresh, merSize)      (src/correctionAlignment.h:7,
 *                                                          body src/correctionAlignment.cpp:47-139)
 *
 * called once per read by processRead (src/CONSENT-correction.cpp:47) right after the
 * per-window consensus calls.  Every window consensus is located on the (progressively
 * corrected) read by a local alignment — StripedSmithWaterman::Aligner defaults: match 2,
 * mismatch 2, gap open 3, gap extend 1, BMEAN/Complete-Striped-Smith-Waterman-Library/
 * src/ssw_cpp.cpp:419-426, ssw.c:788-891 — overlaps with the previous window are
 * arbitrated by solid k-mers (correctionAlignment.cpp:90-118) and the aligned stretch of
 * the read is replaced by the upper-cased consensus.  Windows of one read are processed in
 * order (each alignment sees the replacements of the previous ones); reads are independent.
 *
 * Inputs: the window batch (its first sequence per window = templates[i]), the results of
 * cg_correct_windows on it (consensuses[i]; the solid k-mer lists stand for merCounts[i],
 * whose consumers only test `count >= solidThresh`, correctionAlignment.cpp:6-15), and the
 * reads below.  Read r owns windows [read_win_begin[r], read_win_begin[r+1]) of the batch in
 * pile order; startPos = win_pos of its first window (CONSENT-correction.cpp:47).
 * A read without windows yields an empty string (CONSENT-correction.cpp:23-25). */
typedef struct cg_reads {
    uint32_t        n_reads;
    const uint32_t* read_win_begin;  /* [n_reads + 1], read_win_begin[n_reads] == n_windows          */
    const uint64_t* read_off;        /* [n_reads + 1] byte offsets into read_bases                    */
    const char*     read_bases;      /* the reads (`sequence`; any case, lower-cased internally :57)  */
    const uint32_t* win_pos;         /* [n_windows] pilesPos[i].first                                 */
    uint32_t        window_size;     /* -l windowSize    (CONSENT-correct default 500)                */
    uint32_t        window_overlap;  /* -m windowOverlap (default 50)                                 */
} cg_reads;

/* Corrected reads, owned by the library until cg_free_corrected(): upper case = replaced by
 * a consensus, lower case = untouched raw read (what alignConsensus returns, before the
 * trimRead/dropRead post-filters of processRead). */
typedef struct cg_corrected {
    uint32_t  n_reads;
    uint64_t* read_off;      /* [n_reads + 1] byte offsets into bases */
    char*     bases;
    void*     owner_;        /* private */
} cg_corrected;

/* Host buffers in, host buffers out (H2D, kernel, D2H inside).  `cons` may be the live result
 * of cg_correct_windows on the same handle (it is only read).  Blocking; one call at a time
 * per handle. */
int  cg_reanchor_reads(cg_handle* h, const cg_batch* windows, const cg_results* cons,
                       const cg_reads* reads, cg_corrected* out);
void cg_free_corrected(cg_corrected* c);
/* CUDA-event time (ms) of the kernel of the last cg_reanchor_reads and the DP cells
 * (query x reference, forward + reverse passes of every alignment) it swept. */
int  cg_reanchor_stats(const cg_handle* h, float* kernel_ms, uint64_t* dp_cells);

/* Window extraction ----------------------------------------
 * Phase A of processRead (src/CONSENT-correction.cpp:21-35) on the device: for every read pile, the window positions
 *
 *   getAlignmentWindowsPositions(tplLen, alignments, minSupport, maxSupport, windowSize, windowOverlap)
 *                                                        (src/alignmentWindows.cpp:27-85, getCoverages :5-25)
 * and for every window its pile
 *
 *   getAlignmentWindowsSequences(alignments, ..., sequences, qBeg, end, merSize, ...)   (src/alignmentWindows.cpp:87-149)
 *
 * cut (and reverse-complemented, src/reverseComplement.cpp:6-24) from a read store that is shipped once, instead of
 * W x N strings assembled on the host and uploaded.  The host keeps what is serial and tiny: PAF parsing, grouping by
 * query and the top-maxSupport selection (src/alignmentPiles.cpp:22-58) — a pile's overlaps are given in the order
 * getNextReadPile leaves them.  The result is the resident window batch: cg_run / cg_download follow as after cg_upload,
 * cg_download_windows returns the extracted piles and the cg_reads view cg_reanchor_reads needs. */
typedef struct cg_overlap {          /* one PAF record as src/Overlap.h:26-60 holds it */
    uint32_t t_read;                 /* tName: index of the target read in the store                  */
    uint32_t strand;                 /* 0 '+', 1 '-'                                                  */
    uint32_t q_start, q_end;         /* qStart, qEnd = PAF end - 1 (inclusive, Overlap.h:37)          */
    uint32_t t_start, t_end;         /* tStart, tEnd = PAF end - 1 (inclusive, Overlap.h:48)          */
    uint32_t t_length;               /* tLength (PAF column 7)                                        */
} cg_overlap;

typedef struct cg_piles {
    uint32_t          n_store;       /* reads in the store                                            */
    const uint64_t*   store_off;     /* [n_store + 1] byte offsets into store_bases                   */
    const char*       store_bases;   /* ASCII; stored as the reference stores reads: upper-cased, anything
                                        but A, C, G becomes T (src/utils.cpp:21-32,189)              */
    uint32_t          n_piles;
    const uint32_t*   pile_read;     /* [n_piles] the query read (qName) of pile p                    */
    const uint32_t*   pile_qlen;     /* [n_piles] qLength (PAF column 2)                              */
    const uint32_t*   pile_ov_begin; /* [n_piles + 1] overlaps of pile p = [pile_ov_begin[p], [p+1])  */
    const cg_overlap* overlaps;
    uint32_t          min_support;   /* -s */
    uint32_t          window_size;   /* -l */
    uint32_t          window_overlap;/* -m */
} cg_piles;

/* The extracted windows on the host, owned by the library until cg_free_window_set(). */
typedef struct cg_window_set {
    cg_batch  batch;                 /* the piles (bases == NULL if they were not asked for)          */
    cg_reads  reads;                 /* pile p = read p: its windows, pilesPos[i].first, its sequence */
    uint32_t* win_end;               /* [n_windows] pilesPos[i].second                                */
    void*     owner_;
} cg_window_set;

/* Extracts every window of every pile into the handle's resident batch (what cg_upload would have received from a host
 * running phase A).  Blocking.  Piles without a window yield reads without windows. */
int  cg_upload_piles(cg_handle* h, const cg_piles* piles);
/* A host that streams the PAF in bounded batches of piles (the reference's ring of 100 000 jobs,
 * src/CONSENT-correction.cpp:76-127) ships the read store once: after cg_set_read_store, cg_upload_piles calls whose
 * cg_piles has store_off == NULL and store_bases == NULL (n_store unchanged) cut their windows from the resident store.
 * The store is what indexReads holds (src/utils.cpp:166-204), normalised on the device like cg_upload_piles does. */
int  cg_set_read_store(cg_handle* h, uint32_t n_store, const uint64_t* store_off, const char* store_bases);
int  cg_download_windows(cg_handle* h, int with_bases, cg_window_set* out);
void cg_free_window_set(cg_window_set* s);
/* CUDA-event time (ms) of the extraction kernels of the last cg_upload_piles (host round trips for the window counts
 * excluded), of the copy kernel alone (k_ex_copy: reads and writes one byte per pile base), and the pile bytes written. */
int  cg_extract_stats(const cg_handle* h, float* kernel_ms, float* copy_ms, uint64_t* pile_bytes);

/* PAF ingest --------------------------------------------------
 * The serial front of runCorrection (src/CONSENT-correction.cpp:87,107) on the device: every
 *
 *   std::vector<Overlap> getNextReadPile(std::ifstream& f, unsigned maxSupport)   (src/alignmentPiles.cpp:22-58)
 *
 * of one PAF text at once — Overlap(line) (src/Overlap.h:26-60: 12 tab-separated columns, qEnd / tEnd = column - 1,
 * strand = column 5 != "+"), grouping of consecutive lines with the same qName into a pile (an empty line also ends a
 * pile, :29-37), `std::sort(rbegin, rend)` by resMatches (column 10, Overlap.h:90-96) and the cut to maxSupport (:39-42).
 * std::sort is not stable: the order of overlaps with equal resMatches is the one libstdc++'s introsort (median-of-3
 * quicksort above 16 elements, heapsort below depth 2·log2 n, final insertion sort; bits/stl_algo.h) leaves on the
 * reversed range, and the kernel replays exactly that sequence of comparisons and moves per pile — which overlaps survive
 * the cut, and in which order they enter a window's pile, decides bytes downstream.
 * Read names (PAF columns 1 and 6) become store indices through `names` (the FASTA headers up to the first blank, as
 * indexReads keeps them: src/utils.cpp:163-190; a name listed twice resolves to its last entry, like `index[header] =`).
 * A line with fewer than 12 columns, a numeric column stoi() would not take as a non-negative int, or a name that is
 * not in the table is an error (the reference throws or reads an empty sequence there). */
typedef struct cg_read_names {
    uint32_t        n_reads;
    const uint64_t* name_off;        /* [n_reads + 1] byte offsets into names                          */
    const char*     names;           /* concatenated, no terminators                                   */
} cg_read_names;

/* Piles as cg_piles takes them (pile_* / overlaps fields), owned by the library until cg_free_pile_set(). */
typedef struct cg_pile_set {
    uint32_t    n_piles;
    uint32_t*   pile_read;           /* [n_piles] qName of the pile                                    */
    uint32_t*   pile_qlen;           /* [n_piles] alignments.begin()->qLength (after the sort)         */
    uint32_t*   pile_ov_begin;       /* [n_piles + 1]                                                  */
    cg_overlap* overlaps;            /* in the order getNextReadPile returns them                      */
    uint32_t*   res_matches;         /* [pile_ov_begin[n_piles]] column 10 of every kept overlap       */
    uint64_t    n_lines;             /* non-empty PAF lines parsed                                     */
    void*       owner_;
} cg_pile_set;

/* Host text in, host piles out (H2D, kernels, D2H inside).  Blocking; one call at a time per handle. */
int  cg_ingest_paf(cg_handle* h, const char* paf, uint64_t paf_bytes, const cg_read_names* names,
                   uint32_t max_support, cg_pile_set* out);
void cg_free_pile_set(cg_pile_set* s);
/* CUDA-event time (ms) of all kernels of the last cg_ingest_paf, of the line parser alone (k_paf_parse: reads every
 * byte of the text once), and the bytes of text. */
int  cg_ingest_stats(const cg_handle* h, float* kernel_ms, float* parse_ms, uint64_t* paf_bytes);

/* Post-filters -------------------------------------------------
 * cg_reanchor_reads followed, still on the device, by the tail of processRead (src/CONSENT-correction.cpp:49-59):
 *
 *   correctedRead = trimRead(correctedRead, 1);          (src/utils.cpp:96-128: cut to the first / last run of
 *                                                          `trim_mer` upper-case bases; "" unless end > beg)
 *   if (dropRead(correctedRead)) correctedRead = "";     (src/utils.cpp:71-73: fewer than 10 % upper-case bases,
 *                                                          (float) n / length < 0.1)
 *
 * so that `out` holds exactly the sequence line of every FASTA record runCorrection prints (:100-103; an empty string =
 * no record).  trim_mer = 0 skips both filters (the proof-file mode, doTrimRead = false, :76-79) and equals
 * cg_reanchor_reads.  A read without any upper-case base makes trimRead index before its string (unsigned i >= 0,
 * utils.cpp:111-120); it yields "" here. */
int  cg_finish_reads(cg_handle* h, const cg_batch* windows, const cg_results* cons, const cg_reads* reads,
                     uint32_t trim_mer, cg_corrected* out);
/* The same for the batch cg_upload_piles + cg_run left resident on this handle, without taking anything through the host on the way:
 * window consensuses, solid k-mer lists, templates and the reads (pile p = read p of the store) are read where they lie in HBM.
 * Equals cg_download + cg_download_windows + cg_finish_reads(..., trim_mer, out); only the corrected reads come back. */
int  cg_finish_resident(cg_handle* h, uint32_t trim_mer, cg_corrected* out);
/* CUDA-event time (ms) of the post-filter kernel of the last cg_finish_reads. */
int  cg_finish_stats(const cg_handle* h, float* kernel_ms);

/* Instrumentation -------------------------------------------------------- */
#define CG_STAGE_PACK     0   /* ASCII -> 2-bit                                         */
#define CG_STAGE_INDEX    1   /* k-mer counts, solid list, template anchor table  (a3-a5) */
#define CG_STAGE_CHAIN    2   /* pair scores + chain DP                           (a6-a7) */
#define CG_STAGE_SPLIT    3   /* distance stats + segment split                   (a8-a9) */
#define CG_STAGE_POA      4   /* segmented POA + column vote                     (a11-a14) */
#define CG_STAGE_STITCH   5   /* concatenate region consensuses                     (a11) */
#define CG_STAGE_POLISH   6   /* weight + DBG polish                             (a15-a19) */
#define CG_N_STAGES       7
/* CUDA-event time (ms) each stage spent on the handle's stream in the last cg_run,
 * and the number of kernel launches it made. */
int  cg_stage_ms(const cg_handle* h, float ms[CG_N_STAGES], uint32_t launches[CG_N_STAGES]);
/* CUDA-event time (ms) of the whole last cg_run on the handle's stream, first launch to last result. */
int  cg_run_ms(const cg_handle* h, float* ms);
/* Number of chunks the uploaded batch is processed in (each chunk = one launch of every kernel of the path). */
int  cg_chunk_count(const cg_handle* h);

/* Per-kernel timing of the last cg_run.  Every launch of the kernels below is bracketed by its own pair of CUDA events ON THE
 * STREAM IT IS LAUNCHED ON (the POA tiers of a chunk run side by side on three streams, and two chunks are in flight at a
 * time, so a stage span on one stream — cg_stage_ms — also contains other kernels' time; these do not).
 * ms = sum of the launch durations, launches = their number; poa_cells / poa_pred_cells = the score-matrix cells (V+1)*L and
 * predecessor-row cells E*L each POA tier computed (the algorithmic-bytes model of SURVEY §8d, per kernel). */
#define CG_K_PACK      0   /* k_plan + k_scan + k_pack                  */
#define CG_K_INDEX     1   /* k_index                                   */
#define CG_K_CHAIN     2   /* k_chain (both shared-memory sizes)        */
#define CG_K_SPLIT     3   /* k_split                                   */
#define CG_K_POA_C1    4   /* k_poa2<C1>: graph + matrix in shared memory */
#define CG_K_POA_G     5   /* k_poa2<G>: graph in shared memory, matrix in L2 */
#define CG_K_POA_W1    6   /* k_poa2<W1>: wide tier, <= 1024 nodes      */
#define CG_K_POA_W2    7   /* k_poa2<W2>: wide tier, <= 4096 nodes      */
#define CG_K_POA_LAST  8   /* k_poa: the last resort                    */
#define CG_K_POLISH    9   /* k_polish                                  */
#define CG_K_OUT      10   /* k_stitch_len, k_out_sizes, k_scan, k_gather */
#define CG_N_KERNELS  11
typedef struct cg_kernel_stats {
    float    ms[CG_N_KERNELS];
    uint32_t launches[CG_N_KERNELS];
    uint64_t poa_cells[4], poa_pred_cells[4];      /* C1, G, W1, W2 */
} cg_kernel_stats;
int  cg_get_kernel_stats(const cg_handle* h, cg_kernel_stats* out);

/* Test instrumentation: per-stage text dump (solid list, surviving template k-mers, anchor chain, mean distances, regions with their
 * segments and consensuses, stitched consensus) of window w of the batch the last cg_run processed as ONE chunk — the format of
 * oracle_dump_window / ref_dump_window, so that every stage of the CUDA path can be compared with the reference's, not only the end
 * result.  *text is malloc'ed. */
int  cg_debug_dump_window(cg_handle* h, uint32_t w, char** text);

/* Work counters of the last cg_run, the inputs of the algorithmic-bytes model
 * (SURVEY §8d): alignments, score-matrix cells sum (V+1)*L, predecessor-row cells
 * sum E*L, POA graphs, anchors in chains, solid k-mers, packed input bytes,
 * consensus bytes. */
typedef struct cg_counters {
    uint64_t windows, sequences, bases;
    uint64_t anchors, regions, poa_graphs, alignments;
    uint64_t dp_cells, dp_pred_cells;
    uint64_t solid_kmers, consensus_bytes, fallback_windows;
    uint64_t error_windows;          /* windows returned with CG_WINDOW_ERROR */
} cg_counters;
int  cg_get_counters(const cg_handle* h, cg_counters* out);

#ifdef __cplusplus
}
#endif
#endif /* CONSENT_B200_H */
