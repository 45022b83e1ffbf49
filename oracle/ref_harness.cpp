// ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
//
// Thin C-ABI harness around the UNMODIFIED reference implementation of the hot
// path.  It is compiled by oracle/Makefile together with the reference's own
// source files, taken where they lie under /root/reference (nothing is copied
// into this repository), into oracle/_ref/libconsent_ref_*.so.
//
// What it calls (reference file:line):
//   computeConsensusReadCorrection            src/correctionMSA.cpp:29-49
//   and, for the per-stage dump only, the stage functions of BMEAN/bmean.cpp
//   (fill_index_kmers :43, filter_index_kmers :88, get_template :220,
//    longest_ordered_chain :239, average_distance_next_anchor :331,
//    split_reads :476, easy_consensus :603) and weightConsensus /
//   polishCorrection (src/correctionMSA.cpp:6, src/correctionDBG.cpp:93).
//   alignConsensus                            src/correctionAlignment.cpp:47-139
//   (re-anchoring; pulls in the vendored SSW library,
//    BMEAN/Complete-Striped-Smith-Waterman-Library/src/{ssw.c,ssw_cpp.cpp})
//   getAlignmentWindowsPositions / getAlignmentWindowsSequences   src/alignmentWindows.cpp:27-149
//   (window extraction)
//   getNextReadPile                           src/alignmentPiles.cpp:22-58 (PAF ingest; with
//   Overlap(line) and operator<, src/Overlap.h:26-96)
//   trimRead / dropRead                       src/utils.cpp:71-73,96-128 (post-filters)
//
// Used by: tests/ (to pin oracle/consent_oracle.c and to generate
// tests/golden/*), bench.py's cpu_baseline / --impl reference legs.
// Threading = std::thread workers pulling windows from an atomic counter, the
// faithful equivalent of the reference's CTPL pool (src/CONSENT-correction.cpp:77).

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <thread>
#include <utility>
#include <vector>
#include <algorithm>

#include "robin_hood.h"          // src/robin_hood.h (3.11.3) via -I, see Makefile
#include "correctionMSA.h"       // src/correctionMSA.h
#include "correctionDBG.h"       // src/correctionDBG.h
#include "correctionAlignment.h" // src/correctionAlignment.h (alignConsensus)
#include "alignmentWindows.h"    // src/alignmentWindows.h (window positions / piles; pulls in Overlap.h)
#include "alignmentPiles.h"      // src/alignmentPiles.h (getNextReadPile)
#include <fstream>
#include <unistd.h>
#include "consent_b200.h"        // our ABI structs (cg_batch, cg_results, cg_params)

// ---- prototypes of the reference's stage functions (external linkage in
// ---- BMEAN/bmean.cpp; re-declared here, layout-identical structs) ----------
struct localisation { uint32_t read_id; int32_t position; };
typedef robin_hood::unordered_map<kmer, std::vector<localisation>> kmer2localisation;
void fill_index_kmers(const std::vector<std::string>& Reads, kmer2localisation& kmer_index, uint32_t kmer_size,
                      robin_hood::unordered_map<kmer, unsigned>& merCounts, unsigned solidThresh);
robin_hood::unordered_map<kmer, uint32_t> filter_index_kmers(kmer2localisation& kmer_index, double amount);
std::vector<kmer> get_template(kmer2localisation& kmer_index, const std::string& read, int kmer_size);
std::vector<kmer> longest_ordered_chain(kmer2localisation& kmer_index, const std::vector<kmer>& template_read, double edge_solidity);
std::vector<double> average_distance_next_anchor(kmer2localisation& kmer_index, std::vector<kmer>& anchors,
                                                 robin_hood::unordered_map<kmer, uint32_t>& k_count, bool clean);
std::vector<std::vector<std::string>> split_reads(const std::vector<kmer>& anchors, const std::vector<double>& relative_positions,
                                                  const std::vector<std::string>& Reads, kmer2localisation& kmer_index, uint32_t kmer_size);
std::vector<std::string> easy_consensus(std::vector<std::string> V, unsigned maxMSA, std::string path);
std::vector<std::string> consensus_SPOA(std::vector<std::string>& W, unsigned maxMSA, std::string path);
std::pair<std::vector<std::vector<std::string>>, robin_hood::unordered_map<kmer, unsigned>>
MSABMAAC(const std::vector<std::string>& Reads, uint32_t k, double edge_solidity, unsigned solidThresh,
         unsigned minAnchors, unsigned maxMSA, std::string path);                 // BMEAN/bmean.cpp:738
std::string weightConsensus(std::string& consensus, std::vector<std::string>& pile,
                            robin_hood::unordered_map<kmer, unsigned>& merCounts, unsigned merSize,
                            unsigned windowSize, unsigned solidThresh);

namespace {

struct Owner {
    std::vector<uint64_t> cons_off, solid_off;
    std::string cons;
    std::vector<uint8_t> status;
    std::vector<uint32_t> solid_kmer, solid_count;
};

std::vector<std::string> window_pile(const cg_batch* in, uint32_t w) {
    std::vector<std::string> pile;
    for (uint32_t s = in->win_seq_begin[w]; s < in->win_seq_begin[w + 1]; ++s)
        pile.emplace_back(in->bases + in->seq_off[s], in->bases + in->seq_off[s + 1]);
    return pile;
}

struct WinOut {
    std::string cons;
    std::vector<std::pair<uint32_t, uint32_t>> solid;
    uint8_t status;
};

void run_one(const cg_batch* in, uint32_t w, const cg_params* p, WinOut& o) {
    std::vector<std::string> pile = window_pile(in, w);
    std::string readId = "w";
    std::pair<unsigned, unsigned> pos(0, 0);
    unsigned minSupport = 3, merSize = p->mer_size, commonKMers = p->common_kmers,
             minAnchors = p->min_anchors, solid = p->solid_thresh, windowSize = 500;
    auto r = computeConsensusReadCorrection(readId, pile, pos, minSupport, merSize, commonKMers,
                                            minAnchors, solid, windowSize, 150, std::string());
    o.cons = r.first;
    o.solid.clear();
    for (auto& kv : r.second)
        if (kv.second >= solid) o.solid.emplace_back(kv.first, kv.second);
    std::sort(o.solid.begin(), o.solid.end());
    // "fell back to template" is not observable from the return value alone when the
    // consensus happens to equal the template; recompute it the way the reference
    // decides it: MSABMAAC returned no consensus (src/correctionMSA.cpp:34).
    int bmeanSup = std::min((int)commonKMers, (int)pile.size() / 2);
    auto m = MSABMAAC(pile, merSize, bmeanSup, solid, minAnchors, 150, std::string());
    o.status = m.first.size() == 0 ? CG_WINDOW_TEMPLATE : CG_WINDOW_CONSENSUS;
}

}  // namespace

extern "C" {

// Runs the reference on every window of `in` with `threads` workers.
// *seconds receives the wall time of the compute loop only.
// If with_status == 0 the (costly, second MSABMAAC call) status derivation is
// skipped and status[] is all CG_WINDOW_CONSENSUS — used for timing runs.
int ref_correct_windows(const cg_batch* in, const cg_params* p, int threads, int with_status,
                        cg_results* out, double* seconds) {
    if (!in || !p || !out) return CG_ERR_INVALID_ARG;
    const uint32_t W = in->n_windows;
    std::vector<WinOut> res(W);
    std::atomic<uint32_t> next(0);
    if (threads < 1) threads = 1;
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
        for (;;) {
            uint32_t w = next.fetch_add(1);
            if (w >= W) break;
            if (with_status) {
                run_one(in, w, p, res[w]);
            } else {
                std::vector<std::string> pile = window_pile(in, w);
                std::string readId = "w";
                std::pair<unsigned, unsigned> pos(0, 0);
                unsigned minSupport = 3, merSize = p->mer_size, commonKMers = p->common_kmers,
                         minAnchors = p->min_anchors, solid = p->solid_thresh, windowSize = 500;
                auto r = computeConsensusReadCorrection(readId, pile, pos, minSupport, merSize, commonKMers,
                                                        minAnchors, solid, windowSize, 150, std::string());
                res[w].cons = r.first;
                for (auto& kv : r.second)
                    if (kv.second >= solid) res[w].solid.emplace_back(kv.first, kv.second);
                std::sort(res[w].solid.begin(), res[w].solid.end());
                res[w].status = CG_WINDOW_CONSENSUS;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();

    Owner* ow = new Owner();
    ow->cons_off.resize(W + 1, 0);
    ow->solid_off.resize(W + 1, 0);
    ow->status.resize(W);
    for (uint32_t w = 0; w < W; ++w) {
        ow->cons += res[w].cons;
        ow->cons_off[w + 1] = ow->cons.size();
        for (auto& kv : res[w].solid) { ow->solid_kmer.push_back(kv.first); ow->solid_count.push_back(kv.second); }
        ow->solid_off[w + 1] = ow->solid_kmer.size();
        ow->status[w] = res[w].status;
    }
    out->n_windows = W;
    out->cons_off = ow->cons_off.data();
    out->cons = const_cast<char*>(ow->cons.data());
    out->status = ow->status.data();
    out->solid_off = ow->solid_off.data();
    out->solid_kmer = ow->solid_kmer.data();
    out->solid_count = ow->solid_count.data();
    out->owner_ = ow;
    return CG_OK;
}

void ref_free_results(cg_results* r) {
    if (r && r->owner_) { delete static_cast<Owner*>(r->owner_); r->owner_ = nullptr; }
}

// ---- re-anchoring: the unmodified alignConsensus, once per read -------------------------
// Arguments are rebuilt exactly as processRead hands them over (src/CONSENT-correction.cpp:
// 27-47): consensuses[i] / merCounts[i] = what computeConsensusReadCorrection returned for
// window i (merCounts restricted to its solid k-mers: the callee only tests
// `merCounts[kmer] >= solidThresh`, correctionAlignment.cpp:6-15), templates[i] = pile[0],
// pilesPos[i].first = win_pos[i] (.second is never read by alignConsensus), startPos =
// pilesPos[0].first.
struct CorrectedOwner { std::vector<uint64_t> off; std::string bases; };

int ref_reanchor_reads(const cg_batch* win, const cg_results* cons, const cg_reads* reads,
                       const cg_params* p, int threads, cg_corrected* out, double* seconds) {
    if (!win || !cons || !reads || !p || !out) return CG_ERR_INVALID_ARG;
    const uint32_t R = reads->n_reads;
    std::vector<std::string> res(R);
    std::atomic<uint32_t> next(0);
    if (threads < 1) threads = 1;
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
        for (;;) {
            uint32_t r = next.fetch_add(1);
            if (r >= R) break;
            uint32_t w0 = reads->read_win_begin[r], w1 = reads->read_win_begin[r + 1];
            if (w0 == w1) continue;                               // CONSENT-correction.cpp:23-25
            std::vector<std::string> consensuses, templates;
            std::vector<robin_hood::unordered_map<kmer, unsigned>> merCounts(w1 - w0);
            std::vector<std::pair<unsigned, unsigned>> pilesPos;
            for (uint32_t w = w0; w < w1; ++w) {
                consensuses.emplace_back(cons->cons + cons->cons_off[w], cons->cons + cons->cons_off[w + 1]);
                uint32_t s0 = win->win_seq_begin[w];
                templates.emplace_back(win->bases + win->seq_off[s0], win->bases + win->seq_off[s0 + 1]);
                for (uint64_t i = cons->solid_off[w]; i < cons->solid_off[w + 1]; ++i)
                    merCounts[w - w0][cons->solid_kmer[i]] = cons->solid_count[i];
                pilesPos.emplace_back(reads->win_pos[w], reads->win_pos[w] + reads->window_size - 1);
            }
            std::string seq(reads->read_bases + reads->read_off[r], reads->read_bases + reads->read_off[r + 1]);
            res[r] = alignConsensus(std::string("r"), seq, consensuses, merCounts, pilesPos, templates,
                                    (int)pilesPos[0].first, reads->window_size, reads->window_overlap,
                                    p->solid_thresh, p->mer_size);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    CorrectedOwner* ow = new CorrectedOwner();
    ow->off.resize(R + 1, 0);
    for (uint32_t r = 0; r < R; ++r) { ow->bases += res[r]; ow->off[r + 1] = ow->bases.size(); }
    out->n_reads = R;
    out->read_off = ow->off.data();
    out->bases = const_cast<char*>(ow->bases.data());
    out->owner_ = ow;
    return CG_OK;
}

void ref_free_corrected(cg_corrected* c) {
    if (c && c->owner_) { delete static_cast<CorrectedOwner*>(c->owner_); c->owner_ = nullptr; }
}

// ---- window extraction: the reference's own phase A of processRead (src/CONSENT-correction.cpp:21-35) --------------------
// Overlaps are rebuilt field by field (src/Overlap.h:8-20), reads are named by their store index, `sequences` holds the reads
// a pile mentions exactly as getSequencesMap would have decoded them from the 2-bit index (src/alignmentPiles.cpp:5-20).
struct WindowSetOwner {
    std::vector<uint32_t> wsb, rwb, wpos, wend;
    std::vector<uint64_t> soff, roff;
    std::string bases, rbases;
};

int ref_extract_windows(const cg_piles* p, unsigned merSize, cg_window_set* out) {
    if (!p || !out) return CG_ERR_INVALID_ARG;
    WindowSetOwner* ow = new WindowSetOwner();
    ow->wsb.push_back(0); ow->rwb.push_back(0); ow->soff.push_back(0); ow->roff.push_back(0);
    auto store = [&](uint32_t r) { return std::string(p->store_bases + p->store_off[r], p->store_bases + p->store_off[r + 1]); };
    for (uint32_t pi = 0; pi < p->n_piles; ++pi) {
        std::vector<Overlap> al;
        robin_hood::unordered_map<std::string, std::string> sequences;
        const std::string qName = "r" + std::to_string(p->pile_read[pi]);
        sequences[qName] = store(p->pile_read[pi]);
        for (uint32_t o = p->pile_ov_begin[pi]; o < p->pile_ov_begin[pi + 1]; ++o) {
            const cg_overlap& c = p->overlaps[o];
            Overlap a;
            a.qName = qName; a.qLength = p->pile_qlen[pi]; a.qStart = c.q_start; a.qEnd = c.q_end; a.strand = c.strand != 0;
            a.tName = "r" + std::to_string(c.t_read); a.tLength = c.t_length; a.tStart = c.t_start; a.tEnd = c.t_end;
            a.resMatches = 0; a.alBlockLen = 0; a.mapQual = 0;
            if (sequences[a.tName] == "") sequences[a.tName] = store(c.t_read);
            al.push_back(a);
        }
        if (!al.empty()) {
            std::vector<std::pair<unsigned, unsigned>> pilesPos =
                getAlignmentWindowsPositions(al.begin()->qLength, al, p->min_support, 0, p->window_size, (int)p->window_overlap);
            for (auto& pp : pilesPos) {
                std::vector<std::string> pile = getAlignmentWindowsSequences(al, p->min_support, p->window_size, p->window_overlap, sequences,
                                                                             pp.first, pp.second, merSize, 0, 0);
                if (pile.empty()) { delete ow; return CG_ERR_INVALID_ARG; }      // the reference dereferences curPile[0] (CONSENT-correction.cpp:36)
                for (auto& s : pile) { ow->bases += s; ow->soff.push_back(ow->bases.size()); }
                ow->wsb.push_back((uint32_t)(ow->soff.size() - 1));
                ow->wpos.push_back(pp.first); ow->wend.push_back(pp.second);
            }
        }
        ow->rbases += sequences[qName];
        ow->roff.push_back(ow->rbases.size());
        ow->rwb.push_back((uint32_t)ow->wpos.size());
    }
    if (ow->wpos.empty()) { ow->wpos.push_back(0); ow->wend.push_back(0); }
    out->batch.n_windows = (uint32_t)(ow->wsb.size() - 1);
    out->batch.win_seq_begin = ow->wsb.data(); out->batch.seq_off = ow->soff.data(); out->batch.bases = ow->bases.data();
    out->reads.n_reads = p->n_piles; out->reads.read_win_begin = ow->rwb.data(); out->reads.read_off = ow->roff.data();
    out->reads.read_bases = ow->rbases.data(); out->reads.win_pos = ow->wpos.data();
    out->reads.window_size = p->window_size; out->reads.window_overlap = p->window_overlap;
    out->win_end = ow->wend.data();
    out->owner_ = ow;
    return CG_OK;
}

void ref_free_window_set(cg_window_set* s) {
    if (s && s->owner_) { delete static_cast<WindowSetOwner*>(s->owner_); s->owner_ = nullptr; }
}

// ---- PAF ingest: the reference's own getNextReadPile over the text, driven like runCorrection drives it --------------------
// (src/CONSENT-correction.cpp:87-90,107-110: empty piles are skipped while the stream is not at its end).  The text goes through
// a temporary file because getNextReadPile takes a std::ifstream and seeks in it.  Names become store indices the way
// indexReads fills its map (`index[header] = ...`: the last entry of a name wins, src/utils.cpp:186).
struct PileSetOwner { std::vector<uint32_t> pile_read, pile_qlen, ovb, res; std::vector<cg_overlap> ov; };

int ref_ingest_paf(const char* paf, uint64_t nbytes, const cg_read_names* names, uint32_t max_support, cg_pile_set* out) {
    if (!out || !names || (nbytes && !paf) || max_support == 0) return CG_ERR_INVALID_ARG;
    if (nbytes && paf[nbytes - 1] != '\n') return CG_ERR_INVALID_ARG;             // getNextReadPile never returns on such a text
    robin_hood::unordered_map<std::string, uint32_t> id;
    for (uint32_t i = 0; i < names->n_reads; ++i) id[std::string(names->names + names->name_off[i], names->names + names->name_off[i + 1])] = i;
    char path[] = "/tmp/consent_ref_paf_XXXXXX";
    int fd = mkstemp(path);
    if (fd < 0) return CG_ERR_INVALID_ARG;
    for (uint64_t w = 0; w < nbytes;) { ssize_t k = write(fd, paf + w, nbytes - w); if (k <= 0) { close(fd); unlink(path); return CG_ERR_INVALID_ARG; } w += (uint64_t)k; }
    close(fd);
    PileSetOwner* ow = new PileSetOwner();
    ow->ovb.push_back(0);
    uint64_t lines = 0;
    for (uint64_t i = 0; i + 1 <= nbytes; ++i) if (paf[i] == '\n' && i > 0 && paf[i - 1] != '\n') ++lines;
    int rc = CG_OK;
    try {
        std::ifstream f(path);
        while (!f.eof()) {
            std::vector<Overlap> al = getNextReadPile(f, max_support);
            if (al.size() == 0) continue;
            auto q = id.find(al.begin()->qName);
            if (q == id.end()) { rc = CG_ERR_INVALID_ARG; break; }
            ow->pile_read.push_back(q->second);
            ow->pile_qlen.push_back(al.begin()->qLength);
            for (const Overlap& a : al) {
                auto t = id.find(a.tName);
                if (t == id.end()) { rc = CG_ERR_INVALID_ARG; break; }
                cg_overlap o = {t->second, a.strand ? 1u : 0u, a.qStart, a.qEnd, a.tStart, a.tEnd, a.tLength};
                ow->ov.push_back(o);
                ow->res.push_back(a.resMatches);
            }
            ow->ovb.push_back((uint32_t)ow->ov.size());
        }
    } catch (...) { rc = CG_ERR_INVALID_ARG; }                                       // stoi on a malformed column
    unlink(path);
    if (rc != CG_OK) { delete ow; return rc; }
    if (ow->pile_read.empty()) { ow->pile_read.push_back(0); ow->pile_qlen.push_back(0); }
    if (ow->ov.empty()) { ow->ov.push_back(cg_overlap{}); ow->res.push_back(0); }
    out->n_piles = (uint32_t)(ow->ovb.size() - 1);
    out->pile_read = ow->pile_read.data(); out->pile_qlen = ow->pile_qlen.data(); out->pile_ov_begin = ow->ovb.data();
    out->overlaps = ow->ov.data(); out->res_matches = ow->res.data(); out->n_lines = lines; out->owner_ = ow;
    return CG_OK;
}

void ref_free_pile_set(cg_pile_set* s) {
    if (s && s->owner_) { delete static_cast<PileSetOwner*>(s->owner_); s->owner_ = nullptr; }
}

// std::sort(v.rbegin(), v.rend()) on Overlap records, exactly the call of getNextReadPile (src/alignmentPiles.cpp:40,51):
// order[i] = index of the record that ends up at position i.
void ref_sort_desc(const uint32_t* keys, uint32_t n, uint32_t* order) {
    std::vector<Overlap> v(n);
    for (uint32_t i = 0; i < n; ++i) { v[i].resMatches = keys[i]; v[i].alBlockLen = i; }
    std::sort(v.rbegin(), v.rend());
    for (uint32_t i = 0; i < n; ++i) order[i] = v[i].alBlockLen;
}

// ---- post-filters: the tail of processRead (src/CONSENT-correction.cpp:49-59) on the strings alignConsensus returned ----------
int ref_finish_reads(const cg_corrected* in, uint32_t trim_mer, cg_corrected* out) {
    if (!in || !out) return CG_ERR_INVALID_ARG;
    CorrectedOwner* ow = new CorrectedOwner();
    ow->off.resize((size_t)in->n_reads + 1, 0);
    for (uint32_t r = 0; r < in->n_reads; ++r) {
        std::string s(in->bases + in->read_off[r], in->bases + in->read_off[r + 1]);
        if (trim_mer && !s.empty()) {
            unsigned run = 0, best = 0;
            for (char c : s) { run = isUpperCase(c) ? run + 1 : 0; best = std::max(best, run); }
            if (best < trim_mer) s.clear();                                          // trimRead would index s[-1] (utils.cpp:111-120)
            else {
                s = trimRead(s, trim_mer);
                if (dropRead(s)) s.clear();
            }
        }
        ow->bases += s;
        ow->off[r + 1] = ow->bases.size();
    }
    out->n_reads = in->n_reads; out->read_off = ow->off.data(); out->bases = const_cast<char*>(ow->bases.data()); out->owner_ = ow;
    return CG_OK;
}

// Per-stage text dump of one window, produced by calling the reference's own
// stage functions in the order MSABMAAC does (BMEAN/bmean.cpp:763-821).  The
// oracle emits the same format (oracle_dump_window) so stages can be diffed.
// Returned buffer is malloc'ed; free with ref_free_text.
char* ref_dump_window(const cg_batch* in, uint32_t w, const cg_params* p) {
    std::vector<std::string> Reads = window_pile(in, w);
    std::ostringstream os;
    unsigned k = p->mer_size, solid = p->solid_thresh;
    int S = std::min((int)p->common_kmers, (int)Reads.size() / 2);
    os << "S " << S << "\n";
    kmer2localisation kmer_index;
    robin_hood::unordered_map<kmer, unsigned> merCounts;
    fill_index_kmers(Reads, kmer_index, k, merCounts, solid);
    {
        std::vector<std::pair<uint32_t, uint32_t>> v;
        for (auto& kv : merCounts) if (kv.second >= solid) v.emplace_back(kv.first, kv.second);
        std::sort(v.begin(), v.end());
        os << "M " << v.size();
        for (auto& kv : v) os << " " << kv.first << ":" << kv.second;
        os << "\n";
    }
    auto kmer_count = filter_index_kmers(kmer_index, (double)S);
    auto tpl = get_template(kmer_index, Reads[0], (int)k);
    os << "T " << tpl.size();
    for (auto x : tpl) os << " " << x;
    os << "\n";
    std::vector<kmer> anchors = longest_ordered_chain(kmer_index, tpl, (double)S);
    os << "A " << anchors.size();
    for (auto x : anchors) os << " " << x;
    os << "\n";
    std::vector<double> rel = average_distance_next_anchor(kmer_index, anchors, kmer_count, false);
    os << "R " << rel.size();
    for (auto x : rel) os << " " << (long long)x;
    os << "\n";
    auto regions = split_reads(anchors, rel, Reads, kmer_index, k);
    os << "G " << regions.size() << "\n";
    std::string stacked;
    if (regions.size() >= p->min_anchors) {
        for (size_t i = 0; i < regions.size(); ++i) {
            os << "g " << i << " " << regions[i].size();
            for (auto& s : regions[i]) os << " " << s;
            os << "\n";
            if (regions[i].empty()) continue;
            auto c = easy_consensus(regions[i], 150, std::string());
            os << "c " << i << " " << c[0] << "\n";
            stacked += c[0];
        }
    }
    os << "C " << stacked << "\n";
    std::string cons;
    if (stacked.empty()) {
        cons = Reads[0];
    } else {
        cons = stacked;
        if (cons.length() >= k) {
            cons = weightConsensus(cons, Reads, merCounts, k, 500, solid);
            os << "W " << cons << "\n";
            cons = polishCorrection(cons, merCounts, k, solid);
        }
    }
    os << "P " << cons << "\n";
    std::string s = os.str();
    char* buf = (char*)malloc(s.size() + 1);
    memcpy(buf, s.c_str(), s.size() + 1);
    return buf;
}

// MSA rows of one POA run over the given strings (consensus_SPOA, bmean.cpp:585-599),
// newline separated.  For pinning the oracle's POA restatement in isolation.
char* ref_spoa_msa(const char* const* seqs, uint32_t n) {
    std::vector<std::string> W;
    for (uint32_t i = 0; i < n; ++i) W.emplace_back(seqs[i]);
    auto msa = consensus_SPOA(W, 150, std::string());
    std::string s;
    for (auto& r : msa) { s += r; s += '\n'; }
    char* buf = (char*)malloc(s.size() + 1);
    memcpy(buf, s.c_str(), s.size() + 1);
    return buf;
}

void ref_free_text(char* p) { free(p); }

int ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
