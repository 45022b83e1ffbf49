// consent_main.cpp — bin/CONSENT-correction and bin/CONSENT-polishing of this repository: the reference's two drivers
// (src/main.cpp, src/CONSENT-correction.cpp:62-137, src/CONSENT-polishing.cpp:107-136) with everything between the input files and
// the FASTA records on B200s, through the C ABI of include/consent_b200.h.
//
//   reference                                                         here
//   indexReads (-r, -R)                                               ReadStore::load, shipped once per GPU (cg_set_read_store)
//   main thread: getNextReadPile per read, ring of 100 000 jobs       reader thread: PafStream, bounded batches of whole piles,
//                                                                     at most 2 batches in flight per GPU
//   CTPL pool: processRead / processContig                            one host thread + one cg_handle per GPU:
//     getNextReadPile's parse + std::sort + cut to maxSupport           cg_ingest_paf
//     getSequencesMap, window positions + piles                         cg_upload_piles (resident store)
//     computeConsensusReadCorrection / ...AssemblyPolishing             cg_run
//     alignConsensus, trimRead, dropRead                                cg_finish_resident
//   results[curJob].get() in submission order -> stdout               writer: batches in file order -> stdout
//
// Same command line as the reference binaries (getopt string of src/main.cpp:29; -M, -p, -j, -i, -d, -e, -w, -n are accepted and,
// as in the reference's hot path, without effect on the output: -j sizes the reference's thread pool, here the GPUs do the work).
// Extra options: -g LIST  CUDA devices, e.g. "0,1,2,3" or "all" (default: device 0; also CONSENT_GPUS)
//                -B MB    PAF text per batch (default: file size / (6 x workers), between 256 KB and 48 MB; always whole piles)
//                -W N     host threads (each with its own handle) per GPU (default 2: one batch's host phases run under the other's kernels)
// The binary is the polisher when it is called as *polishing* (or with -P): it never trims (src/CONSENT-polishing.cpp:19).
// Reads are sharded over the GPUs batch by batch; nothing is exchanged between GPUs during the computation, the corrected reads of a
// batch come back to this process over PCIe and are written in input order.  (A multi-PROCESS run — one rank per GPU under torchrun —
// gathers over NCCL instead: consent_b200/shard.py.)
#include <getopt.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "consent_b200.h"
#include "paf_stream.h"
#include "reads.h"

namespace {

struct Options {
    std::string alignments, reads, proof;
    unsigned minSupport = 3, maxSupport = 1000, windowSize = 500, merSize = 9, commonKMers = 8, minAnchors = 10, solidThresh = 4,
             windowOverlap = 50;                                                    // src/main.cpp:17-26
    std::vector<int> gpus;
    unsigned workers_per_gpu = 2;       // host threads (each with its own handle) per GPU: one batch's host phases overlap the other's kernels
    size_t batch_mb = 0;                // 0: sized from the PAF file so that every worker gets several batches (at most 48 MB)
    bool polishing = false, verbose = false;
};

struct Batch { size_t seq; std::string text; };

// Bounded queue between the reader and the GPU threads.
class BatchQueue {
public:
    explicit BatchQueue(size_t cap) : cap_(cap) {}
    void push(Batch&& b) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return q_.size() < cap_ || aborted_; });
        if (aborted_) return;
        q_.push_back(std::move(b));
        cv_.notify_all();
    }
    bool pop(Batch* b) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty() || closed_ || aborted_; });
        if (aborted_ || q_.empty()) return false;
        *b = std::move(q_.front());
        q_.pop_front();
        cv_.notify_all();
        return true;
    }
    void close() { std::lock_guard<std::mutex> lk(mu_); closed_ = true; cv_.notify_all(); }
    void abort() { std::lock_guard<std::mutex> lk(mu_); aborted_ = true; cv_.notify_all(); }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Batch> q_;
    size_t cap_;
    bool closed_ = false, aborted_ = false;
};

// Records of finished batches, written in file order.
class OrderedWriter {
public:
    void put(size_t seq, std::string&& fasta) {
        std::lock_guard<std::mutex> lk(mu_);
        done_[seq] = std::move(fasta);
        while (!done_.empty() && done_.begin()->first == next_) {
            const std::string& s = done_.begin()->second;
            fwrite(s.data(), 1, s.size(), stdout);
            done_.erase(done_.begin());
            ++next_;
        }
        fflush(stdout);
    }

private:
    std::mutex mu_;
    std::map<size_t, std::string> done_;
    size_t next_ = 0;
};

struct Shared {
    const Options* opt;
    const consent::ReadStore* store;
    BatchQueue* queue;
    OrderedWriter* writer;
    std::mutex err_mu;
    std::string error;
    unsigned long long windows = 0, error_windows = 0, piles = 0;
    double t_first = -1, t_last = 0;                 // seconds since start: first batch taken by a GPU, last batch written
    double t_created = 0, t_store = 0, t_first_done = -1;   // last handle created, last read store shipped, first batch finished
    std::chrono::steady_clock::time_point t0;
    double now() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

void fail(Shared* sh, const std::string& what) {
    {
        std::lock_guard<std::mutex> lk(sh->err_mu);
        if (sh->error.empty()) sh->error = what;
    }
    sh->queue->abort();
}

void gpu_main(Shared* sh, int device) {
    const Options& o = *sh->opt;
    const consent::ReadStore& st = *sh->store;
    cg_params prm = {o.merSize, o.solidThresh, o.commonKMers, o.minAnchors};
    cg_handle* h = nullptr;
    if (cg_create(device, &prm, &h) != CG_OK) { fail(sh, std::string("cg_create: ") + cg_last_error(nullptr)); return; }
    { std::lock_guard<std::mutex> lk(sh->err_mu); sh->t_created = sh->now(); }
    const char dummy = 'A';
    if (cg_set_read_store(h, st.size(), st.off.data(), st.bases.empty() ? &dummy : st.bases.data()) != CG_OK) {
        fail(sh, std::string("cg_set_read_store: ") + cg_last_error(h)); cg_destroy(h); return;
    }
    { std::lock_guard<std::mutex> lk(sh->err_mu); sh->t_store = sh->now(); }
    const cg_read_names names = {st.size(), st.name_off.data(), st.name_bytes.empty() ? &dummy : st.name_bytes.data()};
    const bool trim = !o.polishing && o.proof.empty();                               // doTrimRead, CONSENT-correction.cpp:17,70-73
    Batch b;
    while (sh->queue->pop(&b)) {
        {
            std::lock_guard<std::mutex> lk(sh->err_mu);
            if (sh->t_first < 0) sh->t_first = sh->now();
        }
        cg_pile_set ps;
        cg_corrected cor;
        if (cg_ingest_paf(h, b.text.data(), b.text.size(), &names, o.maxSupport, &ps) != CG_OK) { fail(sh, std::string("cg_ingest_paf: ") + cg_last_error(h)); break; }
        cg_piles piles = {st.size(), nullptr, nullptr, ps.n_piles, ps.pile_read, ps.pile_qlen, ps.pile_ov_begin, ps.overlaps,
                          o.minSupport, o.windowSize, o.windowOverlap};
        int rc = cg_upload_piles(h, &piles);
        if (rc == CG_OK) rc = cg_run(h);
        if (rc == CG_OK) rc = cg_finish_resident(h, trim ? 1u : 0u, &cor);
        if (rc != CG_OK) { fail(sh, std::string("batch ") + std::to_string(b.seq) + ": " + cg_last_error(h)); cg_free_pile_set(&ps); break; }
        std::string fasta;
        fasta.reserve((size_t)cor.read_off[ps.n_piles] + 64 * (size_t)ps.n_piles);
        for (uint32_t p = 0; p < ps.n_piles; ++p) {
            const uint64_t a = cor.read_off[p], e = cor.read_off[p + 1];
            if (a == e) continue;                                                    // no window / dropped: no record (:100-103)
            fasta.push_back('>');
            fasta += st.names[ps.pile_read[p]];
            fasta.push_back('\n');
            fasta.append(cor.bases + a, (size_t)(e - a));
            fasta.push_back('\n');
        }
        cg_counters cnt;
        if (cg_get_counters(h, &cnt) == CG_OK) {
            std::lock_guard<std::mutex> lk(sh->err_mu);
            sh->windows += cnt.windows; sh->error_windows += cnt.error_windows; sh->piles += ps.n_piles;
        }
        cg_free_corrected(&cor);
        cg_free_pile_set(&ps);
        sh->writer->put(b.seq, std::move(fasta));
        {
            std::lock_guard<std::mutex> lk(sh->err_mu);
            sh->t_last = sh->now();
            if (sh->t_first_done < 0) sh->t_first_done = sh->t_last;
        }
    }
    cg_destroy(h);
}

std::vector<int> parse_gpus(const char* s) {
    std::vector<int> g;
    if (!s || !*s) return g;
    if (!strcmp(s, "all")) { const int n = cg_device_count(); for (int i = 0; i < n; ++i) g.push_back(i); return g; }
    const char* p = s;
    while (*p) {
        char* e;
        const long v = strtol(p, &e, 10);
        if (e == p) break;
        g.push_back((int)v);
        p = *e == ',' ? e + 1 : e;
    }
    return g;
}

}  // namespace

int main(int argc, char* argv[]) {
    if (argc < 2) {
        fprintf(stderr, "Usage: %s -a alignments.paf -r reads.fasta [-R reads2.fasta] [-s minSupport] [-S maxSupport] [-l windowSize] [-k merSize] "
                        "[-c commonKMers] [-A minAnchors] [-f solidThresh] [-m windowOverlap] [-j threads] [-g gpus] [-B batchMB]\n\n", argv[0]);
        return EXIT_FAILURE;
    }
    Options o;
    const char* self = strrchr(argv[0], '/');
    self = self ? self + 1 : argv[0];
    if (strstr(self, "olish")) o.polishing = true;
    const char* genv = getenv("CONSENT_GPUS");
    if (genv) o.gpus = parse_gpus(genv);
    int opt;
    while ((opt = getopt(argc, argv, "a:A:d:k:s:S:M:l:f:e:p:c:m:j:w:r:R:n:i:g:B:W:Pv")) != -1) {
        switch (opt) {
            case 'a': o.alignments = optarg; break;
            case 's': o.minSupport = atoi(optarg); break;
            case 'S': o.maxSupport = atoi(optarg); break;
            case 'l': o.windowSize = atoi(optarg); break;
            case 'k': o.merSize = atoi(optarg); break;
            case 'c': o.commonKMers = atoi(optarg); break;
            case 'A': o.minAnchors = atoi(optarg); break;
            case 'f': o.solidThresh = atoi(optarg); break;
            case 'm': o.windowOverlap = atoi(optarg); break;
            case 'r': o.reads = optarg; break;
            case 'R': o.proof = optarg; break;
            case 'g': o.gpus = parse_gpus(optarg); break;
            case 'B': o.batch_mb = (size_t)std::max(1, atoi(optarg)); break;
            case 'W': o.workers_per_gpu = (unsigned)std::max(1, atoi(optarg)); break;
            case 'P': o.polishing = true; break;
            case 'v': o.verbose = true; break;
            case 'M': case 'p': case 'j': case 'i': case 'd': case 'e': case 'w': case 'n': break;     // accepted, unused by the path
            default:
                fprintf(stderr, "Usage: %s -a alignments.paf -r reads.fasta [...]\n\n", argv[0]);
                return EXIT_FAILURE;
        }
    }
    if (o.alignments.empty() || o.reads.empty()) { fprintf(stderr, "%s: -a and -r are required\n", self); return EXIT_FAILURE; }
    if (o.gpus.empty()) o.gpus.push_back(0);

    Shared sh;
    sh.t0 = std::chrono::steady_clock::now();
    consent::ReadStore store;
    std::string err;
    if (!store.load(o.reads, &err)) { fprintf(stderr, "%s: %s\n", self, err.c_str()); return EXIT_FAILURE; }
    if (!o.proof.empty() && !store.load(o.proof, &err)) { fprintf(stderr, "%s: %s\n", self, err.c_str()); return EXIT_FAILURE; }
    store.finish();

    size_t batch_bytes = o.batch_mb << 20;
    if (!batch_bytes) {                                                     // ~6 batches per worker, between 256 KB and 48 MB (a batch
        size_t fsize = 0;                                                   // always holds whole piles: at least one)
        if (FILE* pf = fopen(o.alignments.c_str(), "rb")) { fseek(pf, 0, SEEK_END); const long e = ftell(pf); fsize = e > 0 ? (size_t)e : 0; fclose(pf); }
        const size_t workers = o.gpus.size() * o.workers_per_gpu;
        batch_bytes = std::min<size_t>((size_t)48 << 20, std::max<size_t>((size_t)256 << 10, fsize / (6 * workers)));
    }
    if (const char* bb = getenv("CONSENT_BATCH_BYTES")) batch_bytes = (size_t)std::max(1ll, atoll(bb));      // tests: many small batches
    consent::PafStream paf(o.alignments, batch_bytes);
    if (!paf.ok()) { fprintf(stderr, "%s: cannot open %s\n", self, o.alignments.c_str()); return EXIT_FAILURE; }

    const double t_loaded = sh.now();
    BatchQueue queue(2 * o.gpus.size() * o.workers_per_gpu);
    OrderedWriter writer;
    sh.opt = &o; sh.store = &store; sh.queue = &queue; sh.writer = &writer;
    std::vector<std::thread> workers;
    for (unsigned w = 0; w < o.workers_per_gpu; ++w)
        for (int g : o.gpus) workers.emplace_back(gpu_main, &sh, g);
    size_t seq = 0;
    {
        Batch b;
        while (paf.next(&b.text)) {
            b.seq = seq++;
            queue.push(std::move(b));
            b = Batch();
            std::lock_guard<std::mutex> lk(sh.err_mu);
            if (!sh.error.empty()) break;
        }
    }
    queue.close();
    for (std::thread& t : workers) t.join();
    if (!sh.error.empty()) { fprintf(stderr, "%s: %s\n", self, sh.error.c_str()); return EXIT_FAILURE; }
    if (o.verbose || sh.error_windows)
        fprintf(stderr, "%s: %zu batches, %llu piles, %llu windows on %zu GPU(s)%s\n", self, seq, sh.piles, sh.windows, o.gpus.size(),
                sh.error_windows ? (", " + std::to_string(sh.error_windows) + " windows over a limit of this build were left uncorrected").c_str() : "");
    if (o.verbose) {
        const double span = sh.t_last - sh.t_first;
        fprintf(stderr, "{\"gpus\": %zu, \"batches\": %zu, \"piles\": %llu, \"windows\": %llu, \"load_s\": %.3f, \"first_batch_at_s\": %.3f, "
                        "\"handles_created_at_s\": %.3f, \"read_store_shipped_at_s\": %.3f, \"first_batch_done_at_s\": %.3f, "
                        "\"processing_s\": %.3f, \"total_s\": %.3f, \"windows_per_s_processing\": %.0f, \"windows_per_s_total\": %.0f}\n",
                o.gpus.size(), seq, sh.piles, sh.windows, t_loaded, sh.t_first, sh.t_created, sh.t_store, sh.t_first_done, span, sh.now(), span > 0 ? sh.windows / span : 0.0,
                sh.windows / sh.now());
    }
    return EXIT_SUCCESS;
}
