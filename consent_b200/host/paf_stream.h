// paf_stream.h — the PAF file in bounded batches of WHOLE piles.
//
// The reference reads one pile at a time on its main thread (getNextReadPile, src/alignmentPiles.cpp:22-58) and keeps a ring
// of 100 000 read jobs in flight (src/CONSENT-correction.cpp:76-127).  Here the unit handed to a GPU is a batch of consecutive
// piles: the text is cut where getNextReadPile would start a new pile anyway — in front of a line whose query name (column 1)
// differs from the line before, or after an empty line — so cg_ingest_paf sees exactly the piles the reference would, batch
// after batch, and memory stays bounded by the batch size (+ one pile).
#pragma once
#include <cstdio>
#include <string>

namespace consent {

class PafStream {
public:
    PafStream(const std::string& path, size_t target_bytes);
    ~PafStream();
    bool ok() const { return f_ != nullptr; }
    // The next batch (always ends with '\n'); false when the file is exhausted.
    bool next(std::string* out);

private:
    bool line(const char** p, size_t* n);       // next line without its '\n'; false at the end of the file
    void unget(size_t n_with_nl) { pos_ -= n_with_nl; }
    FILE* f_ = nullptr;
    std::string buf_;
    size_t pos_ = 0, target_;
    bool eof_ = false;
};

}  // namespace consent
