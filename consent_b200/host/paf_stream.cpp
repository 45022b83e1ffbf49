// paf_stream.cpp — see paf_stream.h.
#include "paf_stream.h"

#include <cstring>

namespace consent {

PafStream::PafStream(const std::string& path, size_t target_bytes) : target_(target_bytes ? target_bytes : 1) {
    f_ = fopen(path.c_str(), "rb");
}

PafStream::~PafStream() {
    if (f_) fclose(f_);
}

bool PafStream::line(const char** p, size_t* n) {
    for (;;) {
        const char* b = buf_.data() + pos_;
        const char* e = pos_ < buf_.size() ? (const char*)memchr(b, '\n', buf_.size() - pos_) : nullptr;
        if (e) { *p = b; *n = (size_t)(e - b); pos_ += *n + 1; return true; }
        if (eof_) {
            if (pos_ >= buf_.size()) return false;
            buf_.push_back('\n');                       // a last line without a newline: std::getline returns it all the same
            continue;
        }
        // refill: drop what has been consumed, read another block
        buf_.erase(0, pos_);
        pos_ = 0;
        const size_t old = buf_.size(), block = 8u << 20;
        buf_.resize(old + block);
        const size_t got = fread(&buf_[old], 1, block, f_);
        buf_.resize(old + got);
        if (got == 0) eof_ = true;
    }
}

bool PafStream::next(std::string* out) {
    out->clear();
    if (!f_) return false;
    std::string last;                                   // query name of the previous line; empty = a pile boundary
    bool have = false;
    const char* p; size_t n;
    while (line(&p, &n)) {
        const char* tab = (const char*)memchr(p, '\t', n);
        const size_t qn = n == 0 ? 0 : (tab ? (size_t)(tab - p) : n);
        if (out->size() >= target_ && have) {
            const bool same = n != 0 && !last.empty() && qn == last.size() && memcmp(p, last.data(), qn) == 0;
            if (!same) { unget(n + 1); break; }         // this line starts another pile: it opens the next batch
        }
        out->append(p, n);
        out->push_back('\n');
        have = true;
        last.assign(p, qn);                             // empty line -> "": whatever follows is a new pile
    }
    return !out->empty();
}

}  // namespace consent
