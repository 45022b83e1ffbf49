// reads.cpp — see reads.h.  Restates indexReads (reference src/utils.cpp:166-204) over a memory-mapped-style block reader.
#include "reads.h"

#include <cstdio>
#include <cstring>

namespace consent {

namespace {

// getline over a whole file read in one piece (read sets are a few GB at most and are kept in memory anyway).
struct Lines {
    std::string data;
    size_t pos = 0;
    bool eof = false;
    // false once nothing is left (like a failed std::getline, which leaves the string empty)
    bool next(const char** p, size_t* n) {
        if (pos >= data.size()) { eof = true; *p = data.data() + data.size(); *n = 0; return false; }
        const char* b = data.data() + pos;
        const char* e = (const char*)memchr(b, '\n', data.size() - pos);
        if (!e) { *p = b; *n = data.size() - pos; pos = data.size(); return true; }
        *p = b; *n = (size_t)(e - b); pos = (size_t)(e - data.data()) + 1;
        return true;
    }
};

bool slurp(const std::string& path, std::string* out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::string& s = *out;
    s.clear();
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return true;
}

}  // namespace

bool ReadStore::load(const std::string& path, std::string* err) {
    Lines in;
    if (!slurp(path, &in.data)) { if (err) *err = "cannot open " + path; return false; }
    const char* p; size_t n;
    in.next(&p, &n);                                         // getline(f, header)
    std::string header(p, n);
    while (header.length() > 0) {                            // utils.cpp:172
        header.erase(0, 1);
        const size_t blank = header.find(' ');
        const std::string name = blank == std::string::npos ? header : header.substr(0, blank);   // splitString(header, " ")[0]
        std::string sequence;
        in.next(&p, &n);                                     // first sequence line
        sequence.assign(p, n);
        int nbLines = 1;
        in.next(&p, &n);
        while (n > 0 && p[0] != '>' && p[0] != '+') {        // :181-185
            sequence.append(p, n);
            nbLines++;
            in.next(&p, &n);
        }
        // index[header] = sequence: a repeated name replaces the earlier sequence (its bytes stay unused in `bases`)
        auto it = id.find(name);
        if (it == id.end()) {
            id.emplace(name, (uint32_t)names.size());
            names.push_back(name);
            bases += sequence;
            off.push_back(bases.size());
        } else {
            // keep indices stable: append the new bytes and repoint this read at them
            const uint32_t r = it->second;
            if (r + 1 == names.size()) { bases.resize(off[r]); bases += sequence; off[r + 1] = bases.size(); }
            else {
                // rare (duplicate names far apart): rebuild the tail so that offsets stay monotone
                std::vector<std::string> tail;
                for (uint32_t k = r + 1; k < names.size(); ++k) tail.emplace_back(bases, off[k], off[k + 1] - off[k]);
                bases.resize(off[r]); bases += sequence; off[r + 1] = bases.size();
                for (uint32_t k = r + 1; k < names.size(); ++k) { bases += tail[k - r - 1]; off[k + 1] = bases.size(); }
            }
        }
        if (n > 0 && p[0] == '+') {                          // FASTQ: skip the quality lines (:192-199)
            in.next(&p, &n);
            for (int i = 1; i < nbLines; i++) in.next(&p, &n);
            in.next(&p, &n);
        }
        header.assign(p, n);                                 // the next header has been read
        if (in.eof && n == 0) break;
    }
    return true;
}

void ReadStore::finish() {
    name_off.assign(1, 0);
    name_bytes.clear();
    for (const std::string& s : names) { name_bytes += s; name_off.push_back(name_bytes.size()); }
}

}  // namespace consent
