// paf_tools.cpp — the three PAF text helpers the reference's wrappers call, one binary, dispatched on the
// name it is called by (bin/reformatPAF, bin/explode, bin/merge are links to it) or on a first argument of that name:
//
//   reformatPAF in.paf out.paf                 swaps query and target columns (1-4 <-> 6-9; strand stays in column 5; the rest
//                                              follows) so that the contigs of a reads-to-contigs mapping become the piles' queries
//                                              (reference src/reformatPAF.cpp:22-33, CONSENT-polish:189-193)
//   explode in.paf prefix                      cuts the PAF of a split minimap2 index into prefix_1, prefix_2, ...: a new file
//                                              starts when a query comes back after other queries (src/explode.cpp:14-53)
//   merge out.paf headers prefix_1 prefix_2 .. for every read header, in order, the lines of that query from every chunk file
//                                              (src/merge.cpp:32-63), so that a read's overlaps are consecutive again
// Plain sequential text plumbing: it stays on the host.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace {

int reformat_paf(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: reformatPAF in.paf out.paf\n"); return 1; }
    std::ifstream in(argv[1]);
    std::ofstream out(argv[2]);
    if (!in || !out) { fprintf(stderr, "reformatPAF: cannot open the files\n"); return 1; }
    std::string line, token;
    std::vector<std::string> v;
    while (std::getline(in, line)) {
        v.clear();
        std::stringstream iss(line);
        while (std::getline(iss, token, '\t')) v.push_back(token);       // a trailing empty field is dropped, as in the reference
        if (v.size() < 9) { fprintf(stderr, "reformatPAF: a line with fewer than 9 columns\n"); return 1; }
        out << v[5] << '\t' << v[6] << '\t' << v[7] << '\t' << v[8] << '\t' << v[4] << '\t' << v[0] << '\t' << v[1] << '\t' << v[2] << '\t' << v[3];
        for (size_t i = 9; i < v.size(); ++i) out << '\t' << v[i];
        out << '\n';
    }
    return 0;
}

std::string first_column(const std::string& line) {
    const size_t t = line.find('\t');
    return t == std::string::npos ? line : line.substr(0, t);
}

int explode(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: explode in.paf prefix\n"); return 1; }
    std::ifstream f(argv[1]);
    if (!f) { fprintf(stderr, "explode: cannot open %s\n", argv[1]); return 1; }
    const std::string prefix = argv[2];
    std::set<std::string> seen;
    int n_files = 1;
    std::ofstream cur((prefix + "_" + std::to_string(n_files)).c_str());
    std::string line, cur_read, old_read, pending;
    std::getline(f, line);
    while (line.length() > 0) {                                          // an empty line ends the input (src/explode.cpp:24)
        old_read = cur_read;
        cur_read = first_column(line);
        if (old_read.empty() || cur_read == old_read) {
            pending += line; pending += '\n';
            if (!std::getline(f, line)) line.clear();
        } else {
            seen.insert(old_read);
            cur << pending;
            pending = line; pending += '\n';
            if (!std::getline(f, line)) line.clear();
            if (seen.count(cur_read)) {                                  // this query was already written: the index had been split here
                seen.clear();
                cur.close();
                ++n_files;
                cur.open((prefix + "_" + std::to_string(n_files)).c_str());
            }
        }
    }
    if (!pending.empty()) cur << pending;
    return 0;
}

int merge(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: merge out.paf headers chunk_1 [chunk_2 ...]\n"); return 1; }
    std::ofstream out(argv[1]);
    std::ifstream headers(argv[2]);
    if (!out || !headers) { fprintf(stderr, "merge: cannot open the files\n"); return 1; }
    std::vector<std::ifstream> files;
    for (int i = 3; i < argc; ++i) files.emplace_back(argv[i]);
    std::vector<std::string> held(files.size());                        // a line read ahead that belongs to a later query
    std::vector<bool> has(files.size(), false);
    std::string header, line;
    while (std::getline(headers, header)) {
        header = header.empty() ? header : header.substr(1);             // drop '>' / '@'; the reference compares the WHOLE rest of the line
        for (size_t i = 0; i < files.size(); ++i) {
            for (;;) {
                if (!has[i]) { if (!std::getline(files[i], line)) break; held[i] = line; has[i] = true; }
                if (first_column(held[i]) != header) break;              // stays held (the reference seeks back by one line)
                out << held[i] << '\n';
                has[i] = false;
            }
        }
    }
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    const char* self = strrchr(argv[0], '/');
    self = self ? self + 1 : argv[0];
    std::string tool = self;
    if (tool != "reformatPAF" && tool != "explode" && tool != "merge" && argc > 1) { tool = argv[1]; ++argv; --argc; }
    if (tool == "reformatPAF") return reformat_paf(argc, argv);
    if (tool == "explode") return explode(argc, argv);
    if (tool == "merge") return merge(argc, argv);
    fprintf(stderr, "usage: paf_tools reformatPAF|explode|merge ...\n");
    return 1;
}
