// reads.h — the host's read store: what the reference keeps in `readIndex` (src/utils.cpp:166-204, indexReads), as flat arrays
// the C ABI takes (cg_set_read_store, cg_read_names).
//
// Same parsing rules as indexReads, because they decide bytes of the output:
//   * a record starts at a line whose first character is dropped ('>' or '@'); the name is the header up to the first blank;
//   * sequence lines are concatenated until a line that is empty or starts with '>' or '+' (multi-line FASTA / FASTQ);
//   * after a '+' line as many quality lines are skipped as there were sequence lines;
//   * an empty line where a header is expected ends the file;
//   * a name listed twice keeps its LAST sequence (index[header] = ...);
//   * lines end at '\n' only: a '\r' of a CRLF file stays in the sequence (and is stored as T, like any non-ACG byte).
// Bases are kept as they are in the file; the device normalises them the way fullstr2num does (upper case, non-ACG -> T).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace consent {

struct ReadStore {
    std::vector<std::string> names;          // index = store index
    std::vector<uint64_t> off{0};            // [n + 1] offsets into bases
    std::string bases;
    std::vector<uint64_t> name_off{0};       // [n + 1] offsets into name_bytes
    std::string name_bytes;
    std::unordered_map<std::string, uint32_t> id;

    uint32_t size() const { return (uint32_t)names.size(); }
    // Adds every record of a FASTA / FASTQ file; false if the file cannot be opened.
    bool load(const std::string& path, std::string* err);
    // Builds name_off / name_bytes (after the last load).
    void finish();
};

}  // namespace consent
