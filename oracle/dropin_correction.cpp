// dropin_correction.cpp — TEST INFRASTRUCTURE: the binding of INTEGRATION.md made real.
//
// The reference's CONSENT-correction binary with its per-window / per-read hot path replaced by two calls into
// libconsent_b200.so.  Everything else is the reference's OWN, UNMODIFIED host code, compiled from /root/reference where
// it lies (oracle/Makefile, target `dropin`): PAF pile reading (src/alignmentPiles.cpp), window positions and piles
// (src/alignmentWindows.cpp), read index / trimRead / dropRead (src/utils.cpp), reverse complement.
//
// processRead (src/CONSENT-correction.cpp:19-60) is cut in two around the GPU, as INTEGRATION.md describes:
//   phase A  per read: getSequencesMap, getAlignmentWindowsPositions, getAlignmentWindowsSequences  (:21-35)
//            -> the flat cg_batch / cg_reads arrays instead of one computeConsensusReadCorrection call per window
//   GPU      cg_correct_windows  (= every computeConsensusReadCorrection of the batch, :36)
//            cg_reanchor_reads   (= every alignConsensus of the batch, :47)
//   phase B  per read, in PAF order: trimRead / dropRead and the FASTA record (:50-59, :100-103)
// With -x phase A's window cutting moves to the device as well (cg_upload_piles): the host only reads the PAF
// piles (getNextReadPile: parsing, sort, top-maxSupport) and ships the read store once.
// With -X nothing of processRead / getNextReadPile is left on the host: the PAF text goes to the device as it is
// (cg_ingest_paf: parsing, grouping, std::sort + top-maxSupport), and re-anchoring + trimRead / dropRead run on what cg_run left
// in HBM (cg_finish_resident); the host reads two files and prints FASTA records.
// Same command line as bin/CONSENT-correction and bin/CONSENT-polishing (src/main.cpp:27-80; the two reference binaries differ in
// that processContig runs its windows on the CTPL pool and never trims, src/CONSENT-polishing.cpp:19-111): with -R (contigs in -r,
// reads in -R, as CONSENT-polish:197 calls it) this binary is the polisher.  tests/test_dropin_example.py runs it on a B200 and
// compares its FASTA byte for byte with the one the unmodified reference binary printed for the same PAF.
#include <getopt.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>

#include "alignmentPiles.h"      // reference src/
#include "alignmentWindows.h"
#include "utils.h"
#include "consent_b200.h"        // our ABI

static void die(const char* what, const char* why) { fprintf(stderr, "consent_correction_b200: %s: %s\n", what, why ? why : ""); exit(1); }

int main(int argc, char** argv) {
    std::string alignmentFile, readsFile, proofFile;
    unsigned minSupport = 3, maxSupport = 1000, windowSize = 500, merSize = 9, commonKMers = 8, minAnchors = 10,
             solidThresh = 4, windowOverlap = 50;                                       // src/main.cpp:15-24
    int opt, device = 0;
    bool deviceExtraction = false, deviceIngest = false;
    while ((opt = getopt(argc, argv, "a:A:d:k:s:S:M:l:f:e:p:c:m:j:w:r:R:n:i:g:xX")) != -1) {
        switch (opt) {
            case 'a': alignmentFile = optarg; break;
            case 's': minSupport = atoi(optarg); break;
            case 'S': maxSupport = atoi(optarg); break;
            case 'l': windowSize = atoi(optarg); break;
            case 'k': merSize = atoi(optarg); break;
            case 'c': commonKMers = atoi(optarg); break;
            case 'A': minAnchors = atoi(optarg); break;
            case 'f': solidThresh = atoi(optarg); break;
            case 'm': windowOverlap = atoi(optarg); break;
            case 'r': readsFile = optarg; break;
            case 'R': proofFile = optarg; break;
            case 'g': device = atoi(optarg); break;
            case 'x': deviceExtraction = true; break;
            case 'X': deviceIngest = true; break;
            default: break;                                                             // -j -M -p ...: no meaning here
        }
    }
    if (alignmentFile.empty() || readsFile.empty()) die("usage", "-a alignments.paf -r reads.fasta [-s -S -l -k -c -A -f -m as bin/CONSENT-correction] [-g gpu]");

    robin_hood::unordered_map<std::string, std::vector<bool>> readIndex;
    indexReads(readIndex, readsFile);
    // -R: the second read set of bin/CONSENT-polishing (contigs in -r, reads in -R; src/CONSENT-polishing.cpp:113-117, where
    // doTrimRead is false :19) and of bin/CONSENT-correction's proof mode (src/CONSENT-correction.cpp:76-79): no trim, no drop
    bool doTrimRead = true;
    if (!proofFile.empty()) { indexReads(readIndex, proofFile); doTrimRead = false; }
    std::ifstream alignments(alignmentFile);

    cg_params prm = {merSize, solidThresh, commonKMers, minAnchors};
    cg_handle* h = nullptr;
    if (cg_create(device, &prm, &h) != CG_OK) die("cg_create", cg_last_error(nullptr));

    if (deviceIngest) {
        // ---- everything between the two input files and the FASTA records on the device
        std::vector<std::string> names;
        std::vector<uint64_t> nameOff(1, 0), storeOff(1, 0);
        std::string nameBytes, store;
        for (auto& kv : readIndex) {                                                   // any order: indices are internal
            names.push_back(kv.first);
            nameBytes += kv.first; nameOff.push_back(nameBytes.size());
            store += fullnum2str(kv.second); storeOff.push_back(store.size());
        }
        std::string text((std::istreambuf_iterator<char>(alignments)), std::istreambuf_iterator<char>());
        if (nameBytes.empty()) nameBytes.push_back('x');
        if (store.empty()) store.push_back('A');
        cg_read_names rn = {(uint32_t)names.size(), nameOff.data(), nameBytes.data()};
        cg_pile_set ps; cg_corrected cor;
        if (cg_ingest_paf(h, text.data(), text.size(), &rn, maxSupport, &ps) != CG_OK) die("cg_ingest_paf", cg_last_error(h));
        cg_piles piles = {(uint32_t)names.size(), storeOff.data(), store.data(), ps.n_piles, ps.pile_read, ps.pile_qlen, ps.pile_ov_begin, ps.overlaps,
                          minSupport, windowSize, windowOverlap};
        if (cg_upload_piles(h, &piles) != CG_OK) die("cg_upload_piles", cg_last_error(h));
        if (cg_run(h) != CG_OK) die("cg_run", cg_last_error(h));
        if (cg_finish_resident(h, doTrimRead ? 1 : 0, &cor) != CG_OK) die("cg_finish_resident", cg_last_error(h));
        for (uint32_t p = 0; p < ps.n_piles; ++p)
            if (cor.read_off[p + 1] != cor.read_off[p]) {
                std::cout << ">" << names[ps.pile_read[p]] << std::endl;
                std::cout.write(cor.bases + cor.read_off[p], (std::streamsize)(cor.read_off[p + 1] - cor.read_off[p]));
                std::cout << std::endl;
            }
        cg_free_corrected(&cor); cg_free_pile_set(&ps); cg_destroy(h);
        return 0;
    }

    if (deviceExtraction) {
        // ---- phase A on the device: the store = every indexed read, decoded as getSequencesMap decodes them
        robin_hood::unordered_map<std::string, uint32_t> id;
        std::vector<uint64_t> storeOff(1, 0);
        std::string store;
        auto readId = [&](const std::string& name) {
            auto it = id.find(name);
            if (it != id.end()) return it->second;
            uint32_t r = (uint32_t)id.size();
            id[name] = r;
            store += fullnum2str(readIndex[name]);
            storeOff.push_back(store.size());
            return r;
        };
        std::vector<uint32_t> pileRead, pileQlen, pileOvBegin(1, 0);
        std::vector<cg_overlap> ovs;
        std::vector<std::string> names;
        while (!alignments.eof()) {
            std::vector<Overlap> al = getNextReadPile(alignments, maxSupport);
            if (al.size() == 0) continue;
            names.push_back(al.begin()->qName);
            pileRead.push_back(readId(al.begin()->qName));
            pileQlen.push_back(al.begin()->qLength);
            for (const Overlap& a : al) {
                cg_overlap o = {readId(a.tName), a.strand ? 1u : 0u, a.qStart, a.qEnd, a.tStart, a.tEnd, a.tLength};
                ovs.push_back(o);
            }
            pileOvBegin.push_back((uint32_t)ovs.size());
        }
        if (store.empty()) store.push_back('A');
        cg_piles piles = {(uint32_t)id.size(), storeOff.data(), store.data(), (uint32_t)names.size(), pileRead.data(), pileQlen.data(),
                          pileOvBegin.data(), ovs.data(), minSupport, windowSize, windowOverlap};
        cg_results res; cg_window_set ws; cg_corrected cor;
        if (cg_upload_piles(h, &piles) != CG_OK) die("cg_upload_piles", cg_last_error(h));
        if (cg_run(h) != CG_OK) die("cg_run", cg_last_error(h));
        if (cg_download(h, &res) != CG_OK) die("cg_download", cg_last_error(h));
        if (cg_download_windows(h, 0, &ws) != CG_OK) die("cg_download_windows", cg_last_error(h));
        if (cg_reanchor_reads(h, &ws.batch, &res, &ws.reads, &cor) != CG_OK) die("cg_reanchor_reads", cg_last_error(h));
        for (size_t r = 0; r < names.size(); ++r) {
            if (ws.reads.read_win_begin[r + 1] == ws.reads.read_win_begin[r]) continue;
            std::string corrected(cor.bases + cor.read_off[r], cor.bases + cor.read_off[r + 1]);
            if (doTrimRead) { corrected = trimRead(corrected, 1); if (dropRead(corrected)) corrected.clear(); }
            if (corrected.length() != 0) std::cout << ">" << names[r] << std::endl << corrected << std::endl;
        }
        cg_free_corrected(&cor); cg_free_window_set(&ws); cg_free_results(&res); cg_destroy(h);
        return 0;
    }

    // ---- phase A
    std::vector<uint32_t> winSeqBegin(1, 0), readWinBegin(1, 0), winPos;
    std::vector<uint64_t> seqOff(1, 0), readOff(1, 0);
    std::string bases, readBases;
    std::vector<std::string> readIds;
    while (!alignments.eof()) {
        std::vector<Overlap> al = getNextReadPile(alignments, maxSupport);
        if (al.size() == 0) continue;
        std::string readId = al.begin()->qName;
        robin_hood::unordered_map<std::string, std::string> sequences = getSequencesMap(al, readIndex);
        std::vector<std::pair<unsigned, unsigned>> pilesPos =
            getAlignmentWindowsPositions(al.begin()->qLength, al, minSupport, maxSupport, windowSize, windowOverlap);
        for (unsigned i = 0; i < pilesPos.size(); i++) {
            std::vector<std::string> pile = getAlignmentWindowsSequences(al, minSupport, windowSize, windowOverlap, sequences,
                                                                         pilesPos[i].first, pilesPos[i].second, merSize, maxSupport, commonKMers);
            if (pile.empty()) die("empty pile", "the reference dereferences curPile[0] here (CONSENT-correction.cpp:36)");
            for (const std::string& s : pile) { bases += s; seqOff.push_back(bases.size()); }
            winSeqBegin.push_back((uint32_t)(seqOff.size() - 1));
            winPos.push_back(pilesPos[i].first);
        }
        readIds.push_back(readId);
        readBases += sequences[al[0].qName];
        readOff.push_back(readBases.size());
        readWinBegin.push_back((uint32_t)winPos.size());
    }

    // ---- GPU
    if (bases.empty()) bases.push_back('A');
    if (readBases.empty()) readBases.push_back('A');
    if (winPos.empty()) winPos.push_back(0);
    cg_batch in = {(uint32_t)(winSeqBegin.size() - 1), winSeqBegin.data(), seqOff.data(), bases.data()};
    cg_results res;
    if (cg_correct_windows(h, &in, &res) != CG_OK) die("cg_correct_windows", cg_last_error(h));
    cg_reads rd = {(uint32_t)readIds.size(), readWinBegin.data(), readOff.data(), readBases.data(), winPos.data(), windowSize, windowOverlap};
    cg_corrected cor;
    if (cg_reanchor_reads(h, &in, &res, &rd, &cor) != CG_OK) die("cg_reanchor_reads", cg_last_error(h));

    // ---- phase B (CONSENT-correction.cpp:50-59,100-103): doTrimRead is true without a proof file
    for (size_t r = 0; r < readIds.size(); ++r) {
        if (readWinBegin[r + 1] == readWinBegin[r]) continue;                           // :23-25: no window, nothing printed
        std::string corrected(cor.bases + cor.read_off[r], cor.bases + cor.read_off[r + 1]);
        if (doTrimRead) { corrected = trimRead(corrected, 1); if (dropRead(corrected)) corrected.clear(); }
        if (corrected.length() != 0) std::cout << ">" << readIds[r] << std::endl << corrected << std::endl;
    }
    cg_free_corrected(&cor);
    cg_free_results(&res);
    cg_destroy(h);
    return 0;
}
