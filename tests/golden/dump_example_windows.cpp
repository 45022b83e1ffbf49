// Fixture generator (build container only): runs the reference's own PAF parsing and window extraction
// (src/alignmentPiles.cpp, src/alignmentWindows.cpp, src/utils.cpp — unmodified, compiled in place) over the
// shipped example and writes the piles of the first reads' windows, one window per block:
//   W <n_seqs>\n<seq>\n...   (pile[0] = template), exactly what processRead hands to computeConsensusReadCorrection.
#include <fstream>
#include <iostream>
#include "alignmentPiles.h"
#include "alignmentWindows.h"
#include "utils.h"
robin_hood::unordered_map<std::string, std::vector<bool>> readIndex;
int main(int argc, char** argv) {
    std::string paf = argv[1], reads = argv[2];
    unsigned maxWindows = atoi(argv[3]), skipReads = atoi(argv[4]);
    unsigned minSupport = 3, maxSupport = 150, windowSize = 500, merSize = 9, commonKMers = 8, windowOverlap = 50;   // CONSENT-correct:42-50
    indexReads(readIndex, reads);
    std::ifstream alignments(paf);
    unsigned nw = 0, nreads = 0;
    while (!alignments.eof() && nw < maxWindows) {
        std::vector<Overlap> al = getNextReadPile(alignments, maxSupport);
        if (al.size() == 0) continue;
        if (nreads++ < skipReads) continue;
        robin_hood::unordered_map<std::string, std::string> sequences = getSequencesMap(al, readIndex);
        std::vector<std::pair<unsigned, unsigned>> pilesPos = getAlignmentWindowsPositions(al.begin()->qLength, al, minSupport, maxSupport, windowSize, windowOverlap);
        for (unsigned i = 0; i < pilesPos.size() && nw < maxWindows; i++) {
            std::vector<std::string> pile = getAlignmentWindowsSequences(al, minSupport, windowSize, windowOverlap, sequences, pilesPos[i].first, pilesPos[i].second, merSize, maxSupport, commonKMers);
            if (pile.size() == 0) continue;
            std::cout << "W " << pile.size() << "\n";
            for (auto& s : pile) std::cout << s << "\n";
            nw++;
        }
    }
    return 0;
}
