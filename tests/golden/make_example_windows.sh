#!/bin/bash
# Regenerates tests/golden/example_windows.txt.gz (build container only: needs /root/reference).
# The piles are produced by the reference's OWN code: vendored minimap2 with the PacBio flags of CONSENT-correct:185,
# then src/alignmentPiles.cpp + src/alignmentWindows.cpp (getNextReadPile, getSequencesMap,
# getAlignmentWindowsPositions, getAlignmentWindowsSequences) exactly as processRead calls them
# (src/CONSENT-correction.cpp:19-35).  Nothing of the reference is copied into the repo: sources are compiled in place.
set -e
REF=${REF:-/root/reference}
T=$(mktemp -d)
cp -r $REF/minimap2/. $T/mm2 2>/dev/null || { mkdir -p $T/mm2; cp -r $REF/minimap2/* $T/mm2/; }
make -C $T/mm2 -j8 > $T/mm2.log 2>&1
$T/mm2/minimap2 --dual=yes -PD --no-long-join -w5 -g1000 -m30 -n1 -t8 $REF/example/reads.fasta $REF/example/reads.fasta > $T/example.paf 2> $T/mm2.err
g++ -std=c++11 -O2 -w -include cstdint -I$REF/src -o $T/dump_windows "$(dirname "$0")/dump_example_windows.cpp" \
    $REF/src/alignmentPiles.cpp $REF/src/alignmentWindows.cpp $REF/src/utils.cpp $REF/src/reverseComplement.cpp
# 300 windows starting at the 21st read pile of the PAF
$T/dump_windows $T/example.paf $REF/example/reads.fasta 300 20 | gzip -9 > "$(dirname "$0")/example_windows.txt.gz"
rm -rf $T
