#!/bin/bash
# How tests/golden/pipeline_md5.json was made (needs /root/reference; run from the repository root after `make -C oracle ref dropin pipeline`).
# Every FASTA below is printed by an UNMODIFIED reference binary; nothing of this repository's path is involved.
set -e
R=oracle/_ref; W=${1:-/tmp/consent_pipeline_golden}; J=${2:-8}
mkdir -p $W/ex $W/c4 $W/c5
CF="-s 3 -S 150 -l 500 -k 9 -c 8 -A 2 -f 4 -m 50 -M 150"            # CONSENT-correct:42-50,202
PF="-s 1 -S 20000 -l 500 -k 9 -c 8 -A 2 -f 4 -m 50 -M 150"          # CONSENT-polish:42-50,197
MM="--dual=yes -PD --no-long-join -w5 -g1000 -m30 -n1 -I1G"         # CONSENT-correct:185 (PB), CONSENT-polish:186
# config 1: the shipped example
$R/minimap2 $MM -t$J $R/example/reads.fasta $R/example/reads.fasta > $W/ex/ex.paf 2>/dev/null
head -2000 $W/ex/ex.paf > $W/ex/small.paf
$R/consent_correction_ref -a $W/ex/ex.paf $CF -j $J -r $R/example/reads.fasta -p /nonexistent > $W/ex/ref.fasta
$R/consent_correction_ref -a $W/ex/small.paf $CF -j $J -r $R/example/reads.fasta -p /nonexistent > $W/ex/small_ref.fasta
md5sum $W/ex/ex.paf $W/ex/ref.fasta $W/ex/small_ref.fasta; grep -c ">" $W/ex/ref.fasta $W/ex/small_ref.fasta
# config 4: 30x PacBio over 5 Mb
python tools/make_pipeline_data.py config4 $W/c4
$R/minimap2 $MM -t$J $W/c4/reads.fasta $W/c4/reads.fasta > $W/c4/c4.paf 2>/dev/null
$R/consent_correction_ref -a $W/c4/c4.paf $CF -j $J -r $W/c4/reads.fasta -p /nonexistent > $W/c4/ref.fasta
md5sum $W/c4/reads.fasta $W/c4/c4.paf $W/c4/ref.fasta; grep -c ">" $W/c4/ref.fasta
# config 5: 50 contigs x 100 kb polished with 30x ONT
python tools/make_pipeline_data.py config5 $W/c5
$R/minimap2 $MM -t$J $W/c5/contigs.fasta $W/c5/reads.fasta > $W/c5/raw.paf 2>/dev/null
LC_COLLATE=C sort -k6,6 $W/c5/raw.paf > $W/c5/sorted.paf
$R/reformatPAF_ref $W/c5/sorted.paf $W/c5/c5.paf
$R/consent_polishing_ref -a $W/c5/c5.paf $PF -j $J -r $W/c5/contigs.fasta -R $W/c5/reads.fasta -p /nonexistent > $W/c5/ref.fasta
md5sum $W/c5/contigs.fasta $W/c5/reads.fasta $W/c5/ref.fasta; grep -c ">" $W/c5/ref.fasta
