#!/bin/bash
# Counts of the Blackwell-specific SASS instructions in the built library, per cubin (proof that the GEMM is genuinely
# tcgen05 / TMEM / TMA, the Adam exchange TMA-bulk / multimem):  tools/sass_summary.sh > profiles/rNN_sass_summary.txt
set -e
cd "$(dirname "$0")/.."
tmp=$(mktemp -d)
(cd "$tmp" && cuobjdump -xelf all "$OLDPWD/cellbender_b200/libcellbender_b200.so" > /dev/null)
echo "# cuobjdump -sass of cellbender_b200/libcellbender_b200.so (sm_100a), instruction counts per cubin"
echo "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc,"
echo "# UTMALDG / UTMASTG / UTMAREDG = cp.async.bulk.tensor load / store / reduce-add, UBLKCP = cp.async.bulk (1-D),"
echo "# SYNCS.* = mbarrier, LDGMC...ADD = multimem.ld_reduce, STG...STRONG.SYS at the multicast address = multimem.st"
for f in "$tmp"/*.cubin; do
  n=$(basename "$f" .sm_100a.cubin)
  cuobjdump -sass "$f" | grep -E "^\s*/\*[0-9a-f]{4}\*/" | sed -E 's/^\s*\/\*[0-9a-f]+\*\/\s*(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' \
    | grep -E "^(UTC|LDTM|UTMA|UBLKCP|SYNCS|LDGMC|MUFU)" | sed -E 's/^(MUFU)\..*/\1/' | sort | uniq -c | awk -v n="$n" '{printf "%-18s %6d  %s\n", n, $1, $2}'
done
rm -rf "$tmp"
