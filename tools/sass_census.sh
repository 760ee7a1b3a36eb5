#!/bin/bash
# Instruction census of the built library per object file: the tcgen05 / TMA / mbarrier mnemonics (B200_PROFILING.md)
# usage: tools/sass_census.sh > profiles/sass_census.txt
cd "$(dirname "$0")/../plotpointe-gat-recommendation_b200/csrc"
echo "# cuobjdump -sass census of libb200gat.so objects (sm_100a), $(date -u +%Y-%m-%dT%H:%MZ), nvcc $(/usr/local/cuda/bin/nvcc --version | grep -o 'V[0-9][0-9.]*' | tail -1)"
echo "# columns: object | UTCHMMA (tcgen05.mma) | UTCQMMA | LDTM (tcgen05.ld) | STTM (tcgen05.st) | UTCBAR (tcgen05.commit) | UBLKCP (cp.async.bulk) | UTMALDG (TMA tensor load) | SYNCS (mbarrier) | LDGSTS (cp.async) | HMMA/IMMA (legacy mma.sync) | total SASS lines"
for o in _obj/*.o; do
  s=$(/usr/local/cuda/bin/cuobjdump -sass "$o" 2>/dev/null)
  c() { echo "$s" | grep -c -E "$1"; }
  printf "%-16s | %5d | %3d | %4d | %4d | %4d | %4d | %4d | %5d | %5d | %3d | %7d\n" "$(basename $o .o)" "$(c '\bUTCHMMA')" "$(c '\bUTCQMMA')" "$(c '\bLDTM')" "$(c '\bSTTM')" "$(c '\bUTCBAR')" "$(c '\bUBLKCP')" "$(c '\bUTMALDG')" "$(c '\bSYNCS')" "$(c '\bLDGSTS')" "$(c '\b(HMMA|IMMA)\b')" "$(echo "$s" | grep -c -E '^\s+/\*[0-9a-f]{4}\*/')"
done
echo
echo "# kernels per object (name, registers, from -Xptxas -v build logs)"
for l in _obj/*.log; do
  echo "## $(basename $l .log)"
  awk "/Compiling entry function/ {match(\$0, /'[^']+'/); name=substr(\$0, RSTART+1, RLENGTH-2)} /Used [0-9]+ registers/ {match(\$0, /Used [0-9]+ registers/); print name, substr(\$0, RSTART+5, RLENGTH-15)}" "$l" | while read m r; do echo "  $(echo $m | /usr/local/cuda/bin/cu++filt | cut -c1-120)  regs=$r"; done
done
