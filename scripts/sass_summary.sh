#!/bin/bash
# SASS evidence of the Blackwell-native kernels (B200_PROFILING.md "What proves a Blackwell-native kernel"): opcode counts per
# kernel of the built library.  Usage: scripts/sass_summary.sh > profiles/sass_summary.txt   (needs only cuobjdump, no GPU)
set -e
cd "$(dirname "$0")/.."
LIB=nopesac_b200/libnopesac_b200.so
[ -f "$LIB" ] || python -c "from nopesac_b200 import build; build.build()"
echo "# cuobjdump -sass $LIB  (sm_100a) — opcode counts per kernel; UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,"
echo "# UTMALDG/UBLKCP = TMA (cp.async.bulk[.tensor]), HMMA = mma.sync, FFMA2/FADD2/FMUL2 = packed fp32, MUFU = SFU"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3; gsub(/^_Z[0-9]*N?[0-9]*_GLOBAL__N__[0-9a-f_]*/, "", fn); next }
  {
    for (i = 1; i <= NF; i++) {
      op=$i
      if (op ~ /^(UTC[A-Z]*MMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|HMMA|FFMA2|FADD2|FMUL2|MUFU|SYNCS|LDGSTS|UTCATOMSWS|REDG|ATOMG|ST|STG|LDG)(\.|$)/) {
        split(op, a, "."); key=a[1]; cnt[fn SUBSEP key]++; fns[fn]=1; keys[key]=1
      }
    }
  }
  END {
    n=0; for (k in keys) order[n++]=k
    for (f in fns) {
      line=""
      for (i=0;i<n;i++) if ((f SUBSEP order[i]) in cnt) line=line " " order[i] "=" cnt[f SUBSEP order[i]]
      print f ":" line
    }
  }' | sort
