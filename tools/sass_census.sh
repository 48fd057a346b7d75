#!/bin/bash
# SASS census of the product library: Blackwell-native mnemonics per kernel (cuobjdump -sass), so a reviewer need not
# rebuild.  UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA load / store / reduce,
# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.   usage: tools/sass_census.sh > profiles/r2_sass_census.txt
cd "$(dirname "$0")/.."
LIB=${1:-clip-mixer_b200/libmixerclip.so}
echo "# $LIB  ($(stat -c %s $LIB) bytes)  built by: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
echo "# git $(git rev-parse --short HEAD 2>/dev/null)  $(date -u +%F)"
cuobjdump -sass $LIB | awk '
  /Function :/ { fn=$3; next }
  { for (i = 1; i <= NF; ++i) {
      t=$i; sub(/\..*/, "", t);
      if (t=="UTCHMMA"||t=="UTCQMMA"||t=="LDTM"||t=="STTM"||t=="UTMALDG"||t=="UTMASTG"||t=="UTMAREDG"||t=="UTMAPF"||t=="UTCBAR"||t=="SYNCS"||t=="R2UR"||t=="MUFU"||t=="HMMA"||t=="FFMA2"||t=="UCGABAR_WAIT") c[fn" "t]++;
    }
    n[fn]++ }
  END { for (k in n) { printf "%s  instr=%d", k, n[k];
          split("UTCHMMA LDTM STTM UTMALDG UTMASTG UTMAREDG UTMAPF UTCBAR SYNCS R2UR MUFU FFMA2 HMMA UCGABAR_WAIT", m, " ");
          for (j = 1; j <= 14; ++j) if (c[k" "m[j]] > 0) printf "  %s=%d", m[j], c[k" "m[j]];
          printf "\n" } }' | sort | c++filt | sed 's/mc::(anonymous namespace):://; s/CUtensorMap_st/TM/g' | cut -c1-260
