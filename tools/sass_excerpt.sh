#!/bin/bash
# Instruction-mix evidence from the built library: per kernel the count of the SASS mnemonics that show which hardware
# path it uses (TMA bulk copies, cp.async, packed 16x2 min/max, POPC, fp64 tensor-core MMA, ...).
#   bash tools/sass_excerpt.sh > profiles/rNN_vM_sass_excerpt.txt
LIB=stereo-visual-slam_b200/libvslam_b200.so
echo "cuobjdump -sass $LIB  (sm_100a), mnemonic counts per kernel"
cuobjdump -sass $LIB | awk '
/Function :/ { fn=$3; sub(/^_Z[0-9]+/, "", fn); n=split(fn, a, /[0-9]/); name=$3 }
/^[[:space:]]+\/\*[0-9a-f]+\*\// {
  op=$2; if (op ~ /^@/) op=$3; sub(/;.*/, "", op);
  tot[name]++;
  if (op ~ /^UBLKCP/) c[name,"UBLKCP(cp.async.bulk=TMA)"]++;
  if (op ~ /^UTMALDG|^UTMASTG/) c[name,"UTMA*(tensor-map TMA)"]++;
  if (op ~ /^LDGSTS/) c[name,"LDGSTS(cp.async)"]++;
  if (op ~ /^SYNCS/) c[name,"SYNCS(mbarrier)"]++;
  if (op ~ /^VIMNMX3/) c[name,"VIMNMX3"]++;
  if (op ~ /^VIMNMX\./ || op == "VIMNMX") c[name,"VIMNMX"]++;
  if (op ~ /^VIADD\.16x2|^VIADDMNMX/) c[name,"VIADD.16x2"]++;
  if (op ~ /^POPC/) c[name,"POPC"]++;
  if (op ~ /^CREDUX/) c[name,"CREDUX"]++;
  if (op ~ /^DMMA/) c[name,"DMMA(fp64 tensor)"]++;
  if (op ~ /^UTC.*MMA|^TCGEN/) c[name,"tcgen05"]++;
  if (op ~ /^DFMA/) c[name,"DFMA"]++;
  if (op ~ /^LDG\.E\.128|^LDG\.E\.CONSTANT\.128|^LDG.*\.128/) c[name,"LDG.128"]++;
  if (op ~ /^STG.*\.128/) c[name,"STG.128"]++;
  if (op ~ /^IDP/) c[name,"IDP(dp2a/dp4a)"]++;
  if (op ~ /^ATOM|^RED/) c[name,"ATOM/RED"]++;
}
END {
  for (k in tot) {
    line=""; 
    for (key in c) { split(key, kk, SUBSEP); if (kk[1]==k) line=line sprintf("  %s=%d", kk[2], c[key]); }
    printf "%s  [%d instructions]%s\n", k, tot[k], line;
  }
}' | sort | c++filt 2>/dev/null | sed 's/(.*)  \[/  [/'
