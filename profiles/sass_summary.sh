#!/bin/sh
# Blackwell-native evidence per object file: counts of the SASS mnemonics that only tcgen05 / TMEM / TMA code contains.
#   sh profiles/sass_summary.sh > profiles/r02_sass_summary.txt      (after `python __graft_entry__.py build`)
cd "$(dirname "$0")/../aarmvsnet_b200/lib" || exit 1
printf '%-22s %8s %6s %6s %9s %7s %6s %8s\n' object UTCHMMA LDTM UTCBAR UTMALDG UBLKCP REDG 'ACQBULK'
for o in *.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  c() { printf '%s\n' "$s" | grep -c "$1"; }
  printf '%-22s %8s %6s %6s %9s %7s %6s %8s\n' "$o" "$(c UTCHMMA)" "$(c LDTM)" "$(c UTCBAR)" "$(c UTMALDG)" "$(c UBLKCP)" "$(c 'REDG.*F32x4')" "$(c ACQBULK)"
done
echo
echo "UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit -> mbarrier,"
echo "UTMALDG = cp.async.bulk.tensor (TMA tensor load), UBLKCP = cp.async.bulk (1-D bulk copy of the weights),"
echo "REDG F32x4 = red.global.add.v4.f32 (warp / deformable-gather backward scatter), ACQBULK = griddepcontrol.wait (PDL)."
cuobjdump --list-elf libaarmvs_sm100.so | head -3
