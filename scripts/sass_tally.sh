#!/usr/bin/env bash
# Opcode tally of the shipped library (cuobjdump -sass): the SASS mnemonics that prove tcgen05 / TMEM / TMA use.
# usage: scripts/sass_tally.sh > profiles/rNN_sass_tally.md
cd "$(dirname "$0")/.."
SO=longterm360fov_b200/lib/libfov360.so
TMP=$(mktemp)
cuobjdump -sass $SO > $TMP 2>/dev/null
echo "# SASS opcode tally of \`$SO\` (sm_100a), \`cuobjdump -sass\`"
echo
echo "| mnemonic | what | count |"
echo "|---|---|---:|"
row() { printf "| \`%s\` | %s | %s |\n" "$1" "$2" "$(grep -c "$1" $TMP)"; }
row UTCHMMA "tcgen05.mma kind::f16 (bf16 operands, fp32 accumulate in TMEM)"
row UTMALDG "cp.async.bulk.tensor (TMA tensor-map tile loads)"
row UBLKCP "cp.async.bulk (TMA-engine bulk copies: resident weights)"
row LDTM "tcgen05.ld (TMEM -> registers)"
row STTM "tcgen05.st (registers -> TMEM)"
row UTCBAR "tcgen05.commit -> mbarrier"
row SYNCS "mbarrier operations"
row LDGSTS "cp.async (Ampere-style staging)"
row "LDG.E.ENL2.256" "256-bit global loads"
row "STG.E.ENL2.256" "256-bit global stores"
row "LDG.E.128" "128-bit global loads"
row "STG.E.128" "128-bit global stores"
row "REDG.E.ADD.F32" "red.global.add.f32 (weight-gradient partial sums)"
row "MUFU.EX2" "exp2 (tanh / sigmoid)"
row DFMA "float64 FMA (angles, Gaussian tiles: bit-exact sample builders)"
echo
echo "## kernels per source file (template instantiations)"
echo
echo "| source | kernels |"
echo "|---|---:|"
grep "Function :" $TMP | grep -o "_[0-9]*_[a-z0-9_]*_cu_" | sed -E "s/^_+[0-9a-f]+_[0-9]+_//; s/_cu_$/.cu/" | sort | uniq -c | sort -rn | awk '{printf "| `%s` | %s |\n", $2, $1}'
rm -f $TMP
