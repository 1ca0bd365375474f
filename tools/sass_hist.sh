#!/bin/bash
# usage: tools/sass_hist.sh  -> profiles/r2_sass_<kernel>.txt: opcode histogram + resource usage of every hot kernel of the built library
LIB=parasail_rs_b200/libparasail_b200.so
OUT=profiles
declare -A K=(
  [sw16_scan_K25]="sw16_scan_kernelILi25ELb0"
  [sw16_scan_strip_K25]="sw16_scan_kernelILi25ELb1"
  [pairs16_G16_K10_nwsg_trace]="pairs16_kernelILi16ELi10ELb0ELb1"
  [pairs16_G32_K10_nwsg_score]="pairs16_kernelILi32ELi10ELb0ELb0"
  [pairs16_G32_K8_sw_trace]="pairs16_kernelILi32ELi8ELb1ELb1"
  [walk16_cigar]="walk16_kernelILb0"
  [walk16_stats]="walk16_kernelILb1"
  [wave32v3_K8_sw]="wave32v3_kernelILi8ELi4ELb1"
)
for name in "${!K[@]}"; do
  pat=${K[$name]}
  f=$OUT/r2_sass_$name.txt
  {
    echo "# SASS opcode histogram of the first function matching '$pat' in $LIB (cuobjdump -sass; sm_100a)"
    cuobjdump --dump-resource-usage $LIB 2>/dev/null | grep -A1 "$pat" | grep -o "REG:[0-9]*\|STACK:[0-9]*\|SHARED:[0-9]*\|LOCAL:[0-9]*" | tr '\n' ' '; echo
    cuobjdump -sass $LIB 2>/dev/null | awk -v pat="$pat" '/Function :/ {on = index($0, pat) > 0 && !done; if (on) seen=1; else if (seen) done=1} on {print}' \
      | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/\/\* 0x[0-9a-f]* \*\///' | awk '{print $2}' | sed 's/^@!*U*P[0-9T] *//' | grep -v '^$' | sort | uniq -c | sort -rn
  } > $f
  echo "$f: $(head -2 $f | tail -1) $(sed -n 3,5p $f | tr '\n' ';')"
done
