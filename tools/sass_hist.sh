#!/bin/bash
# usage: sass_hist.sh <object/.so> <function-name-substring> [top-n]
# Opcode histogram of one kernel's SASS plus ALU-pipe / FMA-pipe totals.
set -e
obj="$1"; pat="$2"; top="${3:-14}"
fn=$(cuobjdump -sass "$obj" | grep -E '^\s+Function : ' | sed 's/.*Function : //' | grep -- "$pat" | head -1)
[ -z "$fn" ] && { echo "no function matching $pat"; exit 1; }
echo "function: $fn"
cuobjdump -sass -fun "$fn" "$obj" | grep -E '^\s+/\*[0-9a-f]{4,6}\*/' \
  | awk '{ if ($2 ~ /^@/) print $3; else print $2}' | sed 's/;//' | sort | uniq -c | sort -rn > /tmp/sass_hist.txt
head -"$top" /tmp/sass_hist.txt
awk '{n=$1; op=$2;
  if (op ~ /^(IADD3|LOP3|SHF|PRMT|SEL|ISETP|IMNMX|VIADD|LEA|MOV|VIMNMX|BMSK|SGXT|FLO|POPC|IABS)/) alu+=n;
  else if (op ~ /^(IMAD|FFMA|FMUL|FADD)/) { if (op ~ /WIDE/) wide+=n; else fma+=n; }
  else oth+=n; tot+=n}
  END {printf "ALU %d  FMA %d  FMA.WIDE %d  other %d  total %d\n", alu, fma, wide, oth, tot}' /tmp/sass_hist.txt
