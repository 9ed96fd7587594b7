#!/bin/bash
# usage: tools/sass_probe.sh <file.cu under csrc> <L> [extra nvcc flags]   -- compile ONE filter length, print per-kernel
# registers / spills and the instruction mix (DFMA operand forms, shared-memory and constant loads, barriers)
f=$1; L=$2; shift 2
o=/tmp/probe_$(basename $f .cu)_$L.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr \
  -DJWC_ONLY_L=$L "$@" -c jwave-pro_b200/csrc/$f -o $o 2>/dev/null || { echo compile failed; exit 1; }
for fn in $(cuobjdump -res-usage $o 2>/dev/null | grep -oE "Function [^:]*" | awk '{print $2}'); do
  res=$(cuobjdump -res-usage $o 2>/dev/null | grep -A1 "$fn" | grep -oE "REG:[0-9]+ STACK:[0-9]+")
  short=$(echo $fn | c++filt | sed -E 's/\(anonymous namespace\):://g; s/jwc:://g; s/\(.*//; s/^void //')
  mix=$(cuobjdump -sass -fun $fn $o | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | awk '{op=$2; if(op ~ /^@/) op=$3; split(op,a,"."); k=a[1]; if (k=="DFMA"){ if ($0 ~ /UR[0-9]/) k="DFMA_UR"; else if ($0 ~ /c\[/) k="DFMA_C"; else k="DFMA_RRR"}; if(k=="LDS"||k=="STS"||k=="LDC"||k=="LDCU"||k=="LDL"||k=="STL") k=op; c[k]++; n++} END{printf "n=%d ", n; for(k in c) if (k ~ /DFMA|LDS|STS|LDC|R2UR|BAR|LDL|STL/) printf "%s=%d ", k, c[k]}')
  echo "$short | $res | $mix"
done
