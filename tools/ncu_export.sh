#!/bin/bash
# ncu_export.sh <rep-without-extension> : text summaries of a capture, then the (large) .ncu-rep is deleted so that
# gpurun_out/ stays under the 64 MiB that travels back.
set -u
rep="$1"
python tools/ncu_summary.py "$rep.ncu-rep" --top 12 > "$rep.summary.txt" 2>&1
ncu -i "$rep.ncu-rep" --page raw --csv > "$rep.raw.csv" 2>/dev/null
rm -f "$rep.ncu-rep"
