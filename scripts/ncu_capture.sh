#!/bin/bash
# usage: scripts/ncu_capture.sh <name> <kernel-regex> <count> <source:0|1> <command...>
# ncu --set full capture -> gpurun_out/<name>.raw.csv (+ .source.csv when asked); nothing larger than 12 MB is left in
# gpurun_out/ (the whole directory must stay below 64 MiB to be copied back).
name=$1; regex=$2; count=$3; src=$4; shift 4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$regex" -c "$count" -f -o /tmp/$name "$@" > /dev/null 2> gpurun_out/$name.ncu.err
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
if [ "$src" = "1" ]; then ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null; fi
cp /tmp/$name.ncu-rep gpurun_out/ 2>/dev/null
find gpurun_out -type f -size +12M -print -delete
ls -la gpurun_out/$name.* | awk '{print $5, $9}'
