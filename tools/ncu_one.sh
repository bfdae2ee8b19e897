#!/bin/bash
# tools/ncu_one.sh <label> <dir> <prof_one args...>: plain run, then one ncu --set full capture of the 4th wavefront launch
label=$1; dir=$2; shift 2
root=$(pwd)
( cd "$dir" && python "$root/tools/prof_one.py" "$@" > "$root/gpurun_out/${label}_plain.log" 2>&1 ) || { echo "$label: plain run failed"; tail -5 "gpurun_out/${label}_plain.log"; exit 1; }
( cd "$dir" && ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 3 -c 1 -f -o "$root/gpurun_out/${label}" python "$root/tools/prof_one.py" "$@" > "$root/gpurun_out/${label}_ncu.log" 2>&1 )
echo "$label: ncu rc=$? $(ls -la gpurun_out/${label}.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
