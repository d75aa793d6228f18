#!/bin/bash
# usage: profiles/dense_phases.sh <report.ncu-rep> [kernel regex]    (run in the build container, after gpurun brought the report back)
# Prints the per-phase and per-line breakdown of enhance_dense_kernel from one `ncu --set full --import-source on` capture:
#   gpurun -- 'CMD="python bench.py --steps 3 --warmup 3 --batch 8 --no-e2e --no-cpu-baseline --no-verify --no-overlap";
#              $CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on
#              -k regex:enhance_dense -s 3 -c 1 -o gpurun_out/prof_dense $CMD > gpurun_out/ncu.log 2>&1'
set -e
rep=$1; kre=${2:-enhance_dense}
here=$(cd "$(dirname "$0")" && pwd); root=$(dirname "$here")
tmp=$(mktemp -d)
ncu -i "$rep" --page source --csv --kernel-name regex:$kre --launch-skip 0 --launch-count 1 2>/dev/null > $tmp/src.csv
(cd $tmp && cuobjdump -xelf all $root/yolo-mslesseg_b200/mslesseg_b200/libmslesseg.so >/dev/null && nvdisasm -g -c msl_enhance_dense.sm_100a.cubin > dis.txt)
python $here/ncu_lines.py $tmp/src.csv $tmp/dis.txt $kre 60 > $tmp/lines.txt
python $here/ncu_phases.py $tmp/lines.txt
echo; cat $tmp/lines.txt
rm -rf $tmp
