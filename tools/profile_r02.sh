#!/bin/bash
# Round-2 evidence run (on the GPU box, one B200): launch list of the bench command, then one full capture each of the
# bench kernel (pair form, configs[1]), the one-habitat kernel (monolithic form: configs[2]/[3] shape) and the fan-out
# kernel (configs[4] shape).  Every command runs plain first (exit 0) and then under ncu.
cd "$(dirname "$0")/.." || exit 1
O=$PWD/gpurun_out
python bench.py --steps 2 --warmup 3 --no-configs > $O/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-configs > $O/r02_ncu_bench.log 2>&1
cd tests || exit 1
python prof_chain.py 100000 2 fast 1 > $O/r02_plain_ws.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -f -o $O/prof_r02_ws \
    python prof_chain.py 100000 2 fast 1 > $O/r02_ncu_ws.log 2>&1
python prof_chain.py 100000 1 fast 1 > $O/r02_plain_mono.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -f -o $O/prof_r02_mono \
    python prof_chain.py 100000 1 fast 1 > $O/r02_ncu_mono.log 2>&1
python prof_chain.py 9472 64 fast 1 > $O/r02_plain_fan.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -f -o $O/prof_r02_fan \
    python prof_chain.py 9472 64 fast 1 > $O/r02_ncu_fan.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw-metric and per-instruction pages as CSV, drop the reports
for k in ws mono fan; do
  ncu -i $O/prof_r02_$k.ncu-rep --page raw --csv > $O/prof_r02_${k}_raw.csv 2>/dev/null
  ncu -i $O/prof_r02_$k.ncu-rep --page source --csv > $O/prof_r02_${k}_src.csv 2>/dev/null
  rm -f $O/prof_r02_$k.ncu-rep
done
for f in $O/r02_plain_ws.log $O/r02_plain_mono.log $O/r02_plain_fan.log; do tail -n 2 $f; done
du -sh $O
