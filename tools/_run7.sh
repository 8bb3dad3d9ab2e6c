#!/bin/bash
cd /root/repo
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python tools/profile_kernels.py > gpurun_out/r2b_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'step_|nonlocal' -o /tmp/r2b_kernels python tools/profile_kernels.py > gpurun_out/r2b_ncu.log 2>&1
ncu -i /tmp/r2b_kernels.ncu-rep --page raw --csv > gpurun_out/r2b_kernels_raw.csv 2>/dev/null
ncu -i /tmp/r2b_kernels.ncu-rep --page source --csv > /tmp/r2b_src.csv 2>/dev/null
# the last launch of each case (launch index = 3*case + 2)
for idx in 2 5 8 11 14 17 20 23 26 29 32 35; do
  python tools/ncu_source_summary.py /tmp/r2b_src.csv $idx > gpurun_out/r2b_src_$idx.txt 2>&1
done
ls -la /tmp/r2b_kernels.ncu-rep /tmp/r2b_src.csv gpurun_out/
tail -3 gpurun_out/r2b_ncu.log
