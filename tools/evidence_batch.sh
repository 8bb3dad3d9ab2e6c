#!/bin/bash
cd /root/repo
O=gpurun_out
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
python bench.py --impl reference > $O/r02_bench_reference_arm.json 2>> $O/r02_bench_n1.err
MT_LIB=/root/repo/mbrl_traffic_b200/libmt_trace.so python tools/trace_pipe.py --steps 40 > $O/r02_trace_streaming.txt 2>&1
MT_LIB=/root/repo/mbrl_traffic_b200/libmt_trace.so python tools/trace_pipe.py --steps 40 --chained > $O/r02_trace_chained.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-extras > $O/r02_bench_s20.json 2>> $O/r02_bench_n1.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv python bench.py --steps 20 --warmup 5 --no-extras > $O/r02_ncu_launches.log 2>&1
python tools/profile_kernels.py > $O/r02_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'step_|nonlocal' -o /tmp/r02_kernels python tools/profile_kernels.py > $O/r02_ncu_kernels.log 2>&1
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > $O/r02_kernels_raw.csv 2>/dev/null
ls -la $O | tail -12; tail -2 $O/r02_ncu_kernels.log; head -c 600 $O/r02_trace_streaming.txt
