#!/bin/bash
cd /root/repo
{
echo "== per-env U=1"; MT_PIPE_U=1 python tools/quick_bench.py --per-env
echo "== per-env U=1 16384"; MT_PIPE_U=1 python tools/quick_bench.py --per-env --envs 16384
echo "== default U=1"; MT_PIPE_U=1 python tools/quick_bench.py
echo "== r112 default"; MT_LIB=/root/repo/mbrl_traffic_b200/libmt_r112.so python tools/quick_bench.py
echo "== r112 per-env"; MT_LIB=/root/repo/mbrl_traffic_b200/libmt_r112.so python tools/quick_bench.py --per-env
echo "== r112 reward"; MT_LIB=/root/repo/mbrl_traffic_b200/libmt_r112.so python tools/quick_bench.py --reward
echo "== population exact 16384"; python tools/quick_bench.py --population --envs 16384
echo "== population exact 16384 U=1"; MT_PIPE_U=1 python tools/quick_bench.py --population --envs 16384
echo "== population fast 16384 U=1"; MT_PIPE_U=1 python tools/quick_bench.py --population --envs 16384 --fast
} > gpurun_out/r2d_quick.log 2>&1
cat gpurun_out/r2d_quick.log
