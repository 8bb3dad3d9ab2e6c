#!/bin/bash
cd /root/repo
NS=/root/repo/mbrl_traffic_b200/libmt_nosplit.so
run() { echo "== split: $*"; python tools/quick_bench.py "$@"; echo "== nosplit: $*"; MT_LIB=$NS python tools/quick_bench.py "$@"; }
{
run --model lwr --envs 16384
run --dtype float32 --fast --envs 16384
run --dtype float32 --envs 16384
run --model nonlocal --envs 16384 --bc inflow_outflow --sets 2 --steps 40
run --model nonlocal --envs 16384 --bc inflow_outflow --sets 2 --steps 40 --scheme lxf
run --per-env
run --model lwr
echo "== split: long road"; python tools/bench_legs.py --legs long_road
echo "== nosplit: long road"; MT_LIB=$NS python tools/bench_legs.py --legs long_road
echo "== split: fused"; python tools/fused_bench.py
echo "== nosplit: fused"; MT_LIB=$NS python tools/fused_bench.py
} > gpurun_out/r2m_quick.log 2>&1
grep -o '^== .*\|"streaming": {[^}]*}\|"value": [0-9.e+]*\|"fused_gcups": [0-9.]*\|"stepwise_gcups": [0-9.]*' gpurun_out/r2m_quick.log
