#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2i_pytest.log
{
echo "== arz default"; python tools/quick_bench.py
echo "== arz reward"; python tools/quick_bench.py --reward
echo "== arz reward+action"; python tools/quick_bench.py --reward --action
echo "== arz default"; python tools/quick_bench.py
echo "== legs rollout fused"; python tools/bench_legs.py --legs rollout
} > gpurun_out/r2i_quick.log 2>&1
python tools/bench_legs.py --legs rollout --horizon 16 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_rollout_launches.csv python tools/bench_legs.py --legs rollout --horizon 16 > gpurun_out/r2i_ncu.log 2>&1
cat gpurun_out/r2i_pytest.log; grep -o '^== .*\|"streaming": {[^}]*}\|"value": [0-9.e+]*\|"us_per_step": [0-9.]*\|error.*' gpurun_out/r2i_quick.log
