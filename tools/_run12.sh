#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2h_pytest.log
{
echo "== arz default"; python tools/quick_bench.py
echo "== arz reward+action"; python tools/quick_bench.py --reward --action
echo "== arz default"; python tools/quick_bench.py
echo "== legs rollout fused"; python tools/bench_legs.py --legs rollout
echo "== legs rollout plain"; python tools/bench_legs.py --legs rollout --rollout-plain
} > gpurun_out/r2h_quick.log 2>&1
cat gpurun_out/r2h_pytest.log; grep -o '^== .*\|"streaming": {[^}]*}\|"value": [0-9.e+]*\|"us_per_step": [0-9.]*\|"policy_path": "[^"]*"\|error.*' gpurun_out/r2h_quick.log
