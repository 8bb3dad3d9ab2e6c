#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2j_pytest.log
{
echo "== arz default"; python tools/quick_bench.py
echo "== arz reward (reducer)"; python tools/quick_bench.py --reward
echo "== arz reward (no reducer)"; MT_REWARD_REDUCER=0 python tools/quick_bench.py --reward
echo "== arz reward+action (reducer)"; python tools/quick_bench.py --reward --action
echo "== arz reward+action (no reducer)"; MT_REWARD_REDUCER=0 python tools/quick_bench.py --reward --action
echo "== arz reward 16384 (reducer)"; python tools/quick_bench.py --reward --envs 16384
echo "== arz reward 16384 (no reducer)"; MT_REWARD_REDUCER=0 python tools/quick_bench.py --reward --envs 16384
echo "== arz default"; python tools/quick_bench.py
echo "== legs rollout"; python tools/bench_legs.py --legs rollout
} > gpurun_out/r2j_quick.log 2>&1
cat gpurun_out/r2j_pytest.log; grep -o '^== .*\|"streaming": {[^}]*}\|"value": [0-9.e+]*\|"us_per_step": [0-9.]*\|error.*' gpurun_out/r2j_quick.log
