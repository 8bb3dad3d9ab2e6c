#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2l_pytest.log
for i in 1 2; do
echo "== new default"; python tools/quick_bench.py
echo "== prev default"; MT_LIB=/root/repo/mbrl_traffic_b200/libmt_prev.so python tools/quick_bench.py
echo "== new reward"; python tools/quick_bench.py --reward
echo "== new reward no reducer"; MT_REWARD_REDUCER=0 python tools/quick_bench.py --reward
echo "== prev reward"; MT_LIB=/root/repo/mbrl_traffic_b200/libmt_prev.so python tools/quick_bench.py --reward
echo "== new reward+action"; python tools/quick_bench.py --reward --action
echo "== prev reward+action"; MT_LIB=/root/repo/mbrl_traffic_b200/libmt_prev.so python tools/quick_bench.py --reward --action
done > gpurun_out/r2l_quick.log 2>&1
echo "== legs rollout" >> gpurun_out/r2l_quick.log; python tools/bench_legs.py --legs rollout >> gpurun_out/r2l_quick.log 2>&1
cat gpurun_out/r2l_pytest.log; grep -o '^== .*\|"streaming": {[^}]*}\|"us_per_step": [0-9.]*' gpurun_out/r2l_quick.log
