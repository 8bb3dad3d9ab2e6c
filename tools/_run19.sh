#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2o_pytest.log
{
echo "== fused (reducer)"; python tools/fused_bench.py
echo "== fused (no reducer)"; MT_REWARD_REDUCER=0 python tools/fused_bench.py
} > gpurun_out/r2o_quick.log 2>&1
cat gpurun_out/r2o_pytest.log; grep -o '^== .*\|"model": "[a-z+]*"\|"fused_gcups": [0-9.]*\|"stepwise_gcups": [0-9.]*' gpurun_out/r2o_quick.log | paste - - - | head -30
