#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2n_pytest.log
{
echo "== default"; python tools/quick_bench.py
echo "== per-env ring (default)"; python tools/quick_bench.py --per-env
echo "== per-env pipe staged"; MT_KERNEL=pipe python tools/quick_bench.py --per-env
echo "== per-env pipe unstaged"; MT_KERNEL=pipe MT_TABLE_STAGE=0 python tools/quick_bench.py --per-env
echo "== per-env 16384 ring"; python tools/quick_bench.py --per-env --envs 16384
echo "== per-env 16384 pipe staged"; MT_KERNEL=pipe python tools/quick_bench.py --per-env --envs 16384
echo "== per-env 16384 pipe unstaged"; MT_KERNEL=pipe MT_TABLE_STAGE=0 python tools/quick_bench.py --per-env --envs 16384
echo "== reward"; python tools/quick_bench.py --reward
echo "== reward+action"; python tools/quick_bench.py --reward --action
echo "== lwr per-env 16384 staged"; python tools/quick_bench.py --model lwr --per-env --envs 16384
echo "== lwr per-env 16384 unstaged"; MT_TABLE_STAGE=0 python tools/quick_bench.py --model lwr --per-env --envs 16384
echo "== population exact"; python tools/quick_bench.py --population --envs 16384
echo "== population fast"; python tools/quick_bench.py --population --envs 16384 --fast
} > gpurun_out/r2n_quick.log 2>&1
cat gpurun_out/r2n_pytest.log; grep -o '^== .*\|"streaming": {[^}]*}' gpurun_out/r2n_quick.log
