#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2c_pytest.log
{
echo "== arz default"; python tools/quick_bench.py
echo "== arz per-env"; python tools/quick_bench.py --per-env
echo "== arz per-env 16384"; python tools/quick_bench.py --per-env --envs 16384
echo "== arz reward"; python tools/quick_bench.py --reward
echo "== arz reward+action"; python tools/quick_bench.py --reward --action
echo "== lwr per-env 16384"; python tools/quick_bench.py --model lwr --per-env --envs 16384
echo "== arz population fast"; python tools/quick_bench.py --population --fast --envs 16384
echo "== nonlocal godunov"; python tools/quick_bench.py --model nonlocal --envs 16384 --bc inflow_outflow --sets 2 --steps 40
echo "== nonlocal lxf"; python tools/quick_bench.py --model nonlocal --envs 16384 --bc inflow_outflow --sets 2 --steps 40 --scheme lxf
} > gpurun_out/r2c_quick.log 2>&1
cat gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_quick.log
