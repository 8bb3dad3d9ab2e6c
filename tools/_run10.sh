#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream_order.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2e_pytest.log
{
echo "== arz reward"; python tools/quick_bench.py --reward
echo "== arz reward+action"; python tools/quick_bench.py --reward --action
echo "== arz uniform rho_max 0.5"; python tools/quick_bench.py --rho-max 0.5
echo "== arz uniform lam 2"; python tools/quick_bench.py --lam 2
echo "== arz per-env ring"; MT_KERNEL=ring python tools/quick_bench.py --per-env
echo "== arz uniform ring"; MT_KERNEL=ring python tools/quick_bench.py
echo "== fused"; python tools/fused_bench.py
} > gpurun_out/r2e_quick.log 2>&1
cat gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_quick.log
