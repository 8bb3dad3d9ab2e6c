#!/bin/bash
# A/B: range test (FSETP vs integer) and automatic early loads
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/ab1_pytest.log
for i in 1 2; do
  echo "== default (fsetp, auto)"; python tools/quick_bench.py
  echo "== fsetp, MT_AUTO_INDEPENDENT=0"; MT_AUTO_INDEPENDENT=0 python tools/quick_bench.py
  echo "== int test, auto"; MT_LIB=/root/repo/mbrl_traffic_b200/libmt_inttest.so python tools/quick_bench.py
  echo "== int test, no auto"; MT_AUTO_INDEPENDENT=0 MT_LIB=/root/repo/mbrl_traffic_b200/libmt_inttest.so python tools/quick_bench.py
done > gpurun_out/ab1.log 2>&1
echo "== sustained 20000 steps default" >> gpurun_out/ab1.log
python tools/quick_bench.py --steps 20000 --warmup 50 >> gpurun_out/ab1.log 2>&1
MT_AUTO_INDEPENDENT=0 MT_LIB=/root/repo/mbrl_traffic_b200/libmt_inttest.so python tools/quick_bench.py --steps 20000 --warmup 50 >> gpurun_out/ab1.log 2>&1
cat gpurun_out/ab1_pytest.log; cat gpurun_out/ab1.log
