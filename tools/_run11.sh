#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2f_pytest.log
{
echo "== arz default"; python tools/quick_bench.py
echo "== arz reward"; python tools/quick_bench.py --reward
echo "== arz reward no-auto"; MT_AUTO_INDEPENDENT=0 python tools/quick_bench.py --reward
echo "== arz reward+action"; python tools/quick_bench.py --reward --action
echo "== arz per-env (ring)"; python tools/quick_bench.py --per-env
echo "== arz per-env 16384 (ring)"; python tools/quick_bench.py --per-env --envs 16384
echo "== lwr per-env 16384 staged"; python tools/quick_bench.py --model lwr --per-env --envs 16384
echo "== lwr per-env 16384 unstaged"; MT_TABLE_STAGE=0 python tools/quick_bench.py --model lwr --per-env --envs 16384
echo "== arz f32 fast per-env 16384 staged"; python tools/quick_bench.py --dtype float32 --fast --per-env --envs 16384
echo "== arz f32 fast per-env 16384 unstaged"; MT_TABLE_STAGE=0 python tools/quick_bench.py --dtype float32 --fast --per-env --envs 16384
echo "== arz default again"; python tools/quick_bench.py
} > gpurun_out/r2f_quick.log 2>&1
python tools/profile_kernels.py --only arz_f64_reward > /dev/null 2>&1 &&
for c in arz_f64_reward arz_f64_table fused_f64_reward nonlocal_f64 arz_f64; do
  ncu --set full --clock-control none --import-source on -k regex:'step_|nonlocal' -s 2 -c 1 -o /tmp/r2f_$c python tools/profile_kernels.py --only $c > gpurun_out/r2f_ncu_$c.log 2>&1
  ncu -i /tmp/r2f_$c.ncu-rep --page raw --csv > gpurun_out/r2f_raw_$c.csv 2>/dev/null
  ncu -i /tmp/r2f_$c.ncu-rep --page source --csv > /tmp/src_$c.csv 2>/dev/null
  python tools/ncu_source_summary.py /tmp/src_$c.csv 0 > gpurun_out/r2f_src_$c.txt 2>&1
  gzip -c /tmp/src_$c.csv > gpurun_out/r2f_srccsv_$c.csv.gz
done
cat gpurun_out/r2f_pytest.log; grep -o '^== .*\|"streaming": {[^}]*}' gpurun_out/r2f_quick.log
