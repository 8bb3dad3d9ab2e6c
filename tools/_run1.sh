set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_overlap.py -x -q 2>&1 | tail -15
echo "=== A/B quick bench"
for ov in 2 0; do
  echo "--- MT_OVERLAP=$ov"
  MT_OVERLAP=$ov timeout 300 tools/qb.sh "--envs 4096" "--envs 16384" "--envs 4096 --model lwr" "--envs 4096 --dtype float32 --fast" 
done
echo "=== full gpu suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -c 3000 gpurun_out/r02_bench_a.json
