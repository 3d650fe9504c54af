set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_full_parity.py -x -q > gpurun_out/full_parity.log 2>&1; echo "full parity exit $?"
tail -5 gpurun_out/full_parity.log
timeout 1200 python tools/tune.py --workloads roe3d --divisions reference --threads 64,128 --min-blocks 2,3,4,5 --out-stages 1 --spt 1 --out gpurun_out/tune_r02_reference.json > gpurun_out/tune_r02_reference.log 2>&1; echo "tune exit $?"
grep best gpurun_out/tune_r02_reference.log | cut -c1-300
timeout 600 python tools/tune.py --workloads roe3d --orders mixed,rev --divisions reciprocal --threads 128 --min-blocks 3,4 --out-stages 1 --spt 1 --out gpurun_out/tune_r02_fma_check.json > gpurun_out/tune_r02_fma_check.log 2>&1; echo "tune2 exit $?"
grep -A3 best gpurun_out/tune_r02_fma_check.log | cut -c1-300
