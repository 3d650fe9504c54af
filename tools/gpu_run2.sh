set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python tools/tune.py --workloads roe3d --divisions reference --threads 64,128 --min-blocks 2,3,4,5 --out-stages 1 --spt 1 --out gpurun_out/tune_r02_reference.json > gpurun_out/tune_r02_reference.log 2>&1; echo "tune exit $?"
grep -A4 best gpurun_out/tune_r02_reference.log | cut -c1-260
timeout 600 python tools/tune.py --workloads roe3d --orders mixed,rev --divisions reciprocal --threads 128 --min-blocks 3,4 --out-stages 1 --spt 1 --out gpurun_out/tune_r02_fma_check.json > gpurun_out/tune_r02_fma_check.log 2>&1; echo "tune2 exit $?"
grep -A3 best gpurun_out/tune_r02_fma_check.log | cut -c1-260
timeout 900 python -m pytest tests/test_dense.py -x -q -m gpu > gpurun_out/dense_r02.log 2>&1; echo "dense exit $?"; tail -15 gpurun_out/dense_r02.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r02a.log 2>&1; echo "gpu tests exit $?"; tail -8 gpurun_out/gpu_tests_r02a.log
