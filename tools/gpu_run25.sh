cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r02c.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/gpu_tests_r02c.log
timeout 600 python tools/tune.py --workloads roe3d --orders mixed,rev --divisions reference --threads 128 --min-blocks 3,4 --out-stages 1 --spt 1 --slp 0,1 --steps 300 --out gpurun_out/tune_r02c_slp.json > gpurun_out/tune_r02c_slp.log 2>&1; echo "tune exit $?"
grep -E "^ +\{'variant': 'jit'" gpurun_out/tune_r02c_slp.log | cut -c1-200
