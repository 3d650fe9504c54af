set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r02c.log 2>&1; echo "gpu tests exit $?"; tail -8 gpurun_out/gpu_tests_r02c.log
