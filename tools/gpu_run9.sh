set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02b.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_r02b.log
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r02e.log 2>&1; echo "gpu tests exit $?"; tail -8 gpurun_out/gpu_tests_r02e.log
