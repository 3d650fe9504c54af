set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_filtering.py tests/test_deep_learning.py tests/test_rule_registry.py tests/test_structure_nd.py -x -q -m gpu > gpurun_out/gpu_tests_r02d.log 2>&1; echo "exit $?"; tail -30 gpurun_out/gpu_tests_r02d.log
