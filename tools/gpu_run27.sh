cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== dense tests (pair tests incl.)"; timeout 900 python -m pytest tests/test_dense.py -m gpu -x -q 2>&1 | tail -5
echo "== trace"; timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace_r02c.txt 2>&1; echo "exit $?"; cat gpurun_out/gemm_trace_r02c.txt
echo "== trace pair"; timeout 300 python tools/gemm_trace.py --pair > gpurun_out/gemm_trace_r02c_pair.txt 2>&1; echo "exit $?"; tail -50 gpurun_out/gemm_trace_r02c_pair.txt
