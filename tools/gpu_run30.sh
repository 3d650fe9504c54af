cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== dense tests"; timeout 900 python -m pytest tests/test_dense.py tests/test_filtering.py -m gpu -x -q 2>&1 | tail -5
echo "== mlp bench (arena cleared by the first contraction)"; timeout 300 python tools/mlp_bench.py --graph --profile 2>&1 | tail -9
echo "== mlp bench, memset"; GX_DENSE_MEMSET=1 timeout 300 python tools/mlp_bench.py --graph 2>&1 | tail -1
echo "== trace"; timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace_r02c_nomemset.txt 2>&1; echo "exit $?"; grep -E "^launch|phases|epilogue done|entry|wait passed" gpurun_out/gemm_trace_r02c_nomemset.txt | cut -c1-150
