cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== dense tests"; timeout 900 python -m pytest tests/test_dense.py tests/test_filtering.py -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do for v in 1 0; do
  echo "GX_DENSE_DUAL=$v"; GX_DENSE_DUAL=$v timeout 300 python tools/mlp_bench.py --graph --steps 400 2>/dev/null | cut -c90-180
done; done
timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace_r02c_dual.txt 2>&1
grep -E "^launch|phases" gpurun_out/gemm_trace_r02c_dual.txt | cut -c1-150
