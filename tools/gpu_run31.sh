cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== dense tests"; timeout 900 python -m pytest tests/test_dense.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do for v in 0 1; do
  echo "=== rep $rep GX_DENSE_MEMSET=$v"
  GX_DENSE_MEMSET=$v timeout 300 python tools/mlp_bench.py --graph --steps 400 2>/dev/null | cut -c90-180
  GX_DENSE_MEMSET=$v timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace_ab_${v}_$rep.txt 2>&1
  grep -E "phases" gpurun_out/gemm_trace_ab_${v}_$rep.txt | cut -c1-150
  grep -E "epilogue done" gpurun_out/gemm_trace_ab_${v}_$rep.txt | tail -1
done; done
