cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "A=1" "GX_FUSED_SHALLOW=1" "GX_EPI_DEBUG=notanh"; do
  echo "=== $v"
  env $v timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace_x.txt 2>&1; echo "exit $?"
  grep -E "^launch|phases|epilogue done|entry|wait passed" gpurun_out/gemm_trace_x.txt | cut -c1-150
  env $v timeout 300 python tools/mlp_bench.py --graph 2>/dev/null | cut -c1-200
done
