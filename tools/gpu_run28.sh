cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in none notanh noatomic nocsum nostore notanh,nocsum,nostore; do
  echo "=== GX_EPI_DEBUG=$v"
  GX_EPI_DEBUG=$v timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace_dbg_$v.txt 2>&1; echo "exit $?"
  grep -E "^launch|phases|epilogue done" gpurun_out/gemm_trace_dbg_$v.txt | cut -c1-150
done
