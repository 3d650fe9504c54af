cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in none nostore; do
  echo "=== early entry, GX_EPI_DEBUG=$v"
  GX_EPI_DEBUG=$v timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace_y_$v.txt 2>&1
  grep -E "^launch|phases|acc complete|wait passed" gpurun_out/gemm_trace_y_$v.txt | sed -n 1,12p | cut -c1-150
done
