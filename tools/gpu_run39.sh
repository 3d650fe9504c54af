cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 0 1; do
  echo "=== GX_DENSE_MEMSET=$v"
  GX_DENSE_MEMSET=$v timeout 300 python tools/gemm_trace.py --smid > gpurun_out/gemm_trace_smid_$v.txt 2>&1
  sed -n '/^launch 1/,/^launch 2/p' gpurun_out/gemm_trace_smid_$v.txt | cut -c1-230
done
