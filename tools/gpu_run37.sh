cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do for v in 0 2 4; do
  echo "GX_GEMM_CLUSTER_N=$v"; GX_GEMM_CLUSTER_N=$v timeout 300 python tools/mlp_bench.py --graph --steps 400 2>/dev/null | cut -c90-180
done; done
GX_GEMM_CLUSTER_N=4 timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace_x.txt 2>&1
grep -E "^launch|phases" gpurun_out/gemm_trace_x.txt | cut -c1-150
GX_GEMM_CLUSTER_N=2 timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace_x.txt 2>&1
grep -E "^launch|phases" gpurun_out/gemm_trace_x.txt | cut -c1-150
