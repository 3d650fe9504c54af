cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== pair on: gemm_check"; timeout 180 python tools/gemm_check.py > gpurun_out/gemm_check_pair.log 2>&1; echo "exit $?"; tail -40 gpurun_out/gemm_check_pair.log
nvidia-smi --query-gpu=name,temperature.gpu --format=csv,noheader
echo "== pair on: mlp shapes"; timeout 180 python tools/gemm_mlp_shapes.py 2>&1 | tail -6
echo "== pair off: mlp shapes"; GX_GEMM_PAIR=0 timeout 180 python tools/gemm_mlp_shapes.py 2>&1 | tail -6
echo "== dense tests"; timeout 600 python -m pytest tests/test_dense.py tests/test_filtering.py -m gpu -x -q 2>&1 | tail -5
echo "== mlp bench pair"; timeout 300 python tools/mlp_bench.py --graph --profile 2>&1 | tail -9
echo "== mlp bench nopair"; GX_GEMM_PAIR=0 timeout 300 python tools/mlp_bench.py --graph 2>&1 | tail -2
