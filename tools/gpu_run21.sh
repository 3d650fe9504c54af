cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dense.py tests/test_filtering.py -m gpu -x -q 2>&1 | tail -4
python tools/mlp_bench.py --graph --profile > gpurun_out/mlp_r02b_p1.json 2> gpurun_out/mlp_r02b_p1.err; tail -7 gpurun_out/mlp_r02b_p1.err; cat gpurun_out/mlp_r02b_p1.json
