cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/mlp_bench.py --graph --profile > gpurun_out/mlp_r02b_base.json 2> gpurun_out/mlp_r02b_base.err; tail -8 gpurun_out/mlp_r02b_base.err; cat gpurun_out/mlp_r02b_base.json
ncu --set full --clock-control none --import-source on -k regex:gx_gemm -s 15 -c 5 -f -o gpurun_out/prof_r02b_mlp python tools/mlp_bench.py --steps 3 > gpurun_out/ncu_r02b_mlp.log 2>&1
tail -3 gpurun_out/ncu_r02b_mlp.log
