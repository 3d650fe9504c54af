set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/mlp_bench.py --graph --profile > gpurun_out/mlp_r02_2streams.json 2> gpurun_out/mlp_r02_2streams.err; tail -12 gpurun_out/mlp_r02_2streams.err; cat gpurun_out/mlp_r02_2streams.json
GX_DENSE_STREAMS=1 python tools/mlp_bench.py --graph > gpurun_out/mlp_r02_1stream.json 2>/dev/null; cat gpurun_out/mlp_r02_1stream.json
GX_PDL=0 python tools/mlp_bench.py --graph > gpurun_out/mlp_r02_nopdl.json 2>/dev/null; cat gpurun_out/mlp_r02_nopdl.json
GX_DENSE_FUSE=0 python tools/mlp_bench.py --graph > gpurun_out/mlp_r02_nofuse.json 2>/dev/null; cat gpurun_out/mlp_r02_nofuse.json
python tools/mlp_bench.py > gpurun_out/mlp_r02_nograph.json 2>/dev/null; cat gpurun_out/mlp_r02_nograph.json
