cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/tune.py --workloads roe3d --orders mixed,fwd --divisions reference --schedules minpressure,ilp:150:4,ilp:160:6,ilp:140:3,ilp:130:8,elimination --threads 128 --min-blocks 3 --out-stages 1 --spt 1 --steps 300 --out gpurun_out/tune_r02b_ilp.json > gpurun_out/tune_r02b_ilp.log 2>&1
grep best gpurun_out/tune_r02b_ilp.log | cut -c1-60,150-260
