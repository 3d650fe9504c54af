set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python tools/tune.py --workloads roe3d --orders mixed --divisions reference --threads 32,64,96,128 --min-blocks 3,4,5,6,7,13,14,15 --out-stages 1 --spt 1 --schedules pressure,elimination,ilp:120:4,ilp:150:6,ilp:100:3 --steps 100 --out gpurun_out/tune_r02_reference_mixed_shapes.json > gpurun_out/tune_r02_reference_mixed_shapes.log 2>&1; echo "tune exit $?"
grep -A6 best gpurun_out/tune_r02_reference_mixed_shapes.log | cut -c1-250
