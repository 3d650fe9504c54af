set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/tune.py --workloads roe3d --divisions reference --schedules pressure,minpressure --threads 128 --min-blocks 3,4 --out-stages 1 --spt 1 --steps 200 --out gpurun_out/tune_r02b_minpressure_reference.json > gpurun_out/tune_r02b_minpressure_reference.log 2>&1
grep best gpurun_out/tune_r02b_minpressure_reference.log
python tools/tune.py --workloads roe3d --divisions reciprocal --schedules pressure,minpressure --threads 128 --min-blocks 3,4 --out-stages 1 --spt 1 --steps 200 --out gpurun_out/tune_r02b_minpressure_fma.json > gpurun_out/tune_r02b_minpressure_fma.log 2>&1
grep best gpurun_out/tune_r02b_minpressure_fma.log
