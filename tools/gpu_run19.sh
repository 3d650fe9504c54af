cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/tune.py --workloads roe1d,readme --divisions reference,reciprocal,ieee --threads 32,64,128,256 --min-blocks 1,2,3,4,5,6,8 --out-stages 1 --spt 1 --steps 200 --out gpurun_out/tune_r02b_small.json --write > gpurun_out/tune_r02b_small.log 2>&1
echo "exit $?"
grep best gpurun_out/tune_r02b_small.log | cut -c1-210
cp graphax_b200/tuning.json gpurun_out/tuning_r02b.json
