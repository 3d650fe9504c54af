set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# A: inputs prefetched into registers (new default)   B: TMA-staged inputs (GX_INPUT_STAGED=1)
python tools/tune.py --workloads roe3d --divisions reference --schedules minpressure --threads 128 --min-blocks 3,4 --out-stages 1,2 --spt 1 --steps 200 --out gpurun_out/tune_r02b_ldg_reference.json > gpurun_out/tune_r02b_ldg_reference.log 2>&1
grep -A4 best gpurun_out/tune_r02b_ldg_reference.log | cut -c1-250
GX_INPUT_STAGED=1 python tools/tune.py --workloads roe3d --divisions reference --schedules minpressure --threads 128 --min-blocks 3,4 --out-stages 1,2 --spt 1 --steps 200 --out gpurun_out/tune_r02b_staged_reference.json > gpurun_out/tune_r02b_staged_reference.log 2>&1
grep -A4 best gpurun_out/tune_r02b_staged_reference.log | cut -c1-250
for at in 0.4 0.55 0.85; do
GX_PREFETCH_AT=$at python tools/tune.py --workloads roe3d --orders mixed --divisions reference --schedules minpressure --threads 128 --min-blocks 3 --out-stages 1 --spt 1 --steps 200 --out gpurun_out/tune_r02b_ldg_at$at.json 2>&1 | grep best | cut -c1-250
done
