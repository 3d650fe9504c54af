set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02a.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_r02a.log
timeout 900 python bench.py --steps 400 --warmup 10 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_r02a.err; head -c 3000 gpurun_out/bench_r02a.json
SHORT="--steps 20 --warmup 3 --no-cpu-baseline --no-other-orders --no-other-configs --no-verify --e2e-steps 1 --clock-probe-s 0.05"
python bench.py $SHORT > gpurun_out/plain_ref.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_r02_roe3d_mixed_reference.csv python bench.py $SHORT > gpurun_out/ncu_l.log 2>&1
python bench.py $SHORT > gpurun_out/plain_ref2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gx_jacve_kernel -s 8 -c 2 -f -o gpurun_out/prof_r02_mixed_reference python bench.py $SHORT > gpurun_out/ncu_f1.log 2>&1
python bench.py $SHORT --rounding fma > gpurun_out/plain_fma.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gx_jacve_kernel -s 8 -c 2 -f -o gpurun_out/prof_r02_mixed_fma python bench.py $SHORT --rounding fma > gpurun_out/ncu_f2.log 2>&1
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r02b.log 2>&1; echo "gpu tests exit $?"; tail -5 gpurun_out/gpu_tests_r02b.log
ls -la gpurun_out | grep r02
