cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SHORT="--steps 20 --warmup 3"
python bench.py > gpurun_out/bench_r02b_roe3d_mixed.json 2> gpurun_out/bench_r02b.err; echo "bench exit $?"
python bench.py $SHORT > gpurun_out/plain_r02b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_r02b_roe3d_mixed_reference.csv python bench.py $SHORT > gpurun_out/ncu_l.log 2>&1
python bench.py $SHORT > gpurun_out/plain_r02b2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gx_jacve_kernel -s 8 -c 2 -f -o gpurun_out/prof_r02b_mixed_reference python bench.py $SHORT > gpurun_out/ncu_f1.log 2>&1
python bench.py $SHORT --rounding fma > gpurun_out/plain_r02b_fma.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gx_jacve_kernel -s 8 -c 2 -f -o gpurun_out/prof_r02b_mixed_fma python bench.py $SHORT --rounding fma > gpurun_out/ncu_f2.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02b_reference_arm.json 2>> gpurun_out/bench_r02b.err; echo "ref arm exit $?"
ls -la gpurun_out/prof_r02b_mixed_*.ncu-rep gpurun_out/ncu_launches_r02b_roe3d_mixed_reference.csv
tail -c 600 gpurun_out/bench_r02b_reference_arm.json
