cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r02c.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/gpu_tests_r02c.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
S=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_r02c_roe3d_mixed.json 2> gpurun_out/bench_r02c.err; echo "bench exit $? in $(( $(date +%s) - S )) s"; tail -2 gpurun_out/bench_r02c.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02c_roe3d_mixed.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["frac_of_floor"], d["clocks"])
for k, v in d["other_configs"].items(): print(k, round(v["us_per_step"], 2), round(v["roofline_frac"], 3))
PY
python bench.py --steps 20 --warmup 3 > gpurun_out/plain_r02c.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_r02c_roe3d_mixed_reference.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_l.log 2>&1; echo "ncu list exit $?"
