cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r02b.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/gpu_tests_r02b.log
timeout 900 python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r02b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02b.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["e2e"]["value"], d["clocks"])
print({k: (round(v["ms_per_step"]*1e3, 1), round(v["roofline_frac"], 3)) for k, v in d["other_orders"].items()})
print(d["rounding_modes"]["fma"]["ms_per_step"], d["verify"])
print(json.dumps(d.get("other_configs"))[:1500])
PY
