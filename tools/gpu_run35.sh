cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "host or chunk or pinned or large_plan or api or ragged" 2>&1 | tail -3
for rep in 1 2; do for v in 1 0; do
  GX_HOST_RAMP=$v timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ramp_$v.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_ramp_$v.json").read().strip().splitlines()[-1])
print("ramp=$v rep=$rep e2e", round(d["e2e"]["value"]/1e6, 2), "M  frac_of_floor", round(d["e2e"]["frac_of_floor"], 4), "floor", round(d["e2e"]["floor"]["value"]/1e6, 2))
PY
done; done
