cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r02c.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/gpu_tests_r02c.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
S=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_r02c_roe3d_mixed.json 2> gpurun_out/bench_r02c.err; echo "bench exit $? in $(( $(date +%s) - S )) s"; tail -2 gpurun_out/bench_r02c.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02c_roe3d_mixed.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["frac_of_floor"], d["clocks"])
print({k: (round(v["ms_per_step"]*1e3, 1), round(v["roofline_frac"], 3)) for k, v in d["other_orders"].items()})
print(d["rounding_modes"]["fma"]["ms_per_step"], d["verify"]["bit_exact"])
for k, v in d["other_configs"].items(): print(k, round(v["us_per_step"], 2), round(v["roofline_frac"], 3))
PY
S=$(date +%s); python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02c_reference_arm.json 2>> gpurun_out/bench_r02c.err; echo "ref arm exit $? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/bench_r02c_reference_arm.json
echo "== pair mode, lean epilogues"; GX_GEMM_PAIR=1 timeout 300 python tools/mlp_bench.py --graph --steps 400 2>/dev/null | cut -c90-180
timeout 300 python tools/mlp_bench.py --graph --steps 400 --profile > gpurun_out/mlp_r02c.json 2> gpurun_out/mlp_r02c_kernels.txt; cat gpurun_out/mlp_r02c.json | cut -c1-260; tail -7 gpurun_out/mlp_r02c_kernels.txt
timeout 300 python tools/mlp_bench.py --graph --steps 20 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gx_gemm -c 10 -f -o gpurun_out/prof_r02c_mlp python tools/mlp_bench.py --graph --steps 20 > gpurun_out/ncu_mlp_r02c.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/prof_r02c_mlp.ncu-rep
