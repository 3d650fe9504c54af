cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=power.limit,power.max_limit,clocks.max.sm,temperature.gpu --format=csv
python tools/power_probe.py --order mixed --rounding reference --seconds 3 --out gpurun_out/power_mixed_reference.json
python tools/power_probe.py --order mixed --rounding fma --seconds 2 --out gpurun_out/power_mixed_fma.json | grep -E "us_per_step|sm_mhz|watts|window_ms|power_cap" | paste - - - - - 
python tools/power_probe.py --order rev --rounding reference --seconds 2 --out gpurun_out/power_rev_reference.json | grep -E "us_per_step|sm_mhz|watts|window_ms|power_cap" | paste - - - - -
