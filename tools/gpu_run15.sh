cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$PWD/graphax_b200/csrc/build/variants
prof() { name=$1; shift
  env "$@" ncu --set full --clock-control none --import-source on -k regex:gx_jacve_kernel -s 20 -c 1 -f -o gpurun_out/prof_r02b_$name python tools/power_probe.py --order mixed --rounding reference --seconds 0.05 --idle 0.1 > gpurun_out/ncu_r02b_$name.log 2>&1
  tail -2 gpurun_out/ncu_r02b_$name.log
}
prof A_pressure_staged GX_SCHEDULE=pressure GX_INPUT_STAGED=1 GX_LIB=$V/libA.so
prof B_minpressure_staged GX_SCHEDULE=minpressure GX_INPUT_STAGED=1 GX_LIB=$V/libB.so
prof C_minpressure_ldg GX_SCHEDULE=minpressure
ls -la gpurun_out/*.ncu-rep
