cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$PWD/graphax_b200/csrc/build/variants
run() { # name, env...
  name=$1; shift
  env "$@" python tools/tune.py --workloads roe3d --divisions reference --threads 128 --min-blocks 3 --out-stages 1 --spt 1 --steps 300 --out gpurun_out/tune_r02b_aot_$name.json > gpurun_out/tune_r02b_aot_$name.log 2>&1
  echo "== $name"; grep -E "variant': '(aot|jit)'" gpurun_out/tune_r02b_aot_$name.log | grep -v best | cut -c1-200
}
run A_pressure_staged GX_SCHEDULE=pressure GX_INPUT_STAGED=1 GX_LIB=$V/libA.so
run B_minpressure_staged GX_SCHEDULE=minpressure GX_INPUT_STAGED=1 GX_LIB=$V/libB.so
run C_minpressure_ldg GX_SCHEDULE=minpressure
run D_pressure_ldg GX_SCHEDULE=pressure GX_LIB=$V/libD.so
