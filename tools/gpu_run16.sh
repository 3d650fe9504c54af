cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" python tools/tune.py --workloads roe3d --divisions reference --threads 128,384 --min-blocks 1,3 --out-stages 1 --spt 1 --steps 300 --out gpurun_out/tune_r02b_sync_$name.json > gpurun_out/tune_r02b_sync_$name.log 2>&1
  echo "== $name"; grep -E "^ +\{'variant': 'jit'" gpurun_out/tune_r02b_sync_$name.log | cut -c1-175
}
run ldg_nosync GX_TILE_SYNC=0
run ldg_sync GX_TILE_SYNC=1
run staged_sync GX_TILE_SYNC=1 GX_INPUT_STAGED=1
