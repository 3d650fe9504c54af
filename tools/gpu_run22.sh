cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" python tools/tune.py --workloads roe3d --orders mixed,rev --divisions reference --threads 256,384,512 --min-blocks 1 --out-stages 1 --spt 1 --steps 300 --out gpurun_out/tune_r02b_gsync_$name.json > gpurun_out/tune_r02b_gsync_$name.log 2>&1
  echo "== $name"; grep -E "^ +\{'variant': 'jit'" gpurun_out/tune_r02b_gsync_$name.log | cut -c1-175
}
run off GX_GROUP_SYNC=0
run on GX_GROUP_SYNC=1
