set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_r02a_${N}gpu.json 2> gpurun_out/bench_r02a_${N}gpu.err; echo "bench$N exit $?"; tail -c 600 gpurun_out/bench_r02a_${N}gpu.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02a_${N}gpu.json'))
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "verify")}); print(d["e2e"])
PY
