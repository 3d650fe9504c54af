set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_2gpu.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_r02a_2gpu.json 2> gpurun_out/bench_r02a_2gpu.err; echo "bench2 exit $?"; tail -c 800 gpurun_out/bench_r02a_2gpu.err; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r02a_2gpu.json'))
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "verify")}); print(d["e2e"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_r02a_2gpu_reference.json 2> gpurun_out/bench_r02a_2gpu_reference.err; echo "ref arm exit $?"; cut -c1-400 gpurun_out/bench_r02a_2gpu_reference.json
