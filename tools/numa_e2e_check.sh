#!/bin/bash
# e2e (host buffers in / out) at N GPUs with and without binding each rank to its GPU's NUMA node.
N=${1:-8}
nvidia-smi topo -m 2>/dev/null | head -14
for b in 1 0; do
  GX_NUMA_BIND=$b python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2951$b bench.py --gpus $N --steps 200 --warmup 5 --no-other-orders 2>/dev/null > /tmp/numa_$b.json
  python - "$b" <<'PY'
import json, sys
d = json.loads(open(f"/tmp/numa_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("bind", sys.argv[1], "value G/s", d["value"] / 1e9, "e2e G/s", d["e2e"]["value"] / 1e9, d["config"].get("host_affinity"))
PY
done
