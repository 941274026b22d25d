#!/bin/bash
# usage (GPU box, N GPUs): bash scripts/dev_defer.sh N  -- bench line with the near field deferred (1) and not (0)
n=${1:-2}
for d in 1 0; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $n --no-c5 --no-cpu-baseline --plan-option p2p_defer=$d 2>/dev/null | grep "^{" > gpurun_out/defer_${n}_$d.json
  python - <<P
import json
d = json.load(open("gpurun_out/defer_${n}_$d.json"))
print("gpus $n p2p_defer $d: %.4f ms device, %.4f ms e2e" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
P
done
