#!/bin/bash
# scaling sweep of bench.py on one box (run under gpurun --gpus 8)
cd "$(dirname "$0")/.."
for n in ${NS:-1 2 4 8}; do
  if [ "$n" = 1 ]; then
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/scale_n$n.err > gpurun_out/scale_n$n.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 3 --warmup 3 2>gpurun_out/scale_n$n.err > gpurun_out/scale_n$n.json
  fi
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/scale_n$n.json") if x.startswith("{")][-1]
    d=json.loads(l)
    print("N=$n value %.1f Mev/s  e2e %.1f Mev/s  ms/step %.1f  grid %s  act/step %d  digest %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d["config"]["grid"], d["config"]["activations_per_step"], d.get("parity_digest_ok")))
except Exception as ex:
    print("N=$n failed", ex)
PY
done
