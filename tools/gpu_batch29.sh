#!/bin/bash
# N=8 only: activation budget fine-tune and (TIMING ONLY, DFD_PREFETCH=2) the upper bound of what a cheaper release fence could give
cd "$(dirname "$0")/.."
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 3 --warmup 2 2>gpurun_out/b29_$1.err > gpurun_out/b29_$1.json
  python - <<PY
import json
try:
    d=json.loads([x for x in open("gpurun_out/b29_$1.json") if x.startswith("{")][-1])
    print("$2: value %.1f Mev/s  e2e %.1f Mev/s  ms/step %.2f  digest %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d.get("parity_digest_ok")))
except Exception as ex:
    print("$2 failed", ex)
PY
}
run 29601 "default (budget 10)"
DFD_PREFETCH=2 run 29602 "gpu-scope fence (timing only)"
DFD_ACT_BUDGET=13 run 29603 "budget 13"
DFD_ACT_BUDGET=16 run 29604 "budget 16"
