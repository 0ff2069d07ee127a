#!/bin/bash
# round-2 re-entry: full GPU suite on the build with R_SNAPSHOT events + a short bench run (no perf change expected)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/b15_tests.log 2>&1
tail -15 gpurun_out/b15_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/b15_bench.json 2> gpurun_out/b15_bench.err
python - <<'PY'
import json
try:
    d = json.loads([x for x in open("gpurun_out/b15_bench.json") if x.startswith("{")][-1])
    print("bench value %.1f Mev/s e2e %.1f Mev/s ms/step %.2f digest %s cpu %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d.get("parity_digest_ok"), d.get("cpu_baseline", {}).get("value")))
except Exception as ex:
    print("bench failed", ex)
PY
