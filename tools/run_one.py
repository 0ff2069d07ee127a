"""Run one simulation through the C-ABI on the GPU and print the event rate (profiling target for ncu).
usage: python tools/run_one.py <conf> [key=value ...]   e.g. configs/dfdally-1056.conf num_messages=100 traffic=1"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import codes_b200  # noqa: E402

conf = sys.argv[1]
opts = {}
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    opts[k] = float(v) if "." in v else int(v)
eng = codes_b200.Engine()
eng.setup(conf, codes_b200.make_opts(**opts))
eng.upload()
t0 = time.perf_counter()
eng.run()
eng.finish()
dt = time.perf_counter() - t0
s = eng.summary()
print(f"{os.path.basename(conf)} {opts}: events {s.events} in {dt * 1e3:.1f} ms (run+finish) -> {s.events / dt / 1e6:.1f} Mev/s, grid {s.grid_blocks}x{s.block_threads}")
