"""Kernel time of a group-sharded run whose ranks all live on ONE GPU (the setting of run_sharded_local): every cross-rank
channel word goes the way it goes between GPUs (publisher queues / system-scope fence), only NVLink is missing.
usage: [DFD_NPUB=n] python tools/sharded_local_probe.py <world> <radix> <conns> [num_messages]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import codes_b200  # noqa: E402
from codes_b200 import _P  # noqa: E402

world, radix, conns = int(sys.argv[1]), sys.argv[2], sys.argv[3]
msgs = int(sys.argv[4]) if len(sys.argv) > 4 else 10
d = tempfile.mkdtemp()
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_topology.py"), radix, conns, d, "--name", "big", "--routing", "prog-adaptive"],
               check=True, stdout=subprocess.DEVNULL)
conf = os.path.join(d, "big.conf")
engines, streams = [], [torch.cuda.Stream() for _ in range(world)]
for k in range(world):
    e = codes_b200.Engine()
    engines.append(e)
    e.setup(conf, codes_b200.make_opts(traffic=1, num_messages=msgs))
    e.set_partition(k, world)
    e.set_ensemble(world)
    e.upload(streams[k].cuda_stream)
arr = (_P * world)(*[e._h for e in engines])
for e in engines:
    e._ck(e._lib.dfd_engine_attach_local_peers(e._h, arr, world))
best = None
for rep in range(3):
    for k, e in enumerate(engines):
        e.reset(streams[k].cuda_stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(world)]
    for k, e in enumerate(engines):
        with torch.cuda.stream(streams[k]):
            ev[k][0].record()
            e.run(streams[k].cuda_stream)
            ev[k][1].record()
    torch.cuda.synchronize()
    ms = max(a.elapsed_time(b) for a, b in ev)
    best = ms if best is None else min(best, ms)
for k in range(world - 1, -1, -1):
    engines[k].finish(streams[k].cuda_stream)
s = engines[0].summary()
print(f"{world} ranks on one GPU, radix {radix} conns {conns}, NPUB={os.environ.get('DFD_NPUB', 'default')}: events {s.events} kernel {best:.2f} ms -> "
      f"{s.events / best / 1e3:.1f} Mev/s | grid {s.grid_blocks}x{s.block_threads} activations {s.activations}")
for e in engines:
    e.close()
