"""Multi-GPU parity check of the group-sharded engine (run with torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Rank 0 compares the merged lp-io output with the CPU oracle: N-GPU results must be bit-identical to the
sequential run (the reference's analogue: --sync=1 vs --sync=3 equality, tests/CMakeLists.txt:66-68)."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import codes_b200
from tests import oracle_util

CASES = [
    ("configs/dfdally-72.conf", dict(num_messages=20)),
    ("configs/dfdally-72.conf", dict(num_messages=10, traffic=3)),
    ("configs/dfdally-1056.conf", dict(num_messages=10)),
    ("configs/dfdally-8320-par.conf", dict(num_messages=3, traffic=3)),
    ("configs/dfdally-1056.conf", dict(num_messages=5, payload_sz=10000)),   # multi-packet messages across GPUs
    ("configs/dfdally-72-snapshots.conf", dict(num_messages=20, traffic=3)),  # R_SNAPSHOT records gathered from every GPU
]


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    ok = True
    cases = list(CASES)
    # parallel global links between groups (the far-end port pairing of the link lanes), generated on the fly
    import subprocess
    gen = os.path.join(tempfile.gettempdir(), "mgpu_check_shapes")
    if rank == 0:
        for radix, conns, routing in ((15, 2, "prog-adaptive"), (23, 4, "minimal")):
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_topology.py"), str(radix), str(conns), os.path.join(gen, f"r{radix}c{conns}"),
                            "--name", "t", "--routing", routing], check=True, stdout=subprocess.DEVNULL)
    dist.barrier()
    cases.append((os.path.join(gen, "r15c2", "t.conf"), dict(num_messages=4, traffic=3)))
    cases.append((os.path.join(gen, "r23c4", "t.conf"), dict(num_messages=4)))
    for conf, opts in cases:
        conf = os.path.join(ROOT, conf)
        with tempfile.TemporaryDirectory() as tmp:
            gdir, odir = os.path.join(tmp, "gpu"), os.path.join(tmp, "oracle")
            t0 = time.perf_counter()
            s = codes_b200.run_sharded(conf, dist, lp_io_dir=gdir if rank == 0 else None, **opts)
            dt = time.perf_counter() - t0
            if rank == 0:
                o = oracle_util.run(conf, lp_io_dir=odir, **opts)
                g, r = oracle_util.read_dir(gdir), oracle_util.read_dir(odir)
                same = s.events == o.events and all(g.get(k) == r[k] for k in r)
                ok &= same
                print(f"[{dist.get_world_size()} GPUs] {os.path.basename(os.path.dirname(conf)) + '/' + os.path.basename(conf)} {opts}: events {s.events} vs oracle {o.events}, "
                      f"windows {s.windows}, {dt:.2f} s -> {'IDENTICAL' if same else 'DIFFERENT'}", flush=True)
                if not same:
                    for k in r:
                        if g.get(k) != r[k]:
                            print("   differs:", k)
    # ~100K-terminal (and, with MGPU_CHECK_1M=1, ~1M-terminal) scale against the committed digests (tests/golden/)
    import hashlib
    import json
    large = [("131584_uniform_m3", 64, 2, 3)]
    if os.environ.get("MGPU_CHECK_1M"):
        large.append(("1050624_uniform_m1", 128, 4, 1))
    for gold_name, radix, conns, msgs in large:
        net = os.path.join(tempfile.gettempdir(), "mgpu_check_" + gold_name)
        if rank == 0:
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_topology.py"), str(radix), str(conns), net, "--name", "big"], check=True,
                           stdout=subprocess.DEVNULL)
        dist.barrier()
        with tempfile.TemporaryDirectory() as tmp:
            gdir = os.path.join(tmp, "gpu")
            t0 = time.perf_counter()
            s = codes_b200.run_sharded(os.path.join(net, "big.conf"), dist, lp_io_dir=gdir if rank == 0 else None, num_messages=msgs)
            dt = time.perf_counter() - t0
            if rank == 0:
                with open(os.path.join(ROOT, "tests", "golden", gold_name + ".json")) as f:
                    gold = json.load(f)
                g = oracle_util.read_dir(gdir)
                same = s.events == gold["events"] and all(hashlib.sha256(g[k]).hexdigest() == d for k, d in gold["sha256"].items() if k != "summary.txt")
                ok &= same
                print(f"[{dist.get_world_size()} GPUs] {gold_name}: events {s.events} vs golden {gold['events']}, windows {s.windows}, "
                      f"{dt:.2f} s -> {'IDENTICAL' if same else 'DIFFERENT'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU PARITY OK" if ok else "MGPU PARITY FAIL", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
