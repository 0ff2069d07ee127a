"""Quick GPU probe: parity against the oracle on a few cases + device-timed event rate of the event-loop kernel.
usage: python tools/gpu_probe.py [parity] [perf]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import codes_b200  # noqa: E402
from tests import oracle_util  # noqa: E402

C72 = os.path.join(ROOT, "configs", "dfdally-72.conf")
C1056 = os.path.join(ROOT, "configs", "dfdally-1056.conf")
C8K = os.path.join(ROOT, "configs", "dfdally-8320-par.conf")


def parity():
    cases = [(C72, dict(num_messages=1)), (C72, dict(num_messages=20)), (C72, dict(num_messages=10, traffic=3)),
             (C72, dict(num_messages=10, traffic=2)), (C1056, dict(num_messages=10)), (C1056, dict(num_messages=5, traffic=3)),
             (C8K, dict(num_messages=3)), (C8K, dict(num_messages=3, traffic=3))]
    bad = 0
    for conf, opts in cases:
        with tempfile.TemporaryDirectory() as d:
            try:
                s = codes_b200.run_synthetic(conf, lp_io_dir=d + "/g", **opts)
            except codes_b200.EngineError as ex:
                print("ERROR", os.path.basename(conf), opts, ex, flush=True)
                bad += 1
                continue
            o = oracle_util.run(conf, lp_io_dir=d + "/o", **opts)
            g, r = oracle_util.read_dir(d + "/g"), oracle_util.read_dir(d + "/o")
            diff = [n for n in r if g.get(n) != r[n]]
            ok = not diff and s.events == o.events
            bad += not ok
            print("IDENTICAL" if ok else "DIFFERENT", os.path.basename(conf), opts, "events", s.events, o.events, diff, flush=True)
    return bad


def perf(conf, reps=3, **opts):
    eng = codes_b200.Engine()
    eng.setup(conf, codes_b200.make_opts(**opts))
    st = torch.cuda.current_stream().cuda_stream
    eng.upload(st)
    best = None
    for _ in range(reps):
        eng.reset(st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run(st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    eng.finish(st)
    s = eng.summary()
    print(f"{os.path.basename(conf)} {opts}: events {s.events} kernel {best:.2f} ms -> {s.events / best / 1e3:.1f} Mev/s | grid {s.grid_blocks}x{s.block_threads} "
          f"passes {s.passes} activations {s.activations} productive {s.windows} ev/prod {s.events / max(s.windows, 1):.1f} "
          f"busy/idle cyc per warp {s.cyc_busy:.0f}/{s.cyc_idle:.0f} max busy {s.cyc_busy_max_warp:.0f} mem {s.device_bytes / 2**20:.0f} MiB", flush=True)
    eng.close()


def big(radix, conns, routing, traffic, msgs):
    import subprocess
    d = tempfile.mkdtemp()
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_topology.py"), str(radix), str(conns), d, "--name", "big", "--routing", routing], check=True)
    perf(os.path.join(d, "big.conf"), reps=2, num_messages=msgs, traffic=traffic)


if __name__ == "__main__":
    what = sys.argv[1:] or ["parity", "perf"]
    rc = 0
    if "parity" in what:
        rc = parity()
    if "perf" in what:
        perf(C72, num_messages=20)
        perf(C1056, num_messages=100)
        perf(C8K, num_messages=20)
        perf(C8K, num_messages=20, traffic=3)
    if "big" in what or "big131" in what:
        big(64, 2, "minimal", 1, 10)
        big(64, 2, "prog-adaptive", 1, 10)
        big(64, 2, "prog-adaptive", 3, 10)
    if "big" in what or "big1m" in what:
        big(128, 4, "minimal", 1, 5)
    sys.exit(1 if rc else 0)
