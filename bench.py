#!/usr/bin/env python3
"""bench.py -- committed events/s of the dragonfly-dally hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json north_star target / configs[3] scale): dragonfly-dally with 131,584 terminals (a=32 routers per
group, p=16 terminals and h=16 global channels per router, g=257 groups, 2 links between every pair of groups: the
topology tools/gen_topology.py writes in the reference's binary link-file format), progressive-adaptive (PAR) routing,
uniform-random traffic, 10 messages of 2,048 B per terminal -- 30.5 M committed events per step.
A "step" is one complete simulation: every LP's state is re-initialised on the device, then the persistent event-loop
kernel runs to completion.
  value      committed events / s, device-timed (CUDA events on the launching stream) over the event-loop kernel,
             LP state and connection tables already resident in HBM
  e2e        the same metric through the C-ABI call sequence of the reference's main() with host buffers:
             table upload (H2D) + device init + event loop + statistics download (D2H) + *_final handlers
             (N>1: + merge of the ranks' partitions on rank 0)
  roofline   algorithmic bytes (832*sum(N_hop) + 896*N_chunks per run) / kernel time vs the
             measured HBM copy bandwidth (MEASURED_PEAKS.json)
  other_configs (N=1, extra) the other BASELINE configurations on the same GPU, device-timed like `value`: configs[1] (1,056
             terminals, minimal), configs[2] (8,320 terminals, PAR; uniform and nearest-group), configs[3] (the bench network
             under the worst-case group permutation, also as `worst_case`), and the bench workload with 50 instead of 10 messages
             per terminal (a saturated network: lower throughput); configs[1] and configs[2] with the reference's own
             sources (oracle/_ref, kind "reference", 1 core) timed beside them
  cpu_baseline  (N=1) the CPU implementation of the path, single thread, on a bounded sample of the workload (1 message
             per terminal, 3.1 M events).  At this size that is the restatement oracle (kind "port"): oracle/_ref, the
             reference's own sources on the sequential shim engine, needs 14 minutes and 28 GB to set this network up (done once:
             the golden digest of this workload is checked against it, 0.30 M events/s in its event loop).
--impl reference times that CPU implementation alone (the real ROSS + MPI are external and absent here).
N>1: the dragonfly groups are sharded over the GPUs (one process per GPU); chunks, credits and channel words of the global
links are written straight into the owner GPU's memory over NVLink (CUDA-IPC), there is no barrier and no collective.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = dict(radix=64, conns=2, routing="prog-adaptive")  # a=32 p=16 h=16 g=257 -> 8,224 routers, 131,584 terminals
WORKLOAD = dict(traffic=1, num_messages=10, payload_sz=2048)
SAMPLE = dict(traffic=1, num_messages=1, payload_sz=2048)   # bounded sample of the workload for the CPU arm
WORKLOAD_NAME = ("dragonfly-dally 131,584-terminal (a=32,p=16,h=16,g=257, 2 links per group pair) uniform-random, "
                 "progressive-adaptive (PAR) routing, 10 msgs/terminal x 2048 B, arrival 1000 ns")


def make_network(tmp):
    """Writes the link files and the .conf of the bench network (same binary formats the reference reads)."""
    d = os.path.join(tmp, "net")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_topology.py"), str(SHAPE["radix"]), str(SHAPE["conns"]), d,
                    "--name", "bench", "--routing", SHAPE["routing"]], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(d, "bench.conf")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6 and f[0].isdigit():
                self.rows.append(f)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(int(r[0]) for r in self.rows), "sm_max_mhz": int(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


def measured_traffic():
    """DRAM bytes per launch of the window kernel for this workload, from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)["dram_bytes_per_launch"]
    return None


def parity_digest(eng, tmp):
    """lp-io files of the engine's last run (the e2e leg: rank 0 holds the merged result of a sharded run) hashed against
    the committed digests of the restatement oracle for this workload (tests/golden/131584_par_uniform_m10.json,
    `python -m tests.make_golden --large 131584_par_uniform_m10`).  Outside every timed region."""
    import hashlib
    gold_path = os.path.join(ROOT, "tests", "golden", "131584_par_uniform_m10.json")
    if not os.path.exists(gold_path):
        return None
    with open(gold_path) as f:
        gold = json.load(f)
    d = os.path.join(tmp, "lpio")
    eng.lp_io_flush(d)
    ok = eng.summary().events == gold["events"]
    for name, digest in gold["sha256"].items():
        if name == "summary.txt":
            continue
        with open(os.path.join(d, name), "rb") as f:
            ok = ok and hashlib.sha256(f.read()).hexdigest() == digest
    return bool(ok)


def algorithmic_bytes(total_hops, chunks):
    return 832 * total_hops + 896 * chunks


def cpu_sample(conf):
    """One run of the CPU implementation on the bounded sample; returns (events, packets, event-loop seconds, kind)."""
    from tests import oracle_util
    r = oracle_util.run(conf, **SAMPLE)
    return r.events, r.finished_packets, r.run_seconds, "port"


CPU_SAMPLE_TEXT = ("1 of the workload's 10 messages per terminal on the same 131,584-terminal network (3.1 M events per run): the "
                   "restatement oracle (oracle/dfdally_oracle.cpp), single-threaded -- oracle/_ref (the reference's own sources on "
                   "the sequential shim engine) needs 14 min and 28 GB to set this network up (run once: 30,490,527 events at 0.30 M events/s, output "
                   "byte-identical to the golden digest every bench run is compared with); ROSS+MPI are external and absent")


def run_reference(args, rank):
    """The reference arm: the CPU implementation of the path on the box's host cores, one bounded sample per step.
    The algorithm is sequential (one event list), so it uses one core."""
    if rank != 0:
        return
    from tests import oracle_util
    oracle_util.load()
    events = packets = 0
    loop = 0.0
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        conf = make_network(tmp)
        for _ in range(args.warmup):
            cpu_sample(conf)
        for _ in range(args.steps):
            ev, pk, secs, kind = cpu_sample(conf)
            events += ev
            packets += pk
            loop += secs
    wall = time.perf_counter() - t0
    value = events / loop
    line = {"metric": "committed events/sec", "value": value, "unit": "events/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * loop / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "packets_per_s": packets / loop,
                       "note": "event loop only (config parsing / table construction excluded); wall incl. setup %.3f s" % wall},
            "cpu_baseline": {"value": value, "unit": "events/s", "cores": 1, "kind": kind,
                             "sample": "each step: " + CPU_SAMPLE_TEXT},
            "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def device_timed(eng, stream, flush, fence, sharded, steps, warmup):
    """W untimed + K timed runs of the event-loop kernel; returns the summed kernel time in ms (this rank)."""
    import torch

    def one_step(timed):
        if sharded:
            fence()  # nobody may re-initialise its exchange region while a peer's kernel is still running
        eng.reset(stream)
        flush.fill_(1)  # evict LP state from L2 between steps
        if sharded:
            fence()
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run(stream)
            e1.record()
            return e0, e1
        eng.run(stream)
        return None

    for _ in range(max(warmup, 0)):
        one_step(False)
    fence()
    evs = [one_step(True) for _ in range(steps)]
    fence()
    return sum(a.elapsed_time(b) for a, b in evs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the figures for the other BASELINE configurations (N=1 only)")
    ap.add_argument("--replicas", action="store_true", help="N>1: run N independent replicas instead of the group-sharded engine")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import codes_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the dragonfly-dally engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("cpu:gloo,cuda:nccl")

    tmpdir = tempfile.TemporaryDirectory()
    conf = make_network(tmpdir.name)  # every rank writes its own copy (deterministic, 1 MB)
    w = dict(WORKLOAD)
    opts = codes_b200.make_opts(**w)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    sharded = world > 1 and not args.replicas

    def fence():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    eng = codes_b200.Engine()
    eng.setup(conf, opts)
    if sharded:
        eng.set_partition(rank, world)
    eng.upload(stream)
    if sharded:
        handles = [None] * world
        dist.all_gather_object(handles, eng.ipc_export())
        eng.ipc_import(handles)
    fence()
    sampler = ClockSampler(local_rank)
    device_timed(eng, stream, flush, fence, sharded, 0, args.warmup)  # W untimed warm-up steps
    sampler.start()
    kernel_ms = device_timed(eng, stream, flush, fence, sharded, args.steps, 0)
    clocks = sampler.stop()
    fence()  # every rank's event loop is over
    eng.finish(stream)  # sharded: rank 0 reads every partition's statistics over the CUDA-IPC mappings
    fence()
    s = eng.summary()
    meta = [s.events, s.finished_packets, s.total_hops, s.finished_chunks, s.activations] if rank == 0 or not sharded else [0] * 5
    if sharded:
        mt = torch.tensor(meta, dtype=torch.int64, device="cuda")
        dist.broadcast(mt, 0)
        meta = [int(x) for x in mt]
    events_per_step, packets_per_step, hops, chunks, activations = meta

    # e2e: the reference's main() call sequence on an engine whose configuration is loaded: host buffers on both sides,
    # H2D of the connection tables + device init + event loop + D2H of the statistics + *_final inside the timed region
    if sharded:
        codes_b200.sharded_tw_run(eng, dist, stream)  # warm
        fence()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            codes_b200.sharded_tw_run(eng, dist, stream)
        fence()
        e2e_s = time.perf_counter() - t0
        es = eng.summary()
        h2d, d2h = es.h2d_bytes, es.d2h_bytes
        if rank == 0:
            assert es.events == events_per_step, (es.events, events_per_step)
    else:
        eng.tw_run(stream)  # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.tw_run(stream)
        e2e_s = time.perf_counter() - t0
        es = eng.summary()
        assert es.events == events_per_step
        h2d, d2h = es.h2d_bytes, es.d2h_bytes

    digest_ok = parity_digest(eng, tmpdir.name) if rank == 0 else None

    t = torch.tensor([kernel_ms, e2e_s], dtype=torch.float64, device="cuda")
    io = torch.tensor([h2d, d2h], dtype=torch.int64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(io, op=dist.ReduceOp.SUM)  # every rank uploads the tables and downloads its own partition
    kernel_ms, e2e_s = float(t[0]), float(t[1])
    h2d, d2h = int(io[0]), int(io[1])
    jobs = 1 if (sharded or world == 1) else world  # replicas: N independent simulations
    total_events = events_per_step * args.steps * jobs
    value = total_events / (kernel_ms * 1e-3)
    peak, peak_src = peaks()
    ach = algorithmic_bytes(hops, chunks) * args.steps * jobs / (kernel_ms * 1e-3) / 1e9
    if world == 1:
        par = "1 GPU"
    elif sharded:
        par = (f"dragonfly groups sharded over {world} GPUs (one process per GPU, {s.num_routers // world} routers each): chunks, credits and "
               f"channel words of the global links and the ACK statistics are written into the owner GPU's memory over NVLink P2P "
               f"(CUDA-IPC); no barrier, no collective on the data path")
    else:
        par = f"{world} independent replicas"
    line = {"metric": "committed events/sec", "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kernel_ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if jobs > 1 else "strong",
            "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "events_per_step": events_per_step, "packets_per_step": packets_per_step,
                       "packets_per_s": packets_per_step * args.steps * jobs / (kernel_ms * 1e-3),
                       "activations_per_step": activations,
                       "l2": "256 MiB buffer written between timed steps (L2 flush); LP state of this network is 3.4 GB",
                       "grid": [s.grid_blocks, s.block_threads], "device_bytes": s.device_bytes, "parallelism": par},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak * world, "unit": "GB/s",
                         "frac": ach / (peak * world), "traffic": measured_traffic() if world == 1 else None,
                         "algorithmic_bytes_per_launch": algorithmic_bytes(hops, chunks) // (world if sharded else 1),
                         "peak_source": peak_src, "kernel": "dfd_run_kernel (one launch per step per GPU)",
                         "note": "instruction-supply- and latency-bound, not bandwidth-bound: one warp per router group executes branchy event handlers "
                                 "(GPC instruction cache at 95 % of its request rate, DRAM at 8 %; DESIGN.md 5, profiles/r2_summary.md)"},
            "e2e": {"value": events_per_step * args.steps * jobs / e2e_s, "unit": "events/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps, "gpu_launches_per_step": 2 * world},
            # inside the device-timed region: one dfd_run_kernel* launch per step and GPU (the e2e region adds dfd_init_kernel)
            "gpu_launches": args.steps * world, "clocks": clocks,
            # every lp-io file of the (merged) result hashes to the CPU oracle's digest for this workload
            "parity_digest_ok": digest_ok}
    if world == 1 and not args.no_extras:
        # Not the headline: the other BASELINE configurations on the same GPU, device-timed like `value` (1 warm-up + 3
        # timed runs each, L2 flushed in between), each with its own roofline fraction on algorithmic bytes.
        def extra(name, conf_path, workload_text, cpu_ref=False, **kw):
            try:
                e2 = codes_b200.Engine()
                e2.setup(conf_path, codes_b200.make_opts(**kw))
                e2.upload(stream)
                ms = device_timed(e2, stream, flush, fence, False, 3, 1)
                e2.finish(stream)
                s2 = e2.summary()
                e2.close()
                bps = algorithmic_bytes(s2.total_hops, s2.finished_chunks) * 3 / (ms * 1e-3) / 1e9
                out = {"workload": workload_text, "value": 3 * s2.events / (ms * 1e-3), "unit": "events/s",
                       "events_per_step": s2.events, "steps": 3, "ms_per_step": ms / 3, "grid": [s2.grid_blocks, s2.block_threads],
                       "roofline_frac": bps / peak, "achieved_gbs": bps}
                if cpu_ref and not args.no_cpu_baseline:
                    from tests import oracle_util
                    if oracle_util.ref_available():  # the reference's own sources on the sequential shim engine, 1 core
                        ev, secs, _ = oracle_util.run_ref(conf_path, os.path.join(tmpdir.name, "ref_" + name), **kw)
                        out["cpu_reference"] = {"value": ev / secs, "unit": "events/s", "cores": 1, "kind": "reference",
                                                "events_identical": ev == s2.events}
                line.setdefault("other_configs", {})[name] = out
            except Exception as ex:  # an extra figure must never cost the headline
                line.setdefault("other_configs", {})[name] = {"error": repr(ex)[:200]}

        cdir = os.path.join(ROOT, "configs")
        extra("configs1_1056_minimal_uniform", os.path.join(cdir, "dfdally-1056.conf"),
              "BASELINE configs[1]: 1,056 terminals, minimal routing, uniform, 100 msgs/terminal", cpu_ref=True, traffic=1, num_messages=100)
        extra("configs2_8320_par_uniform", os.path.join(cdir, "dfdally-8320-par.conf"),
              "BASELINE configs[2] network: 8,320 terminals, PAR, uniform, 20 msgs/terminal", cpu_ref=True, traffic=1, num_messages=20)
        extra("configs2_8320_par_nearest_group", os.path.join(cdir, "dfdally-8320-par.conf"),
              "same network, nearest-group (worst-case group permutation) traffic", traffic=3, num_messages=20)
        extra("configs3_131584_par_nearest_group", conf,
              "BASELINE configs[3]: the bench network under the worst-case group permutation", **dict(w, traffic=3))
        # the same network and traffic in a longer run: at 95 % injection load the network saturates after ~15 us of simulated time and
        # the credit lanes' lookahead shrinks to the credit delay ("run length") -- the headline workload is the favourable end
        extra("bench_network_50_msgs_per_terminal", conf,
              "the bench network and traffic with 50 instead of 10 msgs/terminal (saturated network)", **dict(w, num_messages=50))
        if "value" in line["other_configs"].get("configs3_131584_par_nearest_group", {}):
            wc = line["other_configs"]["configs3_131584_par_nearest_group"]
            line["worst_case"] = {"workload": wc["workload"], "value": wc["value"], "unit": "events/s",
                                  "events_per_step": wc["events_per_step"], "steps": 3}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # the CPU baseline is timed at N=1 only
        ev, pk, secs, kind = cpu_sample(conf)
        line["cpu_baseline"] = {"value": ev / secs, "unit": "events/s", "cores": 1, "kind": kind,
                                "sample": "one run (%d events, %.2f s event loop): " % (ev, secs) + CPU_SAMPLE_TEXT}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    tmpdir.cleanup()


if __name__ == "__main__":
    main()
