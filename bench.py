#!/usr/bin/env python3
"""bench.py -- committed events/s of the dragonfly-dally hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one complete simulation of the workload (BASELINE.json configs[1]: dragonfly-dally 1,056
terminals, a=8 p=4 h=4 g=33, uniform-random traffic, minimal routing, default credit delay): every LP's
state is re-initialised on the device, then the window kernel runs the event loop to completion.
  value      committed events / s, device-timed (CUDA events on the launching stream) over the window kernel,
             LP state and connection tables already resident in HBM
  e2e        the same metric through the C-ABI call sequence of the reference's main() with host buffers:
             table upload (H2D) + device init + event loop + statistics download (D2H) + *_final handlers
  roofline   algorithmic bytes (832*sum(N_hop) + 896*N_chunks per run) / kernel time vs the
             measured HBM copy bandwidth (MEASURED_PEAKS.json)
  ensemble   (N=1, extra) 32 independent simulations of the workload at once on the GPU: aggregate events / s
  cpu_baseline  (N=1) the reference's own handler sources compiled against the sequential shim engine (oracle/_ref,
             kind "reference"), or the restatement oracle when that binary is absent (kind "port"), single thread,
             on the same workload
--impl reference times that CPU implementation alone (the real ROSS + MPI are external and absent here).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(conf=os.path.join(ROOT, "configs", "dfdally-1056.conf"), traffic=1, num_messages=100, payload_sz=2048)
WORKLOAD_NAME = "dragonfly-dally 1,056-terminal (a=8,p=4,h=4,g=33) uniform-random, minimal routing, 100 msgs/terminal x 2048 B, arrival 1000 ns"


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
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)["dram_bytes_per_launch"]
    return None


def algorithmic_bytes(total_hops, chunks):
    return 832 * total_hops + 896 * chunks


def run_reference(args, rank):
    """The reference arm: the reference's own CPU implementation of the path on the box's host cores.
    oracle/_ref = the reference's handler sources compiled unchanged against a sequential ROSS/MPI shim (ROSS itself
    is external and absent, so this is the reference's model code on a minimal sequential engine: --synch=1 analogue;
    there is no optimistic/MPI mode to run).  Falls back to the restatement oracle when oracle/_ref is missing."""
    if rank != 0:
        return
    import tempfile
    from tests import oracle_util
    w = dict(WORKLOAD)
    conf = w.pop("conf")
    use_ref = oracle_util.ref_available()
    events = packets = 0
    loop = 0.0
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        def once():
            if use_ref:
                ev, secs, _ = oracle_util.run_ref(conf, os.path.join(tmp, "lpio"), **w)
                return ev, 1056 * w["num_messages"], secs
            r = oracle_util.run(conf, **w)
            return r.events, r.finished_packets, r.run_seconds
        if not use_ref:
            oracle_util.load()
        for _ in range(args.warmup):
            once()
        for _ in range(args.steps):
            ev, pk, secs = once()
            events += ev
            packets += pk
            loop += secs
    wall = time.perf_counter() - t0
    value = events / loop
    kind = "reference" if use_ref else "port"
    sample = ("the full workload, %d run(s) of oracle/_ref (the reference's own dragonfly-dally / model-net / synthetic-driver sources, "
              "unchanged, on a single-threaded sequential shim engine; ROSS+MPI are external and absent)" % args.steps) if use_ref else \
             ("the full workload, %d run(s) of the restatement oracle (single-threaded)" % args.steps)
    line = {"metric": "committed events/sec", "value": value, "unit": "events/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * loop / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "packets_per_s": packets / loop,
                       "note": "event loop only (config parsing / table construction excluded); wall incl. setup %.3f s" % wall},
            "cpu_baseline": {"value": value, "unit": "events/s", "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ensemble", action="store_true", help="skip the 32-simulation ensemble figure (N=1 only)")
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

    w = dict(WORKLOAD)
    conf = w.pop("conf")
    opts = codes_b200.make_opts(**w)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    sharded = world > 1 and not args.replicas

    eng = codes_b200.Engine()
    eng.setup(conf, opts)
    if sharded:  # group-sharded over the ranks: P2P inbox writes + cross-GPU window barrier over NVLink
        eng.set_partition(rank, world)
    eng.upload(stream)
    if sharded:
        handles = [None] * world
        dist.all_gather_object(handles, eng.ipc_export())
        eng.ipc_import(handles)

    def fence():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def one_step(timed):
        if sharded:
            fence()  # nobody may reset its sync block while a peer is still closing the last window
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

    for _ in range(max(args.warmup, 0)):
        one_step(False)
    fence()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [one_step(True) for _ in range(args.steps)]
    fence()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    kernel_ms = sum(a.elapsed_time(b) for a, b in evs)
    eng.finish(stream)
    if sharded:
        parts = [None] * world
        dist.gather_object(eng.export_partition(), parts if rank == 0 else None, dst=0)
        if rank == 0:
            for k in range(1, world):
                eng.import_partition(parts[k])
            eng.finalize()
    s = eng.summary()
    meta = [s.events, s.finished_packets, s.total_hops, s.finished_chunks, s.windows] if rank == 0 or not sharded else [0] * 5
    if sharded:
        mt = torch.tensor(meta, dtype=torch.int64, device="cuda")
        dist.broadcast(mt, 0)
        meta = [int(x) for x in mt]
    events_per_step, packets_per_step, hops, chunks, windows = meta

    # e2e: the reference's main() call sequence, host buffers on both sides, H2D + D2H inside the timed region
    if sharded:
        codes_b200.run_sharded(conf, dist, stream=stream, **w)  # warm
        fence()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            es = codes_b200.run_sharded(conf, dist, stream=stream, **w)
        fence()
        e2e_s = time.perf_counter() - t0
        h2d, d2h = (es.h2d_bytes, es.d2h_bytes) if rank == 0 else (0, 0)
    else:
        e2e_eng = codes_b200.Engine()
        e2e_eng.setup(conf, opts)
        e2e_eng.tw_run(stream)  # warm (allocations, module load)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_eng.tw_run(stream)
        e2e_s = time.perf_counter() - t0
        es = e2e_eng.summary()
        assert es.events == events_per_step
        h2d, d2h = es.h2d_bytes, es.d2h_bytes

    t = torch.tensor([kernel_ms, e2e_s], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    kernel_ms, e2e_s = float(t[0]), float(t[1])
    jobs = 1 if (sharded or world == 1) else world  # replicas: N independent simulations
    total_events = events_per_step * args.steps * jobs
    value = total_events / (kernel_ms * 1e-3)
    peak, peak_src = peaks()
    ach = algorithmic_bytes(hops, chunks) * args.steps * jobs / (kernel_ms * 1e-3) / 1e9
    if world == 1:
        par = "1 GPU"
    elif sharded:
        par = f"dragonfly groups sharded over {world} GPUs (one process per GPU): global-link chunks/credits and ACK statistics are " \
              f"written into the owner GPU's inbox over NVLink P2P, one flag+minimum exchange per lookahead window"
    else:
        par = f"{world} independent replicas"
    line = {"metric": "committed events/sec", "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kernel_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "events_per_step": events_per_step, "packets_per_step": packets_per_step,
                       "packets_per_s": packets_per_step * args.steps * jobs / (kernel_ms * 1e-3),
                       "windows_per_step": windows, "window_ns": s.window_ns, "us_per_window": 1e3 * kernel_ms / args.steps / max(windows, 1),
                       "l2": "256 MiB buffer written between timed steps (L2 flush)",
                       "grid": [s.grid_blocks, s.block_threads], "device_bytes": s.device_bytes, "parallelism": par},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak * (world if sharded or jobs > 1 else 1), "unit": "GB/s",
                         "frac": ach / (peak * (world if sharded or jobs > 1 else 1)), "traffic": measured_traffic() if world == 1 else None,
                         "algorithmic_bytes_per_launch": algorithmic_bytes(hops, chunks),
                         "peak_source": peak_src, "kernel": "dfd_run_kernel (one launch per step per GPU)",
                         "note": "the kernel is bound by per-window latency (grid barrier + dependent instruction chain of the slowest LP), not by HBM bandwidth; see DESIGN.md 5"},
            "e2e": {"value": events_per_step * args.steps * jobs / e2e_s, "unit": "events/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": 2 * args.steps, "clocks": clocks}
    if world == 1 and not args.no_ensemble:
        try:
            # Not the headline: the same workload as a parameter sweep -- 32 independent simulations at once on this GPU,
            # one engine and CUDA stream each (codes_b200.run_ensemble), device-timed from the first launch to the last finish
            k = 32
            engs, streams = [], [torch.cuda.Stream() for _ in range(k)]
            for st in streams:
                e = codes_b200.Engine()
                e.setup(conf, opts)
                e.set_ensemble(k)
                e.upload(st.cuda_stream)
                engs.append(e)
            torch.cuda.synchronize()
            for rep in range(2):  # first pass warms up
                for e, st in zip(engs, streams):
                    e.reset(st.cuda_stream)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for e, st in zip(engs, streams):
                    st.wait_event(e0)
                    e.run(st.cuda_stream)
                for st in streams:
                    torch.cuda.current_stream().wait_stream(st)
                e1.record()
                torch.cuda.synchronize()
                ens_ms = e0.elapsed_time(e1)
            ens_events = 0
            for e, st in zip(engs, streams):
                e.finish(st.cuda_stream)
                ens_events += e.summary().events
                e.close()
            assert ens_events == k * events_per_step
            line["ensemble"] = {"simulations": k, "value": ens_events / (ens_ms * 1e-3), "unit": "events/s", "ms": ens_ms,
                                "note": "aggregate of 32 independent simulations of the same workload running concurrently on the one GPU "
                                        "(a parameter sweep); `value` above is a single simulation"}
        except Exception as ex:  # the extra figure must never cost the headline
            line["ensemble"] = {"error": repr(ex)[:200]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # the CPU baseline is timed at N=1 only
        import tempfile
        from tests import oracle_util
        if oracle_util.ref_available():
            with tempfile.TemporaryDirectory() as tmp:
                ev, secs, _ = oracle_util.run_ref(conf, os.path.join(tmp, "lpio"), **w)
            assert ev == events_per_step, (ev, events_per_step)
            line["cpu_baseline"] = {"value": ev / secs, "unit": "events/s", "cores": 1, "kind": "reference",
                                    "sample": "the full workload once (%d events, %.2f s event loop): oracle/_ref = the reference's own handler "
                                              "sources on a single-threaded sequential shim engine (ROSS+MPI are external and absent)" % (ev, secs)}
        else:
            oracle_util.load()
            r = oracle_util.run(conf, **w)
            assert r.events == events_per_step, (r.events, events_per_step)
            line["cpu_baseline"] = {"value": r.events / r.run_seconds, "unit": "events/s", "cores": 1, "kind": "port",
                                    "sample": "the full workload once (%d events, %.2f s): restatement oracle, single-threaded" % (r.events, r.run_seconds)}
    if rank == 0:
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
