#!/usr/bin/env python3
"""Randomised parity campaign (seeded, reproducible): random PARAMS / driver options on small generated dragonflies.

  python tools/fuzz_parity.py ref  [--cases N] [--seed S]   restatement oracle vs oracle/_ref (reference sources)  -- CPU
  python tools/fuzz_parity.py gpu  [--cases N] [--seed S]   CUDA engine vs restatement oracle                      -- GPU box
  python tools/fuzz_parity.py sim  [--cases N] [--seed S]   the engine's scheduler + handlers on the CPU harness (tests/cpusim,
                                                            random interleaving, 1-4 emulated ranks) vs restatement oracle  -- CPU
  python -m torch.distributed.run --nproc-per-node G tools/fuzz_parity.py mgpu ...   group-sharded engine over G GPUs vs oracle

Every case compares all lp-io files (and, on the GPU, the hex-float raw statistics) byte for byte.  A case whose
configuration the reference itself rejects is skipped when both sides reject it."""
import argparse
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_util  # noqa: E402

# (radix, links between groups); (11, 9) / (15, 16) / (15, 8) / (11, 6) have routers with several links to one group
SHAPES = [(7, 1), (11, 1), (11, 2), (15, 1), (15, 2), (19, 2), (23, 4), (11, 9), (15, 16), (15, 8), (11, 6)]


def make_case(rng):
    radix, conns = rng.choice(SHAPES)
    routing = rng.choice(["minimal", "minimal", "nonminimal", "prog-adaptive", "prog-adaptive"])
    chunk = rng.choice([256, 512, 1024, 2048, 4096])
    packet = chunk * rng.choice([1, 1, 2, 4])
    vc_chunks = rng.choice([1, 2, 4, 8])
    params = dict(chunk_size=chunk, packet_size=packet,
                  local_vc_size=chunk * vc_chunks * rng.choice([1, 2]), global_vc_size=chunk * vc_chunks * rng.choice([1, 2]),
                  cn_vc_size=chunk * vc_chunks * rng.choice([1, 2, 4]),
                  local_bandwidth=rng.choice(["2.0", "5.25", "12.5"]), global_bandwidth=rng.choice(["2.0", "4.7", "12.5"]),
                  cn_bandwidth=rng.choice(["2.0", "8.0", "16.0"]))
    if rng.random() < 0.3:
        params["router_delay"] = rng.choice(["0", "50", "300"])
    if rng.random() < 0.3:
        params["credit_delay"] = rng.choice(["10", "50", "100"])
    if rng.random() < 0.2:
        params[rng.choice(["cn_delay", "local_delay", "global_delay"])] = rng.choice(["5", "30", "250"])
    if rng.random() < 0.15:
        params[rng.choice(["cn_credit_delay", "local_credit_delay", "global_credit_delay"])] = rng.choice(["2", "20"])
    if rng.random() < 0.15:
        params["nic_seq_delay"] = rng.choice(["0", "25", "400"])
    if rng.random() < 0.1:
        params["credit_size"] = rng.choice(["4", "16", "64"])
    if rng.random() < 0.1:
        params["global_k_picks"] = rng.choice(["1", "2"])  # k > links to one group hangs the reference (dally:1822-1839)
    if routing == "prog-adaptive":
        if rng.random() < 0.5:
            params["adaptive_threshold"] = str(rng.choice([0, chunk, 4 * chunk]))
        if rng.random() < 0.5:
            params["route_scoring_metric"] = rng.choice(["alpha", "delta"])
    opts = dict(traffic=rng.choice([1, 1, 2, 3, 4, 5]), num_messages=rng.randint(1, 8),
                payload_sz=rng.choice([1, 8, 200, 1024, 2048, 3000, 4096, 4097, 9000]))
    if rng.random() < 0.5:
        opts["arrival_time"] = rng.choice([50.0, 200.0, 1000.0, 5000.0])
    elif rng.random() < 0.3:
        opts["load_per_svr"] = rng.choice([0.2, 0.5, 0.9])
    if rng.random() < 0.2:
        opts["warm_up_time"] = rng.choice([500.0, 3000.0])
    if rng.random() < 0.15:
        opts["end_time"] = rng.choice([4000.0, 20000.0])
    if rng.random() < 0.15:
        params["_switch_conns"] = 2
    if opts["traffic"] == 2 and rng.random() < 0.6:
        # RAND_PERM produces self-sends: below node_eager_limit they take the no-op shortcut of model-net.c:305-309,
        # otherwise (or with codes_noop_bypass) the base LP copies them inside the node (model-net-lp.c:672-704)
        params["node_eager_limit"] = rng.choice(["0", "100", "5000", "100000"])
        if rng.random() < 0.5:
            params["codes_noop_bypass"] = rng.choice(["0", "1"])
    if rng.random() < 0.2:  # router VC-occupancy snapshots (dragonfly-snapshots.csv): unsorted, possibly equal or past the end
        times = [rng.choice(["500", "2500.5", "6000", "6000", "15000", "60000"]) for _ in range(rng.randint(1, 3))]
        params["router_buffer_snapshots"] = "( " + ",".join(f'"{t}"' for t in times) + " )"
    return radix, conns, routing, params, opts


def write_conf(tmp, radix, conns, routing, params):
    net = os.path.join(tmp, "net")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "gen_topology.py"), str(radix), str(conns), net, "--name", "t", "--routing", routing,
           "--chunk-size", str(params["chunk_size"]), "--packet-size", str(params["packet_size"]),
           "--switch-conns", str(params.get("_switch_conns", 1))]
    for k, v in params.items():
        if k not in ("chunk_size", "packet_size", "_switch_conns"):
            cmd += ["--param", f"{k}={v}"]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return os.path.join(net, "t.conf")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["ref", "gpu", "mgpu", "sim"])
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--verbose", action="store_true", help="print every case before it runs")
    args = ap.parse_args()
    rng = random.Random(args.seed)
    oracle_util.load()
    if args.mode in ("gpu", "mgpu", "sim"):
        import codes_b200
    if args.mode == "sim":
        from tests import cpusim_util
    dist, rank = None, 0
    if args.mode == "mgpu":  # every rank generates the same cases from the same seed; rank 0 checks
        import torch
        import torch.distributed as dist
        rank = int(os.environ["RANK"])
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("cpu:gloo,cuda:nccl")
    bad = skipped = 0
    for i in range(args.cases):
        radix, conns, routing, params, opts = make_case(rng)
        tag = f"case {i}: shape {radix}/{conns} {routing} {params} {opts}"
        if args.verbose:
            print(tag, flush=True)
        with tempfile.TemporaryDirectory() as tmp:
            conf = write_conf(tmp, radix, conns, routing, params)
            try:
                res = oracle_util.run(conf, lp_io_dir=os.path.join(tmp, "o"), **opts)
                o = oracle_util.read_dir(os.path.join(tmp, "o"))
                oerr = None
            except oracle_util.OracleError as ex:
                oerr = str(ex)
            if args.mode == "mgpu":
                groups = int(open(conf).read().split('num_groups="')[1].split('"')[0])
                if groups < dist.get_world_size():
                    skipped += 1
                    continue
                flag = [oerr]
                dist.broadcast_object_list(flag, src=0)  # the oracle's verdict of rank 0 decides for everybody
                if flag[0]:
                    skipped += 1
                    continue
                s = codes_b200.run_sharded(conf, dist, lp_io_dir=os.path.join(tmp, "g") if rank == 0 else None, **opts)
                if rank == 0:
                    g = oracle_util.read_dir(os.path.join(tmp, "g"))
                    diffs = [k for k in o if g.get(k) != o[k]]
                    if s.events != res.events or diffs:
                        print("MISMATCH", tag, "events", s.events, res.events, diffs, flush=True)
                        bad += 1
                continue
            if args.mode == "ref":
                try:
                    ev, _, _ = oracle_util.run_ref(conf, os.path.join(tmp, "r"), **opts)
                    r = oracle_util.read_dir(os.path.join(tmp, "r"))
                    rerr = None
                except Exception as ex:  # the reference aborted (tw_error / assert)
                    rerr = str(ex)[:200]
                if oerr or rerr:
                    if oerr and rerr:
                        skipped += 1
                        continue
                    print("MISMATCH (one side rejected)", tag, "| oracle:", oerr, "| ref:", rerr, flush=True)
                    bad += 1
                    continue
                diffs = [k for k in oracle_util.REF_FILES if r.get(k) != o.get(k)]
                if ev != res.events or diffs:
                    print("MISMATCH", tag, "events", ev, res.events, diffs, flush=True)
                    bad += 1
            else:
                try:
                    if args.mode == "sim":
                        world = rng.choice([1, 1, 2, 3, 4])
                        tag += f" [sim seed {i} world {world}]"
                        s = cpusim_util.run(conf, lp_io_dir=os.path.join(tmp, "g"), seed=i, world=world, warps=rng.choice([None, None, 3]), **opts)
                    else:
                        s = codes_b200.run_synthetic(conf, lp_io_dir=os.path.join(tmp, "g"), **opts)
                    g = oracle_util.read_dir(os.path.join(tmp, "g"))
                    gerr = None
                except codes_b200.EngineError as ex:
                    gerr = str(ex)
                if oerr or gerr:
                    if oerr and gerr:
                        skipped += 1
                        continue
                    print("MISMATCH (one side rejected)", tag, "| oracle:", oerr, "| gpu:", gerr, flush=True)
                    bad += 1
                    continue
                diffs = [k for k in o if g.get(k) != o[k]]
                if s.events != res.events or diffs:
                    print("MISMATCH", tag, "events", s.events, res.events, diffs, flush=True)
                    bad += 1
    if dist is not None:
        dist.barrier()
        world = dist.get_world_size()
        dist.destroy_process_group()
        if rank != 0:
            sys.exit(0)
        print(f"mgpu ({world} GPUs): {args.cases} cases, {skipped} skipped (fewer groups than GPUs / rejected), {bad} mismatches")
        sys.exit(1 if bad else 0)
    print(f"{args.mode}: {args.cases} cases, {skipped} rejected by both sides, {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
