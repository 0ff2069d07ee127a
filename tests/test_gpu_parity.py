"""GPU parity tests: the CUDA engine, called through the C-ABI, against the CPU oracle on the same config and
seed.  Bit-exact bar: every lp-io file (integer columns and %lf columns) and the hex-float raw statistics
must be byte-identical; `Net Events Processed` must be equal."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

import codes_b200
from tests import oracle_util
from tests.test_oracle_cpu import CONF72, CONF1056, CONF8K, CONF72SNAP, GOLDEN, GOLDEN_CASES, SNAPSHOTS, parse_cn_stats, parse_link_stats, quoted

pytestmark = pytest.mark.gpu


def run_both(tmp_path, conf, **opts):
    gdir, odir = str(tmp_path / "gpu"), str(tmp_path / "oracle")
    s = codes_b200.run_synthetic(conf, lp_io_dir=gdir, **opts)
    o = oracle_util.run(conf, lp_io_dir=odir, **opts)
    return s, o, oracle_util.read_dir(gdir), oracle_util.read_dir(odir)


def assert_identical(s, o, g, r):
    assert s.events == o.events
    for i, name in enumerate(oracle_util.EV_NAMES):
        assert s.events_by_type[i] == o.ev_by_type[i], name
    assert (s.finished_packets, s.finished_chunks, s.finished_msgs, s.total_hops) == \
           (o.finished_packets, o.finished_chunks, o.finished_msgs, o.total_hops)
    assert (s.minimal_count, s.nonmin_count) == (o.minimal_count, o.nonmin_count)
    assert s.rng_draws == o.rng_draws
    assert s.total_time == o.total_time and s.max_latency == o.max_latency  # exact doubles
    assert s.svr_sum_latency == o.svr_sum_latency and s.max_end_ts == o.max_end_ts
    assert sorted(g) == sorted(r)
    for name in r:
        assert g[name] == r[name], f"{name} differs"


CASES = [
    ("72-m1", CONF72, dict(num_messages=1)),                       # the reference's own test case
    ("72-m20", CONF72, dict(num_messages=20)),                      # driver default
    ("72-nearest-group", CONF72, dict(num_messages=10, traffic=3)),  # congested: overflow queues, stalls, busy time
    ("72-nearest-neighbor", CONF72, dict(num_messages=10, traffic=4)),
    ("72-rand-perm", CONF72, dict(num_messages=10, traffic=2)),      # includes self-sends (noop path)
    ("72-random-other-group", CONF72, dict(num_messages=5, traffic=5)),
    ("72-high-load", CONF72, dict(num_messages=30, arrival_time=200.0)),  # injection throttling (issueIdle)
    ("72-load-per-svr", CONF72, dict(num_messages=10, load_per_svr=0.5)),
    ("72-small-payload", CONF72, dict(num_messages=10, payload_sz=8)),
    ("72-full-chunk", CONF72, dict(num_messages=10, payload_sz=4096)),
    ("72-warmup", CONF72, dict(num_messages=10, warm_up_time=3000.0)),
    ("72-end-time", CONF72, dict(num_messages=20, end_time=12000.0)),  # events past --end are dropped
    ("72-empty", CONF72, dict(num_messages=0)),                         # nothing to send: one terminating KICKOFF per server
    ("72-one-byte", CONF72, dict(num_messages=3, payload_sz=1)),
    ("72-4097-bytes", CONF72, dict(num_messages=2, payload_sz=4097)),   # a full packet + a packet of one byte
    ("72-burst", CONF72, dict(num_messages=1, arrival_time=1.0)),
    ("72-long-burst", CONF72, dict(num_messages=40, arrival_time=1.0)),  # inter-arrival below nic_seq_delay: queued NEW_MSG events pile up
    ("72-long-burst-tiny", CONF72, dict(num_messages=40, payload_sz=8, load_per_svr=0.5)),
    ("72-early-end", CONF72, dict(num_messages=3, end_time=1500.0)),
    ("72-end-before-first-kickoff", CONF72, dict(num_messages=2, arrival_time=5000.0, end_time=4000.0)),  # nothing ever runs    # the run ends before the first packet arrives
    ("1056-uniform", CONF1056, dict(num_messages=10)),
    ("1056-nearest-group", CONF1056, dict(num_messages=5, traffic=3)),
    ("8320-par-uniform", CONF8K, dict(num_messages=3)),
    ("8320-par-nearest-group", CONF8K, dict(num_messages=3, traffic=3)),  # PAR takes non-minimal paths
]


@pytest.mark.parametrize("name,conf,opts", CASES, ids=[c[0] for c in CASES])
def test_engine_matches_oracle(tmp_path, name, conf, opts):
    s, o, g, r = run_both(tmp_path, conf, **opts)
    assert_identical(s, o, g, r)


def variant_conf(tmp_path, base, name, **params):
    with open(base) as f:
        text = f.read()
    d = os.path.dirname(base)
    stem = os.path.basename(base)[:-5]
    text = text.replace(f'"{stem}-intra"', f'"{d}/{stem}-intra"').replace(f'"{stem}-inter"', f'"{d}/{stem}-inter"')
    for k, v in params.items():
        if f"{k}=" in text:
            import re
            text = re.sub(r'^\s*' + re.escape(k) + r'\s*=.*$', f'   {k}={quoted(v)};', text, flags=re.M)
        else:
            text = text.replace("PARAMS\n{", "PARAMS\n{\n   " + f'{k}={quoted(v)};')
    p = tmp_path / (name + ".conf")
    p.write_text(text)
    return str(p)


VARIANTS = [
    ("72-nonminimal", CONF72, dict(routing="nonminimal"), dict(num_messages=10)),
    ("72-par", CONF72, dict(routing="prog-adaptive", adaptive_threshold="0"), dict(num_messages=20, traffic=3)),
    ("72-par-alpha", CONF72, dict(routing="prog-adaptive", route_scoring_metric="alpha"), dict(num_messages=20, arrival_time=300.0)),
    ("72-credit-delay-100", CONF72, dict(credit_delay="100"), dict(num_messages=20)),  # wider lookahead window
    ("72-tiny-vcs", CONF72, dict(local_vc_size="4096", global_vc_size="4096", cn_vc_size="4096"), dict(num_messages=10)),
    ("72-asym-bw", CONF72, dict(local_bandwidth="5.25", global_bandwidth="4.7", cn_bandwidth="8.0", router_delay="50"), dict(num_messages=10)),
    ("72-multichunk", CONF72, dict(chunk_size="1024"), dict(num_messages=10)),                       # packets of 2 chunks
    ("72-multipacket", CONF72, dict(packet_size="1024", chunk_size="1024"), dict(num_messages=10)),  # messages of 2 packets
    ("72-multi-both", CONF72, dict(packet_size="2048", chunk_size="512"), dict(num_messages=6, payload_sz=5000)),
    # RAND_PERM self-sends that do not take the no-op shortcut: copied inside the node by the base LP (model-net-lp.c:672-704)
    ("72-self-copy-bypass", CONF72, dict(codes_noop_bypass="1"), dict(num_messages=12, traffic=2)),
    ("72-self-copy-eager-limit", CONF72, dict(node_eager_limit="100"), dict(num_messages=12, traffic=2, payload_sz=4096, arrival_time=300.0)),
    ("72-multi-par", CONF72, dict(packet_size="2048", chunk_size="512", routing="prog-adaptive"), dict(num_messages=8, payload_sz=6000, traffic=3)),
    ("72-big-payload", CONF72, dict(), dict(num_messages=5, payload_sz=20000)),                        # 5 packets per message
    ("1056-nonminimal", CONF1056, dict(routing="nonminimal"), dict(num_messages=4)),
    ("1056-par", CONF1056, dict(routing="prog-adaptive", adaptive_threshold="16384"), dict(num_messages=6, traffic=3)),
    # router VC-occupancy snapshots (PARAMS router_buffer_snapshots -> dragonfly-snapshots.csv; dally:4363-4437, 4687-4709):
    # unsorted times, one past the last packet, one past --end, equal times, long after the network has drained
    ("72-snapshots", CONF72, dict(router_buffer_snapshots=SNAPSHOTS), dict(num_messages=20)),
    ("72-snapshots-par-nearest-group", CONF72, dict(router_buffer_snapshots=SNAPSHOTS, routing="prog-adaptive", adaptive_threshold="0"),
     dict(num_messages=20, traffic=3)),
    ("72-snapshots-end", CONF72, dict(router_buffer_snapshots=SNAPSHOTS), dict(num_messages=20, end_time=10000.0)),
    ("72-snapshots-far", CONF72, dict(router_buffer_snapshots='( "250000" )'), dict(num_messages=3)),
    ("72-snapshots-equal-times", CONF72, dict(router_buffer_snapshots='( "5000","5000" )'), dict(num_messages=10, traffic=3)),
    ("1056-snapshots", CONF1056, dict(router_buffer_snapshots='( "4000","9000" )'), dict(num_messages=5, traffic=3)),
]


@pytest.mark.parametrize("name,base,params,opts", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_engine_matches_oracle_config_variants(tmp_path, name, base, params, opts):
    conf = variant_conf(tmp_path, base, name, **params)
    s, o, g, r = run_both(tmp_path, conf, **opts)
    assert_identical(s, o, g, r)
    if name in ("72-par", "1056-par"):
        assert s.nonmin_count > 0


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_engine_matches_committed_golden(tmp_path, name):
    case = dict(GOLDEN_CASES[name])
    conf = case.pop("conf")
    gdir = str(tmp_path / "gpu")
    s = codes_b200.run_synthetic(conf, lp_io_dir=gdir, **case)
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        gold = json.load(f)
    assert s.events == gold["events"]
    out = oracle_util.read_dir(gdir)
    for fname, digest in gold["sha256"].items():
        assert hashlib.sha256(out[fname]).hexdigest() == digest, fname


def test_determinism_and_state_reuse(tmp_path):
    """Two runs give identical output, and a reset engine (tables kept on the device) repeats it exactly."""
    a, b, c = str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "c")
    s1 = codes_b200.run_synthetic(CONF72, lp_io_dir=a, num_messages=12)
    with codes_b200.Engine() as e:
        e.setup(CONF72, codes_b200.make_opts(num_messages=12))
        e.upload()
        e.run()
        e.finish()
        e.lp_io_flush(b)
        e.reset()
        e.run()
        e.finish()
        e.lp_io_flush(c)
        assert e.summary().events == s1.events
        assert "Net Events Processed" in e.model_net_report_stats()
    assert oracle_util.read_dir(a) == oracle_util.read_dir(b) == oracle_util.read_dir(c)


def test_full_size_properties_1056(tmp_path):
    """BASELINE config[1] at the bench size: size-independent properties (conservation, credit invariants)."""
    gdir = str(tmp_path / "gpu")
    msgs = 100
    s = codes_b200.run_synthetic(CONF1056, lp_io_dir=gdir, num_messages=msgs)
    n = 1056
    assert s.packet_gen == s.packet_fin == n * msgs == s.finished_msgs == s.svr_msgs_recvd
    assert s.minimal_count == s.finished_chunks and s.nonmin_count == 0
    # committed events: 11 + 3*hops per message + terminating KICKOFF per server + one T_SEND per stalled injection
    out = oracle_util.read_dir(gdir)
    links = parse_link_stats(out["dragonfly-link-stats"])
    stalled_inj = sum(l["stalled"] for l in links if l["st"] == "T")
    assert s.events == 11 * n * msgs + 3 * s.total_hops + n + stalled_inj
    assert sum(l["traffic"] for l in links if l["st"] == "T") == 4096 * s.finished_chunks
    assert sum(l["traffic"] for l in links if l["st"] == "R" and l["dt"] == "T") == 2048 * s.finished_chunks
    cn = parse_cn_stats(out["dragonfly-cn-stats"])
    assert sum(r["recvd"] for r in cn) == 2048 * n * msgs
    assert all(r["avg_hops"] <= 4.0 for r in cn)


NARROW_CASES = [
    ("72-m20", CONF72, dict(num_messages=20)),
    ("72-nearest-group", CONF72, dict(num_messages=10, traffic=3)),
    ("1056-uniform", CONF1056, dict(num_messages=10)),
    ("8320-par-nearest-group", CONF8K, dict(num_messages=3, traffic=3)),
]


@pytest.mark.parametrize("name,conf,opts", NARROW_CASES, ids=[c[0] for c in NARROW_CASES])
def test_large_partition_kernel_matches_oracle(tmp_path, monkeypatch, name, conf, opts):
    """The one-warp-per-block event loop exists in two register budgets; partitions of more than 1,184 groups get the
    128-register instantiation (16 resident warps per SM).  Force it on the small configs the oracle can follow."""
    monkeypatch.setenv("DFD_WIDE_REGS", "0")
    s, o, g, r = run_both(tmp_path, conf, **opts)
    assert_identical(s, o, g, r)


POOL_CASES = NARROW_CASES + [
    ("72-snapshots-nearest-group", CONF72SNAP, dict(num_messages=20, traffic=3)),  # R_SNAPSHOT events in the pick (horizon in its own pass)
    ("72-multichunk", CONF72, dict(num_messages=5, payload_sz=10000)),
    ("8320-par-uniform", CONF8K, dict(num_messages=3)),
]


@pytest.mark.parametrize("name,conf,opts", POOL_CASES, ids=[c[0] for c in POOL_CASES])
def test_pooled_kernel_matches_oracle(tmp_path, monkeypatch, name, conf, opts):
    """Partitions with more groups than resident warps run dfd_run_kernel_pool (16-warp blocks, groups dealt out by ticket,
    the router horizon in a pass of its own).  Force it on the small configs the oracle can follow: the warps of a block then
    compete for few groups, which exercises the per-group locks far harder than a large network does."""
    monkeypatch.setenv("DFD_POOL", "1")
    s, o, g, r = run_both(tmp_path, conf, **opts)
    assert s.block_threads == 512
    assert_identical(s, o, g, r)


LARGE_GOLDEN = [
    # golden, radix, links between groups, terminals, messages per terminal, routing, traffic
    ("131584_uniform_m3", 64, 2, 131584, 3, "minimal", 1),     # BASELINE configs[3] scale: a=32, p=16, h=16, g=257, 8,224 routers
    ("131584_par_nearest_group_m3", 64, 2, 131584, 3, "prog-adaptive", 3),  # configs[3]: PAR, worst-case group permutation
    ("1050624_uniform_m1", 128, 4, 1050624, 1, "minimal", 1),  # BASELINE configs[4] scale: a=64, p=32, h=32, g=513, 32,832 routers
]
if os.environ.get("DFD_LONG_TESTS"):  # the bench network in a long run (saturated: 154.6 M events, 1.2 s of kernel, 6 GB of host text to hash)
    LARGE_GOLDEN.append(("131584_par_uniform_m50", 64, 2, 131584, 50, "prog-adaptive", 1))
    LARGE_GOLDEN.append(("131584_par_nearest_group_m50", 64, 2, 131584, 50, "prog-adaptive", 3))


@pytest.mark.parametrize("gold_name,radix,conns,n,msgs,routing,traffic", LARGE_GOLDEN, ids=[g[0] for g in LARGE_GOLDEN])
def test_large_network_matches_golden(tmp_path, gold_name, radix, conns, n, msgs, routing, traffic):
    """~100K and ~1M terminals on the large-partition kernel: every lp-io file and the hex-float statistics must hash
    to the committed digests (tests/golden/<name>.json, generated by the restatement oracle with
    `python -m tests.make_golden --large`), plus the size-independent conservation / credit invariants."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, "tools", "gen_topology.py"), str(radix), str(conns), str(tmp_path / "net"), "--name", "big",
                    "--routing", routing], check=True, stdout=subprocess.DEVNULL)
    gdir = str(tmp_path / "gpu")
    s = codes_b200.run_synthetic(str(tmp_path / "net" / "big.conf"), lp_io_dir=gdir, num_messages=msgs, traffic=traffic)
    assert s.block_threads == 512  # the pooled large-partition kernel (one 16-warp block per SM)
    out = oracle_util.read_dir(gdir)
    with open(os.path.join(GOLDEN, gold_name + ".json")) as f:
        gold = json.load(f)
    assert s.events == gold["events"] and s.total_hops == gold["total_hops"]
    for fname, digest in gold["sha256"].items():
        if fname != "summary.txt":
            assert hashlib.sha256(out[fname]).hexdigest() == digest, fname
    assert s.packet_gen == s.packet_fin == n * msgs == s.finished_msgs == s.svr_msgs_recvd
    if routing == "minimal":
        assert s.minimal_count == s.finished_chunks and s.nonmin_count == 0
    else:
        assert s.minimal_count + s.nonmin_count == s.finished_chunks and s.nonmin_count > 0
    links = parse_link_stats(out["dragonfly-link-stats"])
    stalled_inj = sum(l["stalled"] for l in links if l["st"] == "T")
    assert s.events == 11 * n * msgs + 3 * s.total_hops + n + stalled_inj
    assert sum(l["traffic"] for l in links if l["st"] == "T") == 4096 * s.finished_chunks
    assert sum(l["traffic"] for l in links if l["st"] == "R" and l["dt"] == "T") == 2048 * s.finished_chunks
    cn = parse_cn_stats(out["dragonfly-cn-stats"])
    assert sum(r["recvd"] for r in cn) == 2048 * n * msgs
    max_hops = 4.0 if routing == "minimal" else 7.0
    assert all(r["avg_hops"] <= max_hops for r in cn if r["avg_hops"] == r["avg_hops"])  # NaN: a terminal that received nothing


from tests.test_oracle_cpu import SWITCH_CONN_CASES, TOPO_CASES, TOPO_K_CASES, make_topology


@pytest.mark.parametrize("name,radix,conns,routing,opts", TOPO_CASES, ids=[c[0] for c in TOPO_CASES])
def test_engine_matches_oracle_other_shapes(tmp_path, name, radix, conns, routing, opts):
    """Dragonflies with 2 and 4 parallel global links between groups and router counts that are not powers of two
    (the candidate lists of the routing functions and the far-end port pairing look different there)."""
    conf = make_topology(tmp_path, radix, conns, routing)
    s, o, g, r = run_both(tmp_path, conf, **opts)
    assert_identical(s, o, g, r)


def test_ensemble_matches_solo_runs(tmp_path):
    """Several simulations at once on one GPU (parameter sweep): every member's output is byte-identical to the same
    simulation run alone -- concurrent engines share nothing but the device."""
    variants = [dict(num_messages=10), dict(num_messages=6, traffic=3), dict(num_messages=8, arrival_time=500.0),
                dict(num_messages=5, payload_sz=10000), dict(num_messages=10, traffic=2), dict(num_messages=3, load_per_svr=0.7),
                dict(num_messages=4, traffic=4), dict(num_messages=7, warm_up_time=2000.0), dict(num_messages=2, traffic=5),
                dict(num_messages=9, arrival_time=2000.0)]
    dirs = [str(tmp_path / f"e{i}") for i in range(len(variants))]
    sums = codes_b200.run_ensemble(CONF1056, variants, lp_io_dirs=dirs)
    for i, v in enumerate(variants):
        solo = str(tmp_path / f"s{i}")
        s = codes_b200.run_synthetic(CONF1056, lp_io_dir=solo, **v)
        assert sums[i].events == s.events and sums[i].total_hops == s.total_hops
        assert oracle_util.read_dir(dirs[i]) == oracle_util.read_dir(solo), i
        assert sums[i].grid_blocks * len(variants) <= 16 * sums[i].num_sms  # all members are co-resident (16 one-warp blocks per SM)


def test_terminal_runs_ahead_of_its_link(tmp_path):
    """Found by tools/fuzz_parity.py: small chunks and a deep terminal VC let terminal_available_time run 32 injection
    times ahead of the clock (packet_generate starts T_SEND at +0 whenever the send loop is idle, dally:5749-5760),
    so chunk arrivals lie 3.8 us in the future -- the router timing wheel has to span that."""
    import subprocess
    import sys
    net = str(tmp_path / "net")
    cmd = [sys.executable, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "gen_topology.py"), "11", "1", net,
           "--name", "t", "--chunk-size", "256", "--packet-size", "256"]
    for kv in ("local_vc_size=2048", "global_vc_size=2048", "cn_vc_size=8192", "local_bandwidth=12.5", "global_bandwidth=4.7", "cn_bandwidth=2.0"):
        cmd += ["--param", kv]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    s, o, g, r = run_both(tmp_path, os.path.join(net, "t.conf"), num_messages=7, payload_sz=9000, arrival_time=1000.0)
    assert_identical(s, o, g, r)


def test_random_configurations_match_oracle():
    """A slice of the seeded campaign of tools/fuzz_parity.py on the GPU (engine vs oracle, every file byte for byte)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "gpu", "--cases", "40", "--seed", "11"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "40 cases" in r.stdout and "0 mismatches" in r.stdout


@pytest.mark.parametrize("name,radix,conns,k,opts", TOPO_K_CASES, ids=[c[0] for c in TOPO_K_CASES])
def test_engine_matches_oracle_k_picks(tmp_path, name, radix, conns, k, opts):
    """global_k_picks other than 2 (dfdally_poll_k_connections dally:1796-1849, sampling without replacement)."""
    conf = make_topology(tmp_path, radix, conns, "prog-adaptive", global_k_picks=k)
    s, o, g, r = run_both(tmp_path, conf, **opts)
    assert_identical(s, o, g, r)


def test_k_picks_larger_than_the_link_count_is_an_error_not_a_hang(tmp_path):
    """The reference spins forever when global_k_picks exceeds the links a router has to one group (its guard is
    commented out, dally:1822-1823).  The engine reports a device error instead of hanging the GPU."""
    conf = make_topology(tmp_path, 11, 9, "prog-adaptive", global_k_picks=3)
    with pytest.raises(codes_b200.EngineError, match="dead end"):
        codes_b200.run_synthetic(conf, num_messages=8, arrival_time=300.0)


@pytest.mark.parametrize("name,radix,conns,routing,opts", SWITCH_CONN_CASES, ids=[c[0] for c in SWITCH_CONN_CASES])
def test_engine_matches_oracle_parallel_local_links(tmp_path, name, radix, conns, routing, opts):
    """num_parallel_switch_conns=2: every local hop has two candidate ports (one more draw per hop)."""
    conf = make_topology(tmp_path, radix, conns, routing, switch_conns=2)
    s, o, g, r = run_both(tmp_path, conf, **opts)
    assert_identical(s, o, g, r)


@pytest.mark.parametrize("world,conf,opts", [(2, CONF72, dict(num_messages=20)), (3, CONF72, dict(num_messages=10, traffic=3)),
                                             (4, CONF1056, dict(num_messages=10)), (2, CONF1056, dict(num_messages=5, payload_sz=10000)),
                                             (8, CONF8K, dict(num_messages=3, traffic=3)),
                                             (3, CONF72SNAP, dict(num_messages=20, traffic=3))])  # snapshot records gathered from every rank
def test_group_sharded_ranks_on_one_gpu_match_oracle(tmp_path, world, conf, opts):
    """The group-sharded engine with all ranks on this one GPU: cross-rank lanes, channel words, ACK statistics,
    termination detection over several control blocks and rank 0's gather of the partitions must give the sequential
    result bit for bit (the reference's analogue: --sync=1 vs --sync=3, tests/CMakeLists.txt:66-68).  The multi-process
    CUDA-IPC variant of the same path is tests/mgpu_check.py (torchrun, needs >= 2 GPUs)."""
    gdir, odir = str(tmp_path / "gpu"), str(tmp_path / "oracle")
    s = codes_b200.run_sharded_local(conf, world, lp_io_dir=gdir, **opts)
    o = oracle_util.run(conf, lp_io_dir=odir, **opts)
    assert s.events == o.events
    g, r = oracle_util.read_dir(gdir), oracle_util.read_dir(odir)
    for name in r:
        assert g[name] == r[name], f"{name} differs"


def test_group_sharded_ranks_without_publisher_queues(tmp_path, monkeypatch):
    """Sharded runs hand the channel words for other ranks to publisher blocks (one system-scope fence per batch);
    DFD_NPUB=0 is the direct path: every activation ends in its own system-scope fence (what the pooled kernel always does)."""
    monkeypatch.setenv("DFD_NPUB", "0")
    gdir, odir = str(tmp_path / "gpu"), str(tmp_path / "oracle")
    opts = dict(num_messages=10, traffic=3)
    s = codes_b200.run_sharded_local(CONF1056, 4, lp_io_dir=gdir, **opts)
    o = oracle_util.run(CONF1056, lp_io_dir=odir, **opts)
    assert s.events == o.events
    g, r = oracle_util.read_dir(gdir), oracle_util.read_dir(odir)
    for name in r:
        assert g[name] == r[name], f"{name} differs"


def test_multi_process_sharding_when_two_gpus_are_present():
    """torchrun with one process per GPU (CUDA-IPC exchange regions over NVLink): tests/mgpu_check.py."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(root, "tests", "mgpu_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MGPU PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
