"""CPU tests of the oracle: known-answer checks derived from the reference's closed forms and
regression against the committed golden vectors.  No GPU needed."""
import math
import os
import struct

import pytest

from tests import oracle_util

ROOT = oracle_util.ROOT
CONF72 = os.path.join(ROOT, "configs", "dfdally-72.conf")
CONF1056 = os.path.join(ROOT, "configs", "dfdally-1056.conf")
CONF8K = os.path.join(ROOT, "configs", "dfdally-8320-par.conf")
CONF72SNAP = os.path.join(ROOT, "configs", "dfdally-72-snapshots.conf")
GOLDEN = os.path.join(ROOT, "tests", "golden")

M = [2147483647, 2147483543, 2147483423, 2147483323]
A = [45991, 207707, 138556, 49689]
SEED0 = [11111111, 22222222, 33333333, 44444444]


def py_bytes_to_ns(nbytes, bw):
    t = float(nbytes) / (1024.0 * 1024.0 * 1024.0)
    t = t / bw
    return t * 1000.0 * 1000.0 * 1000.0


def test_bytes_to_ns_known_answers(oracle):
    # dragonfly-dally.cxx:1588-1599; the 72-terminal config numbers
    assert oracle.dfdally_oracle_bytes_to_ns(4096, 2.0) == 1907.3486328125
    assert oracle.dfdally_oracle_bytes_to_ns(2048, 2.0) == 953.67431640625
    assert oracle.dfdally_oracle_bytes_to_ns(8, 2.0) == 3.725290298461914
    for n, bw in [(1, 5.25), (4096, 4.7), (512, 5.25), (123457, 12.5)]:
        assert oracle.dfdally_oracle_bytes_to_ns(n, bw) == py_bytes_to_ns(n, bw)


def test_clcg4_stream_seeding_and_steps():
    # stream k starts at seed0 * (a^(2^(31+41)))^k mod m  (L'Ecuyer-Andres jump-ahead)
    for stream in [0, 1, 2, 3, 7, 180 * 3 + 2, 2134080 * 3 + 2]:
        u, st = oracle_util.rng_stream(stream, 5)
        exp = [SEED0[j] * pow(pow(A[j], 2 ** 72, M[j]), stream, M[j]) % M[j] for j in range(4)]
        assert st == exp
        s = list(exp)
        for i in range(5):
            s = [s[j] * A[j] % M[j] for j in range(4)]
            v = 0.0
            v = v + 4.65661287524579692e-10 * s[0]
            v = v - 4.65661310075985993e-10 * s[1]
            if v < 0:
                v += 1.0
            v = v + 4.65661336096842131e-10 * s[2]
            if v >= 1.0:
                v -= 1.0
            v = v - 4.65661357780891134e-10 * s[3]
            if v < 0:
                v += 1.0
            assert u[i] == v
            assert 0.0 <= u[i] < 1.0
    # stream 0 is the generator's default seed
    assert oracle_util.rng_stream(0, 1)[1] == SEED0


def parse_cn_stats(blob):
    rows = []
    for line in blob.decode().splitlines():
        if line.startswith("#") or not line.strip():
            continue
        f = line.split()
        rows.append(dict(gid=int(f[0]), tid=int(f[1]), sent=int(f[2]), recvd=int(f[3]), avg_lat=float(f[4]),
                         max_lat=float(f[5]), min_lat=float(f[6]), packets=int(f[7]), avg_hops=float(f[8]),
                         busy=float(f[9])))
    return rows


def parse_link_stats(blob):
    rows = []
    for line in blob.decode().splitlines():
        if line.startswith("#") or not line.strip():
            continue
        f = line.split()
        rows.append(dict(src=int(f[0]), st=f[1], dst=int(f[2]), dt=f[3], kind=f[4], traffic=int(f[5]),
                         busy=float(f[6]), stalled=int(f[7])))
    return rows


def test_single_message_idle_network_latency(tmp_path):
    """One message per server on an idle 72-terminal network: every chunk latency follows the closed form
    inj(tail, cn) + cn_delay + hops * (inj(tail) + router_delay + link_delay)  (dally:5898-5916, 7811-7833)."""
    res = oracle_util.run(CONF72, lp_io_dir=str(tmp_path), num_messages=1)
    n = 72
    assert res.packet_gen == n and res.packet_fin == n and res.finished_msgs == n and res.finished_chunks == n
    # event count: 11 + 3*hops per message + one terminating KICKOFF per server
    assert res.events == 11 * n + 3 * res.total_hops + n
    tail = py_bytes_to_ns(2048, 2.0)
    link = py_bytes_to_ns(4096, 2.0)
    per_hop = tail + 100.0 + link
    rows = parse_cn_stats(oracle_util.read_dir(str(tmp_path))["dragonfly-cn-stats"])
    assert len(rows) == n
    seen = 0
    for r in rows:
        if r["packets"] == 0:
            assert math.isnan(r["avg_lat"]) and r["min_lat"] == 2147483647.0  # dally:6850-6853, :4766
            continue
        seen += r["packets"]
        for lat in (r["max_lat"], r["min_lat"]):
            hops = (lat - tail - link) / per_hop
            # with one message per node a few chunks share a link and wait; the uncontended ones sit on the grid
            assert hops > 0.99
        assert 1.0 <= r["avg_hops"] <= 4.0  # minimal routing: at most local-global-local + ejection
    assert seen == n
    # the minimum over all terminals is an uncontended chunk: exactly on the closed form
    best = min(r["min_lat"] for r in rows if r["packets"])
    hops = round((best - tail - link) / per_hop)
    assert abs(best - (tail + link + hops * per_hop)) < 1e-6


@pytest.mark.parametrize("conf,msgs,traffic", [(CONF72, 20, 1), (CONF72, 5, 3), (CONF72, 5, 2), (CONF1056, 3, 1)])
def test_conservation_and_credit_invariants(tmp_path, conf, msgs, traffic):
    res = oracle_util.run(conf, lp_io_dir=str(tmp_path), num_messages=msgs, traffic=traffic)
    out = oracle_util.read_dir(str(tmp_path))
    cn = parse_cn_stats(out["dragonfly-cn-stats"])
    nterm = len(cn)
    # every generated packet is delivered (dally:3186-3189)
    assert res.packet_gen == res.packet_fin
    assert res.svr_msgs_recvd == nterm * msgs
    if traffic != 2:
        assert res.packet_gen == nterm * msgs
    else:
        # RAND_PERM may draw the sender itself: those messages take model_net_noop_event (model-net.c:252-286)
        assert res.packet_gen <= nterm * msgs and (nterm * msgs - res.packet_gen) % msgs == 0
    assert res.finished_msgs == res.packet_gen
    assert sum(r["sent"] for r in cn) == 2048 * res.packet_gen
    assert sum(r["recvd"] for r in cn) == 2048 * res.packet_gen
    links = parse_link_stats(out["dragonfly-link-stats"])
    # terminals inject one full chunk_size of link traffic per chunk (dally:5975)
    assert sum(l["traffic"] for l in links if l["st"] == "T") == 4096 * res.finished_chunks
    # router->terminal ports carry the delivered bytes (tail chunk size, dally:7926-7954)
    assert sum(l["traffic"] for l in links if l["st"] == "R" and l["dt"] == "T") == 2048 * res.finished_chunks
    assert res.minimal_count + res.nonmin_count == res.finished_chunks
    # no "leftover" warnings: all queues drained (dally:6855-6858, 6988-6997) -> summary has none
    assert b"leftover" not in out["summary.txt"] and b"lefover" not in out["summary.txt"]
    # hops: minimal routing never exceeds 4 router traversals
    for r in cn:
        if r["packets"]:
            assert r["avg_hops"] <= 4.0 + 1e-9


def test_prog_adaptive_runs_and_uses_nonminimal_paths(tmp_path):
    res = oracle_util.run(CONF8K, lp_io_dir=str(tmp_path), num_messages=2, traffic=3)  # worst-case group permutation
    assert res.packet_fin == 8320 * 2
    assert res.nonmin_count > 0  # PAR diverts around the single minimal global link
    assert res.minimal_count + res.nonmin_count == res.finished_chunks


def test_determinism_two_runs_identical(tmp_path):
    # the reference's own test strategy: two runs of the same config are identical (tests/CMakeLists.txt:654-660)
    a, b = tmp_path / "a", tmp_path / "b"
    r1 = oracle_util.run(CONF72, lp_io_dir=str(a), num_messages=7)
    r2 = oracle_util.run(CONF72, lp_io_dir=str(b), num_messages=7)
    assert r1.events == r2.events
    assert oracle_util.read_dir(str(a)) == oracle_util.read_dir(str(b))


GOLDEN_CASES = {
    "72_uniform_m1": dict(conf=CONF72, num_messages=1, traffic=1),
    "72_uniform_m20": dict(conf=CONF72, num_messages=20, traffic=1),
    "72_nearest_group_m10": dict(conf=CONF72, num_messages=10, traffic=3),
    "1056_uniform_m5": dict(conf=CONF1056, num_messages=5, traffic=1),
    "8320_par_uniform_m2": dict(conf=CONF8K, num_messages=2, traffic=1),
    # router VC-occupancy snapshots (PARAMS router_buffer_snapshots -> dragonfly-snapshots.csv), congested traffic
    "72_snapshots_nearest_group_m20": dict(conf=CONF72SNAP, num_messages=20, traffic=3),
}


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_matches_golden(tmp_path, name):
    import hashlib
    import json
    case = dict(GOLDEN_CASES[name])
    conf = case.pop("conf")
    res = oracle_util.run(conf, lp_io_dir=str(tmp_path), **case)
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        gold = json.load(f)
    assert res.events == gold["events"]
    out = oracle_util.read_dir(str(tmp_path))
    for fname, digest in gold["sha256"].items():
        assert hashlib.sha256(out[fname]).hexdigest() == digest, fname
    small = os.path.join(GOLDEN, name)
    if os.path.isdir(small):
        for fname in os.listdir(small):
            with open(os.path.join(small, fname), "rb") as f:
                assert out[fname] == f.read(), fname


def test_config_errors():
    with pytest.raises(oracle_util.OracleError):
        oracle_util.run("/nonexistent.conf")


# ------------------------------------------------------------------------------------------------------------------
# Pinning the restatement: the reference's OWN handler sources (compiled unchanged against the ROSS/MPI shim,
# oracle/ref/) must produce byte-identical lp-io files, the same `Net Events Processed` and the same stdout statistics.
# ------------------------------------------------------------------------------------------------------------------
REF_CASES = [
    ("72-m1", CONF72, dict(num_messages=1)),
    ("72-m20", CONF72, dict(num_messages=20)),
    ("72-nearest-group", CONF72, dict(num_messages=10, traffic=3)),
    ("72-rand-perm", CONF72, dict(num_messages=10, traffic=2)),
    ("72-nearest-neighbor", CONF72, dict(num_messages=10, traffic=4)),
    ("72-random-other-group", CONF72, dict(num_messages=5, traffic=5)),
    ("72-high-load", CONF72, dict(num_messages=30, arrival_time=200.0)),
    ("72-load-per-svr", CONF72, dict(num_messages=10, load_per_svr=0.5)),
    ("72-small-payload", CONF72, dict(num_messages=10, payload_sz=8)),
    ("72-full-chunk", CONF72, dict(num_messages=10, payload_sz=4096)),
    ("72-warmup", CONF72, dict(num_messages=10, warm_up_time=3000.0)),
    ("72-end-time", CONF72, dict(num_messages=20, end_time=12000.0)),
    ("72-empty", CONF72, dict(num_messages=0)),                         # nothing to send: one terminating KICKOFF per server
    ("72-one-byte", CONF72, dict(num_messages=3, payload_sz=1)),
    ("72-4097-bytes", CONF72, dict(num_messages=2, payload_sz=4097)),   # a full packet + a packet of one byte
    ("72-burst", CONF72, dict(num_messages=1, arrival_time=1.0)),
    ("72-long-burst", CONF72, dict(num_messages=40, arrival_time=1.0)),  # inter-arrival below nic_seq_delay: queued NEW_MSG events pile up
    ("72-long-burst-tiny", CONF72, dict(num_messages=40, payload_sz=8, load_per_svr=0.5)),
    ("72-early-end", CONF72, dict(num_messages=3, end_time=1500.0)),
    ("72-end-before-first-kickoff", CONF72, dict(num_messages=2, arrival_time=5000.0, end_time=4000.0)),  # nothing ever runs    # the run ends before the first packet arrives
    ("1056-uniform", CONF1056, dict(num_messages=5)),
    ("8320-par-nearest-group", CONF8K, dict(num_messages=2, traffic=3)),
]


@pytest.fixture(scope="session")
def ref_oracle():
    if not oracle_util.ref_available() and not oracle_util.build_ref():
        pytest.skip("oracle/_ref is not built (needs /root/reference)")
    return oracle_util


@pytest.mark.parametrize("name,conf,opts", REF_CASES, ids=[c[0] for c in REF_CASES])
def test_restatement_matches_reference_sources(tmp_path, ref_oracle, name, conf, opts):
    rdir, odir = str(tmp_path / "ref"), str(tmp_path / "or")
    ev, _, out = oracle_util.run_ref(conf, rdir, **opts)
    res = oracle_util.run(conf, lp_io_dir=odir, **opts)
    assert ev == res.events
    r, o = oracle_util.read_dir(rdir), oracle_util.read_dir(odir)
    for f in oracle_util.REF_FILES:
        assert r.get(f) == o.get(f), f
    assert oracle_util.stdout_stats_block(out) == oracle_util.stdout_stats_block(o["summary.txt"].decode())


def quoted(v):
    """a list value ( "a","b" ) is written as it is, a scalar in quotes"""
    return v if v.startswith("(") else f'"{v}"'


SNAPSHOTS = '( "3000","12000.5","8000","16000" )'


def ref_variant(tmp_path, base, **params):
    import re
    with open(base) as f:
        text = f.read()
    d, stem = os.path.dirname(base), os.path.basename(base)[:-5]
    text = text.replace(f'"{stem}-intra"', f'"{d}/{stem}-intra"').replace(f'"{stem}-inter"', f'"{d}/{stem}-inter"')
    for k, v in params.items():
        if re.search(r'^\s*' + re.escape(k) + r'\s*=', text, flags=re.M):
            text = re.sub(r'^\s*' + re.escape(k) + r'\s*=.*$', f'   {k}={quoted(v)};', text, flags=re.M)
        else:
            text = text.replace("PARAMS\n{", "PARAMS\n{\n   " + f'{k}={quoted(v)};')
    p = tmp_path / "variant.conf"
    p.write_text(text)
    return str(p)


REF_VARIANTS = [
    ("72-nonminimal", CONF72, dict(routing="nonminimal"), dict(num_messages=10)),
    ("72-par", CONF72, dict(routing="prog-adaptive", adaptive_threshold="0"), dict(num_messages=20, traffic=3)),
    ("72-par-alpha", CONF72, dict(routing="prog-adaptive", route_scoring_metric="alpha"), dict(num_messages=20, arrival_time=300.0)),
    ("72-credit-delay-100", CONF72, dict(credit_delay="100"), dict(num_messages=20)),
    ("72-tiny-vcs", CONF72, dict(local_vc_size="4096", global_vc_size="4096", cn_vc_size="4096"), dict(num_messages=10)),
    ("72-asym-bw", CONF72, dict(local_bandwidth="5.25", global_bandwidth="4.7", cn_bandwidth="8.0", router_delay="50"), dict(num_messages=10)),
    ("72-multichunk", CONF72, dict(chunk_size="1024"), dict(num_messages=10)),                 # packets of 2 chunks
    ("72-multipacket", CONF72, dict(packet_size="1024", chunk_size="1024"), dict(num_messages=10)),  # messages of 2 packets
    ("72-multi-both", CONF72, dict(packet_size="2048", chunk_size="512"), dict(num_messages=6, payload_sz=5000)),
    # RAND_PERM self-sends that do not take the no-op shortcut: copied inside the node by the base LP (model-net-lp.c:672-704)
    ("72-self-copy-bypass", CONF72, dict(codes_noop_bypass="1"), dict(num_messages=12, traffic=2)),
    ("72-self-copy-eager-limit", CONF72, dict(node_eager_limit="100"), dict(num_messages=12, traffic=2, payload_sz=4096, arrival_time=300.0)),
    ("1056-par", CONF1056, dict(routing="prog-adaptive", adaptive_threshold="16384"), dict(num_messages=4, traffic=3)),
    # router VC-occupancy snapshots (dally:4363-4437, 4687-4709): unsorted times, one past the last packet, one past --end
    ("72-snapshots", CONF72, dict(router_buffer_snapshots=SNAPSHOTS), dict(num_messages=20)),
    ("72-snapshots-par-nearest-group", CONF72, dict(router_buffer_snapshots=SNAPSHOTS, routing="prog-adaptive", adaptive_threshold="0"),
     dict(num_messages=20, traffic=3)),
    ("72-snapshots-end", CONF72, dict(router_buffer_snapshots=SNAPSHOTS), dict(num_messages=20, end_time=10000.0)),
    ("72-snapshots-far", CONF72, dict(router_buffer_snapshots='( "250000" )'), dict(num_messages=3)),  # long after the network has drained
    ("72-snapshots-equal-times", CONF72, dict(router_buffer_snapshots='( "5000","5000" )'), dict(num_messages=10, traffic=3)),
    ("1056-snapshots", CONF1056, dict(router_buffer_snapshots='( "4000","9000" )'), dict(num_messages=5, traffic=3)),
    ("8320-par-snapshots", CONF8K, dict(router_buffer_snapshots='( "4000","2000","9000.25" )'), dict(num_messages=2, traffic=3)),
]


@pytest.mark.parametrize("name,base,params,opts", REF_VARIANTS, ids=[v[0] for v in REF_VARIANTS])
def test_restatement_matches_reference_sources_config_variants(tmp_path, ref_oracle, name, base, params, opts):
    conf = ref_variant(tmp_path, base, **params)
    rdir, odir = str(tmp_path / "ref"), str(tmp_path / "or")
    ev, _, out = oracle_util.run_ref(conf, rdir, **opts)
    res = oracle_util.run(conf, lp_io_dir=odir, **opts)
    assert ev == res.events
    r, o = oracle_util.read_dir(rdir), oracle_util.read_dir(odir)
    for f in oracle_util.REF_FILES:
        assert r.get(f) == o.get(f), f
    assert oracle_util.stdout_stats_block(out) == oracle_util.stdout_stats_block(o["summary.txt"].decode())


# ---- other dragonfly shapes: parallel global links between groups, odd router counts (tools/gen_topology.py) ----
TOPO_CASES = [
    # name, radix, links between groups, routing, options       -> a, p=h, groups, terminals
    ("r11c1-minimal", 11, 1, "minimal", dict(num_messages=5)),                          # 6, 3, 19, 342
    ("r15c2-minimal", 15, 2, "minimal", dict(num_messages=5)),                          # 8, 4, 17, 544
    ("r15c2-par-nearest-group", 15, 2, "prog-adaptive", dict(num_messages=4, traffic=3)),
    ("r19c2-nonminimal", 19, 2, "nonminimal", dict(num_messages=4)),                    # 10, 5, 26, 1300
    ("r23c4-minimal-nearest-group", 23, 4, "minimal", dict(num_messages=4, traffic=3)),  # 12, 6, 19, 1368
    ("r23c4-par", 23, 4, "prog-adaptive", dict(num_messages=4, arrival_time=300.0)),
    # routers with two links to the same group: the k-connection poll of progressive-adaptive routing draws twice
    ("r11c9-par-two-links-per-group", 11, 9, "prog-adaptive", dict(num_messages=8, arrival_time=300.0)),   # 6, 3, 3 groups, 54
    ("r15c16-par-two-links-per-group", 15, 16, "prog-adaptive", dict(num_messages=6, traffic=3)),           # 8, 4, 3 groups, 96
    ("r15c8-nonminimal", 15, 8, "nonminimal", dict(num_messages=6)),                                       # 8, 4, 5 groups, 160
]
TOPO_K_CASES = [  # global_k_picks other than the default 2
    ("r11c9-par-k1", 11, 9, 1, dict(num_messages=8, arrival_time=300.0)),
    ("r15c8-par-k3", 15, 8, 3, dict(num_messages=6, arrival_time=300.0)),
    ("r11c6-par-k3", 11, 6, 3, dict(num_messages=6, traffic=3)),
]


def make_topology(tmp_path, radix, conns, routing, switch_conns=1, **params):
    import subprocess
    import sys
    d = str(tmp_path / "net")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "gen_topology.py"), str(radix), str(conns), d, "--name", "t", "--routing", routing,
           "--switch-conns", str(switch_conns)]
    for k, v in params.items():
        cmd += ["--param", f"{k}={v}"]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return os.path.join(d, "t.conf")


@pytest.mark.parametrize("name,radix,conns,routing,opts", TOPO_CASES, ids=[c[0] for c in TOPO_CASES])
def test_restatement_matches_reference_sources_other_shapes(tmp_path, ref_oracle, name, radix, conns, routing, opts):
    conf = make_topology(tmp_path, radix, conns, routing)
    rdir, odir = str(tmp_path / "ref"), str(tmp_path / "or")
    ev, _, out = oracle_util.run_ref(conf, rdir, **opts)
    res = oracle_util.run(conf, lp_io_dir=odir, **opts)
    assert ev == res.events
    r, o = oracle_util.read_dir(rdir), oracle_util.read_dir(odir)
    for f in oracle_util.REF_FILES:
        assert r.get(f) == o.get(f), f
    assert oracle_util.stdout_stats_block(out) == oracle_util.stdout_stats_block(o["summary.txt"].decode())


def test_random_configurations_match_reference_sources(ref_oracle):
    """A slice of the seeded campaign of tools/fuzz_parity.py (random shapes, routing, chunk / packet / VC sizes,
    bandwidths, delays, traffic, payloads, loads): restatement vs reference sources, every lp-io file byte for byte.
    The full campaigns (310 cases here, 400 on the GPU) are run with the tool itself."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "ref", "--cases", "20", "--seed", "3"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "20 cases" in r.stdout and "0 mismatches" in r.stdout


@pytest.mark.parametrize("name,radix,conns,k,opts", TOPO_K_CASES, ids=[c[0] for c in TOPO_K_CASES])
def test_restatement_matches_reference_sources_k_picks(tmp_path, ref_oracle, name, radix, conns, k, opts):
    conf = make_topology(tmp_path, radix, conns, "prog-adaptive", global_k_picks=k)
    rdir, odir = str(tmp_path / "ref"), str(tmp_path / "or")
    ev, _, out = oracle_util.run_ref(conf, rdir, **opts)
    res = oracle_util.run(conf, lp_io_dir=odir, **opts)
    assert ev == res.events
    r, o = oracle_util.read_dir(rdir), oracle_util.read_dir(odir)
    for f in oracle_util.REF_FILES:
        assert r.get(f) == o.get(f), f


SWITCH_CONN_CASES = [  # two parallel local links between every pair of routers of a group (num_parallel_switch_conns=2)
    ("r7c1-minimal-2-local-links", 7, 1, "minimal", dict(num_messages=8, arrival_time=300.0)),
    ("r11c1-par-2-local-links", 11, 1, "prog-adaptive", dict(num_messages=8, traffic=3)),
    ("r11c2-nonminimal-2-local-links", 11, 2, "nonminimal", dict(num_messages=6)),
]


@pytest.mark.parametrize("name,radix,conns,routing,opts", SWITCH_CONN_CASES, ids=[c[0] for c in SWITCH_CONN_CASES])
def test_restatement_matches_reference_sources_parallel_local_links(tmp_path, ref_oracle, name, radix, conns, routing, opts):
    conf = make_topology(tmp_path, radix, conns, routing, switch_conns=2)
    rdir, odir = str(tmp_path / "ref"), str(tmp_path / "or")
    ev, _, out = oracle_util.run_ref(conf, rdir, **opts)
    res = oracle_util.run(conf, lp_io_dir=odir, **opts)
    assert ev == res.events
    r, o = oracle_util.read_dir(rdir), oracle_util.read_dir(odir)
    for f in oracle_util.REF_FILES:
        assert r.get(f) == o.get(f), f
