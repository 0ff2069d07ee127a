"""The asynchronous scheduler of the CUDA engine, checked on the CPU.

tests/cpusim compiles codes_b200/csrc/dfd_core.h -- the channel-word bounds, the group activation and the LP handlers, the
very code the kernel runs -- as plain C++ and drives the activations from a seeded random sequential scheduler.  These
tests need no GPU and prove three things for any interleaving the scheduler throws at the code:
  * results are byte-identical to the oracle (lp-io files, event counts),
  * no LP ever receives an event older than one it has executed (DFD_ERR_LATE_EVENT would fail the run),
  * the bounds cannot deadlock (a sweep that changes nothing while events are pending is DFD_ERR_STALL).
They also run the group-sharded code path (peer addressing, per-rank termination counters) with several emulated
ranks in one address space.  Test infrastructure only: the product never loads the harness library."""
import pytest

from tests import cpusim_util, oracle_util
from tests.test_gpu_parity import CASES, VARIANTS, variant_conf
from tests.test_oracle_cpu import CONF72, CONF72SNAP, CONF1056, CONF8K


def check(tmp_path, conf, seed, world, warps=None, **opts):
    gdir, odir = str(tmp_path / "sim"), str(tmp_path / "oracle")
    s = cpusim_util.run(conf, lp_io_dir=gdir, seed=seed, world=world, warps=warps, **opts)
    o = oracle_util.run(conf, lp_io_dir=odir, **opts)
    assert s.events == o.events
    for i, name in enumerate(oracle_util.EV_NAMES):
        assert s.events_by_type[i] == o.ev_by_type[i], name
    assert s.rng_draws == o.rng_draws
    g, r = oracle_util.read_dir(gdir), oracle_util.read_dir(odir)
    assert sorted(g) == sorted(r)
    for name in r:
        assert g[name] == r[name], f"{name} differs"


SMALL = [c for c in CASES if c[1] == CONF72]


@pytest.mark.parametrize("name,conf,opts", SMALL, ids=[c[0] for c in SMALL])
def test_scheduler_matches_oracle(tmp_path, name, conf, opts):
    check(tmp_path, conf, seed=1, world=1, **opts)


@pytest.mark.parametrize("name,base,params,opts", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_scheduler_matches_oracle_config_variants(tmp_path, name, base, params, opts):
    conf = variant_conf(tmp_path, base, name, **params)
    check(tmp_path, conf, seed=2, world=1, **opts)


@pytest.mark.parametrize("seed", [3, 4, 5, 6])
def test_any_interleaving_gives_the_same_result(tmp_path, seed):
    check(tmp_path, CONF72, seed=seed, world=1, num_messages=10, traffic=3)


@pytest.mark.parametrize("world,warps", [(2, None), (3, 2), (8, None)])
def test_sharded_code_path_emulated_ranks(tmp_path, world, warps):
    check(tmp_path, CONF72, seed=7, world=world, warps=warps, num_messages=12, traffic=3)


@pytest.mark.parametrize("world,warps", [(1, 5), (2, None), (4, 3)])
def test_router_snapshots_sharded_and_pooled(tmp_path, world, warps):
    """R_SNAPSHOT events (PARAMS router_buffer_snapshots): every rank's records end up in rank 0's dragonfly-snapshots.csv"""
    check(tmp_path, CONF72SNAP, seed=15 + world, world=world, warps=warps, num_messages=20, traffic=3)


def test_several_groups_per_warp(tmp_path):
    check(tmp_path, CONF72, seed=8, world=1, warps=5, num_messages=10)


def test_1056_uniform_and_sharded(tmp_path):
    check(tmp_path, CONF1056, seed=9, world=4, num_messages=4)


def test_8320_prog_adaptive(tmp_path):
    check(tmp_path, CONF8K, seed=10, world=1, num_messages=1, traffic=3)


def test_long_burst_of_messages(tmp_path):
    """inter-arrival below nic_seq_delay piles up second-stage NEW_MSG events (more than the old fixed
    self-event list of 24 could hold); they now live in their own FIFO sized by num_messages."""
    check(tmp_path, CONF72, seed=11, world=1, num_messages=60, arrival_time=1.0)
    check(tmp_path, CONF72, seed=12, world=1, num_messages=40, payload_sz=8, load_per_svr=0.5)


def test_run_ends_before_the_first_kickoff(tmp_path):
    """Found by the round-2 random campaign on the GPU (3 of 300 cases): --end below the inter-arrival time means the first
    KICKOFF of every server is never sent; the engine used to execute it."""
    check(tmp_path, CONF72, seed=13, world=1, num_messages=2, arrival_time=5000.0, end_time=4000.0)
    check(tmp_path, CONF72, seed=14, world=2, num_messages=2, arrival_time=5000.0, end_time=4000.0)
