"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol, the
.conf / PARAMS / LPGROUPS handling mirrors the reference's behaviour, and compute calls fail loudly without
a GPU.  No compute is run here."""
import os
import re

import pytest

import codes_b200
from tests import oracle_util

ROOT = oracle_util.ROOT
CONF72 = os.path.join(ROOT, "configs", "dfdally-72.conf")


def declared_symbols():
    with open(os.path.join(ROOT, "include", "dfdally_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(dfd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(engine_lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(engine_lib, n), f"{n} is declared in include/dfdally_b200.h but not exported"
    assert sorted(n for n, _, _ in codes_b200.ABI) == names


def test_bytes_to_ns_matches_reference_formula(engine_lib):
    assert codes_b200.bytes_to_ns(4096, 2.0) == 1907.3486328125
    assert codes_b200.bytes_to_ns(2048, 2.0) == 953.67431640625
    assert codes_b200.bytes_to_ns(8, 2.0) == 3.725290298461914


def test_call_sequence_and_mapping(engine_lib):
    with codes_b200.Engine() as e:
        # out-of-order calls are refused (the reference would crash on its NULL globals)
        with pytest.raises(codes_b200.EngineError):
            e.model_net_register()
        e.configuration_load(CONF72)
        assert e.configuration_get_value("PARAMS", "routing") == "minimal"
        assert e.configuration_get_value("PARAMS", "no_such_key") is None
        with pytest.raises(codes_b200.EngineError):
            e.codes_mapping_setup()
        e.model_net_register()
        assert e.lp_type_lookup("modelnet_dragonfly_dally") and e.lp_type_lookup("nw-lp")
        assert not e.lp_type_lookup("modelnet_torus")
        e.codes_mapping_setup()
        # codes_mapping.c:149-378 -- repetition r holds [nw-lp x2][terminal x2][router]
        assert e.codes_mapping_get_lp_count("nw-lp") == 72
        assert e.codes_mapping_get_lp_count("modelnet_dragonfly_dally") == 72
        assert e.codes_mapping_get_lp_count("modelnet_dragonfly_dally_router") == 36
        assert e.codes_mapping_get_lpid_from_relative(0, "nw-lp") == 0
        assert e.codes_mapping_get_lpid_from_relative(3, "nw-lp") == 6
        assert e.codes_mapping_get_lpid_from_relative(3, "modelnet_dragonfly_dally") == 8
        assert e.codes_mapping_get_lpid_from_relative(1, "modelnet_dragonfly_dally_router") == 9
        for k in range(72):
            for t in ("nw-lp", "modelnet_dragonfly_dally"):
                assert e.codes_mapping_get_lp_relative_id(e.codes_mapping_get_lpid_from_relative(k, t)) == k
        # model_net_configure returns ids in modelnet_order order (codes/model-net.h:74-97)
        assert e.model_net_configure() == [15, 16]
        e.synthetic_configure(codes_b200.make_opts(num_messages=3))


def write_conf(tmp_path, **changes):
    with open(CONF72) as f:
        text = f.read()
    d = os.path.join(ROOT, "configs")
    text = text.replace('"dfdally-72-intra"', f'"{d}/dfdally-72-intra"').replace('"dfdally-72-inter"', f'"{d}/dfdally-72-inter"')
    for k, v in changes.items():
        key = k.replace("__", "-")
        pat = re.compile(r'^\s*' + re.escape(key) + r'\s*=.*$', re.M)
        if v is None:
            text = pat.sub("", text)
        elif pat.search(text):
            text = pat.sub(f'   {key}="{v}";', text)
        else:
            text = text.replace("PARAMS\n{", "PARAMS\n{\n   " + f'{key}="{v}";')
    p = tmp_path / "c.conf"
    p.write_text(text)
    return str(p)


def configure(conf, **opts):
    with codes_b200.Engine() as e:
        e.setup(conf, codes_b200.make_opts(**opts))


@pytest.mark.parametrize("changes,msg", [
    (dict(routing="adaptive"), "Standard Adaptive Routing not implemented"),  # dally:7190-7193
    (dict(num_groups=None), "num_groups not specified"),                        # dally:2414-2417
    (dict(num_routers=None), "num_routers not specified"),
    (dict(route_scoring_metric="beta"), "Beta scoring not implemented"),
    (dict(credit_delay="10", local_credit_delay="1", global_credit_delay="1", cn_credit_delay="1"),
     "Cannot set both a general credit delay"),                                 # dally:2694-2698
    (dict(credit_delay="10", auto_credit_delay="1"), "auto credit delay"),     # dally:2700-2704
    (dict(num_routers="5"), "Config error"),                                    # dally:4925-4931
    (dict(message_size="512"), "ROSS is configured for events of size"),       # model-net.c:294-300
    (dict(intra__group__connections="/nonexistent"), "intra-group file not found"),
    (dict(num_qos_levels="2"), "priority needs to be specified with qos_levels>1"),  # the reference's own abort (dally:3254)
    (dict(modelnet_scheduler="round-robin"), "modelnet_scheduler"),
])
def test_config_errors_match_reference_behaviour(engine_lib, tmp_path, changes, msg):
    conf = write_conf(tmp_path, **changes)
    with pytest.raises(codes_b200.EngineError, match=re.escape(msg)):
        configure(conf, num_messages=1)
    # the oracle rejects the same configurations
    with pytest.raises(oracle_util.OracleError):
        oracle_util.run(conf, num_messages=1)


def test_defaults_follow_reference(engine_lib, tmp_path):
    # explicit credit_delay and missing optional keys are accepted (dally:2663-2756)
    configure(write_conf(tmp_path, credit_delay="100"), num_messages=1)
    configure(write_conf(tmp_path, routing=None), num_messages=1)  # default: minimal (dally:2300-2305)
    configure(CONF72, num_messages=1, payload_sz=8192)  # multi-packet messages are accepted
    with pytest.raises(codes_b200.EngineError, match="unknown traffic"):
        configure(CONF72, traffic=9)


def test_compute_fails_loudly_without_gpu(engine_lib):
    if codes_b200.cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(codes_b200.EngineError, match="no CUDA device"):
        codes_b200.run_synthetic(CONF72, num_messages=1)


def test_product_never_touches_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "codes_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cc", ".cu", ".h", ".c", "Makefile")):
                with open(os.path.join(dirpath, fn), errors="replace") as f:
                    text = f.read()
                assert "oracle" not in text.lower().replace("oracle/ (test", ""), os.path.join(dirpath, fn)
