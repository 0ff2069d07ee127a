"""The drop-in boundary under the reference's own names (include/codes_compat.h, libcodes_compat_b200.so).

CPU: the library loads and exports every entry point the header declares; where the reference tree exists, the header
agrees with the reference's declarations (one translation unit includes both) and the reference's UNMODIFIED
model-net-synthetic-dragonfly-all.c compiles and links against it; without a GPU the binary fails loudly.
GPU: that binary -- the reference's main(), command line and call sequence, tw_run() on the B200 -- and the C twin driver
write lp-io directories that are byte-identical to the golden vectors produced by the reference's own sources
(tests/golden, the lpio-equivalence pattern of the reference's tests/CMakeLists.txt:654-660)."""
import ctypes
import hashlib
import json
import os
import re
import subprocess

import pytest

from tests import oracle_util

ROOT = oracle_util.ROOT
LIB = os.path.join(ROOT, "codes_b200", "libcodes_compat_b200.so")
REFMAIN = os.path.join(ROOT, "codes_b200", "_refmain", "model-net-synthetic-dragonfly-all")
TWIN = os.path.join(ROOT, "codes_b200", "model-net-synthetic-dragonfly-all")
REF = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def declared_functions():
    with open(os.path.join(ROOT, "include", "codes_compat.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    text = text[text.index('extern "C"'):]
    text = re.sub(r"struct model_net_method \{.*?\n\};", "", text, flags=re.S)  # its members are pointers, not exports
    names = set(re.findall(r"\b([a-z][a-z0-9_]+)\s*\(", text))
    return sorted(names - {"defined", "sizeof"})


def test_compat_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    names = declared_functions()
    for must in ("model_net_register", "model_net_configure", "model_net_event", "model_net_report_stats", "lp_type_register",
                 "lp_type_lookup", "codes_mapping_setup", "codes_mapping_get_lp_count", "lp_io_prepare", "lp_io_flush",
                 "configuration_load"):
        assert must in names
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/codes_compat.h but not exported"
    for obj in ("dragonfly_dally_method", "dragonfly_dally_router_method", "config", "g_tw_ts_end", "MPI_COMM_CODES"):
        ctypes.c_void_p.in_dll(lib, obj)
    for n in ("tw_run", "tw_init", "tw_opt_add", "tw_end", "tw_event_new", "MPI_Reduce"):  # include/codes_compat/{ross,mpi}.h
        assert hasattr(lib, n)


def test_method_structs_have_the_reference_field_order():
    """codes/model-net-method.h:24-66: packet_size, mn_configure, mn_register, packet_event, packet_event_rc, recv_msg_event,
    recv_msg_event_rc, mn_get_lp_type, mn_get_msg_sz, mn_report_stats, ...; the terminal method has configure / register /
    packet_event / get_lp_type / get_msg_sz / report_stats, the router method only register / get_lp_type / get_msg_sz
    (dragonfly-dally.cxx:10972-11024)."""
    lib = ctypes.CDLL(LIB)
    words = ctypes.c_void_p * 24
    term = words.in_dll(lib, "dragonfly_dally_method")
    rtr = words.in_dll(lib, "dragonfly_dally_router_method")
    assert [bool(term[i]) for i in range(1, 10)] == [True, True, True, True, False, False, True, True, True]
    assert [bool(rtr[i]) for i in range(1, 10)] == [False, True, False, False, False, False, True, True, False]
    assert not any(term[i] for i in range(10, 24)) and not any(rtr[i] for i in range(10, 24))


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (authoring container)")
def test_header_agrees_with_the_reference_declarations(tmp_path):
    src = tmp_path / "both.c"
    src.write_text('#include "codes/model-net.h"\n#include "codes/lp-io.h"\n#include "codes/codes.h"\n#include "codes/codes_mapping.h"\n'
                   '#include "codes/configuration.h"\n#include "codes/lp-type-lookup.h"\n#include "codes/model-net-method.h"\n'
                   '#include "codes_compat.h"\n'
                   '_Static_assert(DRAGONFLY_DALLY == 15 && DRAGONFLY_DALLY_ROUTER == 16, "enum NETWORKS moved");\n'
                   'int main(void) { return 0; }\n')
    subprocess.run(["gcc", "-std=gnu11", "-fsyntax-only", "-I", os.path.join(ROOT, "include", "codes_compat"), "-I", REF,
                    "-I", os.path.join(ROOT, "include"), str(src)], check=True)


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (authoring container)")
def test_unmodified_reference_driver_builds_against_the_compat_library():
    subprocess.run(["make", "-C", os.path.join(ROOT, "codes_b200", "csrc"), "refmain"], check=True, stdout=subprocess.DEVNULL)
    assert os.access(REFMAIN, os.X_OK)
    und = subprocess.run(["nm", "-u", REFMAIN], capture_output=True, text=True, check=True).stdout
    for n in ("model_net_register", "model_net_configure", "model_net_event", "lp_type_register", "codes_mapping_setup", "tw_run",
              "lp_io_prepare", "lp_io_flush", "model_net_report_stats"):
        assert re.search(rf"\bU {n}\b", und), n  # resolved by libcodes_compat_b200.so at load time


def has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(not os.path.exists(REFMAIN), reason="codes_b200/_refmain is built where the reference tree exists")
def test_reference_driver_fails_loudly_without_a_gpu(tmp_path):
    if has_gpu():
        pytest.skip("a GPU is present")
    r = subprocess.run([REFMAIN, "--sync=1", "--num_messages=2", f"--lp-io-dir={tmp_path}/out", "--", os.path.join(ROOT, "configs", "dfdally-72.conf")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def check_against_golden(out_dir, gold_name):
    with open(os.path.join(GOLDEN, gold_name + ".json")) as f:
        gold = json.load(f)
    for fname, digest in gold["sha256"].items():
        if fname in ("summary.txt", "raw-lp-stats"):  # written by the oracle / the Python binding only
            continue
        with open(os.path.join(out_dir, fname), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, fname
    return gold


CONF72 = os.path.join(ROOT, "configs", "dfdally-72.conf")
CASES = [("72_uniform_m1", CONF72, ["--traffic=1", "--num_messages=1"]), ("72_uniform_m20", CONF72, ["--traffic=1", "--num_messages=20"]),
         ("72_nearest_group_m10", CONF72, ["--traffic=3", "--num_messages=10"]),
         # PARAMS router_buffer_snapshots: dragonfly-snapshots.csv through the drivers' own lp_io_flush
         ("72_snapshots_nearest_group_m20", os.path.join(ROOT, "configs", "dfdally-72-snapshots.conf"), ["--traffic=3", "--num_messages=20"])]


@pytest.mark.gpu
@pytest.mark.parametrize("gold_name,conf,args", CASES, ids=[c[0] for c in CASES])
def test_reference_main_on_the_engine_matches_golden(tmp_path, gold_name, conf, args):
    """The reference's own main() (unmodified source, compiled in the authoring container) with tw_run() on the B200."""
    if not os.path.exists(REFMAIN):
        pytest.skip("codes_b200/_refmain is built where the reference tree exists and travels with the snapshot")
    out = str(tmp_path / "lpio")
    r = subprocess.run([REFMAIN, "--sync=1", *args, f"--lp-io-dir={out}", "--", conf], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    gold = check_against_golden(out, gold_name)
    m = re.search(r"Net Events Processed\s+(\d+)", r.stdout)
    assert m and int(m.group(1)) == gold["events"]
    assert "Synthetic Generator: Detected Dragonfly Dally" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("gold_name,conf,args", CASES, ids=[c[0] for c in CASES])
def test_c_twin_driver_matches_golden(tmp_path, gold_name, conf, args):
    """codes_b200/model-net-synthetic-dragonfly-all (csrc/synthetic_main.c): same command line through the dfd_* C-ABI."""
    out = str(tmp_path / "lpio")
    r = subprocess.run([TWIN, "--sync=1", *args, f"--lp-io-dir={out}", "--", conf], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    check_against_golden(out, gold_name)
