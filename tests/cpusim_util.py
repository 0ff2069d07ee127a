"""Binding of tests/cpusim/libdfdally_cpusim.so -- TEST INFRASTRUCTURE ONLY.

The library is codes_b200/csrc/dfd_host.cc plus the scheduler / handler code of codes_b200/csrc/dfd_core.h compiled for
the CPU and driven by a seeded random sequential scheduler (tests/cpusim/dfd_cpusim.cc).  It exists so that the
asynchronous scheduling logic can be checked against the oracle without a GPU; the product (`codes_b200`) never loads it."""
import os
import subprocess

import codes_b200

HERE = os.path.dirname(os.path.abspath(__file__))
DIR = os.path.join(HERE, "cpusim")
LIB = os.path.join(DIR, "libdfdally_cpusim.so")

_lib = None


def load():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-s", "-C", DIR], check=True)
        _lib = codes_b200.bind_library(LIB)
    return _lib


def run(conf, lp_io_dir=None, seed=1, world=1, warps=None, **opts):
    os.environ["DFD_CPUSIM_SEED"] = str(seed)
    os.environ["DFD_CPUSIM_WORLD"] = str(world)
    if warps:
        os.environ["DFD_CPUSIM_WARPS"] = str(warps)
    else:
        os.environ.pop("DFD_CPUSIM_WARPS", None)
    return codes_b200.run_synthetic(conf, lp_io_dir=lp_io_dir, lib=load(), **opts)
