"""Regenerates tests/golden/*: outputs of the CPU oracle on the committed configs.

The reference pins no numeric results for this path and cannot be built here (ROSS, MPI,
flex/bison are absent), so these vectors are produced by oracle/dfdally_oracle.cpp, which restates the
reference's handlers.  They pin (i) the oracle against accidental change and (ii) the GPU engine on the
GPU box, where /root/reference does not exist.  Run:  python -m tests.make_golden"""
import hashlib
import json
import os
import shutil
import tempfile

from tests import oracle_util
from tests.test_oracle_cpu import GOLDEN, GOLDEN_CASES


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for name, case in sorted(GOLDEN_CASES.items()):
        case = dict(case)
        conf = case.pop("conf")
        with tempfile.TemporaryDirectory() as tmp:
            res = oracle_util.run(conf, lp_io_dir=tmp, **case)
            out = oracle_util.read_dir(tmp)
            meta = dict(conf=os.path.relpath(conf, oracle_util.ROOT), options=case, events=res.events,
                        finished_packets=res.finished_packets, total_hops=res.total_hops,
                        sha256={k: hashlib.sha256(v).hexdigest() for k, v in sorted(out.items())})
            with open(os.path.join(GOLDEN, name + ".json"), "w") as f:
                json.dump(meta, f, indent=1, sort_keys=True)
            # full text of the small cases
            if name.startswith("72_"):
                d = os.path.join(GOLDEN, name)
                shutil.rmtree(d, ignore_errors=True)
                os.makedirs(d)
                for k in ("dragonfly-cn-stats", "dragonfly-link-stats", "synthetic-stats", "summary.txt"):
                    with open(os.path.join(d, k), "wb") as f:
                        f.write(out[k])
        print(name, res.events)


if __name__ == "__main__":
    main()
