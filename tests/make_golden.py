"""Regenerates tests/golden/*: golden output vectors for the committed configs.

The reference pins no numeric results for this path.  The vectors are produced in the authoring
container by the reference-source oracle -- the reference's OWN sources for this path compiled unchanged against the
ROSS/MPI shim (oracle/ref -> oracle/_ref) -- and every file is checked to be byte-identical to the output of the
restatement oracle (oracle/dfdally_oracle.cpp) before it is written.  They pin (i) the restatement against accidental
change and (ii) the GPU engine on the GPU box, where /root/reference does not exist.
Run (in the authoring container):  python -m tests.make_golden   (and  python -m tests.make_golden --large)"""
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
            if not oracle_util.ref_available() and not oracle_util.build_ref():
                raise SystemExit("oracle/_ref is needed to generate golden vectors (build it where /root/reference exists)")
            ev, _, _ = oracle_util.run_ref(conf, os.path.join(tmp, "ref"), **case)
            ref = oracle_util.read_dir(os.path.join(tmp, "ref"))
            assert ev == res.events
            for k in oracle_util.REF_FILES:
                assert ref.get(k) == out.get(k), (name, k)
            meta = dict(conf=os.path.relpath(conf, oracle_util.ROOT), options=case, events=res.events,
                        generated_by="oracle/_ref (reference sources + ROSS/MPI shim); identical to oracle/dfdally_oracle.cpp",
                        finished_packets=res.finished_packets, total_hops=res.total_hops,
                        sha256={k: hashlib.sha256(v).hexdigest() for k, v in sorted(out.items())})
            with open(os.path.join(GOLDEN, name + ".json"), "w") as f:
                json.dump(meta, f, indent=1, sort_keys=True)
            # full text of the small cases
            if name.startswith("72_"):
                d = os.path.join(GOLDEN, name)
                shutil.rmtree(d, ignore_errors=True)
                os.makedirs(d)
                for k in ("dragonfly-cn-stats", "dragonfly-link-stats", "synthetic-stats", "summary.txt", "dragonfly-snapshots.csv"):
                    if k not in out:
                        continue
                    with open(os.path.join(d, k), "wb") as f:
                        f.write(out[k])
        print(name, res.events)


LARGE = {
    # name: (radix, links between groups, terminals, messages per terminal, routing, traffic)
    "131584_uniform_m3": (64, 2, 131584, 3, "minimal", 1),     # a=32, p=16, h=16, g=257   (~20 s of oracle time)
    "131584_par_nearest_group_m3": (64, 2, 131584, 3, "prog-adaptive", 3),  # BASELINE configs[3]: PAR, worst-case group permutation
    "131584_par_uniform_m10": (64, 2, 131584, 10, "prog-adaptive", 1),  # bench.py's workload (parity_digest_ok)
    "131584_par_uniform_m50": (64, 2, 131584, 50, "prog-adaptive", 1),  # the same in a long run: saturated network (154.6 M events, ~8 min)
    "131584_par_nearest_group_m50": (64, 2, 131584, 50, "prog-adaptive", 3),  # configs[3] in a long run (173.2 M events)
    "1050624_uniform_m1": (128, 4, 1050624, 1, "minimal", 1),  # a=64, p=32, h=32, g=513   (~2 min, 1 GB of lp-io text)
}


def large(which):
    """Large goldens (BASELINE configs[3] / configs[4] scale): topology from tools/gen_topology.py, uniform traffic.
    Digests only; checked on the GPU box by tests/test_gpu_parity.py::test_large_network_matches_golden."""
    import subprocess
    import sys
    radix, conns, _, msgs, routing, traffic = LARGE[which]
    with tempfile.TemporaryDirectory() as tmp:
        net = os.path.join(tmp, "net")
        subprocess.run([sys.executable, os.path.join(oracle_util.ROOT, "tools", "gen_topology.py"), str(radix), str(conns), net, "--name", "big",
                        "--routing", routing], check=True, stdout=subprocess.DEVNULL)
        conf = os.path.join(net, "big.conf")
        case = dict(num_messages=msgs, traffic=traffic)
        res = oracle_util.run(conf, lp_io_dir=os.path.join(tmp, "out"), **case)
        out = oracle_util.read_dir(os.path.join(tmp, "out"))
        # oracle/_ref needs 14 minutes and 20-28 GB per 131,584-terminal case (tools/ref_check_large.py runs it and records the check);
        # the restatement is pinned to it on the smaller cases of tests/test_oracle_cpu.py
        meta = dict(conf=f"tools/gen_topology.py {radix} {conns} <dir> --name big --routing {routing}", options=case, events=res.events,
                    generated_by="oracle/dfdally_oracle.cpp (restatement oracle)",
                    finished_packets=res.finished_packets, total_hops=res.total_hops,
                    sha256={k: hashlib.sha256(v).hexdigest() for k, v in sorted(out.items())})
        with open(os.path.join(GOLDEN, which + ".json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        print(which, res.events)


if __name__ == "__main__":
    import sys
    if "--large" in sys.argv:
        for name in LARGE:
            if len(sys.argv) < 3 or name in sys.argv[2:]:
                large(name)
    else:
        main()
