"""Checks a LARGE golden vector (tests/golden/<name>.json, written by the restatement oracle) against the reference's own
sources: oracle/_ref (the reference's handler code compiled unchanged on the sequential shim engine) is run on the same
case -- about 14 minutes of set-up and 20 GB for the 131,584-terminal network -- and its `Net Events Processed` and every
lp-io file it writes are compared with the committed digests.  On success the golden's JSON records the check.
Authoring container only (needs /root/reference for oracle/_ref).   usage: python tools/ref_check_large.py <golden name> [out dir]"""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_util  # noqa: E402
from tests.make_golden import LARGE  # noqa: E402

name = sys.argv[1]
radix, conns, _, msgs, routing, traffic = LARGE[name]
gold_path = os.path.join(ROOT, "tests", "golden", name + ".json")
gold = json.load(open(gold_path))
work = sys.argv[2] if len(sys.argv) > 2 else tempfile.mkdtemp()
out = os.path.join(work, "out")
if not os.path.isdir(out):  # an existing directory is taken as the output of an earlier run of exactly this case
    net = os.path.join(work, "net")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_topology.py"), str(radix), str(conns), net, "--name", "big", "--routing", routing],
                   check=True, stdout=subprocess.DEVNULL)
    t0 = time.time()
    with open(os.path.join(work, "stdout.log"), "w") as so, open(os.path.join(work, "stderr.log"), "w") as se:
        subprocess.run([oracle_util.REF_BIN, "--sync=1", f"--traffic={traffic}", f"--num_messages={msgs}", "--payload_sz=2048", f"--lp-io-dir={out}", "--",
                        "big.conf"], cwd=net, stdout=so, stderr=se, check=True)
    print(f"oracle/_ref finished in {time.time() - t0:.0f} s")
stdout = open(os.path.join(work, "stdout.log")).read()
events = int(re.search(r"Net Events Processed\s+(\d+)", stdout).group(1))
ok = events == gold["events"]
print("events", events, "golden", gold["events"])
checked = []
for fname, digest in sorted(gold["sha256"].items()):
    p = os.path.join(out, fname)
    if not os.path.exists(p):
        continue  # raw-lp-stats and summary.txt are written by the restatement oracle only
    same = hashlib.sha256(open(p, "rb").read()).hexdigest() == digest
    print(fname, "identical" if same else "DIFFERENT")
    ok = ok and same
    checked.append(fname)
if ok:
    gold["reference_sources_check"] = ("oracle/_ref (the reference's own sources on the sequential shim engine) run on this case by tools/ref_check_large.py: "
                                       "Net Events Processed and " + ", ".join(checked) + " byte-identical to these digests")
    json.dump(gold, open(gold_path, "w"), indent=1, sort_keys=True)
print("REFERENCE SOURCES CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
