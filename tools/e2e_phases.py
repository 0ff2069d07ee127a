"""Host wall clock of the three phases of dfd_tw_run (upload + device init, event loop, download + final handlers) on the
bench network: DFD_TIMING=1 python tools/e2e_phases.py [num_messages]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import codes_b200  # noqa: E402

d = tempfile.mkdtemp()
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_topology.py"), "64", "2", d, "--name", "bench", "--routing", "prog-adaptive"],
               check=True, stdout=subprocess.DEVNULL)
eng = codes_b200.Engine()
eng.setup(os.path.join(d, "bench.conf"), codes_b200.make_opts(traffic=1, num_messages=int(sys.argv[1]) if len(sys.argv) > 1 else 10, payload_sz=2048))
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    eng.tw_run(st)
s = eng.summary()
print("events", s.events, "h2d", s.h2d_bytes, "d2h", s.d2h_bytes)
