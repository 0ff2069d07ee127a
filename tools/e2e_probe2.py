import os, sys, tempfile, subprocess, time
sys.path.insert(0, "/root/repo")
import torch, codes_b200
d = tempfile.mkdtemp()
subprocess.run([sys.executable, "tools/gen_topology.py", "64", "2", d, "--name", "bench", "--routing", "prog-adaptive"], check=True, stdout=subprocess.DEVNULL)
eng = codes_b200.Engine()
eng.setup(os.path.join(d, "bench.conf"), codes_b200.make_opts(traffic=1, num_messages=10, payload_sz=2048))
st = torch.cuda.current_stream().cuda_stream
eng.tw_run(st)
os.environ["DFD_TIMING"] = "1"
for _ in range(3):
    t0 = time.perf_counter(); eng.tw_run(st); print("tw_run %.2f ms" % (1e3 * (time.perf_counter() - t0)))
