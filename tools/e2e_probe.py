import sys, os, time
sys.path.insert(0, "/root/repo")
import torch, codes_b200
e = codes_b200.Engine()
e.setup("configs/dfdally-1056.conf", codes_b200.make_opts(num_messages=100))
for i in range(3):
    t0 = time.perf_counter(); e.tw_run(); print("tw_run %.2f ms" % ((time.perf_counter() - t0) * 1e3))
