"""Dev helper: run one workload on the GPU and print the window kernel's cycle accounting."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import codes_b200
if os.environ.get("DFD_LIB"):  # A/B runs: load another build of the library
    codes_b200.LIB_PATH = os.environ["DFD_LIB"]
conf = sys.argv[1]
opts = dict(kv.split("=") for kv in sys.argv[2:])
opts = {k: (float(v) if "." in v else int(v)) for k, v in opts.items()}
e = codes_b200.Engine()
e.setup(conf, codes_b200.make_opts(**opts))
e.upload(); e.run(); e.finish()
import torch
torch.cuda.init()
e.reset(); torch.cuda.synchronize(); t0 = time.perf_counter(); e.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0; e.finish()
s = e.summary()
w = max(s.windows, 1)
print(f"events {s.events} windows {s.windows} grid {s.grid_blocks}x{s.block_threads} kernel {dt*1e3:.1f} ms -> {s.events/dt/1e6:.2f} Mev/s, {dt/w*1e6:.2f} us/window, {s.events/w:.1f} ev/window")
print(f"cycles/window avg-block: scan {s.cyc_scan/w:.0f} process {s.cyc_process/w:.0f} barrier-wait {s.cyc_barrier/w:.0f} | max-block process {s.cyc_process_max_block/w:.0f} | longest single process phase {s.cyc_process_longest:.0f}")
