"""Dev/measurement helper: K independent simulations of one workload at once on ONE GPU, one CUDA stream each
(the window kernel of a 1,056-terminal network occupies 83 of the GPU's 296 block slots).
   python tools/ensemble_probe.py <conf> <K> key=value ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import codes_b200

conf, K = sys.argv[1], int(sys.argv[2])
opts = dict(kv.split("=") for kv in sys.argv[3:])
opts = {k: (float(v) if "." in v else int(v)) for k, v in opts.items()}
engines, streams = [], []
for k in range(K):
    e = codes_b200.Engine()
    e.setup(conf, codes_b200.make_opts(**opts))
    e.set_ensemble(K)
    s = torch.cuda.Stream()
    e.upload(s.cuda_stream)
    engines.append(e)
    streams.append(s)
torch.cuda.synchronize()
for rep in range(2):
    for e, s in zip(engines, streams):
        e.reset(s.cuda_stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for e, s in zip(engines, streams):
        e.run(s.cuda_stream)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
ev = 0
for e, s in zip(engines, streams):
    e.finish(s.cuda_stream)
    ev += e.summary().events
g = engines[0].summary()
print(f"{K} concurrent simulations ({g.grid_blocks}x{g.block_threads} each): {ev} events in {dt*1e3:.1f} ms -> {ev/dt/1e6:.2f} M events/s aggregate")
