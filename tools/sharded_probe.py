"""Dev/measurement helper (torchrun): run one config group-sharded over the ranks and print the kernel time.
   python -m torch.distributed.run --nproc-per-node N tools/sharded_probe.py <conf> key=value ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import codes_b200

conf = sys.argv[1]
opts = dict(kv.split("=") for kv in sys.argv[2:])
opts = {k: (float(v) if "." in v else int(v)) for k, v in opts.items()}
rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("cpu:gloo,cuda:nccl")
world = dist.get_world_size()
eng = codes_b200.Engine()
eng.setup(conf, codes_b200.make_opts(**opts))
eng.set_partition(rank, world)
eng.upload()
handles = [None] * world
dist.all_gather_object(handles, eng.ipc_export())
eng.ipc_import(handles)
times = []
for it in range(2):
    torch.cuda.synchronize(); dist.barrier()
    eng.reset()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    eng.run()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    times.append(float(dt[0]))
torch.cuda.synchronize(); dist.barrier()
eng.finish()
dist.barrier()
if rank == 0:
    s = eng.summary()
    dt = times[-1]
    print(f"[{world} GPUs sharded] {os.path.basename(conf)} {opts}: events {s.events} packets {s.finished_packets} windows {s.windows} "
          f"kernel {dt*1e3:.1f} ms -> {s.events/dt/1e6:.1f} Mev/s, {s.finished_packets/dt/1e6:.2f} Mpkt/s, {dt/max(s.windows,1)*1e6:.2f} us/window", flush=True)
dist.barrier()
dist.destroy_process_group()
