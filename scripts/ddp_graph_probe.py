"""Probe: does an NCCL all-reduce inside a captured CUDA graph replay correctly on this box?"""
import os, sys, time, torch, torch.distributed as dist
def log(*a):
    print("[r%s %.1fs]" % (os.environ.get("RANK"), time.time() - T0), *a, file=sys.stderr, flush=True)
T0 = time.time()
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev); log("init")
x = torch.ones(1 << 20, device=dev) * (rank + 1)
dist.all_reduce(x); torch.cuda.synchronize(); log("eager allreduce", float(x[0]))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        y = x * 2; dist.all_reduce(y)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); log("warm")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    y = x * 2; dist.all_reduce(y); z = y / dist.get_world_size()
torch.cuda.synchronize(); log("captured")
for i in range(5):
    g.replay()
torch.cuda.synchronize(); log("replayed", float(z[0]))
dist.barrier(); log("barrier (destroy_process_group would hang here after a captured collective)"); os._exit(0)
