"""Why is the host-buffer pipeline slow?  (GPU box, scratch tool)"""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
from emoportraits_b200.config import shipped_config
from emoportraits_b200.infer import Model, DriverPipeline
S = 512
cfg = shipped_config(S)
model = Model(cfg, synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0), "cuda")
g = torch.Generator().manual_seed(0)
st = model.source_pass(torch.rand(1, 3, S, S, generator=g).cuda())
for depth in (1, 2):
    pipe = DriverPipeline(model, st, depth=depth)
    fh = [torch.rand(1, 3, S, S, generator=g).pin_memory() for _ in range(4)]
    fd = [f.cuda() for f in fh]
    oh = [torch.empty(1, 3, S, S).pin_memory() for _ in range(depth)]
    print("pinned:", fh[0].is_pinned(), oh[0].is_pinned())
    def loop(K, host_in, host_out, wait):
        torch.cuda.synchronize(); t0 = time.perf_counter(); tw = 0.0; ts = 0.0
        for i in range(K):
            sl = pipe.slots[i % depth]
            a = time.perf_counter()
            if wait and sl.busy: sl.done.synchronize()
            b = time.perf_counter()
            pipe.submit(fh[i % 4] if host_in else fd[i % 4], host_out=oh[i % depth] if host_out else None)
            c = time.perf_counter(); tw += b - a; ts += c - b
        pipe.drain(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"depth {depth} host_in {host_in} host_out {host_out} wait {wait}: {dt / K * 1e3:.2f} ms/frame (wait {tw / K * 1e3:.2f} submit {ts / K * 1e3:.2f})", flush=True)
    for args in [(0, 0, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1), (1, 1, 0)]:
        loop(5, *args); loop(30, *args)
