"""Per-stage device times of one north-star training step (eager launches bracketed by CUDA events: dsdgp_profile).
usage: python tools/stage_profile.py ["opt=val,..."]"""
import sys
sys.path[:0] = ['/root/repo', '/root/repo/doubly-stochastic-dgp_b200']
import numpy as np
import torch
from doubly_stochastic_dgp import _lib
from workloads import build_model
from workloads import make_problem

prob = make_problem(seed=3000, dims=[8, 8, 8, 8, 8, 1], N=1000, M=100, S=20, num_data=8192)
X = torch.from_numpy(np.float32(prob['X'])).cuda()
Y = torch.from_numpy(np.float32(prob['Y'])).cuda()
m = build_model(prob)
ctx = m._ensure_ctx(1000, 20)
for kv in filter(None, (sys.argv[1] if len(sys.argv) > 1 else "").split(",")):
    k, v = kv.split("=")
    ctx.set_option(k, float(v))
m.adam_init(0.01)
step = lambda i: ctx.train_step(X.data_ptr(), Y.data_ptr(), 1000, 20, 8192, 100 + i, flags=_lib.FLAG_DEVICE_PTRS, want_elbo=True)
for i in range(5):
    step(i)
ctx.set_option("profile", 1)
acc = None
for i in range(6):
    step(10 + i)
    p = np.array(ctx.profile())
    if i:
        acc = p if acc is None else acc + p
ctx.set_option("profile", 0)
p = acc / 5
L = 5
names = ["prep", "likelihood", "grad-assembly(tail)", "allreduce", "tail(result+adam)"]
for l in range(L):
    names += [f"L{l + 1}.fwd", f"L{l + 1}.bwd_rows", f"L{l + 1}.rowred"]
for n, v in zip(names, p):
    print(f"{n:24s} {v * 1000:8.1f} us")
print(f"{'sum':24s} {p.sum() * 1000:8.1f} us")
