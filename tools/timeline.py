"""In-graph timeline of one north-star training step: %globaltimer stamps written by one-thread kernels behind each stage of
the 3-stream step DAG (option "timeline"), so the times include the real overlap of the side branches.
usage: python tools/timeline.py ["opt=val,..."]"""
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
ctx.set_option("timeline", 1)
for i in range(8):
    ctx.train_step(X.data_ptr(), Y.data_ptr(), 1000, 20, 8192, 100 + i, flags=_lib.FLAG_DEVICE_PTRS, want_elbo=True)
for rep in range(3):
    ctx.train_step(X.data_ptr(), Y.data_ptr(), 1000, 20, 8192, 200 + rep, flags=_lib.FLAG_DEVICE_PTRS, want_elbo=True)
    print(f"--- step {rep}", file=sys.stderr)
    ctx.set_option("timeline_dump", 1)
