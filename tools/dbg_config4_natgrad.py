"""Repeated NatGrad steps on BASELINE config 4 (M=512, Matern52): which step fails, and the ELBO trajectory."""
import sys
sys.path[:0] = ['/root/repo', '/root/repo/doubly-stochastic-dgp_b200']
import numpy as np
import torch
from doubly_stochastic_dgp import _lib
from workloads import build_model, make_problem

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 32
gamma = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
prob = make_problem(seed=4000, dims=[9, 9, 9, 1], N=N, M=512, S=S, kern='matern52', num_data=45730)
m = build_model(prob)
ctx = m._ensure_ctx(N, S)
X = torch.from_numpy(np.float32(prob['X'])).cuda()
Y = torch.from_numpy(np.float32(prob['Y'])).cuda()
last = 2
for i in range(12):
    try:
        e = ctx.natgrad_step(X.data_ptr(), Y.data_ptr(), N, S, 45730, 1000 + i, [last], gamma, flags=_lib.FLAG_DEVICE_PTRS)
    except Exception as ex:
        print("step", i, "FAILED:", str(ex)[:120])
        break
    snap = m._layer_snapshot(last) if hasattr(m, "_layer_snapshot") else None
    qs = np.asarray(m.layers[last].q_sqrt) if snap is None else None
    print("step", i, "elbo", e, flush=True)
