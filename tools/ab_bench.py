"""A/B timing of step-level tuning options on the north-star workload (device-resident inputs, back-to-back graph
replays timed with CUDA events on the ctx stream).  usage: python tools/ab_bench.py "opt=val,opt=val" ["..."]"""
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
for spec in sys.argv[1:] or [""]:
    m = build_model(prob)
    ctx = m._ensure_ctx(1000, 20)
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        ctx.set_option(k, float(v))
    m.adam_init(0.01)
    step = lambda i, sync: ctx.train_step(X.data_ptr(), Y.data_ptr(), 1000, 20, 8192, 100 + i,
                                          flags=_lib.FLAG_DEVICE_PTRS | (0 if sync else _lib.FLAG_NO_SYNC), want_elbo=sync)
    for i in range(10):
        e = step(i, True)
    best = 1e9
    for rep in range(3):
        ctx.sync()
        ctx.timer_start()
        for i in range(200):
            step(1000 + i, False)
        best = min(best, ctx.timer_stop() / 200)
    e2 = step(5000, True)
    print(f"{spec or 'default':45s} {best * 1000:8.1f} us/step   elbo {e:.3f} -> {e2:.3f}", flush=True)
    ctx.set_option("dbg_prep", 1)
    m._ctx.close()
