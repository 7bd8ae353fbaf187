import sys
sys.path[:0]=['/root/repo','/root/repo/doubly-stochastic-dgp_b200']
import numpy as np
from workloads import make_problem
from workloads import build_model
prob=make_problem(seed=3000,dims=[8,8,8,8,8,1],N=1000,M=100,S=20,num_data=8192)
m=build_model(prob)
ctx=m._ensure_ctx(1000,20)
for layer in [int(a) for a in sys.argv[1:]]:
    ctx.set_option("dbg_layer",layer)
    for i in range(3): ctx.elbo_grad(prob['X'],prob['Y'],20,8192,seed=i)
    print("layer",layer,file=sys.stderr)
    ctx.set_option("dbg_dump",1)
ctx.set_option("dbg_layer",-1)
ctx.elbo_grad(prob['X'],prob['Y'],20,8192,seed=5)
ctx.set_option("dbg_prep",1)
