import sys
sys.path[:0]=['/root/repo','/root/repo/doubly-stochastic-dgp_b200']
import numpy as np
from workloads import make_problem
from workloads import build_model
prob=make_problem(seed=3000,dims=[8,8,8,8,8,1],N=1000,M=100,S=20,num_data=8192)
m=build_model(prob); ctx=m._ensure_ctx(1000,20)
ctx.set_option("graph",0)
m.adam_init(0.01)
for i in range(3): m.train_step()
ctx.sync()
