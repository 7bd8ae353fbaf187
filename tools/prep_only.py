import sys
sys.path[:0]=['/root/repo','/root/repo/doubly-stochastic-dgp_b200']
from workloads import make_problem
from workloads import build_model
prob=make_problem(seed=3000,dims=[8,8,8,8,8,1],N=1000,M=100,S=20,num_data=8192)
m=build_model(prob); ctx=m._ensure_ctx(1000,20)
ctx.set_option("graph",0)
for kv in sys.argv[1:]:
    k,v=kv.split("="); ctx.set_option(k,float(v))
for i in range(4): ctx.kl()
ctx.sync()
