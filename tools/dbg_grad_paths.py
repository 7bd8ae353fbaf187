import sys
sys.path[:0]=['/root/repo','/root/repo/doubly-stochastic-dgp_b200']
import numpy as np
from workloads import make_problem, round_f32
from workloads import build_model
for kw in [dict(dims=[8,8,1],N=150,M=100,S=2), dict(dims=[3,3,2],N=70,M=37,S=2)]:
    prob=round_f32(make_problem(seed=5,inner_q_scale=0.3,num_data=500,**kw))
    res={}
    for path in (0,1):
        m=build_model(prob); m._ensure_ctx(prob['N'],prob['S']).set_option("path",path)
        e,g,gl=m.compute_log_likelihood_and_grad(zs=prob['zs']); res[path]=(e,g)
    print(kw, res[0][0], res[1][0])
    for l,(g0,g1) in enumerate(zip(res[0][1],res[1][1])):
        for k in g0:
            a,b=np.asarray(g0[k]),np.asarray(g1[k])
            print(l,k,"relerr %.3g"%(np.abs(a-b).max()/(np.abs(a).max()+1e-30)))
        if 'q_sqrt' in g0:
            a,b=g0['q_sqrt'],g1['q_sqrt']
            d=np.abs(a-b)/ (np.abs(a).max()+1e-30)
            idx=np.unravel_index(np.argmax(d),d.shape); print("   worst q_sqrt idx",idx, a[idx], b[idx])
            print("   ratio sample", (b[0,:4,:4]/(a[0,:4,:4]+1e-30)).round(3))
