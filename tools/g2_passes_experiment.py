import sys
sys.path[:0] = ['/root/repo', '/root/repo/doubly-stochastic-dgp_b200']
import numpy as np
from workloads import build_model, make_problem
from oracle.problems import build_oracle
from oracle import reference_dgp as R
NORTH = dict(dims=[8, 8, 8, 8, 8, 1], N=1000, M=100, S=20)
def trained(prob, seed):
    rng = np.random.default_rng(seed)
    for lay in prob['layers']:
        M, D = lay['q_mu'].shape
        lay['q_sqrt'] = np.tril(0.1 * rng.normal(size=(D, M, M))) + 0.3 * np.eye(M)[None]
    return prob
for white in (False, True):
    prob = trained(make_problem(seed=3100, white=white, num_data=8192, **NORTH), 1)
    e_ref = build_oracle(prob).compute_log_likelihood(zs=prob['zs'])
    m = build_model(prob); ctx = m._ensure_ctx(1000, 20)
    for p in (1, 2, 3):
        ctx.set_option("g2_passes", p)
        e = m.compute_log_likelihood(zs=prob['zs'])
        print("trained white=%d passes=%d rel=%.2e" % (white, p, abs(e - e_ref) / abs(e_ref)))
# the I6-like case: q_sqrt = 0.3 I on a single layer (natgrad test regime)
for M in (32, 100):
    prob = make_problem(seed=77, dims=[8, 1], N=400, M=M, S=1, num_data=400, max_cond=None)
    prob['layers'][0]['q_sqrt'] = 0.3 * np.eye(M)[None]
    e_ref = build_oracle(prob).compute_log_likelihood(zs=prob['zs'])
    m = build_model(prob); ctx = m._ensure_ctx(400, 1)
    for p in (1, 2, 3):
        ctx.set_option("g2_passes", p)
        e = m.compute_log_likelihood(zs=prob['zs'])
        print("I6-like M=%d passes=%d rel=%.2e" % (M, p, abs(e - e_ref) / abs(e_ref)))
