"""2-GPU check of the NCCL path (skipped on single-GPU boxes): row-sharded ELBO + gradient with the in-graph
all-reduce == single-GPU result on the whole minibatch; Philox draws are shard-invariant so a train step agrees too."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, tmp, kw):
    import torch
    from doubly_stochastic_dgp import _lib
    from tests.gpu_common import build_model
    from tests.synth import make_problem
    torch.cuda.set_device(rank)
    prob = make_problem(**kw)
    N = prob['N']
    lo, hi = rank * N // world, (rank + 1) * N // world
    idf = os.path.join(tmp, "id.bin")
    if rank == 0:
        with open(idf + ".tmp", "wb") as f:
            f.write(_lib.comm_unique_id())
        os.rename(idf + ".tmp", idf)
    while not os.path.exists(idf):
        time.sleep(0.05)
    m = build_model(prob, device=rank)
    m.comm_init(open(idf, "rb").read(), rank, world)
    zs = [z[:, lo:hi] for z in prob['zs']]
    e, grads, glik = m.compute_log_likelihood_and_grad(zs=zs, X=prob['X'][lo:hi], Y=prob['Y'][lo:hi])
    # Philox path: same seed on both ranks
    ctx = m._ctx
    e2 = ctx.elbo(prob['X'][lo:hi], prob['Y'][lo:hi], prob['S'], prob['num_data'], seed=4242)
    if rank == 0:
        np.savez(os.path.join(tmp, "out.npz"), e=e, e2=e2, glik=glik,
                 **{f"g{l}_{k}": v for l, g in enumerate(grads) for k, v in g.items()})


def test_two_gpu_allreduce_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from tests.gpu_common import build_model
    from tests.synth import make_problem
    kw = dict(seed=91, dims=[8, 8, 1], N=256, M=100, S=4, inner_q_scale=0.3, num_data=2560)
    mp.spawn(_worker, args=(2, str(tmp_path), kw), nprocs=2, join=True)
    out = np.load(str(tmp_path / "out.npz"))
    prob = make_problem(**kw)
    m = build_model(prob)
    e, grads, glik = m.compute_log_likelihood_and_grad(zs=prob['zs'])
    assert abs(out['e'] - e) <= 2e-6 * abs(e)
    for l, g in enumerate(grads):
        for k, v in g.items():
            sc = np.abs(v).max() + 1e-12
            np.testing.assert_allclose(out[f"g{l}_{k}"], v, atol=2e-3 * sc, rtol=0, err_msg=f"{k} l={l}")
    e2 = m._ctx.elbo(prob['X'], prob['Y'], prob['S'], prob['num_data'], seed=4242)
    assert abs(out['e2'] - e2) <= 2e-6 * abs(e2)
