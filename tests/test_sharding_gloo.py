"""Multi-process (world_size 2, gloo, CPU) check of the data-parallel protocol implemented in csrc/api.cu:
each rank gets a contiguous shard of the minibatch rows (all S samples of them), scales its likelihood term by
num_data / (N_global * S), weights the KL value and gradient by 1/world, and ONE sum all-reduce of
[grad || ELBO] reproduces the single-process ELBO and gradient.  The per-rank arithmetic is the NumPy mirror of
the kernels (tests/algo_mirror.py); the reduction is a real torch.distributed all_reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import algo_mirror as A
from tests.synth import make_problem


def _layers(prob):
    return [A.LayerP(prob['kern'], l['Z'], l['q_mu'], l['q_sqrt'], l['ls'], l['var'], l['white'], l['mean'], W=l['W'],
                     bvec=None if l['W'] is None else np.zeros(l['dout'])) for l in prob['layers']]


def _flat(e, grads, lv):
    parts = []
    for g in grads:
        parts += [np.ravel(g['Z']), np.ravel(g['q_mu']), np.ravel(g['q_sqrt']), np.ravel(g['ls']), np.ravel(g['var'])]
    return np.concatenate(parts + [np.array([lv, e])])


def _worker(rank, world, port, kw, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = make_problem(**kw)
    N = prob['N']
    lo, hi = rank * N // world, (rank + 1) * N // world
    zs = [z[:, lo:hi] for z in prob['zs']]
    e, grads, lv = A.elbo_and_grad(_layers(prob), prob['X'][lo:hi], prob['Y'][lo:hi], prob['lik_var'], prob['S'], zs,
                                   prob['num_data'], prob['jitter'], n_global=N, klw=1.0 / world)
    buf = torch.from_numpy(_flat(e, grads, lv))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out, buf.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("white", [False, True])
def test_row_sharded_allreduce_reproduces_single_process(tmp_path, white):
    kw = dict(seed=77, dims=[3, 3, 2], N=24, M=6, S=3, white=white, inner_q_scale=0.3, num_data=240)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, port, kw, out), nprocs=2, join=True)
    got = np.load(out)
    prob = make_problem(**kw)
    ref = _flat(*A.elbo_and_grad(_layers(prob), prob['X'], prob['Y'], prob['lik_var'], prob['S'], prob['zs'],
                                 prob['num_data'], prob['jitter']))
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)


def test_uneven_shards_need_explicit_n_global():
    """With uneven shards (N not divisible by world) each rank must be told N_global (dsdgp_set_option "n_global"):
    the partial sums are still exact."""
    kw = dict(seed=78, dims=[2, 2, 1], N=11, M=4, S=2, inner_q_scale=0.3, num_data=110)
    prob = make_problem(**kw)
    full = A.elbo_and_grad(_layers(prob), prob['X'], prob['Y'], prob['lik_var'], prob['S'], prob['zs'], prob['num_data'],
                           prob['jitter'])
    acc = 0.0
    for lo, hi in ((0, 4), (4, 11)):
        e, _, _ = A.elbo_and_grad(_layers(prob), prob['X'][lo:hi], prob['Y'][lo:hi], prob['lik_var'], prob['S'],
                                  [z[:, lo:hi] for z in prob['zs']], prob['num_data'], prob['jitter'], n_global=11, klw=0.5)
        acc += e
    np.testing.assert_allclose(acc, full[0], rtol=1e-10)


def _ng_worker(rank, world, port, kw, out):
    """dsdgp_natgrad_step with a communicator: each rank all-reduces the layer's [P_d | G | qmubar] block (partial sums over
    its rows; the KL part of the natural gradient is analytic and not reduced), then applies the update locally."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = make_problem(**kw)
    N = prob['N']
    lo, hi = rank * N // world, (rank + 1) * N // world
    layers = _layers(prob)
    aux = {}
    A.elbo_and_grad(layers, prob['X'][lo:hi], prob['Y'][lo:hi], prob['lik_var'], prob['S'], [z[:, lo:hi] for z in prob['zs']],
                    prob['num_data'], prob['jitter'], n_global=N, klw=1.0 / world, aux=aux)
    l = len(layers) - 1
    Pd, qb = torch.from_numpy(aux[l]['Pd'].copy()), torch.from_numpy(aux[l]['qmubar'].copy())
    dist.all_reduce(Pd, op=dist.ReduceOp.SUM)
    dist.all_reduce(qb, op=dist.ReduceOp.SUM)
    mu, sq = A.natgrad_update(layers[l], aux[l]['Kinv'], Pd.numpy(), qb.numpy(), 1.0)
    np.savez(f"{out}.{rank}.npz", mu=mu, sq=sq)
    dist.destroy_process_group()


def test_natgrad_on_allreduced_accumulators_matches_single_process(tmp_path):
    kw = dict(seed=79, dims=[3, 3, 1], N=24, M=6, S=3, inner_q_scale=0.3, num_data=240)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "ng")
    mp.spawn(_ng_worker, args=(2, port, kw, out), nprocs=2, join=True)
    prob = make_problem(**kw)
    layers = _layers(prob)
    aux = {}
    A.elbo_and_grad(layers, prob['X'], prob['Y'], prob['lik_var'], prob['S'], prob['zs'], prob['num_data'], prob['jitter'],
                    aux=aux)
    l = len(layers) - 1
    mu, sq = A.natgrad_update(layers[l], aux[l]['Kinv'], aux[l]['Pd'], aux[l]['qmubar'], 1.0)
    for rank in range(2):
        got = np.load(f"{out}.{rank}.npz")
        np.testing.assert_allclose(got['mu'], mu, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(got['sq'], sq, rtol=1e-8, atol=1e-10)
