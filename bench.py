#!/usr/bin/env python
"""bench.py -- ELBO training steps/sec of the doubly-stochastic DGP hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference restatement (torch-CPU float64) on host cores

Workload (default, --config 3) = BASELINE.json configs[2] (the config the metric is quoted on): 5-layer RBF DGP, kin8nm
shape, N=1000 minibatch rows, M=100 inducing points, S=20 samples, dims 8->8->8->8->8->1, synthetic data.
One step = minibatch in + ELBO forward + full backward + Adam update (= one session.run(minimize_op) of the
reference, demos/run_regression.py:83,138).  --config 2 / 4 / 5 run the other BASELINE configs (4: the step is one
natural-gradient step on the final layer's q(U), as BASELINE.json words it).

N > 1 (torchrun): --scaling strong (default) is the BASELINE metric itself -- S_total = 20 fixed, the S*N sample rows are
sharded, `value` = steps/s of that one problem; --scaling weak keeps S=20 per GPU (S_total = 20 N) and reports it in
`other_scaling`.  Every run prints a `parity` block (ELBO of the sharded evaluation vs a single-GPU evaluation of the same
global minibatch with the same Philox seed, and vs the float64 oracle at N=1) and exits non-zero beyond 1e-4.

Timing: W (>=3) warm-up steps; K timed steps bracketed by barrier + synchronize; every timed step is measured
with CUDA events on the stream the kernels are launched on (the ctx stream), L2 is flushed (256 MiB write)
between timed steps, the per-step device times are summed and the MAX over ranks is reported.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "doubly-stochastic-dgp_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# BASELINE.json configs (1-based ids as in SURVEY 8(d)); config 1 is the CPU-runnable parity case (tests only)
CONFIGS = {
    2: dict(dims=[8, 8, 1], N=1000, M=100, S=20, kern='rbf', n_classes=0, step='adam', num_data=8192, seed=2000,
            name="BASELINE configs[1]: 2-layer RBF DGP N=1000 M=100 S=20 dims 8-8-1"),
    3: dict(dims=[8, 8, 8, 8, 8, 1], N=1000, M=100, S=20, kern='rbf', n_classes=0, step='adam', num_data=8192, seed=3000,
            name="BASELINE configs[2]: 5-layer RBF DGP N=1000 M=100 S=20 dims 8-8-8-8-8-1"),
    4: dict(dims=[9, 9, 9, 1], N=4096, M=512, S=32, kern='matern52', n_classes=0, step='natgrad', num_data=45730, seed=4000,
            name="BASELINE configs[3]: 3-layer Matern52 DGP N=4096 M=512 S=32 dims 9-9-9-1, natural-gradient step on the final q(U)"),
    5: dict(dims=[784, 30, 10], N=1000, M=100, S=10, kern='rbf', n_classes=10, step='adam', num_data=60000, seed=5000,
            name="BASELINE configs[4]: 2-layer MNIST-shape multiclass DGP N=1000 M=100 S=10 dims 784-30-10"),
}
CFG_ID = 3
WORKLOAD = dict(dims=[8, 8, 8, 8, 8, 1], N=1000, M=100, S=20)
NUM_DATA = 8192          # kin8nm size (demos/datasets.py:133)
METRIC = "ELBO training steps/sec (N=1000,M=100,S=20,L=5; fwd+bwd+Adam)"


def select_config(cid):
    global CFG_ID, WORKLOAD, NUM_DATA, METRIC
    c = CONFIGS[cid]
    CFG_ID = cid
    WORKLOAD = dict(dims=c['dims'], N=c['N'], M=c['M'], S=c['S'])
    NUM_DATA = c['num_data']
    if cid != 3:
        METRIC = (f"ELBO training steps/sec (N={c['N']},M={c['M']},S={c['S']},L={len(c['dims']) - 1}; fwd+bwd+"
                  f"{'NatGrad' if c['step'] == 'natgrad' else 'Adam'})")


def config_dict():
    return {"workload": CONFIGS[CFG_ID]['name'], **WORKLOAD}


def algorithmic_flops(dims, N, M, S, white=False):
    """dense-GEMM accounting of the reference's arithmetic as written, layer 1 on N rows.
    Returns (per-layer forward flops list, fixed flops, step flops = 3*(fwd+fixed))."""
    cw = 1 if white else 2
    fwd, fixed = [], 0.0
    for l in range(len(dims) - 1):
        din, dout = dims[l], dims[l + 1]
        rows = N if l == 0 else N * S
        f = 2 * M * din + cw * M * M + 2 * M * dout + 2 * dout * M * M + 2 * dout * M
        fwd.append(rows * f)
        fixed += 2 * M * M * din + M ** 3 / 3 + 2 * dout * M ** 3 + dout * M ** 3 + 4 * M * M * dout
    return fwd, fixed, 3.0 * (sum(fwd) + fixed)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload():
    from workloads import make_problem
    c = CONFIGS[CFG_ID]
    kw = dict(kern=c['kern'], n_classes=c['n_classes'])
    if CFG_ID == 4:
        kw['max_cond'] = None          # (a 512 x 512 condition-number search is not worth the start-up time; timing only)
    prob = make_problem(seed=c['seed'], num_data=NUM_DATA, **WORKLOAD, **kw)
    if c['step'] == 'natgrad':
        from workloads import well_conditioned_q
        prob = well_conditioned_q(prob)      # a gamma < 1 natural-gradient step needs S = q_sqrt q_sqrt^T numerically PD
    return prob


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle in its reference-faithful form (float64, D_out-tiled temporaries,
# autograd backward, Adam on unconstrained variables) on the host cores.
# ----------------------------------------------------------------------------------------------------------------
def run_cpu_reference(steps, warmup, max_seconds=None):
    """Returns (steps/s of the FULL workload, steps done, seconds, threads).  Config 4 is timed on an S=1 sample of its S=32
    (the per-row work is linear in S; the M^3 work is counted S times, which favours the GPU side slightly less than the
    truth) and scaled; the others run whole steps."""
    import torch
    from oracle.problems import build_oracle
    from oracle import reference_dgp as R
    ncpu = os.cpu_count() or 1
    prob = make_workload()
    scale = 1.0
    if CFG_ID == 4:
        scale = prob['S'] / 1.0
        prob = dict(prob, S=1, zs=[z[:1] for z in prob['zs']])
    o = build_oracle(prob, faithful=True)
    o.num_samples = prob['S']
    st = R.AdamState(o, lr=0.01)
    rng = np.random.default_rng(0)

    def one():
        zs = [rng.normal(size=(prob['S'], prob['N'], lay['dout'])) for lay in prob['layers']]   # tf.random_normal
        return st.step(zs=zs)

    # "all the host threads it can use": the graph is many medium-sized float64 ops, which stop scaling (and then get
    # slower) well before 100+ threads -- probe a few thread counts with one step each and keep the fastest.
    best, cores = None, 1
    for nt in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(nt)
        one()
        t0 = time.perf_counter()
        one()
        dt1 = time.perf_counter() - t0
        if best is None or dt1 < best:
            best, cores = dt1, nt
        if dt1 > 3.0 * best:
            break
    torch.set_num_threads(cores)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    done, per_step = 0, []
    for _ in range(steps):
        t1 = time.perf_counter()
        one()
        per_step.append(time.perf_counter() - t1)
        done += 1
        if max_seconds and time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    run_cpu_reference.last_stats = {"median_ms": 1e3 * scale * float(np.median(per_step)), "min_ms": 1e3 * scale * float(np.min(per_step))}
    run_cpu_reference.sample_note = "" if scale == 1.0 else f" (timed on an S=1 sample of S={int(scale)}, time scaled by {int(scale)})"
    return done / (dt * scale), done, dt, cores


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sps, done, dt, cores = run_cpu_reference(args.steps, args.warmup)
    out = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "steps/s", "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1000.0 / sps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(),
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port", **run_cpu_reference.last_stats,
                         "sample": f"{done} steps of the same workload{run_cpu_reference.sample_note}; reference restatement (torch-CPU "
                                   "float64, reference-faithful D_out tiling, autograd, Adam), not TF 1.8"},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------------------------
class Watchdog:
    """Bounds a stalled run: if the bench has not finished after `seconds`, say where it was (rank 0 prints a JSON error line)
    and leave with os._exit, which tears the CUDA context down -- a multi-rank run that stalls in a collective or a spinning
    kernel otherwise holds its GPUs until the caller's own limit.  DSDGP_BENCH_WATCHDOG=<seconds> (0 disables)."""
    def __init__(self, rank, world, seconds=None):
        self.rank, self.world, self.stage = rank, world, "start"
        self.seconds = float(os.environ.get("DSDGP_BENCH_WATCHDOG", "420")) if seconds is None else float(seconds)
        self.timer = None
        if self.seconds > 0:
            self.timer = threading.Timer(self.seconds, self._fire)
            self.timer.daemon = True
            self.timer.start()

    def at(self, stage):
        self.stage = stage

    def _fire(self):
        msg = f"bench watchdog: not finished after {self.seconds:.0f} s (rank {self.rank} of {self.world} was in stage '{self.stage}')"
        sys.stderr.write(msg + "\n")
        sys.stderr.flush()
        if self.rank == 0:
            print(json.dumps({"error": msg, "n_gpus": self.world, "stage": self.stage}), flush=True)
        os._exit(3)

    def done(self):
        if self.timer is not None:
            self.timer.cancel()


def main_b200(args):
    import torch
    import torch.distributed as dist
    from doubly_stochastic_dgp import _lib
    from workloads import build_model

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    dog = Watchdog(rank, world)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    K, W = args.steps, max(3, args.warmup)
    cfg = CONFIGS[CFG_ID]
    prob = make_workload()
    N, S = prob['N'], prob['S']
    natgrad = cfg['step'] == 'natgrad'
    Dy = 1 if cfg['n_classes'] else cfg['dims'][-1]
    if N % world:
        raise SystemExit("N must divide by the number of GPUs")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def new_model(with_comm):
        m = build_model(prob, device=local_rank)
        if with_comm and world > 1:
            ids = [_lib.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            m.comm_init(ids[0], rank, world)
        return m

    def shard_options(ctx, strong, N_loc):
        if world > 1:
            if strong:
                ctx.set_option("n_global", N); ctx.set_option("n_offset", rank * N_loc)
            else:
                ctx.set_option("n_global", N); ctx.set_option("n_offset", 0)
                ctx.set_option("s_world", world); ctx.set_option("s_offset", rank * S)

    def minibatch_pool(n, strong, N_loc):
        """a pool of different minibatches: pinned host copies (e2e leg) and device-resident copies (value leg);
        same pool on every rank, strong mode takes this rank's rows"""
        rng = np.random.default_rng(100)
        hostX, hostY, devX, devY = [], [], [], []
        for _ in range(n):
            xf = rng.normal(size=(N, WORKLOAD['dims'][0])).astype(np.float32)
            if cfg['n_classes']:
                yf = rng.integers(0, cfg['n_classes'], size=(N, 1)).astype(np.float32)
            else:
                yf = (np.sin(xf.sum(1, keepdims=True)) + 0.1 * rng.normal(size=(N, 1))).astype(np.float32)
                yf = np.tile(yf, (1, Dy))
            lo = rank * N_loc if strong else 0
            x = torch.from_numpy(np.ascontiguousarray(xf[lo:lo + N_loc])).pin_memory()
            y = torch.from_numpy(np.ascontiguousarray(yf[lo:lo + N_loc])).pin_memory()
            hostX.append(x); hostY.append(y)
            devX.append(x.cuda()); devY.append(y.cuda())
        return hostX, hostY, devX, devY

    def run_mode(mode, want_profile, want_e2e):
        """mode 'strong': the BASELINE problem itself -- S_total fixed, the S*N sample rows sharded (N/world minibatch rows x all
        S per GPU), one all-reduce of [grad || ELBO]; value = steps/s of that problem.
        mode 'weak': every GPU evaluates S samples of the SAME minibatch -- the S-shards of one ELBO with S_total = S*world are
        all-reduced; per-GPU work fixed; value counts (N, S) evaluations/s (= steps/s x world)."""
        strong = mode == "strong"
        N_loc = N // world if strong else N
        m = new_model(True)
        ctx = m._ensure_ctx(N_loc, S)
        if not natgrad:
            m.adam_init(0.01)
        shard_options(ctx, strong, N_loc)
        POOL = 8 if CFG_ID != 4 else 2
        hostX, hostY, devX, devY = minibatch_pool(POOL, strong, N_loc)
        last = len(cfg['dims']) - 2

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ctx.sync()
            torch.cuda.synchronize()

        def dev_step(i, sync):
            j = i % POOL
            if natgrad:     # one NatGradOptimizer(gamma).minimize(maxiter=1) on the final layer (always synchronous)
                return ctx.natgrad_step(devX[j].data_ptr(), devY[j].data_ptr(), N_loc, S, NUM_DATA, 1000 + i, [last], 0.1,
                                        flags=_lib.FLAG_DEVICE_PTRS)
            return ctx.train_step(devX[j].data_ptr(), devY[j].data_ptr(), N_loc, S, NUM_DATA, 1000 + i,
                                  flags=_lib.FLAG_DEVICE_PTRS | (0 if sync else _lib.FLAG_NO_SYNC), want_elbo=sync)

        for i in range(W):
            dev_step(i, True)
        # ---- value leg: device-resident inputs, per-step CUDA events on the ctx stream, L2 flushed between steps
        clocks = ClockSampler(local_rank)
        launches0 = ctx.launch_count()
        barrier()
        clocks.start()
        tot_ms = 0.0
        for i in range(K):
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
            dev_step(W + i, False)
            tot_ms += ctx.last_step_ms()
        barrier()
        launches = ctx.launch_count() - launches0
        tot_ms = max_over_ranks(tot_ms)
        barrier()
        ctx.timer_start()
        for i in range(K):
            dev_step(W + K + i, False)
        b2b_ms = max_over_ranks(ctx.timer_stop())
        barrier()
        # keep the clock sampler running over a longer loaded stretch (the timed legs above last only ~0.1 s)
        # (a FIXED number of steps, the same on every rank -- about one second's worth by the max-over-ranks step time above:
        # every step carries an all-reduce, so a wall-clock-bounded loop would let ranks enqueue different numbers of
        # collectives and hang the slower ones)
        n_extra = max(2, min(20000, int(1000.0 / max(tot_ms / K, 1e-3))))
        for i in range(n_extra):
            dev_step(9000 + i, False)
            if (i + 1) % 32 == 0:
                ctx.sync()
        barrier()
        clk = clocks.stop()
        units = 1 if strong else world        # weak: every rank does a full (N, S) evaluation per step
        res = {"value": units * K / (tot_ms / 1e3), "ms_per_step": tot_ms / K, "ms_per_step_back_to_back": b2b_ms / K,
               "gpu_launches": int(launches), "clocks": clk, "rows_per_gpu": N_loc * S, "N_loc": N_loc, "e2e": None,
               "stage_ms": None, "roofline": None}
        # ---- e2e leg: the public Python API with host (pinned) buffers; H2D of the minibatch + D2H of the ELBO per step
        if want_e2e:
            Xh = [x.numpy() for x in hostX]
            Yh = [y.numpy() for y in hostY]
            api_step = (lambda x, y: m.natgrad_step(gamma=0.1, X=x, Y=y)) if natgrad else (lambda x, y: m.train_step(x, y))
            for i in range(3):
                api_step(Xh[i % POOL], Yh[i % POOL])
            barrier()
            t0 = time.perf_counter()
            for i in range(K):
                api_step(Xh[i % POOL], Yh[i % POOL])
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            res["e2e"] = {"value": units * K / dt, "unit": "steps/s",
                          "h2d_bytes_per_step": int(Xh[0].nbytes + Yh[0].nbytes), "d2h_bytes_per_step": 16,
                          "timing": "host wall clock around K public-API calls (model.train_step / natgrad_step), each returning the ELBO"}
        if want_profile:
            try:
                profile_leg(ctx, dev_step, res, N_loc, tot_ms)
            except Exception as ex:      # an auxiliary leg must not cost the headline line
                res["roofline_error"] = f"{type(ex).__name__}: {ex}"
        m._ctx.close()
        return res

    def profile_leg(ctx, dev_step, res, N_loc, tot_ms):
        """Per-stage device times (eager launches bracketed by CUDA events) -> dominant kernel and its roofline entry."""
        ctx.set_option("profile", 1)
        acc = None
        reps = 5 if CFG_ID != 4 else 2
        for i in range(reps + 1):
            dev_step(5000 + i, True)
            p = np.array(ctx.profile())
            if i > 0:
                acc = p if acc is None else acc + p
        ctx.set_option("profile", 0)
        p = acc / reps
        L = len(WORKLOAD['dims']) - 1
        names = ["prep(Kuu,chol,KL)", "likelihood", "grad-assembly", "allreduce", "tail(result+adam)"]
        for l in range(L):
            names += [f"layer{l + 1}.fwd", f"layer{l + 1}.bwd_rows", f"layer{l + 1}.rowred"]
        fwd_fl, fixed, step_fl = algorithmic_flops(WORKLOAD['dims'], N_loc, WORKLOAD['M'], S)
        chained = L > 1 and all(p[5 + 3 * l] == 0 for l in range(1, L))     # persistent kernel: all layers' forward in one launch
        if chained:
            names[5] = "fwd_chain(all layers, one persistent launch)"
            fwd_fl = list(fwd_fl)
            fwd_fl[0] = float(sum(fwd_fl))
        res["stage_ms"] = {n: round(float(v), 4) for n, v in zip(names, p)}
        top = int(np.argmax(p[5:])) + 5
        l = (top - 5) // 3
        peaks, how = measured_peaks()
        peak_tf32 = peaks["bf16_tflops"] / 2.0
        ach = fwd_fl[l] / (p[top] * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
        if os.path.exists(tpath) and CFG_ID == 3:
            with open(tpath) as f:
                traffic = json.load(f).get("fwd_chain" if names[top].startswith("fwd_chain") else names[top].split(".")[1])
        res["roofline"] = {
            "bound": "tensor", "kernel": names[top], "achieved": ach, "peak": peak_tf32, "unit": "TFLOP/s",
            "frac": ach / peak_tf32, "traffic": traffic,
            "note": f"algorithmic flops/launch = rows*f(l) = {fwd_fl[l]:.4g} (SURVEY 8(d): each of forward, row-backward "
                    f"and row-reduction kernels of a layer carries rows*f(l)); kernel time from CUDA events around the "
                    f"launch in an eager (non-graph) pass; peak = {how} bf16_tflops/2 = TF32-dense equivalent (the tcgen05 "
                    "kernels issue kind::tf32, ~1.3-1.4x the algorithmic MMA work because of the 3xTF32 stages; shapes outside "
                    "their range run the fp32 SIMT kernels, see DESIGN.md); traffic = dram read+write bytes per launch from the "
                    "ncu --set full capture summarised in profiles/",
            "step_achieved": step_fl / (tot_ms / K * 1e-3) / 1e12, "step_algorithmic_gflop": step_fl / 1e9}

    def parity_leg():
        """ELBO of the sharded evaluation (all-reduced inside the step) vs ONE GPU evaluating the same global minibatch with the
        same Philox seed (draws are keyed by global (s, n): shard-invariant), and -- at N=1 -- vs the float64 oracle on injected
        draws.  BASELINE.md section 3: parity gates the speed number."""
        out = {"tolerance_rel": 1e-4}
        strong = args.scaling == "strong"
        N_loc = N // world if strong else N
        _, _, devX, devY = minibatch_pool(1, strong, N_loc)
        if world > 1:
            m = new_model(True)
            ctx = m._ensure_ctx(N_loc, S)
            shard_options(ctx, strong, N_loc)
            e_sharded = _elbo_dev(ctx, devX[0], devY[0], N_loc, S, 777)
            m._ctx.close()
            if rank == 0:
                m1 = build_model(prob, device=local_rank)
                S1 = S if strong else S * world
                ctx1 = m1._ensure_ctx(N, S1)
                _, _, fX, fY = minibatch_pool(1, False, N)
                e_single = _elbo_dev(ctx1, fX[0], fY[0], N, S1, 777)
                m1._ctx.close()
                out["sharded_vs_single_gpu"] = {"elbo_sharded": e_sharded, "elbo_single": e_single,
                                                "rel_err": abs(e_sharded - e_single) / abs(e_single)}
        if rank == 0 and world == 1 and not args.no_cpu:
            from oracle.problems import build_oracle
            from workloads import round_f32
            p32 = round_f32(prob)
            if CFG_ID == 4:      # one sample shard of the full-size problem (the float64 oracle at S=32 needs > 100 GB)
                p32 = dict(p32, S=1, zs=[z[:1] for z in p32['zs']])
            m = build_model(p32, device=local_rank)
            e_dev = m.compute_log_likelihood(zs=p32['zs'])
            m._ctx.close()
            o = build_oracle(p32)
            o.num_samples = p32['S']
            e_ref = float(o.compute_log_likelihood(zs=p32['zs']))
            out["vs_oracle"] = {"elbo_device": e_dev, "elbo_oracle_f64": e_ref, "rel_err": abs(e_dev - e_ref) / abs(e_ref),
                                "sample": "full workload, injected draws" if CFG_ID != 4 else "S=1 shard of the full-size workload, injected draws"}
        return out

    def _elbo_dev(ctx, xd, yd, n, s, seed):
        import ctypes as C
        e = C.c_double()
        _lib.check(ctx.lib.dsdgp_elbo(ctx.h, C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr()), n, s, float(NUM_DATA), None,
                                      seed, _lib.FLAG_DEVICE_PTRS, C.byref(e)))
        return e.value

    dog.at("parity leg")
    parity = parity_leg()
    dog.at(f"{args.scaling} leg (value, e2e, stage profile)")
    primary = run_mode(args.scaling, want_profile=True, want_e2e=not args.no_e2e)
    other = None
    if world > 1 and not natgrad:
        other_mode = "strong" if args.scaling == "weak" else "weak"
        dog.at(f"{other_mode} leg (other_scaling)")
        o = run_mode(other_mode, want_profile=False, want_e2e=False)
        other = {"scaling": other_mode, "value": o["value"], "ms_per_step": o["ms_per_step"], "rows_per_gpu": o["rows_per_gpu"],
                 "note": "weak: S=20 per GPU, S_total = 20 x n_gpus; value counts (N, S) evaluations/s" if other_mode == "weak"
                         else "strong: S_total fixed, rows sharded"}

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only), bounded sample
    cpu = None
    dog.at("cpu baseline")
    if rank == 0 and world == 1 and not args.no_cpu:
        sps, done, dt, cores = run_cpu_reference(steps=args.cpu_steps, warmup=1, max_seconds=25)
        cpu = {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port", **run_cpu_reference.last_stats,
               "sample": f"{done} steps ({dt:.1f} s) of the same workload{run_cpu_reference.sample_note}: reference restatement "
                         "(torch-CPU float64, reference-faithful tiling, autograd backward, Adam), not TF 1.8"}

    bad = [k for k, v in parity.items() if isinstance(v, dict) and not (v["rel_err"] <= parity["tolerance_rel"])]
    if rank == 0:
        weak = args.scaling == "weak"
        out = {
            "metric": METRIC, "value": primary["value"], "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": primary["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "tf32", "data": "synthetic", "config": config_dict(),
            "run": {"rows_per_gpu": primary["rows_per_gpu"],
                    "parallelism": (f"dp{world}: S sharded -- every GPU draws S={S} samples of the N={N} minibatch, S_total={S * world}, "
                                    f"one NCCL all-reduce of [grad || ELBO]; value counts (N={N},S={S}) evaluations/s"
                                    if weak else
                                    f"dp{world}: S_total={S} fixed, the S*N sample rows sharded ({primary['N_loc']} minibatch rows x all S per GPU), "
                                    "one NCCL all-reduce of [grad || ELBO]"),
                    "l2": "flushed (256 MiB write) between timed steps; per-step CUDA events on the launch stream",
                    "precision": "tcgen05 kind::tf32 (3xTF32 for the whitened projections / solves, weights-split 2xTF32 for the variance product once q_sqrt is non-negligible, 1xTF32 elsewhere), "
                                 "fp32 epilogues, fp64 MxM factorisation + KL"},
            "ms_per_step_back_to_back": primary["ms_per_step_back_to_back"], "gpu_launches": primary["gpu_launches"],
            "clocks": primary["clocks"], "e2e": primary["e2e"], "roofline": primary["roofline"], "cpu_baseline": cpu,
            "stage_ms": primary["stage_ms"], "other_scaling": other, "parity": parity,
        }
        if primary.get("roofline_error"):
            out["roofline_error"] = primary["roofline_error"]
        print(json.dumps(out))
    dog.at("final barrier")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    dog.done()
    if bad:
        raise SystemExit(f"parity gate failed: {bad} beyond {parity['tolerance_rel']}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=20)
    a = ap.parse_args()
    select_config(a.config)
    if a.config == 4:
        a.steps = min(a.steps, 10); a.cpu_steps = min(a.cpu_steps, 3)      # (a step is ~100x the north-star's)
    if a.impl == "reference":
        if a.steps > 30:
            a.steps = 30          # each step is a bounded sample: one full CPU step (~0.5 s); keep the run to minutes
        main_reference(a)
    else:
        main_b200(a)
