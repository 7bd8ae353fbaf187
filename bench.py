#!/usr/bin/env python
"""bench.py -- ELBO training steps/sec of the doubly-stochastic DGP hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference restatement (torch-CPU float64) on host cores

Workload = BASELINE.json configs[2] (the config the metric is quoted on): 5-layer RBF DGP, kin8nm shape,
N=1000 minibatch rows, M=100 inducing points, S=20 samples, dims 8->8->8->8->8->1, synthetic data.
One step = minibatch in + ELBO forward + full backward + Adam update (= one session.run(minimize_op) of the
reference, demos/run_regression.py:83,138).

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

WORKLOAD = dict(dims=[8, 8, 8, 8, 8, 1], N=1000, M=100, S=20)
NUM_DATA = 8192          # kin8nm size (demos/datasets.py:133)
METRIC = "ELBO training steps/sec (N=1000,M=100,S=20,L=5; fwd+bwd+Adam)"


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


def make_workload(seed=3000):
    from tests.synth import make_problem
    return make_problem(seed=seed, num_data=NUM_DATA, **WORKLOAD)


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
    import torch
    from tests.synth import build_oracle
    from oracle import reference_dgp as R
    ncpu = os.cpu_count() or 1
    prob = make_workload()
    o = build_oracle(prob, faithful=True)
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
    run_cpu_reference.last_stats = {"median_ms": 1e3 * float(np.median(per_step)), "min_ms": 1e3 * float(np.min(per_step))}
    return done / dt, done, dt, cores


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sps, done, dt, cores = run_cpu_reference(args.steps, args.warmup)
    out = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "steps/s", "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / done, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: 5-layer RBF DGP N=1000 M=100 S=20 dims 8-8-8-8-8-1", **WORKLOAD},
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port", **run_cpu_reference.last_stats,
                         "sample": f"{done} full steps of the same workload; reference restatement (torch-CPU float64, "
                                   "reference-faithful D_out tiling, autograd, Adam), not TF 1.8"},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------------------------
def main_b200(args):
    import torch
    import torch.distributed as dist
    from doubly_stochastic_dgp import _lib
    from tests.gpu_common import build_model

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    K, W = args.steps, max(3, args.warmup)
    prob = make_workload()
    N, S = prob['N'], prob['S']
    if N % world:
        raise SystemExit("N must divide by the number of GPUs")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    uid = [None]

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run_mode(mode, want_profile, want_e2e):
        """mode 'weak': every GPU evaluates S=20 samples of the SAME (N=1000) minibatch -- the S-shards of one ELBO with
        S_total = 20*world are all-reduced (BASELINE: "the S Monte-Carlo samples shard across the GPUs"); per-GPU work fixed.
        mode 'strong': S_total = 20 fixed; the S*N sample rows are sharded (N/world minibatch rows x all S per GPU)."""
        strong = mode == "strong"
        N_loc = N // world if strong else N
        m = build_model(prob, device=local_rank)
        if world > 1:
            ids = [_lib.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            m.comm_init(ids[0], rank, world)
        ctx = m._ensure_ctx(N_loc, S)
        m.adam_init(0.01)
        if world > 1:
            if strong:
                ctx.set_option("n_global", N); ctx.set_option("n_offset", rank * N_loc)
            else:
                ctx.set_option("n_global", N); ctx.set_option("n_offset", 0)
                ctx.set_option("s_world", world); ctx.set_option("s_offset", rank * S)
        # a pool of different minibatches: pinned host copies (e2e leg) and device-resident copies (value leg)
        POOL = 8
        rng = np.random.default_rng(100)              # same pool on every rank; strong mode takes this rank's rows
        hostX, hostY, devX, devY = [], [], [], []
        for _ in range(POOL):
            xf = rng.normal(size=(N, WORKLOAD['dims'][0])).astype(np.float32)
            yf = (np.sin(xf.sum(1, keepdims=True)) + 0.1 * rng.normal(size=(N, 1))).astype(np.float32)
            lo = rank * N_loc if strong else 0
            x = torch.from_numpy(np.ascontiguousarray(xf[lo:lo + N_loc])).pin_memory()
            y = torch.from_numpy(np.ascontiguousarray(yf[lo:lo + N_loc])).pin_memory()
            hostX.append(x); hostY.append(y)
            devX.append(x.cuda()); devY.append(y.cuda())

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ctx.sync()
            torch.cuda.synchronize()

        def dev_step(i, sync):
            j = i % POOL
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
        t_end = time.perf_counter() + 1.0
        i = 0
        while time.perf_counter() < t_end:
            dev_step(9000 + i, False); i += 1
            if i % 32 == 0:
                ctx.sync()
        barrier()
        clk = clocks.stop()
        units = 1 if strong else world        # weak: every rank does a full (N=1000, S=20) evaluation per step
        res = {"value": units * K / (tot_ms / 1e3), "ms_per_step": tot_ms / K, "ms_per_step_back_to_back": b2b_ms / K,
               "gpu_launches": int(launches), "clocks": clk, "rows_per_gpu": N_loc * S, "N_loc": N_loc, "e2e": None,
               "stage_ms": None, "roofline": None}
        # ---- e2e leg: the public Python API with host (pinned) buffers; H2D of the minibatch + D2H of the ELBO per step
        if want_e2e:
            Xh = [x.numpy() for x in hostX]
            Yh = [y.numpy() for y in hostY]
            for i in range(3):
                m.train_step(Xh[i % POOL], Yh[i % POOL])
            barrier()
            t0 = time.perf_counter()
            for i in range(K):
                m.train_step(Xh[i % POOL], Yh[i % POOL])
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            res["e2e"] = {"value": units * K / dt, "unit": "steps/s",
                          "h2d_bytes_per_step": int(Xh[0].nbytes + Yh[0].nbytes), "d2h_bytes_per_step": 16,
                          "timing": "host wall clock around K public-API calls (model.train_step), each returning the ELBO"}
        if want_profile:
            try:
                profile_leg(ctx, dev_step, res, N_loc, tot_ms)
            except Exception as ex:      # an auxiliary leg must not cost the headline line
                res["roofline_error"] = f"{type(ex).__name__}: {ex}"
        m._ctx.close()
        return res

    def profile_leg(ctx, dev_step, res, N_loc, tot_ms):
        """Per-stage device times (eager launches bracketed by CUDA events) -> dominant kernel and its roofline entry."""
        # ---- per-stage profile (eager launches bracketed by events) -> dominant kernel and its roofline
        ctx.set_option("profile", 1)
        acc = None
        reps = 5
        for i in range(reps + 1):
            dev_step(5000 + i, True)
            p = np.array(ctx.profile())
            if i > 0:
                acc = p if acc is None else acc + p
        ctx.set_option("profile", 0)
        p = acc / reps
        L = len(WORKLOAD['dims']) - 1
        names = ["prep(Kuu,chol,KL)", "likelihood", "grad-assembly", "allreduce", "adam"]
        for l in range(L):
            names += [f"layer{l + 1}.fwd", f"layer{l + 1}.bwd_rows", f"layer{l + 1}.rowred"]
        fwd_fl, fixed, step_fl = algorithmic_flops(WORKLOAD['dims'], N_loc, WORKLOAD['M'], S)
        chained = all(p[5 + 3 * l] == 0 for l in range(1, L))     # persistent kernel: all layers' forward in one launch
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
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("fwd_chain" if names[top].startswith("fwd_chain") else names[top].split(".")[1])
        res["roofline"] = {
            "bound": "tensor", "kernel": names[top], "achieved": ach, "peak": peak_tf32, "unit": "TFLOP/s",
            "frac": ach / peak_tf32, "traffic": traffic,
            "note": f"algorithmic flops/launch = rows*f(l) = {fwd_fl[l]:.4g} (SURVEY 8(d): each of forward, row-backward "
                    f"and row-reduction kernels of a layer carries rows*f(l)); kernel time from CUDA events around the "
                    f"launch in an eager (non-graph) pass; peak = {how} bf16_tflops/2 = TF32-dense equivalent (the kernel "
                    "issues tcgen05 kind::tf32, ~1.3-1.4x the algorithmic MMA work because of the 3xTF32 stages); "
                    "traffic = dram read+write bytes per launch from the ncu --set full capture summarised in profiles/",
            "step_achieved": step_fl / (tot_ms / K * 1e-3) / 1e12, "step_algorithmic_gflop": step_fl / 1e9}

    primary = run_mode(args.scaling, want_profile=True, want_e2e=not args.no_e2e)
    other = None
    if world > 1:
        other_mode = "strong" if args.scaling == "weak" else "weak"
        o = run_mode(other_mode, want_profile=False, want_e2e=False)
        other = {"scaling": other_mode, "value": o["value"], "ms_per_step": o["ms_per_step"], "rows_per_gpu": o["rows_per_gpu"]}

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only), bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sps, done, dt, cores = run_cpu_reference(steps=args.cpu_steps, warmup=1, max_seconds=25)
        cpu = {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port", **run_cpu_reference.last_stats,
               "sample": f"{done} full steps ({dt:.1f} s) of the same workload: reference restatement (torch-CPU float64, "
                         "reference-faithful tiling, autograd backward, Adam), not TF 1.8"}

    if rank == 0:
        weak = args.scaling == "weak"
        out = {
            "metric": METRIC, "value": primary["value"], "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": primary["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "tf32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: 5-layer RBF DGP N=1000 M=100 S=20 dims 8-8-8-8-8-1", **WORKLOAD,
                       "rows_per_gpu": primary["rows_per_gpu"],
                       "parallelism": (f"dp{world}: S sharded -- every GPU draws S=20 samples of the N=1000 minibatch, S_total={S * world}, "
                                       "one NCCL all-reduce of [grad || ELBO]; value counts (N=1000,S=20) evaluations/s"
                                       if weak else
                                       f"dp{world}: S_total=20 fixed, the S*N sample rows sharded ({primary['N_loc']} minibatch rows x all S per GPU), "
                                       "one NCCL all-reduce of [grad || ELBO]"),
                       "l2": "flushed (256 MiB write) between timed steps; per-step CUDA events on the launch stream",
                       "precision": "tcgen05 kind::tf32 (3xTF32 for the whitened projections / solves, 1xTF32 elsewhere), "
                                    "fp32 epilogues, fp64 MxM factorisation + KL"},
            "ms_per_step_back_to_back": primary["ms_per_step_back_to_back"], "gpu_launches": primary["gpu_launches"],
            "clocks": primary["clocks"], "e2e": primary["e2e"], "roofline": primary["roofline"], "cpu_baseline": cpu,
            "stage_ms": primary["stage_ms"], "other_scaling": other,
        }
        if primary.get("roofline_error"):
            out["roofline_error"] = primary["roofline_error"]
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["strong", "weak"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=20)
    a = ap.parse_args()
    if a.impl == "reference":
        if a.steps > 30:
            a.steps = 30          # each step is a bounded sample: one full CPU step (~0.5 s); keep the run to minutes
        main_reference(a)
    else:
        main_b200(a)
