#!/usr/bin/env python
"""bench.py -- NSF training pairs/s (headline) and posterior samples/s / log_prob/s on the BASELINE.json workloads.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the restated sbi path (oracle/) on the host CPU cores

A "step" is one optimiser step of ``SNPE_C.train`` (sbi PosteriorEstimator.train, /root/reference/floppity.py:219) on
one batch: gather the batch from the HBM-resident dataset -> flow forward -> loss -> hand-written backward ->
[NCCL all-reduce] -> clip_grad_norm_(5) + Adam.

Headline workload (``value``): BASELINE.json configs[2] (c3: NSF T15 K16 H512 Bk4, 30 parameters / 2000-bin spectra) --
the largest configuration that fits one GPU -- with the maximum-likelihood loss (configs[2] names no atoms, and it is the
branch the reference takes on every round: force_first_round_loss=True, floppity.py:220).  configs[1] (c2, atoms 10),
configs[0] (c1, the reference-default hidden 50) and configs[3] (c4, atoms 20) ride along as ``extras`` with their own
roofline blocks; configs[4] (sampling / log_prob from a trained flow, sharded over the ranks) gives
``posterior_samples_per_s`` and ``log_prob_per_s``.  Synthetic spectra (ARCiS is external Fortran), random-init
weights.  Weak scaling: the per-GPU batch is fixed.  ``e2e`` goes through the public call a user makes --
``SNPE_C.append_simulations(pinned host tensors)`` + ``SNPE_C.train()`` (upload, epochs, validation, returned copy).
One JSON line on stdout (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# per-config bench settings: loss, per-GPU batch, resident rows per GPU
#   c3: 18 944 rows = 148 row tiles of 128 -> whole waves of the 128x128 GEMM tiles on 148 SMs
#   c1 / c2 / c4: 9 472 pairs x atoms = 740 / 1480 row tiles = 5 / 10 per SM
WORKLOADS = {
    "c1": dict(loss="atomic", batch=9472, rows=65536),
    "c2": dict(loss="atomic", batch=9472, rows=65536),
    "c3": dict(loss="mle", batch=18944, rows=65536),
    "c4": dict(loss="atomic", batch=9472, rows=32768),
}
CONFIG_INDEX = {"c1": 0, "c2": 1, "c3": 2, "c4": 3}
# operand format of each tensor-core kernel: passes of the fp32-accurate split and the rate of one pass relative to bf16
TENSOR_KERNELS = {"k_gemm_tc_nt": ("3xFP16", 3.0), "k_gemm_tc_tn": ("3xFP16", 3.0), "k_cond_fwd": ("3xFP16", 3.0),
                  "k_cond_small_fwd": ("3xFP16 mma.sync", 3.0), "k_cond_small_bwd": ("3xFP16 mma.sync", 3.0),
                  "k_small_wgrad": ("3xFP16 mma.sync", 3.0)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c4"], help="headline workload")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (0 = the workload's default)")
    ap.add_argument("--loss", default="", choices=["", "atomic", "mle"])
    ap.add_argument("--rows", type=int, default=0, help="resident dataset rows per GPU (0 = the workload's default)")
    ap.add_argument("--ref-batch", type=int, default=0, help="bounded per-step sample of the CPU arm (0 = sized per config)")
    ap.add_argument("--dropout", type=float, default=0.0)
    ap.add_argument("--graph", type=int, default=0, help="replay the optimiser step from a CUDA graph")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--only-sampling", action="store_true",
                    help="evidence runs: only the configs[4] leg (posterior samples/s, log_prob/s); the headline keys are null")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_data(cfg_name, rows, seed, base_rows=4096):
    """Synthetic (theta, x) rows of the named shape.  The spectrum generator is evaluated for ``base_rows`` distinct
    parameter vectors; the resident set is those spectra tiled with FRESH noise draws (what -naug does in the reference,
    floppityFUN.py:383-428): throughput does not depend on the content, and the generator's exp() cost stays bounded."""
    from floppity_deprecated_b200 import synthetic
    cfg = synthetic.CONFIGS[cfg_name]
    rng = np.random.default_rng(seed)
    nb = min(rows, base_rows)
    theta = synthetic.swap_hotter_first(cfg, rng.uniform(-3, 3, (nb, cfg.D)))   # prior draws in the -ynorm box
    spec = synthetic.simulate(cfg, theta)
    sigma = 0.01 * np.median(spec, axis=0)
    reps = (rows + nb - 1) // nb
    th = np.tile(theta, (reps, 1))[:rows].astype(np.float32)
    x = np.tile(spec.astype(np.float32), (reps, 1))[:rows]
    x += (sigma[None, :] * rng.standard_normal((rows, cfg.C), dtype=np.float32)).astype(np.float32)
    mean, std = x.mean(axis=0), x.std(axis=0)
    std[std == 0] = 1.0
    x = (x - mean) / std
    return cfg, torch.from_numpy(th), torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))


def workload_desc(name, cfg, loss, B, rows, world, dropout):
    """The ``config`` object: identical in both arms (``--impl reference`` prints the same one)."""
    atoms = f", atoms {cfg.atoms}" if loss == "atomic" else ""
    return {"workload": f"{name}: NSF T{cfg.transforms} K{cfg.bins} H{cfg.hidden} Bk{cfg.blocks}, {cfg.D} params / {cfg.C} bins"
                        f"{atoms} (BASELINE.json configs[{CONFIG_INDEX[name]}])",
            "loss": loss + (" (SNPE-C APT)" if loss == "atomic" else " (force_first_round_loss=True, floppity.py:220)"),
            "per_gpu_batch": B, "global_batch": B * world, "resident_rows_per_gpu": rows, "dropout": dropout,
            "parallelism": f"dp{world}",
            "l2": "no flush: the per-step activation workspace (GBs) is far larger than the 126 MB L2"}


def settings(args, name):
    w = dict(WORKLOADS[name])
    if name == args.config:
        if args.batch:
            w["batch"] = args.batch
        if args.loss:
            w["loss"] = args.loss
        if args.rows:
            w["rows"] = args.rows
    return w


# ------------------------------------------------------------------------------------------------- reference arm
def oracle_steps(cfg, loss, Bc, theta, x, warmup, steps=None, seconds=None, dropout=0.0):
    """Optimiser steps of the restated sbi path (oracle/nsf_oracle.py, eager fp32 PyTorch) on all host cores; returns
    (pairs/s, steps timed, seconds)."""
    from oracle import nsf_oracle as O
    net = O.build_nsf(theta[:2048], x[:2048], hidden_features=cfg.hidden, num_transforms=cfg.transforms, num_bins=cfg.bins,
                      num_blocks=cfg.blocks, dropout_probability=dropout)
    prior = O.BoxUniform(-3 * torch.ones(cfg.D), 3 * torch.ones(cfg.D))
    opt = torch.optim.Adam(net.parameters(), lr=5e-4)
    masks = torch.zeros(Bc, 1)
    net.train()

    def step():
        idx = torch.randint(0, theta.shape[0], (Bc,))
        opt.zero_grad()
        lp = (O.log_prob_proposal_posterior_atomic(net, prior, theta[idx], x[idx], masks, cfg.atoms, True)
              if loss == "atomic" else net.log_prob(theta[idx], x[idx]))
        (-lp.mean()).backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 5.0)
        opt.step()

    for _ in range(warmup):
        step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        dt = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (seconds is not None and dt > seconds and n >= 2):
            break
    return Bc * n / dt, n, dt


def ref_batch_for(args, cfg):
    if args.ref_batch:
        return args.ref_batch
    return 64 if cfg.hidden >= 512 else 256


def run_reference(args):
    """Restated sbi path (oracle/, "port": sbi / nflows cannot be installed here) on ALL host cores.  Same config object
    as our arm; each step is a bounded sample (batch --ref-batch) of that workload's per-GPU batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from floppity_deprecated_b200 import synthetic
    w = settings(args, args.config)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, theta, x = make_data(args.config, 4096, 0)
    Bc = ref_batch_for(args, cfg)
    torch.manual_seed(0)
    v, n, dt = oracle_steps(cfg, w["loss"], Bc, theta, x, args.warmup, steps=args.steps, dropout=args.dropout)
    sample = (f"{n} optimiser steps at batch {Bc} (a bounded sample of the per-GPU batch {w['batch']}), {w['loss']} loss, "
              f"oracle/nsf_oracle.py eager fp32 on {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": "NSF train pairs/s", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / n, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_desc(args.config, cfg, w["loss"], w["batch"], w["rows"], max(args.gpus, 1), args.dropout),
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------- our arm
class Ctx:
    pass


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    # stdout carries exactly ONE line (rank 0's JSON): library chatter such as NCCL's version banner goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import floppity_deprecated_b200 as fb
    from floppity_deprecated_b200 import _lib, dist as fdist
    from floppity_deprecated_b200.inference import _TrainEngine
    L = _lib.lib()
    PK = peaks()
    K, W = args.steps, max(args.warmup, 3)
    NT = L.fp_prof_ntags()
    TAGS = [L.fp_prof_tag_name(i).decode() for i in range(NT)]

    def log(msg):
        print(f"[bench r{rank}] {msg}", file=sys.stderr, flush=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        """n calls of fn between two CUDA events on the current stream, barrier + synchronize on both sides, max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    def tensor_fmt(name):
        """(operand format, bf16-rate divisor): passes per fp32-accurate product x the rate of one pass relative to bf16."""
        if name in ("k_gemm_tc_nt", "k_gemm_tc_tn"):
            f16 = L.fp_set_gemm_format(-1) == 1
            if L.fp_set_precision(-1) == 1:
                return ("1xFP16 high pieces (reduced precision)", 1.0) if f16 else ("1xTF32 (reduced precision)", 2.0)
            return ("3xFP16", 3.0) if f16 else ("3xTF32", 6.0)
        return TENSOR_KERNELS[name]

    def kernel_table(fn, n):
        """Per-kernel CUDA-event timing of n calls of fn (events recorded around every launch on the launching stream)."""
        arr = lambda t: (t * NT)()
        pms, pl, pf, pb = arr(ctypes.c_double), arr(ctypes.c_int64), arr(ctypes.c_double), arr(ctypes.c_double)
        L.fp_prof_enable(1)
        L.fp_prof_collect(pms, pl, pf, pb)
        for i in range(n):
            fn(i)
        _lib.check(L.fp_prof_collect(pms, pl, pf, pb), "prof")
        L.fp_prof_enable(0)
        tot = sum(pms) or 1.0
        table = {}
        for i, name in enumerate(TAGS):
            if pl[i] == 0:
                continue
            sec = pms[i] * 1e-3
            d = {"share": pms[i] / tot, "launches_per_step": pl[i] / n, "avg_us": 1e3 * pms[i] / pl[i]}
            if pb[i] > 0:
                d["algorithmic_bytes_per_launch"] = pb[i] / pl[i]
                d["achieved_gbs"] = pb[i] / sec / 1e9
                d["frac_of_hbm_peak"] = d["achieved_gbs"] / PK["hbm"]
            if pf[i] > 0:
                d["algorithmic_flops_per_launch"] = pf[i] / pl[i]
                d["achieved_tflops"] = pf[i] / sec / 1e12
                if name in TENSOR_KERNELS:
                    fmt, div = tensor_fmt(name)
                    d["operands"] = fmt
                    d["frac_of_tensor_ceiling"] = d["achieved_tflops"] / (PK["tf_sust"] / div)
                    d["frac_of_bf16_peak"] = d["achieved_tflops"] / PK["tf_sust"]
            table[name] = d
        return table

    def roofline_of(table):
        """The roofline block of the kernel with the largest share of the step: ONE kernel, its own CUDA-event time."""
        named = {k: v for k, v in table.items() if k != "other"}
        if not named:
            return None
        name = max(named, key=lambda k: named[k]["share"])
        d = named[name]
        hb = d.get("frac_of_hbm_peak", 0.0)
        tc = d.get("frac_of_tensor_ceiling", 0.0)
        if name in TENSOR_KERNELS and tc >= hb:
            fmt, div = tensor_fmt(name)
            roof = {"bound": "tensor", "achieved": d["achieved_tflops"], "peak": PK["tf_sust"] / div, "unit": "TFLOP/s",
                    "frac": tc, "traffic": None, "kernel": name, "share_of_step": d["share"], "avg_us": d["avg_us"],
                    "peak_source": f"{PK['source']}: bf16 sustained {PK['tf_sust']:.0f} TFLOP/s / {div:.0f} -- {fmt} operands: "
                                   f"{'one tensor-core pass' if div <= 2.0 else 'three tensor-core passes per fp32-accurate product'}"
                                   f"{' at the TF32 rate (half of bf16)' if div in (2.0, 6.0) else ' at the fp16 rate (= bf16)'}",
                    "frac_of_bf16_peak": d["frac_of_bf16_peak"], "hbm_frac": hb}
        elif name == "k_gemm":
            fma = 148 * 128 * 2 * 1.965e9 / 1e12
            roof = {"bound": "fma", "achieved": d.get("achieved_tflops", 0.0), "peak": fma, "unit": "TFLOP/s",
                    "frac": d.get("achieved_tflops", 0.0) / fma, "traffic": None, "kernel": name, "share_of_step": d["share"],
                    "avg_us": d["avg_us"], "peak_source": "148 SMs x 128 fp32 lanes x 2 x 1.965 GHz (FMA issue)", "hbm_frac": hb}
        else:
            roof = {"bound": "hbm", "achieved": d.get("achieved_gbs", 0.0), "peak": PK["hbm"], "unit": "GB/s", "frac": hb,
                    "traffic": None, "kernel": name, "share_of_step": d["share"], "avg_us": d["avg_us"],
                    "peak_source": PK["source"]}
            if name in TENSOR_KERNELS:
                roof["tensor_frac"] = tc
        try:        # measured DRAM traffic of that kernel (one ncu --set full capture, profiles/traffic.json), per launch
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
            if tr:
                roof["traffic"] = tr["dram_bytes_per_launch"]
                roof["traffic_note"] = (f"{tr['kernel']}: algorithmic {tr['algorithmic_bytes_per_launch']} B per launch in that "
                                        f"capture; {tr['source']}")
        except Exception:
            pass
        return roof

    def bcast_net(net):
        dist.broadcast(net._flat.data, src=0)
        for b in (net._theta_shift, net._theta_scale, net._x_mean, net._x_std):
            dist.broadcast(b, src=0)

    def build_workload(name, dropout=0.0):
        w = settings(args, name)
        c = Ctx()
        c.name, c.w = name, w
        c.cfg, c.theta, c.x = make_data(name, w["rows"], rank)
        c.B, c.loss, c.rows = w["batch"], w["loss"], w["rows"]
        torch.manual_seed(0)
        c.net = fb.build_nsf(c.theta[:8192], c.x[:8192], hidden_features=c.cfg.hidden, num_transforms=c.cfg.transforms,
                             num_bins=c.cfg.bins, num_blocks=c.cfg.blocks, dropout_probability=dropout,
                             generator=torch.Generator().manual_seed(0)).to(dev)
        if world > 1:          # one replica: rank 0's weights AND z-score statistics (every rank generated its own data)
            bcast_net(c.net)
        c.net.train()
        c.prior = fb.BoxUniform(-3 * torch.ones(c.cfg.D), 3 * torch.ones(c.cfg.D), device=dev)
        c.eng = _TrainEngine(c.net, c.prior, c.B, c.loss == "atomic", c.cfg.atoms, True, 5e-4, 5.0, use_graph=bool(args.graph))
        c.theta_dev = c.theta.to(dev).contiguous()
        c.x_dev = c.net._standardize(c.x.to(dev))
        c.masks_dev = torch.zeros(c.rows, device=dev)
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        c.batches = [torch.randint(0, c.rows, (c.B,), device=dev, generator=g) for _ in range(8)]
        c.step = lambda i: c.eng.step(c.batches[i % len(c.batches)], c.theta_dev, c.x_dev, c.masks_dev)
        return c

    def measure(c, k, w, with_clocks=False):
        for i in range(w):
            c.step(i)
        clocks = ClockSampler(local) if with_clocks else None
        if clocks:
            clocks.start()
        n0 = L.fp_launch_count()
        ms = timed(c.step, k)
        launches = int(L.fp_launch_count() - n0)
        if c.eng.graph is not None:      # replays do not pass through the launch counter: kernels per captured step x k
            launches = c.eng.launches_per_step * k
        clk = clocks.stop() if clocks else None
        c.eng.check_status()
        value = world * c.B * k / (ms * 1e-3)

        def plain(i):       # plain launches (a graph replay would bypass the per-launch event hooks)
            c.eng.idx.copy_(c.batches[i % len(c.batches)])
            c.eng._step_body(c.theta_dev, c.x_dev, c.masks_dev)

        table = kernel_table(plain, min(k, 10))
        return dict(value=value, ms_per_step=ms / k, launches=launches, clocks=clk, kernels=table, roofline=roofline_of(table))

    # ============================================================ headline workload
    if args.only_sampling:
        args.config, args.no_e2e, args.no_cpu = "c1", True, True           # (the cheapest workload stands in for the headline)
    t_start = time.time()
    main_w = dict(settings(args, args.config))
    main_c = build_workload(args.config, args.dropout)
    cfg = main_c.cfg
    log(f"{args.config}: data + network ready after {time.time() - t_start:.1f} s")
    res = measure(main_c, K, W, with_clocks=True)
    log(f"{args.config}: {res['ms_per_step']:.3f} ms/step, {res['value']:.0f} pairs/s")
    out_extra = {}

    # ============================================================ e2e: the public call, host tensors in, estimator out
    e2e = None
    if not args.no_e2e:
        B = main_c.B
        n_steps_e2e = 8
        # weak scaling: 8 optimiser steps per rank and epoch (every rank holds the whole host dataset and takes its shard,
        # as SNPE_C.train does under data parallelism), capped at 6 GB of host rows per rank
        n_rows = min(int(B * n_steps_e2e * world / 0.9) + 64, int(6e9 / (4 * (cfg.D + cfg.C + 1))))
        reps = (n_rows + main_c.rows - 1) // main_c.rows
        th_host = main_c.theta.repeat(reps, 1)[:n_rows].contiguous().pin_memory()
        x_host = main_c.x.repeat(reps, 1)[:n_rows].contiguous().pin_memory()
        builder = fb.posterior_nn(model="nsf", hidden_features=cfg.hidden, num_transforms=cfg.transforms, num_bins=cfg.bins,
                                  num_blocks=cfg.blocks, dropout_probability=args.dropout)
        inf = fb.SNPE_C(main_c.prior, density_estimator=builder, device=str(dev))
        atomic = main_c.loss == "atomic"
        if atomic:      # a later round (proposal != prior) is what selects the APT loss in SNPE_C.train
            inf.append_simulations(th_host, x_host, proposal=object())
        else:
            inf.append_simulations(th_host, x_host)
        kw = dict(training_batch_size=B, num_atoms=cfg.atoms, use_combined_loss=True, stop_after_epochs=1000,
                  force_first_round_loss=not atomic, use_cuda_graph=False, seed=0)
        inf.train(max_num_epochs=0, **kw)                    # builds the estimator (z-score statistics, initial weights) + warm-up epoch
        del main_c.eng
        torch.cuda.empty_cache()
        epochs = 2
        barrier()
        t0 = time.perf_counter()
        est = inf.train(max_num_epochs=epochs - 1, **kw)     # epochs 0 .. epochs-1
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tsec = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
        n_train = int(0.9 * n_rows) // world
        steps_run = epochs * (n_train // B)
        h2d = (th_host.numel() + x_host.numel() + n_rows) * 4 // max(world, 1)
        e2e = {"value": world * steps_run * B / float(tsec.item()), "unit": "pairs/s",
               "h2d_bytes_per_step": int(h2d / max(steps_run, 1)), "d2h_bytes_per_step": int(16 * epochs / max(steps_run, 1)) + 1,
               "call": f"SNPE_C.append_simulations(pinned host tensors, {n_rows} rows) + SNPE_C.train(max_num_epochs={epochs - 1}): "
                       f"dataset upload, {steps_run} optimiser steps, {epochs} validation passes, best-state snapshots, returned "
                       f"deepcopy; wall clock {float(tsec.item()):.3f} s (max over ranks)",
               "final_validation_log_prob": inf._summary["validation_log_probs"][-1]}
        log(f"e2e: {e2e['value']:.0f} pairs/s through SNPE_C.train ({steps_run} steps in {float(tsec.item()):.2f} s)")
        del inf, est, th_host, x_host
    ws_gb = main_c.net.workspace_floats(main_c.B * (cfg.atoms if main_c.loss == "atomic" else 1), main_c.B, True) * 4 / 1e9
    main_desc = workload_desc(args.config, cfg, main_c.loss, main_c.B, main_c.rows, world, args.dropout)
    out_extra["activation_workspace_gb_per_step"] = ws_gb
    theta_cpu, x_cpu = main_c.theta, main_c.x
    del main_c
    torch.cuda.empty_cache()

    # ============================================================ configs[4]: samples/s and log_prob/s from a trained c2 flow
    def run_sampling_and_extras():
        t1 = time.time()
        c2 = build_workload("c2")
        # the flow of this leg is a COMMITTED one (tests/golden/c5_flow_c2.npz, trained to early stopping by
        # tools/make_c5_flow.py): training is not bit-reproducible (atomics in the weight gradients), and a flow trained
        # here had acceptance rates between 0.34 and 0.85 from run to run -- accepted samples/s would not be comparable
        # across GPU counts.  Every rank loads the same weights, z-score statistics and observation.
        fixture = os.path.join(ROOT, "tests", "golden", "c5_flow_c2.npz")
        x_fix = None
        if os.path.exists(fixture):
            fx = np.load(fixture)
            with torch.no_grad():
                c2.net._flat.copy_(torch.from_numpy(fx["flat"].astype(np.float32)).to(dev))
                for name, key in (("_theta_shift", "theta_shift"), ("_theta_scale", "theta_scale"), ("_x_mean", "x_mean"), ("_x_std", "x_std")):
                    getattr(c2.net, name).copy_(torch.from_numpy(fx[key]).to(dev))
            x_fix = torch.from_numpy(fx["x_obs"])
            flow_desc = "committed c2 flow (tests/golden/c5_flow_c2.npz, 61 epochs of SNPE_C.train)"
        else:
            pre = _TrainEngine(c2.net, c2.prior, 4096, False, c2.cfg.atoms, True, 2e-3, 5.0)
            gp = torch.Generator(device=dev).manual_seed(99)
            for i in range(400):
                pre.step(torch.randint(0, c2.rows, (4096,), device=dev, generator=gp), c2.theta_dev, c2.x_dev, c2.masks_dev)
            pre.check_status()
            del pre
            if world > 1:
                bcast_net(c2.net)
            flow_desc = "c2 flow after 400 seeded maximum-likelihood steps (rank 0's weights broadcast)"
        net = c2.net.eval()
        x_obs = x_fix if x_fix is not None else make_data("c2", 4096, 0)[2][:1]      # ONE observation for every rank
        post = fb.DirectPosterior(net, c2.prior, max_sampling_batch_size=1 << 22).set_default_x(x_obs)
        n_total = (1 << 24) * world            # weak scaling, like the training leg: 2^24 accepted samples per GPU
        lo, hi = fdist.shard_rows(n_total) if world > 1 else (0, n_total)
        post.sample((1 << 16,))
        acc = post.leakage_correction(x_obs, num_rejection_samples=1 << 18)
        try:
            ms_s = timed(lambda i: post.sample((hi - lo,)), 1)
            out_extra["posterior_samples_per_s"] = n_total / (ms_s * 1e-3)
            out_extra["posterior_raw_draws_per_s"] = n_total / max(acc, 1e-9) / (ms_s * 1e-3)
        except AssertionError as exc:
            # posterior.sample raises, like upstream's asserts, when the flow produces a non-finite draw (a negative
            # spline discriminant in fp32): report it, and time the raw inverse pass instead
            out_extra["posterior_sampling_error"] = str(exc)
            ctx1 = net._standardize(x_obs.to(dev))
            zr = torch.randn(1 << 21, c2.cfg.D, device=dev)
            ms_r = timed(lambda i: net.inverse_rows(zr, ctx1), 8)
            out_extra["posterior_raw_draws_per_s"] = world * zr.shape[0] * 8 / (ms_r * 1e-3)
            out_extra["posterior_samples_per_s"] = out_extra["posterior_raw_draws_per_s"] * acc
        out_extra["posterior_acceptance_rate"] = acc
        th_q = net.inverse_rows(torch.randn(1 << 21, c2.cfg.D, device=dev), net._standardize(x_obs.to(dev)))      # raw flow draws
        th_q = torch.nan_to_num(th_q, nan=0.0, posinf=3.0, neginf=-3.0)
        ms_l = timed(lambda i: post.log_prob(th_q, norm_posterior=False), 8)
        out_extra["log_prob_per_s"] = world * th_q.shape[0] * 8 / (ms_l * 1e-3)
        out_extra["sampling_workload"] = (f"{flow_desc}, {n_total} accepted "
                                          f"samples sharded over {world} rank(s), posterior.sample / posterior.log_prob "
                                          "(floppity.py:252, floppityFUN.py:35), no collective on the data path")
        log(f"c5: {out_extra['posterior_samples_per_s']:.3g} accepted samples/s (acceptance {acc:.3f}), "
            f"{out_extra['log_prob_per_s']:.3g} log_prob/s  [{time.time() - t1:.1f} s]")
        del post, th_q
        # ---- the same two calls on a c3-SHAPED flow (hidden 512 x 4 blocks, 15 transforms, 2000-bin context): its 114 M
        # weights cannot be committed, so this is the seeded initialisation (identical on every rank and box) -- the
        # acceptance rate of an untrained flow means nothing, raw draws/s and log_prob/s are what this states
        try:
            cfg3, th3, x3 = make_data("c3", 8192, 0)
            torch.manual_seed(0)
            net3 = fb.build_nsf(th3, x3, hidden_features=cfg3.hidden, num_transforms=cfg3.transforms, num_bins=cfg3.bins,
                                num_blocks=cfg3.blocks, generator=torch.Generator().manual_seed(0)).to(dev).eval()
            ctx3 = net3._standardize(x3[:1].to(dev))
            z3 = torch.randn(1 << 20, cfg3.D, device=dev)
            ms_r3 = timed(lambda i: net3.inverse_rows(z3, ctx3), 4)
            th_q3 = torch.nan_to_num(net3.inverse_rows(z3, ctx3), nan=0.0, posinf=3.0, neginf=-3.0)
            prior3 = fb.BoxUniform(-3 * torch.ones(cfg3.D), 3 * torch.ones(cfg3.D), device=dev)
            post3 = fb.DirectPosterior(net3, prior3).set_default_x(x3[:1])
            ms_l3 = timed(lambda i: post3.log_prob(th_q3, norm_posterior=False), 4)
            out_extra["c3_shape_sampling"] = {
                "raw_draws_per_s": world * z3.shape[0] * 4 / (ms_r3 * 1e-3), "log_prob_per_s": world * z3.shape[0] * 4 / (ms_l3 * 1e-3),
                "fraction_inside_prior": float(((th_q3.abs() < 3).all(dim=1)).float().mean().item()),
                "workload": f"seeded-initialisation c3 flow, {z3.shape[0]} rows per call and rank, one observation; "
                            "flow.sample rows / posterior.log_prob(norm_posterior=False)"}
            log(f"c5 on the c3 shape: {out_extra['c3_shape_sampling']['raw_draws_per_s']:.3g} raw draws/s, "
                f"{out_extra['c3_shape_sampling']['log_prob_per_s']:.3g} log_prob/s")
            del net3, post3, z3, th_q3
            torch.cuda.empty_cache()
        except Exception as exc:      # noqa: BLE001
            out_extra["c3_shape_sampling"] = {"error": f"{type(exc).__name__}: {exc}"}
        # ======================================================== extras: the other configs, each with its own roofline
        extras = {}
        if world == 1 and not args.only_sampling:
            c2.net.train()
            r2 = measure(c2, max(K // 2, 5), 3)
            extras["c2"] = dict(config=workload_desc("c2", c2.cfg, c2.loss, c2.B, c2.rows, world, 0.0), **r2)
            del c2
            torch.cuda.empty_cache()
            for name, drop, loss in (("c1", 0.0, None), ("c4", 0.0, None), ("c2", 0.1, None), ("c1", 0.1, None),
                                     ("c3", 0.0, "atomic")):
                if name == args.config and loss is None and drop == args.dropout:
                    continue
                if loss is not None:          # the headline config under the SNPE-C loss too (10 atoms per pair: ~46 GB of
                    WORKLOADS[name] = dict(WORKLOADS[name], loss=loss, batch=9472)       # activations, the round-1 setting)
                c = build_workload(name, drop)
                r = measure(c, max(K // 4, 3) if loss else max(K // 2, 5), 3)
                key = (name if drop == 0.0 else f"{name}_dropout{drop}") + (f"_{loss}" if loss else "")
                extras[key] = dict(config=workload_desc(name, c.cfg, c.loss, c.B, c.rows, world, drop), **r)
                log(f"{key}: {r['ms_per_step']:.3f} ms/step, {r['value']:.0f} pairs/s")
                del c
                torch.cuda.empty_cache()
            WORKLOADS["c3"] = dict(loss="mle", batch=18944, rows=65536)
            # the same step with 3xTF32 operand pieces in the stand-alone GEMMs (the format of the first half of round 2)
            L.fp_set_gemm_format(0)
            try:
                c = build_workload("c3", 0.0)
                r = measure(c, max(K // 4, 3), 3)
                extras["c3_3xtf32_gemms"] = dict(config=workload_desc("c3", c.cfg, c.loss, c.B, c.rows, world, 0.0), **r)
                log(f"c3 with 3xTF32 GEMMs: {r['ms_per_step']:.3f} ms/step, {r['value']:.0f} pairs/s")
                del c
            finally:
                L.fp_set_gemm_format(1)
            torch.cuda.empty_cache()
            # the REDUCED-PRECISION variant of the wide conditioner, stated apart (one product per GEMM instead of the
            # fp32-accurate three: ~1e-3 relative on log_prob, outside the 1e-4 parity bar -- never the default)
            L.fp_set_precision(1)
            try:
                c = build_workload("c3", 0.0)
                r = measure(c, max(K // 4, 3), 3)
                extras["c3_reduced_precision_single_pass"] = dict(
                    config=workload_desc("c3", c.cfg, c.loss, c.B, c.rows, world, 0.0), **r,
                    precision="ONE product of the high fp16 pieces per tensor-core GEMM (11-bit mantissa operands, fp32 accumulation); "
                              "log_prob within 5e-3 relative of the fp32 oracle (tests/test_gpu_dropout_modes.py), NOT within the "
                              "1e-4 bar")
                log(f"c3 single pass: {r['ms_per_step']:.3f} ms/step, {r['value']:.0f} pairs/s")
                del c
            finally:
                L.fp_set_precision(0)
            torch.cuda.empty_cache()
            for k, v in extras.items():
                v.pop("clocks", None)
            out_extra["extras"] = extras
        else:
            del c2

    if not args.no_extras:
        # a failure in an extra (e.g. the seeded c2 flow producing a non-finite draw, which raises as upstream's
        # assert does) must not take the headline measurement with it: it is reported in the line instead
        try:
            run_sampling_and_extras()
        except Exception as exc:      # noqa: BLE001
            out_extra["extras_error"] = f"{type(exc).__name__}: {exc}"
            log(f"extras failed: {type(exc).__name__}: {exc}")
            if world > 1:
                raise

    # ============================================================ CPU baseline beside it (rank 0, N == 1 only)
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        w = main_w
        Bc = ref_batch_for(args, cfg)
        v, n, dt = oracle_steps(cfg, w["loss"], Bc, theta_cpu[:4096], x_cpu[:4096], 1, seconds=12.0, dropout=args.dropout)
        cpu_base = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                    "sample": f"{n} optimiser steps at batch {Bc} in {dt:.1f} s (a bounded sample of the per-GPU batch {w['batch']}), "
                              f"{w['loss']} loss, oracle/nsf_oracle.py eager fp32"}

    if rank == 0 and args.only_sampling:
        line = {"metric": "posterior samples/s (configs[4] leg only)", "n_gpus": world, "unit": "samples/s",
                "value": out_extra.get("posterior_samples_per_s")}
        line.update(out_extra)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    elif rank == 0:
        line = {
            "metric": "NSF train pairs/s", "value": res["value"], "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": main_desc, "e2e": e2e, "gpu_launches": res["launches"], "clocks": res["clocks"],
            "roofline": res["roofline"], "cpu_baseline": cpu_base, "kernels": res["kernels"],
        }
        line.update(out_extra)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
