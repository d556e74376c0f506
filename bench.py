#!/usr/bin/env python
"""bench.py -- NSF training pairs/s (headline) + posterior samples/s on the BASELINE.json workload.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the restated sbi path (oracle/) on the host CPU cores

A "step" is one optimiser step of ``SNPE_C.train`` (sbi PosteriorEstimator.train, /root/reference/floppity.py:219)
on one batch: gather batch from the HBM-resident dataset -> flow forward -> loss -> hand-written backward ->
[NCCL all-reduce] -> clip_grad_norm_(5) + Adam.  Workload = BASELINE.json configs[1] (c2: NSF T10 K10 H128 Bk2,
20 params / 500 bins, atoms 10), synthetic spectra (ARCiS is external Fortran), random-init weights.
Weak scaling: the per-GPU batch is fixed.  One JSON line on stdout (rank 0).
"""
import argparse
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

CLASS_NAMES = ["gemm_simt_fp32", "gemm_tcgen05", "rqs_forward", "rqs_inverse", "rqs_backward", "other"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--batch", type=int, default=9472,
                    help="per-GPU batch (rows of (theta, x) pairs).  Default 9472 = 148 SMs x 64: with atoms 10 a step is "
                         "94 720 flow rows = 740 row tiles of 128 = exactly 5 per SM, so the persistent kernels run whole "
                         "waves (4096 leaves 2.16 tiles per SM: 3 waves of work for 2.16 waves of rows)")
    ap.add_argument("--loss", default="atomic", choices=["atomic", "mle"],
                    help="atomic = SNPE-C APT loss with `atoms` contrastive proposals (subsystem 3); mle = the branch "
                         "the reference actually takes (force_first_round_loss=True, floppity.py:220)")
    ap.add_argument("--rows", type=int, default=65536, help="resident dataset rows per GPU (c2: 50k pairs/round -> 65 536)")
    ap.add_argument("--ref-batch", type=int, default=256, help="bounded per-step sample for the CPU arm")
    ap.add_argument("--graph", type=int, default=0, help="replay the optimiser step from a CUDA graph")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
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


def make_data(cfg_name, rows, seed):
    from floppity_deprecated_b200 import synthetic
    cfg = synthetic.CONFIGS[cfg_name]
    rng = np.random.default_rng(seed)
    theta = rng.uniform(-3, 3, (rows, cfg.D))          # prior draws in the -ynorm box (same support as the Sobol round 0)
    th, x, scaler, _ = synthetic.make_dataset(cfg_name, theta=theta, noise_seed=4321 + seed)
    return cfg, torch.tensor(th), torch.tensor(x)


# ------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """Restated sbi path (oracle/nsf_oracle.py, plain eager PyTorch fp32) on ALL host cores.  sbi/nflows cannot be
    installed here, so `kind` is "port".  Each step is a bounded sample (batch --ref-batch) of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import nsf_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, theta, x = make_data(args.config, max(2048, args.ref_batch * 4), 0)
    torch.manual_seed(0)
    net = O.build_nsf(theta, x, hidden_features=cfg.hidden, num_transforms=cfg.transforms, num_bins=cfg.bins,
                      num_blocks=cfg.blocks)
    prior = O.BoxUniform(-3 * torch.ones(cfg.D), 3 * torch.ones(cfg.D))
    opt = torch.optim.Adam(net.parameters(), lr=5e-4)
    B = args.ref_batch
    masks = torch.zeros(B, 1)
    net.train()

    def step(i):
        idx = torch.randint(0, theta.shape[0], (B,))
        opt.zero_grad()
        if args.loss == "atomic":
            lp = O.log_prob_proposal_posterior_atomic(net, prior, theta[idx], x[idx], masks, cfg.atoms, True)
        else:
            lp = net.log_prob(theta[idx], x[idx])
        loss = -lp.mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 5.0)
        opt.step()
        return float(loss)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    v = B * args.steps / dt
    sample = f"{args.steps} optimiser steps at batch {B} (bounded sample of the per-GPU batch {args.batch}), {args.loss} loss"
    print(json.dumps({
        "impl": "reference", "metric": "NSF train pairs/s", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: SNPE-C NSF T{cfg.transforms} K{cfg.bins} H{cfg.hidden} Bk{cfg.blocks}, "
                               f"{cfg.D} params / {cfg.C} bins, atoms {cfg.atoms}", "loss": args.loss, "batch": B},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------- our arm
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
    from floppity_deprecated_b200 import _lib
    from floppity_deprecated_b200.inference import _TrainEngine
    L = _lib.lib()
    PK = peaks()

    cfg, theta, x = make_data(args.config, args.rows, rank)
    torch.manual_seed(0)
    net = fb.build_nsf(theta, x, hidden_features=cfg.hidden, num_transforms=cfg.transforms, num_bins=cfg.bins,
                       num_blocks=cfg.blocks, generator=torch.Generator().manual_seed(0)).to(dev)
    if world > 1:
        dist.broadcast(net._flat.data, src=0)
    net.train()
    prior = fb.BoxUniform(-3 * torch.ones(cfg.D), 3 * torch.ones(cfg.D), device=dev)
    B, K, W = args.batch, args.steps, args.warmup
    atomic = args.loss == "atomic"
    eng = _TrainEngine(net, prior, B, atomic, cfg.atoms, True, 5e-4, 5.0, use_graph=bool(args.graph))
    theta_dev = theta.to(dev).contiguous()
    x_dev = net._standardize(x.to(dev))
    masks_dev = torch.zeros(args.rows, device=dev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    batches = [torch.randint(0, args.rows, (B,), device=dev, generator=g) for _ in range(8)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
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

    def step(i):
        eng.step(batches[i % len(batches)], theta_dev, x_dev, masks_dev)

    for i in range(max(W, 3)):
        step(i)
    clocks = ClockSampler(local)
    clocks.start()
    n0 = L.fp_launch_count()
    ms = timed(step, K)
    launches = int(L.fp_launch_count() - n0)
    if eng.graph is not None:            # replays do not pass through the launch counter: kernels per captured step x K
        launches = eng.launches_per_step * K
    clk = clocks.stop()
    eng.check_status()
    value = world * B * K / (ms * 1e-3)
    print(f"[bench] rank {rank}: {ms / K:.3f} ms/step, {value:.0f} pairs/s", file=sys.stderr, flush=True)

    # ---- e2e: host (pinned) batch -> H2D -> standardise -> step -> D2H loss, through the host-facing step call
    th_pin = [theta[b.cpu()].pin_memory() for b in batches[:4]]
    x_pin = [x[b.cpu()].pin_memory() for b in batches[:4]]
    # two device staging buffers: the H2D copy of step i+1 (copy stream) runs under the kernels of step i, the way a
    # prefetching loader feeds a training loop; every step still copies its own inputs and reads its own loss back
    stage = [(torch.empty(B, cfg.D, device=dev), torch.empty(B, cfg.C, device=dev)) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    losses = []

    def issue_copy(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[b])                 # the step that last used this buffer is done with it
            stage[b][0].copy_(th_pin[i % 4], non_blocking=True)
            stage[b][1].copy_(x_pin[i % 4], non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_step(i):
        b = i % 2
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[b])
        issue_copy(i + 1)
        eng.loss_sum.zero_()
        eng.step_on_batch(stage[b][0], net._standardize(stage[b][1]))
        free[b].record(cur)
        losses.append(float(eng.loss_sum.item()) / B)      # device -> host read of the step's loss

    def e2e_run(n):
        for b in range(2):
            free[b].record(torch.cuda.current_stream())
        issue_copy(0)
        for i in range(n):
            e2e_step(i)

    e2e_run(2)
    torch.cuda.synchronize()

    def timed_e2e(n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(n)                                          # includes the first step's exposed copy
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item())

    ms_e2e = timed_e2e(K)
    e2e = {"value": world * B * K / (ms_e2e * 1e-3), "unit": "pairs/s",
           "h2d_bytes_per_step": int(B * (cfg.D + cfg.C) * 4), "d2h_bytes_per_step": 4}

    # ---- per-kernel-class CUDA-event timing inside a second pass over the same K steps
    out_extra = {}
    L.fp_prof_enable(1)
    for i in range(K):      # plain launches (a graph replay would bypass the per-launch event hooks)
        eng.idx.copy_(batches[i % len(batches)])
        eng._step_body(theta_dev, x_dev, masks_dev)
    import ctypes
    pms, pl, pw, pb = (ctypes.c_double * 6)(), (ctypes.c_int64 * 6)(), (ctypes.c_double * 6)(), (ctypes.c_double * 6)()
    _lib.check(L.fp_prof_collect2(pms, pl, pw, pb), "prof")
    L.fp_prof_enable(0)
    tot = sum(pms) or 1.0
    # 3xTF32: three TF32 passes per fp32-accurate product; TF32 dense peak = half the measured bf16 peak
    tf32x3_peak = PK["tf_sust"] / 2.0 / 3.0
    classes = {}
    for i, name in enumerate(CLASS_NAMES):
        if pl[i] == 0:
            continue
        d = {"share": pms[i] / tot, "launches_per_step": pl[i] / K, "avg_us": 1e3 * pms[i] / pl[i]}
        if i in (0, 1):
            d["achieved_tflops_fp32_equiv"] = pw[i] / (pms[i] * 1e-3) / 1e12
            d["frac_of_3xtf32_tensor_peak"] = d["achieved_tflops_fp32_equiv"] / tf32x3_peak
            d["achieved_gbs"] = pb[i] / (pms[i] * 1e-3) / 1e9
            d["frac_of_hbm_peak"] = d["achieved_gbs"] / PK["hbm"]
        elif i in (2, 3, 4):
            d["achieved_gbs"] = pw[i] / (pms[i] * 1e-3) / 1e9
            d["frac_of_hbm_peak"] = d["achieved_gbs"] / PK["hbm"]
        classes[name] = d
    dom = max(range(6), key=lambda i: pms[i])
    if dom in (0, 1):
        # The conditioner GEMMs of this workload are skinny (hidden 128): arithmetic intensity ~32 flop/B in fp32, i.e.
        # below the ridge even at three TF32 passes.  Both fractions are measured; the binding one is reported.
        ach_t = pw[dom] / (pms[dom] * 1e-3) / 1e12
        ach_b = pb[dom] / (pms[dom] * 1e-3) / 1e9
        if ach_b / PK["hbm"] >= ach_t / tf32x3_peak:
            roof = {"bound": "hbm", "achieved": ach_b, "peak": PK["hbm"], "unit": "GB/s", "frac": ach_b / PK["hbm"],
                    "traffic": None, "kernel": CLASS_NAMES[dom], "peak_source": PK["source"],
                    "tensor_frac_3xtf32": ach_t / tf32x3_peak}
        else:
            roof = {"bound": "tensor", "achieved": ach_t, "peak": tf32x3_peak, "unit": "TFLOP/s", "frac": ach_t / tf32x3_peak,
                    "traffic": None, "kernel": CLASS_NAMES[dom],
                    "peak_source": PK["source"] + ": bf16 sustained / 2 (TF32) / 3 (passes of the fp32-accurate split)",
                    "hbm_frac": ach_b / PK["hbm"]}
    else:
        ach = pw[dom] / (pms[dom] * 1e-3) / 1e9 if pw[dom] else 0.0
        roof = {"bound": "hbm", "achieved": ach, "peak": PK["hbm"], "unit": "GB/s", "frac": ach / PK["hbm"], "traffic": None,
                "kernel": CLASS_NAMES[dom], "peak_source": PK["source"]}

    # measured DRAM traffic of the dominant class's main kernel (one ncu --set full capture, profiles/traffic.json)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(roof["kernel"])
        if tr:
            roof["traffic"] = tr["dram_bytes_per_launch"]
            roof["traffic_note"] = f"{tr['kernel']}: algorithmic {tr['algorithmic_bytes_per_launch']} B per launch; {tr['source']}"
    except Exception:
        pass

    # ---- extras: posterior sampling / log_prob throughput (sharded per GPU, no collective), MLE-mode training
    if not args.no_extras:
        net.eval()
        post = fb.DirectPosterior(net, prior).set_default_x(x[:1])
        n_draw = 1 << 20
        ctx_std = net._standardize(x[:1].to(dev))
        z = torch.randn(n_draw, cfg.D, device=dev)
        for _ in range(2):
            net.inverse_rows(z, ctx_std)
        ms_s = timed(lambda i: net.inverse_rows(z, ctx_std), 5)
        out_extra["posterior_draws_per_s"] = world * n_draw * 5 / (ms_s * 1e-3)
        th_q = theta_dev[torch.randint(0, args.rows, (n_draw,), device=dev)]
        x1 = x[:1].to(dev)

        def lp_eval(i):
            with torch.no_grad():       # posterior.log_prob (floppityFUN.py:35): no gradient, rows chunked by the flow
                net.log_prob(th_q, x1)

        for _ in range(2):
            lp_eval(0)
        ms_l = timed(lp_eval, 5)
        out_extra["log_prob_per_s"] = world * n_draw * 5 / (ms_l * 1e-3)
        net.train()
        eng2 = _TrainEngine(net, prior, B, not atomic, cfg.atoms, True, 5e-4, 5.0, use_graph=False)
        for i in range(3):
            eng2.step(batches[i], theta_dev, x_dev, masks_dev)
        ms2 = timed(lambda i: eng2.step(batches[i % 8], theta_dev, x_dev, masks_dev), K)
        out_extra["train_pairs_per_s_" + ("mle" if atomic else "atomic")] = world * B * K / (ms2 * 1e-3)
        if atomic:
            # the same MLE loss at the row count the atomic step works on (B x atoms flow rows): the launch-bound small
            # step above is a property of the batch size, not of the loss
            Bb = B * cfg.atoms
            del eng2
            eng3 = _TrainEngine(net, prior, Bb, False, cfg.atoms, True, 5e-4, 5.0, use_graph=False)
            big = [torch.randint(0, args.rows, (Bb,), device=dev, generator=g) for _ in range(2)]
            for i in range(3):
                eng3.step(big[i % 2], theta_dev, x_dev, masks_dev)
            ms3 = timed(lambda i: eng3.step(big[i % 2], theta_dev, x_dev, masks_dev), max(K // 2, 2))
            out_extra["train_pairs_per_s_mle_batch_x_atoms"] = world * Bb * max(K // 2, 2) / (ms3 * 1e-3)
            del eng3

    # ---- CPU baseline beside it (rank 0, N == 1 only): the oracle on the host cores, bounded sample
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import nsf_oracle as O
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        onet = O.build_nsf(theta[:2048], x[:2048], hidden_features=cfg.hidden, num_transforms=cfg.transforms,
                           num_bins=cfg.bins, num_blocks=cfg.blocks)
        oprior = O.BoxUniform(-3 * torch.ones(cfg.D), 3 * torch.ones(cfg.D))
        opt = torch.optim.Adam(onet.parameters(), lr=5e-4)
        Bc, nsteps = args.ref_batch, 0
        t_end = time.perf_counter() + 12.0
        t0 = None
        while True:
            idx = torch.randint(0, 2048, (Bc,))
            opt.zero_grad()
            lp = (O.log_prob_proposal_posterior_atomic(onet, oprior, theta[idx], x[idx], torch.zeros(Bc, 1), cfg.atoms, True)
                  if atomic else onet.log_prob(theta[idx], x[idx]))
            (-lp.mean()).backward()
            torch.nn.utils.clip_grad_norm_(onet.parameters(), 5.0)
            opt.step()
            if t0 is None:
                t0 = time.perf_counter()       # first step is warm-up
                continue
            nsteps += 1
            if time.perf_counter() > t_end and nsteps >= 2:
                break
        dt = time.perf_counter() - t0
        cpu_base = {"value": Bc * nsteps / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
                    "sample": f"{nsteps} optimiser steps at batch {Bc} (bounded sample), {args.loss} loss, oracle/nsf_oracle.py eager fp32"}

    if rank == 0:
        ws_gb = net.workspace_floats(eng.R, B, True) * 4 / 1e9
        line = {
            "metric": "NSF train pairs/s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: SNPE-C NSF T{cfg.transforms} K{cfg.bins} H{cfg.hidden} Bk{cfg.blocks}, "
                                   f"{cfg.D} params / {cfg.C} bins, atoms {cfg.atoms} "
                                   f"(BASELINE.json configs[{['c1', 'c2', 'c3', 'c4'].index(args.config)}])",
                       "loss": args.loss, "per_gpu_batch": B, "global_batch": B * world, "flow_rows_per_step": eng.R * world,
                       "resident_rows_per_gpu": args.rows, "parallelism": f"dp{world}",
                       "l2": f"no flush: per-step activation workspace {ws_gb:.2f} GB >> 126 MB L2"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clk, "roofline": roof, "cpu_baseline": cpu_base,
            "kernel_classes": classes, "final_loss": losses[-1] if losses else None,
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
