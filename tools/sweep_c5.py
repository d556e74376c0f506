"""BASELINE.json configs[4]: posterior sampling + log_prob sweep, 10^4 .. 10^8 draws from a TRAINED NSF, sharded
across the ranks of one node (no collective on the data path; `torchrun --nproc-per-node N tools/sweep_c5.py`).

The flow is the c2 network trained for a few epochs on the synthetic retrieval task through the public
`SNPE_C.append_simulations / train / build_posterior` surface (floppity.py:216-225), then
`posterior.sample((n,))` (floppity.py:252) and `posterior.log_prob(theta)` (floppityFUN.py:35) are timed on the
device (CUDA events, max over ranks).  One JSON line per sweep point on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--train-rows", type=int, default=16384)
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--max-exp", type=int, default=8)
    ap.add_argument("--chunk", type=int, default=1 << 22, help="rows per kernel batch (posterior max_sampling_batch_size)")
    args = ap.parse_args()

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import floppity_deprecated_b200 as fb
    from floppity_deprecated_b200 import synthetic, dist as fdist

    cfg = synthetic.CONFIGS[args.config]
    theta, x, scaler, _ = synthetic.make_dataset(args.config, args.train_rows)
    prior = fb.BoxUniform(-3 * torch.ones(cfg.D), 3 * torch.ones(cfg.D), device=str(dev))
    builder = fb.posterior_nn(model="nsf", hidden_features=cfg.hidden, num_transforms=cfg.transforms, num_bins=cfg.bins,
                              num_blocks=cfg.blocks)
    inf = fb.SNPE_C(prior, density_estimator=builder, device=str(dev))
    inf.append_simulations(torch.tensor(theta), torch.tensor(x))
    net = inf.train(training_batch_size=1024, max_num_epochs=args.epochs, stop_after_epochs=args.epochs, seed=0)
    # The SAME trained flow at every N (so accepted-samples/s are comparable: the acceptance rate is a property of the
    # model): every rank trains the identical single-process job above, then the group is formed and rank 0's weights win.
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        dist.broadcast(net._flat.data, src=0)
    post = inf.build_posterior(net).set_default_x(torch.tensor(x[:1]))
    post.max_sampling_batch_size = args.chunk
    val_lp = inf._summary["validation_log_probs"][-1]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sync()
        return float(ms.item()), out

    post.sample((10000,))      # warm-up (also fills the workspace caches)
    acc_rate = post.leakage_correction(torch.tensor(x[:1]), num_rejection_samples=200000)
    for e in range(4, args.max_exp + 1):
        n = 10 ** e
        lo, hi = fdist.shard_rows(n)
        mine = hi - lo

        def draw():
            # accepted draws, chunked so the peak footprint stays bounded at 10^8; the result of each chunk is kept
            # only long enough to be counted (a caller would consume / copy it out chunk by chunk the same way)
            got, acc = 0, 0
            while got < mine:
                m = min(args.chunk, mine - got)
                s = post.sample((m,))
                got += s.shape[0]
                acc = s
            return acc

        if os.environ.get("SWEEP_PROFILE") and e == args.max_exp and rank == 0:
            import cProfile, pstats
            pr = cProfile.Profile(); pr.enable()
            ms_s, last = timed(draw)
            pr.disable()
            pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(18)
        else:
            ms_s, last = timed(draw)
        th = last[: min(last.shape[0], args.chunk)].contiguous()

        def lps():
            done = 0
            out = None
            while done < mine:
                m = min(th.shape[0], mine - done)
                out = post.log_prob(th[:m], norm_posterior=False)
                done += m
            return out

        ms_l, lp = timed(lps)
        if rank == 0:
            print(json.dumps({"config": args.config, "n": n, "n_gpus": world, "samples_per_s": n / (ms_s * 1e-3),
                              "sample_ms": ms_s, "log_prob_per_s": n / (ms_l * 1e-3), "log_prob_ms": ms_l,
                              "trained_epochs": args.epochs, "val_log_prob": val_lp,
                              "acceptance_rate": acc_rate, "raw_draws_per_s": n / max(acc_rate, 1e-9) / (ms_s * 1e-3),
                              "mean_log_prob_of_draws": float(lp.mean().item()), "chunk": args.chunk}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
