#!/usr/bin/env python
"""A FlopPITy-style multi-round retrieval on the synthetic forward model, through this package's drop-in surface only.

The body follows /root/reference/floppity.py:107-285 step by step (the ARCiS forward model is replaced by
``synthetic.simulate``; nothing else is simplified):

    Normalizer / create_prior              floppity.py:123,127
    density_builder                        floppity.py:132   (posterior_nn -> SNPE_C)
    per round:  sample_proposal            floppity.py:179   (Sobol in round 0, proposal.sample afterwards)
                compute (simulate)         floppity.py:183
                preprocess                 floppity.py:192   (noise augmentation x naug, StandardScaler -> fp32, on the GPU)
                IS / evidence_from_IS      floppity.py:204-212
                append_simulations, train  floppity.py:216-221
                build_posterior            floppity.py:225
                posterior.sample, pickle   floppity.py:244-260
    input_MAP                              floppity.py:264   (posterior.map)

    python examples/synthetic_retrieval.py --config c1 --rounds 3 --samples 4096 --out /tmp/retrieval
"""
import argparse
import os
import pickle
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import floppity_deprecated_b200 as fb                                    # noqa: E402
from floppity_deprecated_b200 import synthetic                            # noqa: E402
from floppity_deprecated_b200.retrieval import IS, Normalizer, evidence_from_IS, preprocess   # noqa: E402
from scipy.stats.qmc import Sobol                                         # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1", choices=list(synthetic.CONFIGS))
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--samples", type=int, default=4096, help="samples per round (power of two: Sobol, modules.py:106)")
    ap.add_argument("--naug", type=int, default=2)
    ap.add_argument("--atoms", type=int, default=10)
    ap.add_argument("--patience", type=int, default=5)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--npew", type=int, default=5000)
    ap.add_argument("--embedding", action="store_true", help="-embedding -embedding_type FC (modules.py:25)")
    ap.add_argument("--force-first-round-loss", type=int, default=1, help="1 = what floppity.py:220 passes")
    ap.add_argument("--out", default="/tmp/floppity_b200_retrieval")
    args = ap.parse_args(argv)
    os.makedirs(args.out, exist_ok=True)
    cfg = synthetic.CONFIGS[args.config]
    tail_bound = 3.0
    rng = np.random.default_rng(0)

    # the "physical" prior box and the observation (read_input of the reference)
    prior_bounds = np.stack([np.full(cfg.D, 0.0), np.linspace(1.0, 5.0, cfg.D)], axis=1)
    yscaler = Normalizer(prior_bounds, tail_bound)                                                  # floppity.py:123
    theta_true = yscaler.transform(rng.uniform(prior_bounds[:, 0] * 0.7 + prior_bounds[:, 1] * 0.3,
                                               prior_bounds[:, 0] * 0.3 + prior_bounds[:, 1] * 0.7)[None, :])
    theta_true = synthetic.swap_hotter_first(cfg, theta_true)
    obs_spec = synthetic.simulate(cfg, theta_true)[0]
    noise_spec = 0.01 * np.abs(np.median(obs_spec)) * np.ones_like(obs_spec)
    obs_spec = obs_spec + noise_spec * rng.standard_normal(obs_spec.shape)

    low = torch.tensor(yscaler.transform(prior_bounds[:, 0].reshape(1, -1)).reshape(-1))            # modules.py:48-59
    high = torch.tensor(yscaler.transform(prior_bounds[:, 1].reshape(1, -1)).reshape(-1))
    prior = fb.utils.BoxUniform(low=low, high=high, device="cuda")
    summary = fb.FCEmbedding(input_dim=cfg.C, num_hiddens=256, num_layers=4, output_dim=64) if args.embedding \
        else torch.nn.Identity()                                                                     # modules.py:21-45
    inference = fb.SNPE_C(prior=prior, density_estimator=fb.posterior_nn(
        model="nsf", num_transforms=cfg.transforms, hidden_features=cfg.hidden, num_bins=cfg.bins, embedding_net=summary,
        num_blocks=cfg.blocks, dropout_probability=0.0), device="cuda")                             # modules.py:61-66

    proposal, posteriors, logZs, samples = prior, [], [], []
    pp_args = types.SimpleNamespace(res_scale=False)
    for r in range(args.rounds):
        t0 = time.time()
        if r > 0:                                                                                    # modules.py:98-118
            np_theta = proposal.sample((args.samples,)).cpu().detach().numpy().reshape([-1, cfg.D])
        else:
            np_theta = -tail_bound + (2 * tail_bound) * Sobol(cfg.D, seed=0).random_base2(int(np.log2(args.samples)))
        np_theta = synthetic.swap_hotter_first(cfg, np_theta)
        arcis_spec = synthetic.simulate(cfg, np_theta)                                               # floppity.py:183
        t_sim = time.time() - t0
        theta_aug, x, xscaler, pca = preprocess(np_theta, arcis_spec, r, args.samples, obs_spec, noise_spec, args.naug,
                                                False, 0, True, cfg.C, False, args.out, "cuda", pp_args)   # floppity.py:192
        if r > 0:                                                                                    # floppity.py:204-212
            w_i, eff = IS(proposal, obs_spec, noise_spec, arcis_spec, np_theta)
            logZ, logZ_err = evidence_from_IS(w_i, eff)
            logZs.append((logZ, logZ_err))
        inference = inference.append_simulations(theta_aug, x, proposal=proposal)                    # floppity.py:216
        t1 = time.time()
        est = inference.train(show_train_summary=False, stop_after_epochs=args.patience, num_atoms=args.atoms,
                              force_first_round_loss=bool(args.force_first_round_loss), retrain_from_scratch=False,
                              use_combined_loss=True, training_batch_size=args.batch)                # floppity.py:219-221
        t_train = time.time() - t1
        default_x = xscaler.transform(obs_spec.reshape(1, -1))                                       # floppity.py:223
        posterior = inference.build_posterior(est).set_default_x(default_x)                          # floppity.py:225
        posteriors.append(posterior)
        proposal = posterior
        with open(os.path.join(args.out, "posteriors.pt"), "wb") as f:                               # floppity.py:244-248
            torch.save(posteriors, f)
        with open(os.path.join(args.out, "inference.p"), "wb") as f:
            pickle.dump(inference, f)
        t2 = time.time()
        tsamples = proposal.sample((args.npew,))                                                     # floppity.py:252
        samples.append(yscaler.inverse_transform(tsamples.cpu().detach().numpy()))                   # floppity.py:255
        t_samp = time.time() - t2
        err = np.abs(np.median(tsamples.cpu().numpy(), axis=0) - theta_true[0]).mean()
        print(f"round {r}: simulate {t_sim:.2f}s, train {t_train:.2f}s ({inference._summary['epochs_trained'][-1]} epochs, "
              f"best val log-prob {inference._summary['best_validation_log_prob'][-1]:.3f}), {args.npew} posterior draws "
              f"{t_samp * 1e3:.1f} ms, mean |median - truth| = {err:.3f}"
              + (f", logZ = {logZs[-1][0]:.2f} +- {logZs[-1][1]:.2f}" if r > 0 else ""), flush=True)
    MAP = proposal.map(x=None, num_iter=50, num_to_optimize=50, learning_rate=0.01, init_method="posterior",
                       num_init_samples=500, save_best_every=10, show_progress_bars=False, force_update=False)   # modules.py:173
    MAP = yscaler.inverse_transform(MAP.detach().cpu().numpy().reshape(1, -1))[0]
    np.save(os.path.join(args.out, "map_params.npy"), MAP)
    with open(os.path.join(args.out, "samples.p"), "wb") as f:
        pickle.dump(samples, f)
    return dict(samples=samples, logZs=logZs, MAP=MAP, truth=yscaler.inverse_transform(theta_true)[0],
                val=inference._summary["best_validation_log_prob"])


if __name__ == "__main__":
    main()
