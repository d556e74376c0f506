"""CPU suite, part 2: host logic, the C-ABI surface, the g++ build of the spline arithmetic, gloo data-parallel."""
import ctypes
import copy
import os
import pickle
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import floppity_deprecated_b200 as fb
from floppity_deprecated_b200 import _lib, synthetic
from oracle import nsf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "floppity_b200.h")).read()
    declared = set(re.findall(r"\b(fp_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"fp_param_kind", "fp_gemm_flags"}
    assert len(declared) >= 25
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib.SIGNATURES) == declared          # the ctypes table and the header must not drift apart
    assert b"sm_100a" in _lib.lib().fp_version()


def test_library_contains_sm100a_sass_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_layout_matches_upstream_parameter_count():
    for (D, C, T, K, H, Bk, want) in [(10, 100, 5, 8, 50, 2, 157_875), (7, 33, 3, 5, 20, 1, None)]:
        net = fb.build_nsf(torch.randn(64, D), torch.randn(64, C), hidden_features=H, num_transforms=T, num_bins=K, num_blocks=Bk)
        onet = O.build_nsf(torch.randn(64, D), torch.randn(64, C), hidden_features=H, num_transforms=T, num_bins=K, num_blocks=Bk)
        n_up = sum(p.numel() for p in onet.parameters())
        if want:
            assert n_up == want
        packed = net.unpack(net._flat.detach())
        assert sum(v.numel() for v in packed.values()) == n_up
        assert {k: tuple(v.shape) for k, v in packed.items()} == {k: tuple(p.shape) for k, p in onet.named_parameters()}


def test_state_dict_roundtrip_with_oracle_and_pickle():
    th, x = torch.randn(64, 7), torch.randn(64, 33)
    net = fb.build_nsf(th, x, hidden_features=20, num_transforms=3, num_bins=5, num_blocks=1)
    onet = O.build_nsf(th, x, hidden_features=20, num_transforms=3, num_bins=5, num_blocks=1)
    osd = onet.state_dict()
    assert set(osd) == set(net.state_dict())
    net.load_state_dict(osd)
    sd = net.state_dict()
    assert all(torch.equal(sd[k].cpu().to(osd[k].dtype), osd[k]) for k in osd)
    onet.load_state_dict(sd)                                   # and back into the upstream-shaped module
    for clone in (pickle.loads(pickle.dumps(net)), copy.deepcopy(net)):
        assert torch.equal(clone._flat, net._flat) and clone.cfg == net.cfg
    with pytest.raises(KeyError):
        net.load_state_dict({k: v for k, v in osd.items() if "final_layer.bias" not in k})


def test_zscore_statistics_follow_upstream():
    th, x = torch.randn(200, 4) * 3 + 1, torch.randn(200, 6) * 0.5 - 2
    x[:, 2] = 7.0                                              # zero variance -> clamp 1e-7
    net = fb.build_nsf(th, x, hidden_features=8, num_transforms=2, num_bins=4)
    onet = O.build_nsf(th, x, hidden_features=8, num_transforms=2, num_bins=4)
    osd = onet.state_dict()
    assert torch.allclose(net._theta_shift, osd["_transform._transforms.0._shift"])
    assert torch.allclose(net._theta_scale, osd["_transform._transforms.0._scale"])
    assert torch.allclose(net._x_std, osd["_embedding_net.0._std"]) and net._x_std[2].item() == pytest.approx(1e-7)


def test_no_cpu_fallback():
    net = fb.build_nsf(torch.randn(64, 4), torch.randn(64, 6), hidden_features=8, num_transforms=2, num_bins=4)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            net.log_prob(torch.randn(3, 4), torch.randn(3, 6))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            net.sample(3, context=torch.randn(1, 6))


def test_package_never_imports_oracle():
    """The product must not import, load, link or execute anything under oracle/ (comments may mention it)."""
    pkg = os.path.join(ROOT, "floppity_deprecated_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|nsf_oracle|oracle/_ref|hostmath", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), f


def test_unsupported_options_raise():
    with pytest.raises(NotImplementedError):
        fb.posterior_nn("maf")
    with pytest.raises(NotImplementedError):
        fb.build_nsf(torch.randn(8, 4), torch.randn(8, 6), embedding_net=torch.nn.Linear(6, 3))
    with pytest.raises(NotImplementedError):
        fb.build_nsf(torch.randn(8, 1), torch.randn(8, 6))


def test_box_uniform_matches_oracle():
    low, high = torch.tensor([-3.0, 0.0, 1.0]), torch.tensor([3.0, 2.0, 5.0])
    a, b = fb.BoxUniform(low, high), O.BoxUniform(low, high)
    pts = torch.tensor([[0.0, 1.0, 2.0], [3.0, 1.0, 2.0], [-3.0, 0.0, 1.0], [0.0, 2.5, 2.0]])
    assert torch.equal(a.log_prob(pts), b.log_prob(pts))
    assert a.within(pts).tolist() == [True, True, True, False]
    s = a.sample((1000,))
    assert s.shape == (1000, 3) and a.within(s).all()
    assert fb.utils.BoxUniform is fb.BoxUniform


def test_append_simulations_round_logic():
    prior = fb.BoxUniform(-3 * torch.ones(3), 3 * torch.ones(3))
    inf = fb.SNPE_C(prior, fb.posterior_nn("nsf"), device="cuda")
    x = torch.randn(10, 5)
    x[3, 1] = float("nan")
    with pytest.warns(UserWarning):
        inf.append_simulations(torch.randn(10, 3), x, proposal=None)
    assert inf._theta_roundwise[0].shape[0] == 9 and inf._prior_masks[0].sum() == 9
    inf.append_simulations(torch.randn(4, 3), torch.randn(4, 5), proposal=object())
    assert inf._data_round_index == [0, 1] and inf._prior_masks[1].sum() == 0
    inf2 = pickle.loads(pickle.dumps(inf))                      # floppity.py:247 pickles the inference object
    assert inf2._data_round_index == [0, 1]
    with pytest.raises(ValueError):
        inf.append_simulations(torch.randn(4, 3), torch.randn(5, 5))


def test_synthetic_configs_have_baseline_shapes():
    for name, (D, C) in {"c1": (10, 100), "c2": (20, 500), "c3": (30, 2000), "c4": (40, 3000)}.items():
        th, x, scaler, sigma = synthetic.make_dataset(name, 64)
        assert th.shape == (64, D) and x.shape == (64, C) and th.dtype == np.float32
        assert np.abs(th).max() <= 3.0 and abs(x.mean()) < 1e-3 and abs(x.std() - 1) < 1e-2
    th4 = synthetic.make_dataset("c4", 64)[0]
    assert (th4[:, 2] >= th4[:, 20]).all()                      # hotter component first (modules.py:110-116)
    xo, tt = synthetic.observation("c1", scaler=synthetic.make_dataset("c1", 64)[2])
    assert xo.shape == (1, 100)


def test_host_build_of_spline_math_matches_oracle(tmp_path):
    """g++ build of csrc/rqs_math.cuh (the arithmetic the kernels inline) vs the oracle: bins exact."""
    so = tmp_path / "hostmath.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                           os.path.join(ROOT, "tests", "hostmath", "hostmath.cpp")])
    lib = ctypes.CDLL(str(so))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    torch.manual_seed(1)
    for K, H in [(8, 50), (10, 128), (16, 512)]:
        n, Pn = 20000, 3 * K - 1
        x = torch.rand(n) * 7 - 3.5
        x[:4] = torch.tensor([-3.0, 3.0, 3.0000002, -3.0000002])
        params = torch.randn(n, Pn)
        s = np.sqrt(H)
        xr, pr = x.clone().requires_grad_(), params.clone().requires_grad_()
        y, lad, bins, _ = O.unconstrained_rational_quadratic_spline(xr, pr[:, :K] / s, pr[:, K:2 * K] / s, pr[:, 2 * K:],
                                                                  tail_bound=3.0, return_bins=True)
        gy, gl = torch.randn(n), torch.randn(n)
        (y * gy + lad * gl).sum().backward()
        xn, pn = x.numpy().copy(), params.numpy().copy()
        yo, lo, bo = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        lib.hm_rqs_forward(P(xn), P(pn), n, K, H, ctypes.c_float(3.0), P(yo), P(lo), P(bo))
        assert (bo == bins.numpy()).all()
        assert np.abs(yo - y.detach().numpy()).max() < 2e-5 and np.abs(lo - lad.detach().numpy()).max() < 2e-4
        gp, gx = np.zeros((n, Pn), np.float32), np.zeros(n, np.float32)
        lib.hm_rqs_backward(P(xn), P(pn), P(gy.numpy().copy()), P(gl.numpy().copy()), n, K, H, ctypes.c_float(3.0), P(gp), P(gx))
        rg = pr.grad.numpy()
        assert np.abs(gp - rg).max() / np.abs(rg).max() < 2e-4
        assert np.abs(gx - xr.grad.numpy()).max() / np.abs(xr.grad.numpy()).max() < 2e-4
        yi = y.detach()
        xi, li, bi, _ = O.unconstrained_rational_quadratic_spline(yi, params[:, :K] / s, params[:, K:2 * K] / s, params[:, 2 * K:],
                                                                inverse=True, tail_bound=3.0, return_bins=True)
        lib.hm_rqs_inverse(P(yi.numpy().copy()), P(pn), n, K, H, ctypes.c_float(3.0), P(yo), P(lo), P(bo))
        assert (bo == bi.numpy()).all() and np.abs(yo - xi.numpy()).max() < 1e-4


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from floppity_deprecated_b200 import dist as fdist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r, w = fdist.rank_world()
assert (r, w) == (int(sys.argv[3]), 2) and fdist.is_dist()
idx = torch.arange(11)
mine = fdist.shard_indices(idx)
assert mine.numel() == 5 and mine.tolist() == list(range(r, 10, 2))
g = torch.full((7,), float(r + 1)); fdist.allreduce_sum_(g); assert g.tolist() == [3.0] * 7       # gradient all-reduce
p = torch.full((4,), float(r)); fdist.broadcast_(p); assert p.tolist() == [0.0] * 4                # initial weights
v, c = fdist.agree_sum(float(r + 1), 10.0, "cpu"); assert (v, c) == (3.0, 20.0)                    # early-stop agreement
assert fdist.shared_seed(100 + r, "cpu") == 100
lo, hi = fdist.shard_rows(11); assert (lo, hi) == ((0, 6) if r == 0 else (6, 11))
rows = torch.arange(lo, hi, dtype=torch.float32)[:, None]
allr = fdist.gather_rows(rows); assert allr[:, 0].tolist() == list(range(11))
dist.destroy_process_group()
print("ok", r)
'''


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_early_stopping_restores_flow_and_embedding_together():
    """sbi deep-copies the whole state_dict at the best epoch; with an FC embedding in front of the flow (-embedding FC,
    /root/reference/modules.py:21-45) both parameter buffers must come back from the SAME epoch."""
    import floppity_deprecated_b200 as fb
    theta, x = torch.rand(64, 4) * 2 - 1, torch.randn(64, 12)
    net = fb.build_nsf(theta, x, hidden_features=16, num_transforms=2, num_bins=4, embedding_net=fb.FCEmbedding(12, 5, 2, 8))
    inf = fb.SNPE_C(None, "nsf")
    inf._neural_net = net
    inf._val_log_prob = 1.0
    assert not inf._converged(0, 2)
    best = [p.detach().clone() for p in net.parameters()]
    assert len(best) == 2
    for epoch in (1, 2):
        with torch.no_grad():
            for p in net.parameters():
                p.add_(0.1)                       # training moves on, validation gets worse
        inf._val_log_prob = 1.0 - 0.1 * epoch
        done = inf._converged(epoch, 2)
    assert done
    for p, b in zip(net.parameters(), best):
        assert torch.equal(p.detach(), b)
