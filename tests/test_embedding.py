"""Row N1 of the scope table: the FC embedding net in front of the flow (modules.py:21-45, -embedding FC),
trained jointly.  CPU: construction / checkpoint keys against the oracle; GPU: log_prob, samples and one training
step's gradients (flow AND embedding parameters) against torch autograd on the oracle."""
import numpy as np
import pytest
import torch

import floppity_deprecated_b200 as fb
from conftest import rel_err
from oracle import nsf_oracle as O

D, C, E = 6, 40, 12
HP = dict(hidden_features=32, num_transforms=3, num_bins=6, num_blocks=2)


def _pair(seed=0, hidden=HP["hidden_features"]):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    theta = torch.rand(256, D, generator=g) * 6 - 3
    x = torch.randn(256, C, generator=g) * 2.0 + 0.5
    hp = dict(HP, hidden_features=hidden)
    onet = O.build_nsf(theta, x, embedding_net=O.FCEmbedding(C, E, 3, 24), **hp)
    O.randomize_lu(onet, g, 0.1)
    net = fb.build_nsf(theta, x, embedding_net=fb.FCEmbedding(C, E, 3, 24), **hp)
    net.load_state_dict(onet.state_dict())
    return onet, net, theta, x


def test_embedded_flow_has_upstream_checkpoint_layout():
    onet, net, theta, x = _pair()
    assert isinstance(net, fb.EmbeddedNSF) and net.C == C and net.flow.C == E and net.D == D
    want = {k: tuple(v.shape) for k, v in onet.state_dict().items() if not k.endswith("_log_z")}
    got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert got == want
    sd = net.state_dict()
    for k, v in onet.state_dict().items():
        if not k.endswith("_log_z"):
            assert torch.equal(sd[k].cpu().float(), v.float()), k
    import copy
    import pickle
    net2 = pickle.loads(pickle.dumps(copy.deepcopy(net)))
    assert torch.equal(net2.embedding._flat, net.embedding._flat)
    with pytest.raises(NotImplementedError):
        fb.build_nsf(theta, x, embedding_net=torch.nn.Linear(C, E), **HP)
    with pytest.raises(ValueError):
        fb.build_nsf(theta, x, embedding_net=fb.FCEmbedding(C + 1, E), **HP)


@pytest.mark.gpu
@pytest.mark.parametrize("hidden", [32, 128])        # 128: the fused tcgen05 conditioner behind the embedding
def test_embedded_log_prob_and_samples_match_oracle(hidden):
    onet, net, theta, x = _pair(1, hidden)
    net = net.cuda().eval(); onet.eval()
    with torch.no_grad():
        want = onet.log_prob(theta[:64], x[:64]).numpy()
        got = net.log_prob(theta[:64].cuda(), x[:64].cuda()).cpu().numpy()
    assert rel_err(got, want).max() < 1e-4
    z = torch.randn(50, D, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ws = onet.sample(50, context=x[:1], noise=z)[0].numpy()
        s = net.sample(50, context=x[:1].cuda(), noise=z.cuda())[0].cpu().numpy()
    assert rel_err(s, ws).max() < 1e-4
    # the posterior object drives it the same way (floppity.py:225,252)
    prior = fb.BoxUniform(-3 * torch.ones(D), 3 * torch.ones(D), device="cuda")
    post = fb.DirectPosterior(net, prior).set_default_x(x[:1])
    assert post.sample((200,)).shape == (200, D)
    lp = post.log_prob(theta[:64], norm_posterior=False).cpu().numpy()          # every row against the default x
    with torch.no_grad():
        want_o = onet.log_prob(theta[:64], x[:1].repeat(64, 1)).numpy()
    assert rel_err(lp[np.isfinite(lp)], want_o[np.isfinite(lp)]).max() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("hidden", [32, 128])
def test_joint_training_step_gradients_match_autograd(hidden):
    from floppity_deprecated_b200.inference import _TrainEngine
    onet, net, theta, x = _pair(3, hidden)
    net = net.cuda().train(); onet.train()
    B = 128
    prior = fb.BoxUniform(-3 * torch.ones(D), 3 * torch.ones(D), device="cuda")
    eng = _TrainEngine(net, prior, B, False, 10, True, 5e-4, 5.0)
    before_flow, before_emb = net.flow._flat.detach().clone(), net.embedding._flat.detach().clone()
    eng.step_on_batch(theta[:B].cuda(), net._standardize(x[:B].cuda()))
    loss = -onet.log_prob(theta[:B], x[:B]).mean()
    loss.backward()
    assert abs(eng.loss_sum.item() / B - loss.item()) < 1e-4 * max(1.0, abs(loss.item()))
    got = dict(net.flow.unpack(eng.grads[: net.flow.param_count]))
    base = "_embedding_net.1."
    got.update({base + k: v for k, v in net.embedding.unpack(eng.emb_grads).items()})
    worst = 0.0
    for k, p in onet.named_parameters():
        assert k in got, k
        scale = max(float(p.grad.abs().max()), 1e-4)
        worst = max(worst, float((got[k].cpu() - p.grad).abs().max()) / scale)
    assert worst < 3e-4, worst          # (measured 3e-5; no knife-edge bookkeeping through the embedding net)
    assert float((net.flow._flat.detach() - before_flow).abs().max()) > 1e-5
    assert float((net.embedding._flat.detach() - before_emb).abs().max()) > 1e-5     # the embedding is trained too


@pytest.mark.gpu
def test_snpe_c_trains_with_an_fc_embedding():
    """modules.py:25,63-64 + floppity.py:216-225 with -embedding -embedding_type FC"""
    from floppity_deprecated_b200 import synthetic
    theta, x, _, _ = synthetic.make_dataset("c1", 2048)
    prior = fb.BoxUniform(-3 * torch.ones(10), 3 * torch.ones(10), device="cuda")
    summary = fb.FCEmbedding(input_dim=100, num_hiddens=256, num_layers=4, output_dim=64)
    inf = fb.SNPE_C(prior, density_estimator=fb.posterior_nn(model="nsf", hidden_features=50, num_transforms=5, num_bins=8,
                                                            embedding_net=summary, num_blocks=2), device="cuda")
    inf.append_simulations(torch.tensor(theta), torch.tensor(x))
    est = inf.train(training_batch_size=256, max_num_epochs=6, stop_after_epochs=6, seed=0)
    lps = inf._summary["validation_log_probs"]
    assert lps[-1] > lps[0] + 0.5, lps            # it learns
    post = inf.build_posterior(est).set_default_x(torch.tensor(x[:1]))
    assert post.sample((1000,)).shape == (1000, 10)


# ---------------------------------------------------------------------------------------------------------------------
# -embedding_type CNN (the reference's default type, floppity.py:46-47; modules.py:26-33) and -embedding_type multi
# ---------------------------------------------------------------------------------------------------------------------
CNN_KW = dict(out_channels_per_layer=[4, 8], num_conv_layers=2, num_linear_layers=2, num_linear_units=64, output_dim=E,
              kernel_size=5)          # unroll_embed_hypers('2, 4, 2, 64, 5') -- floppity.py:47's default


def _cnn_pair(seed=0, C=C, hidden=32, kw=CNN_KW):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    theta = torch.rand(256, D, generator=g) * 6 - 3
    x = torch.randn(256, C, generator=g) * 2.0 + 0.5
    hp = dict(HP, hidden_features=hidden)
    onet = O.build_nsf(theta, x, embedding_net=O.CNNEmbedding(input_shape=(C,), **kw), **hp)
    O.randomize_lu(onet, g, 0.1)
    net = fb.build_nsf(theta, x, embedding_net=fb.CNNEmbedding(input_shape=(C,), **kw), **hp)
    net.load_state_dict(onet.state_dict())
    return onet, net, theta, x


def test_cnn_embedding_has_upstream_shapes_and_checkpoint_keys():
    onet, net, theta, x = _cnn_pair()
    emb = net.embedding
    assert isinstance(emb, fb.CNNEmbedding) and emb.input_dim == C and emb.output_dim == E
    # (40 + 2 - 4) = 38 -> pool 19 -> (19 + 2 - 4) = 17 -> pool 8: 8 channels x 8 = 64 flattened features
    assert emb.lengths == [40, 19, 8] and emb.core.input_dim == 64
    want = {k: tuple(v.shape) for k, v in onet.state_dict().items() if not k.endswith("_log_z")}
    got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert got == want
    sd = net.state_dict()
    for k, v in onet.state_dict().items():
        if not k.endswith("_log_z"):
            assert torch.equal(sd[k].cpu().float(), v.float()), k
    import pickle
    net2 = pickle.loads(pickle.dumps(net))
    assert torch.equal(net2.embedding._flat, net.embedding._flat)
    with pytest.raises(NotImplementedError):
        fb.CNNEmbedding(input_shape=(8, 8))
    with pytest.raises(ValueError):
        fb.CNNEmbedding(input_shape=(6,), out_channels_per_layer=[2, 2, 2], num_conv_layers=3)


@pytest.mark.gpu
@pytest.mark.parametrize("C_,pool", [(40, 2), (101, 3)])
def test_conv_layer_kernels_match_torch(C_, pool):
    """One Conv1d(padding 1) -> ReLU -> MaxPool1d layer, forward and all three gradients, against torch autograd."""
    from floppity_deprecated_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(C_)
    B, Cin, Cout, k = 37, 3, 5, 5
    x = torch.randn(B, Cin, C_, generator=g)
    conv = torch.nn.Conv1d(Cin, Cout, k, padding=1)
    xr = x.clone().requires_grad_(True)
    y = torch.nn.functional.max_pool1d(torch.relu(conv(xr)), pool)
    dy = torch.randn(y.shape, generator=g)
    (y * dy).sum().backward()
    Lc = C_ + 2 - k + 1
    xd, w, b = x.cuda(), conv.weight.detach().cuda().contiguous(), conv.bias.detach().cuda()
    z = torch.empty(B, Cout, Lc, device="cuda"); yo = torch.empty(B, Cout, Lc // pool, device="cuda")
    _lib.check(L.fp_conv1d_relu_pool_fwd(_lib.ptr(xd), _lib.ptr(w), _lib.ptr(b), _lib.ptr(z), _lib.ptr(yo), B, Cin, C_, Cout, k, 1,
                                         pool, _lib.stream()), "fwd")
    assert rel_err(yo.cpu().numpy(), y.detach().numpy()).max() < 1e-5
    dz, dx = torch.empty_like(z), torch.empty_like(xd)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    _lib.check(L.fp_conv1d_relu_pool_bwd(_lib.ptr(xd), _lib.ptr(w), _lib.ptr(z), _lib.ptr(dy.cuda()), _lib.ptr(dz), _lib.ptr(dx),
                                         _lib.ptr(dw), _lib.ptr(db), B, Cin, C_, Cout, k, 1, pool, _lib.stream()), "bwd")
    assert rel_err(dx.cpu().numpy(), xr.grad.numpy()).max() < 1e-5
    assert rel_err(dw.cpu().numpy(), conv.weight.grad.numpy()).max() < 1e-4
    assert rel_err(db.cpu().numpy(), conv.bias.grad.numpy()).max() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("hidden", [32, 128])
def test_cnn_embedded_log_prob_samples_and_joint_gradients(hidden):
    from floppity_deprecated_b200.inference import _TrainEngine
    onet, net, theta, x = _cnn_pair(3, hidden=hidden)
    net, onet = net.cuda().eval(), onet.eval()
    with torch.no_grad():
        got = net.log_prob(theta[:100].cuda(), x[:100].cuda()).cpu().numpy()
        want = onet.log_prob(theta[:100], x[:100]).numpy()
    assert rel_err(got, want).max() < 1e-4, rel_err(got, want).max()
    z = torch.randn(50, D, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        s = net.sample(50, context=x[:1].cuda(), noise=z.cuda())[0].cpu().numpy()
        ws = onet.sample(50, context=x[:1], noise=z)[0].numpy()
    assert rel_err(s, ws).max() < 1e-4
    # one joint training step: gradients of the flow AND of every convolution / linear tensor of the embedding
    net.train(); onet.train()
    B = 128
    prior = fb.BoxUniform(-3 * torch.ones(D), 3 * torch.ones(D), device="cuda")
    eng = _TrainEngine(net, prior, B, False, 10, True, 5e-4, 5.0)
    before = net.embedding._flat.detach().clone()
    eng.step_on_batch(theta[:B].cuda(), net._standardize(x[:B].cuda()))
    loss = -onet.log_prob(theta[:B], x[:B]).mean()
    loss.backward()
    assert abs(eng.loss_sum.item() / B - loss.item()) < 1e-4 * max(1.0, abs(loss.item()))
    got_g = dict(net.flow.unpack(eng.grads[: net.flow.param_count]))
    got_g.update({"_embedding_net.1." + k: v for k, v in net.embedding.unpack(eng.emb_grads).items()})
    worst = 0.0
    for k, p in onet.named_parameters():
        assert k in got_g, k
        worst = max(worst, float((got_g[k].cpu() - p.grad).abs().max()) / max(float(p.grad.abs().max()), 1e-4))
    assert worst < 5e-4, worst          # (a max-pool arg-max or relu on a knife edge moves a whole term: no bookkeeping here)
    assert float((net.embedding._flat.detach() - before).abs().max()) > 1e-5


@pytest.mark.gpu
def test_multi_embedding_is_a_fixed_feature_map_like_the_reference():
    """-embedding_type multi (floppityFUN.py:93-134): the reference keeps its per-observation CNNEmbeddings in a plain list,
    so they are never registered or trained.  Same here: evaluated as they stand, only the flow trains."""
    g = torch.Generator().manual_seed(9)
    torch.manual_seed(9)
    nwvl = [24, 30]
    theta = torch.rand(256, D, generator=g) * 6 - 3
    x = torch.randn(256, sum(nwvl), generator=g)
    kw = dict(out_channels_per_layer=[2, 4], num_conv_layers=2, num_linear_layers=2, num_linear_units=16, kernel_size=3)
    omulti = O.MultiNet(nwvl, lambda i, n: O.CNNEmbedding(input_shape=(n,), output_dim=5 + i, **kw))
    multi = O.MultiNet(nwvl, lambda i, n: fb.CNNEmbedding(input_shape=(n,), output_dim=5 + i, **kw))
    for a, b in zip(multi.nets, omulti.nets):
        a.load_state_dict(b.state_dict())
    assert len(list(multi.parameters())) == 0                    # the reference's bug, reproduced: nothing registered
    hp = dict(HP)
    onet = O.build_nsf(theta, x, embedding_net=omulti, **hp)
    net = fb.build_nsf(theta, x, embedding_net=multi, **hp)
    assert net.flow.C == 11
    net.load_state_dict(onet.state_dict())
    net, onet = net.cuda().eval(), onet.eval()
    with torch.no_grad():
        got = net.log_prob(theta[:64].cuda(), x[:64].cuda()).cpu().numpy()
        want = onet.log_prob(theta[:64], x[:64]).numpy()
    assert rel_err(got, want).max() < 1e-4
    prior = fb.BoxUniform(-3 * torch.ones(D), 3 * torch.ones(D), device="cuda")
    inf = fb.SNPE_C(prior, lambda t, xx: net, device="cuda")
    inf.append_simulations(theta, x)
    frozen = [n._flat.detach().clone() for n in multi.nets]
    inf.train(training_batch_size=64, max_num_epochs=2, stop_after_epochs=5)
    assert all(torch.equal(a, n._flat.detach()) for a, n in zip(frozen, multi.nets))     # untouched, as upstream
    assert len(inf._summary["validation_log_probs"]) == 3
