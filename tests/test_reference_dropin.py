"""Drop-in check against the REFERENCE'S OWN glue code (only where /root/reference exists: the build container; the GPU
box has no copy, and nothing marked `gpu` reads it).  `modules.py` is imported unmodified with this package standing in
for the three sbi imports it makes (modules.py:1,3,7,8) -- exactly the import swap INTEGRATION.md describes -- and its
builder functions are driven the way floppity.py:118-132,179 drives them.  No CUDA work happens here: building the
prior, the embedding net, the estimator builder and the inference object is host-side."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import floppity_deprecated_b200 as fb

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "modules.py")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref_modules():
    saved = {k: sys.modules.get(k) for k in ("sbi", "sbi.utils", "sbi.inference", "sbi.utils.get_nn_models",
                                             "sbi.neural_nets", "sbi.neural_nets.embedding_nets", "modules")}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("sbi", utils=mod("sbi.utils", BoxUniform=fb.BoxUniform))
    mod("sbi.inference", SNPE_C=fb.SNPE_C)
    mod("sbi.utils.get_nn_models", posterior_nn=fb.posterior_nn)
    mod("sbi.neural_nets")
    from floppity_deprecated_b200 import embedding as E
    mod("sbi.neural_nets.embedding_nets", FCEmbedding=E.FCEmbedding, CNNEmbedding=E.CNNEmbedding,
        PermutationInvariantEmbedding=E.PermutationInvariantEmbedding)
    sys.path.insert(0, REF)
    sys.modules.pop("modules", None)
    import modules
    yield modules
    sys.path.remove(REF)
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_reference_builders_run_against_this_package(ref_modules, tmp_path):
    M = ref_modules
    D = 6
    bounds = np.stack([np.full(D, -2.0), np.linspace(1.0, 4.0, D)], axis=1)
    from floppity_deprecated_b200.retrieval import Normalizer
    yscaler = Normalizer(bounds, 3.0)                                   # floppity.py:123
    prior = M.create_prior(bounds, True, yscaler, 3.0, "cpu", str(tmp_path))   # floppity.py:127 -> modules.py:48-59
    assert isinstance(prior, fb.BoxUniform)
    assert torch.allclose(prior.low.cpu().double(), torch.full((D,), -3.0, dtype=torch.float64))
    assert torch.allclose(prior.high.cpu().double(), torch.full((D,), 3.0, dtype=torch.float64))

    summary = M.createEmbedding(False, "CNN", 64, "2, 4, 2, 64, 5", False, 100)     # floppity.py:118, default: no embedding
    assert isinstance(summary, torch.nn.Identity)
    inference = M.density_builder("nsf", 5, 50, 8, summary, 2, 0.1, prior, "cpu")  # floppity.py:132 -> modules.py:61-66
    assert isinstance(inference, fb.SNPE_C)

    fc = M.createEmbedding(True, "FC", 64, "2, 4, 2, 64, 5", False, 100)           # -embedding -embedding_type FC
    assert isinstance(fc, fb.FCEmbedding) and fc.input_dim == 100 and fc.output_dim == 64 and fc.num_layers == 4
    inference_fc = M.density_builder("nsf", 5, 50, 8, fc, 2, 0.0, prior, "cpu")
    net = inference_fc._build_neural_net(torch.rand(64, D) * 6 - 3, torch.randn(64, 100))
    assert isinstance(net, fb.EmbeddedNSF) and net.flow.C == 64
    # the CNN branch cannot run from modules.py as shipped, even upstream: it calls unroll_embed_hypers, which lives in
    # floppityFUN.py and is never imported by modules.py (NameError at modules.py:28)
    with pytest.raises(NameError):
        M.createEmbedding(True, "CNN", "64", "2, 4, 2, 64, 5", False, 100)
    # with the one-line fix a maintainer would make (import the helper), it builds THIS package's CNNEmbedding with the
    # reference's default hyper-parameters ('2, 4, 2, 64, 5': two conv layers of 4 and 8 channels, kernel 5, 2 x 64 linear)
    import floppityFUN as FUN
    M.unroll_embed_hypers = FUN.unroll_embed_hypers
    try:
        cnn = M.createEmbedding(True, "CNN", "64", "2, 4, 2, 64, 5", False, 100)
    finally:
        del M.unroll_embed_hypers
    assert isinstance(cnn, fb.CNNEmbedding) and cnn.channels == [1, 4, 8] and cnn.kernel_size == 5 and cnn.output_dim == 64
    assert cnn.lengths == [100, 49, 23] and cnn.core.shapes == [(64, 8 * 23), (64, 64)]
    net_cnn = M.density_builder("nsf", 5, 50, 8, cnn, 2, 0.0, prior, "cpu")._build_neural_net(torch.rand(64, D) * 6 - 3,
                                                                                          torch.randn(64, 100))
    assert isinstance(net_cnn, fb.EmbeddedNSF) and net_cnn.flow.C == 64
    # -embedding_type multi: the reference's own multiNet class over this package's CNNEmbedding -- its sub-nets sit in a
    # plain list (floppityFUN.py:126-136), so nothing is registered: a fixed feature map
    multi = FUN.multiNet([40, 60], input_shape=(100,), num_conv_layers=2, out_channels_per_layer=[4, 8], kernel_size=5,
                         num_linear_layers=2, num_linear_units=64, output_dim=[8, 8])
    assert len(multi.nets) == 2 and all(isinstance(n, fb.CNNEmbedding) for n in multi.nets) and not list(multi.parameters())

    theta0 = M.sample_proposal(0, None, 256, bounds, 3.0, "aintnothinhere", 0, 0)  # floppity.py:179, round 0: Sobol
    assert theta0.shape == (256, D) and np.abs(theta0).max() <= 3.0
    # the builder the reference stored is picklable (floppity.py:247 pickles the inference object every round)
    import pickle
    inference.append_simulations(torch.tensor(theta0, dtype=torch.float32), torch.randn(256, 100))
    pickle.loads(pickle.dumps(inference))
