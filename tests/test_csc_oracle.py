"""The CSC oracle against itself: the conv-by-conv transcription of model.jl (NNlib conventions) and the
position-space form used by the CUDA kernels must agree on loss and gradients; frozen golden vectors pin both.
CPU only; small hyper-parameters keep the literal form (which really builds the CS_vlen-long correlations of
model.jl:270-273) fast."""
import os

import numpy as np
import pytest
import torch

from motifs_jl_b200 import synth
from oracle import csc_oracle as co, scan_oracle as so

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL = dict(filter_len=4, M=5, h=3, K=4, q=3, batch_size=3, num_pass_xyz=2, num_pass_df=2)


@pytest.mark.parametrize("seed,Lb", [(1, 24), (2, 31)])
def test_literal_equals_position_space_fp64(seed, Lb):
    hp = co.Hyperparam(**SMALL)
    flat = co.init_params(hp, seed)
    codes = so.ascii_to_codes(synth.random_ascii(hp.batch_size, Lb, seed))
    l1, g1, a1 = co.loss_and_grad(codes, flat, hp, "literal", torch.float64)
    l2, g2, a2 = co.loss_and_grad(codes, flat, hp, "pos", torch.float64)
    assert l1 == pytest.approx(l2, rel=1e-12)
    assert np.abs(g1 - g2).max() <= 1e-11 * max(1.0, np.abs(g1).max())
    assert (g1[: co.n_params(hp)] != 0).any() and np.all(g1[co.n_params(hp):] == 0)      # warm-up scalars are not parameters
    # Z in Julia layout (C, M, B) rows 4p  ==  z (B, c, M)
    assert np.allclose(a1["Z"].detach().numpy()[0::4].transpose(2, 0, 1), a2["z"].detach().numpy(), atol=1e-12)
    assert np.allclose(a1["X"].detach().numpy()[:, 0].transpose(2, 0, 1), a2["x"].detach().numpy(), atol=1e-12)


def test_default_shape_fp32_forms_agree():
    hp = co.Hyperparam(num_pass_xyz=2, num_pass_df=1)                    # default sizes, fewer passes (keeps the literal form quick)
    flat = co.init_params(hp, 3)
    codes = so.ascii_to_codes(synth.planted_gapped(6, 100, 3))
    l1, g1, _ = co.loss_and_grad(codes, flat, hp, "literal", torch.float32)
    l2, g2, _ = co.loss_and_grad(codes, flat, hp, "pos", torch.float32)
    assert l1 == pytest.approx(l2, rel=1e-5)
    assert np.abs(g1 - g2).max() <= 1e-4 * np.abs(g1).max()


def test_selection_semantics():
    v = torch.tensor([0.0, 3.0, -1.0, 1.0, 2.0, 0.0])
    assert float(co.median_of_positives(v)) == 2.0                        # odd count: the middle positive
    assert float(co.median_of_positives(torch.tensor([4.0, 1.0, 0.0, 2.0, 8.0]))) == 3.0   # even: a/2 + b/2
    assert co.median_of_positives(torch.tensor([0.0, -1.0])) is None      # no positives -> no mask (model.jl:197-198)
    x = torch.tensor([[[5.0, 1.0], [3.0, 3.0], [0.0, -2.0]]])
    assert torch.equal(co.topq(x, 2), torch.tensor([[[5.0, 0.0], [3.0, 3.0], [0.0, 0.0]]]))   # ties with the q-th largest are kept


def test_init_and_flat_layout():
    hp = co.Hyperparam()
    flat = co.init_params(hp, 1)
    assert flat.shape == (30433 + 3,) and co.n_params(hp) == 30433
    P = co.unpack(torch.tensor(flat), hp)
    D = P["D"].numpy()
    assert D.shape == (32, 50) and np.allclose((D ** 2).reshape(8, 4, 50).sum(axis=1), 1.0, atol=1e-5)   # sqrt of simplex columns
    assert P["F"].shape == (12, 100, 24) and float(P["F"].min()) >= 0


def test_golden_csc():
    g = np.load(os.path.join(GOLD, "csc_config1.npz"))
    hp = co.Hyperparam()
    loss, grad, aux = co.loss_and_grad(g["codes"], g["flat"], hp, "pos", torch.float32)
    assert loss == pytest.approx(float(g["loss"]), rel=1e-6)
    assert np.abs(grad - g["grad"]).max() <= 1e-5 * np.abs(g["grad"]).max()
    assert np.array_equal(aux["x"].detach().numpy() != 0, g["x"] != 0)
