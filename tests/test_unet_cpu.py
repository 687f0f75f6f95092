"""CPU tests pinning the test-only torch U-Net checker (tests/unet_ref.py) to the oracle, operator by operator: the
checker is only trusted for the device U-Net's parity tests because its conv / group norm / softmax agree with the
restatement of lib/conv.c, lib/norm.c and lib/util.c (itself pinned to the compiled reference)."""
import numpy as np
import pytest
import torch

from helpers import load_oracle, ptr, rel_err
import unet_ref


@pytest.mark.parametrize("C,H,W,F,k,s", [(5, 8, 8, 7, 3, 1), (6, 9, 7, 4, 3, 2), (8, 16, 16, 8, 1, 1), (4, 8, 8, 6, 3, 2)])
def test_checker_conv_matches_oracle(C, H, W, F, k, s):
    o = load_oracle(np.float64)
    rng = np.random.default_rng(C + H + F + k + s)
    x = rng.normal(size=(C, H, W)); w = rng.normal(size=(F, C, k, k))
    Ho, Wo = -(-H // s), -(-W // s)
    want = np.empty((F, Ho, Wo))
    o.orc_conv(C, H, W, F, k, s, ptr(x), ptr(w), ptr(want))
    got = unet_ref.conv(torch.tensor(x)[None], torch.tensor(w), s)[0].numpy()
    assert rel_err(got, want) <= 1e-12


@pytest.mark.parametrize("C,HW,gs,quirk", [(64, 16, 32, 1), (3, 64, 32, 1), (40, 16, 32, 1), (64, 16, 32, 0)])
def test_checker_group_norm_matches_oracle(C, HW, gs, quirk):
    o = load_oracle(np.float64)
    rng = np.random.default_rng(C + HW)
    x = rng.normal(1.0, 2.0, size=(C, HW))
    G = -(-C // gs)
    want = np.empty_like(x); sd = np.empty(G); mu = np.empty(G)
    o.orc_group_norm(C, HW, gs, ptr(x), ptr(want), ptr(sd), ptr(mu), quirk)
    side = int(round(HW ** 0.5))
    got = unet_ref.group_norm(torch.tensor(x).view(1, C, side, side), gs, quirk)[0].reshape(C, HW).numpy()
    assert rel_err(got, want) <= 1e-12


def test_checker_softmax_rows_matches_oracle():
    o = load_oracle(np.float64)
    rng = np.random.default_rng(3)
    a = rng.normal(size=(16, 16))
    want = a.copy()
    o.orc_softmax_rows(16, 16, ptr(want))
    assert rel_err(torch.softmax(torch.tensor(a), dim=-1).numpy(), want) <= 1e-12
