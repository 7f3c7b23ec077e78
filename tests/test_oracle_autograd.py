"""Independent second opinion on the oracle's analytic gradient and Adam: torch fp64 autograd."""
import numpy as np
import torch

from oracle import c_oracle as CO
from oracle import ppo_oracle as O


def _case(rng, nb=9, nhe=5, apa=4, nf=8, H=16, L=2, dtype=np.float64):
    A = nhe * apa
    pol = O.Policy(nf, H, L, apa, rng=rng, dtype=dtype)
    for b in pol.b:
        b[...] = rng.normal(0, 0.1, b.shape)
    feat = rng.integers(-3, 9, (nb, nhe, nf)).astype(dtype)
    mask = np.where(rng.random((nb, A)) < 0.3, -np.inf, 0.0).astype(dtype)
    mask[:, 0] = 0
    act = np.array([rng.choice(np.flatnonzero(mask[b] == 0)) + 1 for b in range(nb)])
    old = rng.uniform(0.01, 0.5, nb).astype(dtype)
    adv = rng.normal(size=nb).astype(dtype)
    return pol, feat, mask, act, old, adv


def _torch_loss(pol, feat, mask, act, old, adv, eps, w):
    Ws = [torch.tensor(np.asarray(x, np.float64), requires_grad=True) for x in pol.W]
    bs = [torch.tensor(np.asarray(x, np.float64), requires_grad=True) for x in pol.b]
    nb = feat.shape[0]
    h = torch.tensor(np.asarray(feat, np.float64))
    for l, (W, b) in enumerate(zip(Ws, bs)):
        h = h @ W + b
        if l < len(Ws) - 1:
            h = torch.nn.functional.leaky_relu(h, 0.01)
    z = h.reshape(nb, -1) + torch.tensor(np.asarray(mask, np.float64))
    p = torch.softmax(z, -1)
    sel = p[torch.arange(nb), torch.tensor(act - 1)]
    advt = torch.tensor(np.asarray(adv, np.float64))
    gain = sel / torch.tensor(np.asarray(old, np.float64)) * advt
    clip = torch.where(advt >= 0, (1 + eps) * advt, (1 - eps) * advt)
    ppo = -torch.minimum(gain, clip).mean()
    A = p.shape[1]
    sp = p + 1e-8 / A     # (1 - 1f-8) == 1 in Float32; keep the same form in fp64
    ent = (sp * sp.log()).sum(-1).mean()
    (ppo + w * ent).backward()
    return ppo.item(), ent.item(), [W.grad.numpy() for W in Ws], [b.grad.numpy() for b in bs]


def test_gradient_matches_autograd_fp64():
    rng = np.random.default_rng(3)
    for L in (1, 2, 3):
        pol, feat, mask, act, old, adv = _case(rng, L=L)
        pl, entw, dW, db = O.policy_gradient(pol, feat, mask, act, old, adv, 0.05, 0.01)
        tp, te, tW, tb = _torch_loss(pol, feat, mask, act, old, adv, 0.05, 0.01)
        assert abs(pl - tp) < 1e-12 and abs(entw - 0.01 * te) < 1e-9
        for a, b in zip(dW + db, tW + tb):
            assert np.max(np.abs(a - b)) < 1e-9


def test_fp32_oracle_close_to_fp64_truth():
    rng = np.random.default_rng(4)
    pol, feat, mask, act, old, adv = _case(rng, nb=64, dtype=np.float64)
    pl, entw, dW, db = O.policy_gradient(pol, feat, mask, act, old, adv, 0.05, 0.01)
    p32 = pol.copy()
    p32.W = [w.astype(np.float32) for w in pol.W]
    p32.b = [b.astype(np.float32) for b in pol.b]
    pl32, entw32, dW32, db32 = O.policy_gradient(p32, feat.astype(np.float32), mask.astype(np.float32), act,
                                                 old.astype(np.float32), adv.astype(np.float32), 0.05, 0.01)
    assert abs(pl - pl32) <= 1e-5 * abs(pl) + 1e-6
    for a, b in zip(dW + db, dW32 + db32):
        assert np.max(np.abs(a - b)) <= 1e-4 * np.max(np.abs(a)) + 1e-6


def test_adam_numpy_vs_c_and_torch():
    rng = np.random.default_rng(5)
    n = 1000
    x = rng.normal(size=n).astype(np.float32)
    opt = O.Adam(1e-3)
    x_np, x_c = x.copy(), x.copy()
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    xt = torch.tensor(x.astype(np.float64), requires_grad=True)
    topt = torch.optim.Adam([xt], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    b1p, b2p = 0.9, 0.999
    for step in range(5):
        g = rng.normal(size=n).astype(np.float32)
        opt.apply("x", x_np, g)
        CO.adam(x_c, m, v, g, 1e-3, 0.9, 0.999, 1e-8, b1p, b2p)
        b1p *= 0.9; b2p *= 0.999
        xt.grad = torch.tensor(g.astype(np.float64)); topt.step()
        assert np.array_equal(x_np, x_c)
    # Flux's Adam == the textbook (torch) Adam up to Float32 rounding of the state
    assert np.max(np.abs(x_np - xt.detach().numpy())) < 1e-6
