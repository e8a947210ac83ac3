"""Torch emulation of the h36m-backbone kernels (LayerNorm, ReLU, plain MSE) on top of tests/emu_backend.py.
TEST INFRASTRUCTURE ONLY."""
import torch
import torch.nn.functional as F

from tests.emu_backend import EmuKernels, _flat

ACT_RELU = 4


class EmuKernelsMLP(EmuKernels):
    def act_fwd(self, x, n, act):
        if act == ACT_RELU:
            v = _flat(x, n)
            v.copy_(torch.relu(v))
        else:
            super().act_fwd(x, n, act)

    def act_bwd(self, dy, y, dx, n, act):
        if act == ACT_RELU:
            _flat(dx, n).copy_(_flat(dy, n) * (_flat(y, n) > 0).float())
        else:
            super().act_bwd(dy, y, dx, n, act)

    def layernorm_fwd(self, x, gamma, beta, y, mean, rstd, rows, C, eps=1e-5):
        xv = _flat(x, rows * C).reshape(rows, C)
        m = xv.mean(1)
        v = ((xv - m[:, None]) ** 2).mean(1)
        r = torch.rsqrt(v + eps)
        _flat(mean, rows).copy_(m)
        _flat(rstd, rows).copy_(r)
        _flat(y, rows * C).copy_(((xv - m[:, None]) * r[:, None] * gamma[:C] + beta[:C]).reshape(-1))

    def layernorm_bwd(self, dy, x, mean, rstd, gamma, dx, dgamma, dbeta, rows, C):
        d = _flat(dy, rows * C).reshape(rows, C)
        xh = (_flat(x, rows * C).reshape(rows, C) - _flat(mean, rows)[:, None]) * _flat(rstd, rows)[:, None]
        if dgamma is not None:
            dgamma[:C].copy_((d * xh).sum(0))
            dbeta[:C].copy_(d.sum(0))
        g = d * gamma[:C]
        r = _flat(rstd, rows)[:, None] * (g - g.mean(1, keepdim=True) - xh * (g * xh).mean(1, keepdim=True))
        _flat(dx, rows * C).copy_(r.reshape(-1))

    def mse_plain(self, pred, x, tgt, coef, G, E, d_pred, partial):
        p = _flat(pred, G * E).reshape(G, E)
        xt = _flat(x, x.numel()).reshape(-1, E)[tgt[:G].long()]
        d = p - xt
        pp = _flat(partial, G * 32).reshape(G, 32)
        pp.zero_()
        pp[:, 0] = (d.double() ** 2).sum(1).float()
        if d_pred is not None:
            _flat(d_pred, G * E).copy_((coef[:G].reshape(G, 1) * 2 * d).reshape(-1))
