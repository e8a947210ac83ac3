"""Torch emulation of the vgg_64 kernels (p2pvg_b200/csrc/vgg.cu and p2pvg_conv_gemm kinds 3-5) on top of
tests/emu_backend.py.  TEST INFRASTRUCTURE ONLY."""
import torch
import torch.nn.functional as F

from tests.emu_backend import EmuKernels, _flat


def _nhwc(t, N, H, W, C):
    return _flat(t, N * H * W * C).reshape(N, H, W, C).float()


def _taps(xv, sgn):
    """[N,H,W,C] -> [N,H,W,9,C]: tap (kh,kw) reads pixel + sgn*(kh-1, kw-1), zero outside."""
    N, H, W, C = xv.shape
    xp = F.pad(xv, (0, 0, 1, 1, 1, 1))
    out = []
    for kh in range(3):
        for kw in range(3):
            oy, ox = 1 + sgn * (kh - 1), 1 + sgn * (kw - 1)
            out.append(xp[:, oy:oy + H, ox:ox + W, :])
    return torch.stack(out, 3)


class EmuKernelsVGG(EmuKernels):
    def im2col3(self, x, col, N, H, W, C, ld, sgn=1):
        t = _taps(_nhwc(x, N, H, W, C), sgn).reshape(N * H * W, 9 * C)
        full = torch.zeros(N * H * W, ld, device=t.device)
        full[:, :9 * C] = t
        _flat(col, N * H * W * ld).copy_(full.reshape(-1).to(col.dtype))

    def col2im3(self, col, y, N, H, W, C, ld, bias=None):
        c = _flat(col, N * H * W * ld).reshape(N, H, W, ld).float()
        out = torch.zeros(N, H + 2, W + 2, C, device=c.device)
        for kh in range(3):
            for kw in range(3):
                tap = kh * 3 + kw
                out[:, kh:kh + H, kw:kw + W, :] += c[..., tap * C:(tap + 1) * C]
        out = out[:, 1:-1, 1:-1, :]
        if bias is not None:
            out = out + bias[:C].float()
        _flat(y, N * H * W * C).copy_(out.reshape(-1).to(y.dtype))

    @staticmethod
    def _windows(xv):
        N, H, W, C = xv.shape
        return xv.reshape(N, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 5, 2, 4).reshape(N, H // 2, W // 2, C, 4)

    def maxpool2_fwd(self, x, y, N, H, W, C):
        w = self._windows(_nhwc(x, N, H, W, C))
        _flat(y, N * (H // 2) * (W // 2) * C).copy_(w.max(-1).values.reshape(-1).to(y.dtype))

    def maxpool2_bwd(self, x, dy, dx, N, H, W, C):
        w = self._windows(_nhwc(x, N, H, W, C))
        is_max = w == w.max(-1, keepdim=True).values
        first = is_max & (is_max.long().cumsum(-1) == 1)
        g = _nhwc(dy, N, H // 2, W // 2, C).unsqueeze(-1) * first.float()
        out = g.reshape(N, H // 2, W // 2, C, 2, 2).permute(0, 1, 4, 2, 5, 3).reshape(N, H, W, C)
        _flat(dx, N * H * W * C).copy_(out.reshape(-1).to(dx.dtype))

    def upsample2_fwd(self, x, y, N, H, W, C):
        xv = _nhwc(x, N, H, W, C)
        out = xv.repeat_interleave(2, 1).repeat_interleave(2, 2)
        _flat(y, N * 4 * H * W * C).copy_(out.reshape(-1).to(y.dtype))

    def upsample2_bwd(self, dy, dx, N, H, W, C):
        d = _nhwc(dy, N, 2 * H, 2 * W, C).reshape(N, H, 2, W, 2, C)
        out = (d[:, :, 0, :, 0] + d[:, :, 0, :, 1]) + (d[:, :, 1, :, 0] + d[:, :, 1, :, 1])
        _flat(dx, N * H * W * C).copy_(out.reshape(-1).to(dx.dtype))

    def gather_add(self, dst, src, grp_src, G, n):
        d = _flat(dst, G * n).reshape(G, n)
        s = src.reshape(-1)[:(int(grp_src[:G].max().item()) + 1) * n].reshape(-1, n).float()
        d.copy_((d.float() + s[grp_src[:G].long()]).to(dst.dtype))

    def conv_gemm(self, kind, a, b, c, N, H, W, Ck, Cn, Cm=0, ldb=None, ldc=None, bias=None, addend=None, grp_src=None,
                  imgs_per_group=0, accumulate=False):
        """kinds 3-5 only (3x3 / stride 1); bf16 operands, fp32 accumulation."""
        assert kind in (3, 4, 5) and not accumulate
        M = N * H * W
        if kind in (3, 5):
            col = _taps(_nhwc(a, N, H, W, Ck), 1 if kind == 3 else -1).reshape(M, 9 * Ck)
            r = col @ _flat(b, Cn * 9 * Ck).reshape(Cn, 9 * Ck).float().t()
            if bias is not None:
                r = r + bias[:Cn].float()
            if addend is not None:
                n = torch.arange(N, device=r.device)
                n2 = grp_src.long()[n // imgs_per_group] * imgs_per_group + n % imgs_per_group
                add = addend.reshape(-1)[:(int(n2.max().item()) + 1) * H * W * Cn].reshape(-1, H * W * Cn).float()
                r = (r.reshape(N, H * W * Cn) + add[n2]).reshape(M, Cn)
            _flat(c, M * Cn).copy_(r.reshape(-1).to(c.dtype))
        else:
            col = _taps(_nhwc(b, N, H, W, Cn), 1).reshape(M, 9 * Cn)
            r = _flat(a, M * Cm).reshape(M, Cm).float().t() @ col
            _flat(c, Cm * 9 * Cn).copy_(r.reshape(-1).to(c.dtype))
