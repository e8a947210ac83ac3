"""Torch emulation of the kernel ABI (p2pvg_b200/_lib.py:CudaKernels).  TEST INFRASTRUCTURE ONLY.

Two uses: (1) on CPU it lets the host-side schedule in p2pvg_b200/engine.py be validated against the
oracle without a GPU; (2) on the GPU box every CUDA kernel is compared against the method of the same
name here.  Never imported by the product path.
"""
import torch
import torch.nn.functional as F

ACT_NONE, ACT_LRELU, ACT_TANH = 0, 1, 2


def _act(z, act):
    if act == ACT_LRELU:
        return torch.where(z > 0, z, 0.2 * z)
    if act == ACT_TANH:
        return torch.tanh(z)
    return z


def _act_grad(y, act):
    if act == ACT_LRELU:
        return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, 0.2))
    if act == ACT_TANH:
        return 1 - y * y
    return torch.ones_like(y)


def _flat(t, n):
    return t.reshape(-1)[:n]


def _mat(t, rows, cols, ld):
    """View the first rows x cols block of a row-major matrix with leading dimension ld."""
    return torch.as_strided(t, (rows, cols), (ld, 1), t.storage_offset())


class EmuKernels:
    name = "emu"

    def __init__(self, device="cpu"):
        self.device = torch.device(device)
        self.launches = 0

    def set_gemm_impl(self, impl):
        pass

    def gemm(self, A, B, C, M, N, K, a_mn=False, b_mn=False, lda=None, ldb=None, ldc=None, accumulate=False, bias=None,
             addend=None, ldd=None):
        lda = lda if lda is not None else (M if a_mn else K)
        ldb = ldb if ldb is not None else (N if b_mn else K)
        ldc = ldc if ldc is not None else N
        ldd = ldd if ldd is not None else N
        a = (_mat(A, K, M, lda).t() if a_mn else _mat(A, M, K, lda)).float()
        b = (_mat(B, K, N, ldb) if b_mn else _mat(B, N, K, ldb).t()).float()
        r = a @ b
        if bias is not None:
            r = r + bias[:N].float()
        if addend is not None:
            r = r + _mat(addend, M, N, ldd).float()
        c = _mat(C, M, N, ldc)
        if accumulate:
            r = r + c.float()
        c.copy_(r.to(C.dtype))

    def im2col(self, x, col, N, H, W, C):
        xv = _flat(x, N * H * W * C).reshape(N, H, W, C).float()
        xp = F.pad(xv, (0, 0, 1, 1, 1, 1))
        Ho, Wo = H // 2, W // 2
        taps = []
        for kh in range(4):
            for kw in range(4):
                taps.append(xp[:, kh:kh + 2 * Ho:2, kw:kw + 2 * Wo:2, :])
        out = torch.stack(taps, 3)  # N,Ho,Wo,16,C
        _flat(col, N * Ho * Wo * 16 * C).copy_(out.reshape(-1).to(col.dtype))

    def col2im(self, col, y, N, Hi, Wi, C, bias=None, col2=None, grp_src=None, imgs_per_group=0, accumulate=False):
        c = _flat(col, N * Hi * Wi * 16 * C).reshape(N, Hi, Wi, 4, 4, C).float()
        if col2 is not None:
            n = torch.arange(N, device=col.device)
            n2 = grp_src.long()[n // imgs_per_group] * imgs_per_group + n % imgs_per_group
            nimg2 = int(n2.max().item()) + 1
            c2 = _flat(col2, nimg2 * Hi * Wi * 16 * C).reshape(nimg2, Hi, Wi, 4, 4, C).float()
            c = c + c2[n2]
        out = torch.zeros(N, 2 * Hi + 2, 2 * Wi + 2, C, device=col.device)
        for kh in range(4):
            for kw in range(4):
                out[:, kh:kh + 2 * Hi:2, kw:kw + 2 * Wi:2, :] += c[:, :, :, kh, kw, :]
        out = out[:, 1:-1, 1:-1, :]
        if bias is not None:
            out = out + bias[:C].float()
        yv = _flat(y, N * 4 * Hi * Wi * C)
        if accumulate:
            out = out.reshape(-1) + yv.float()
        yv.copy_(out.reshape(-1).to(y.dtype))

    def permute4(self, src, dst, dims, strides, accumulate=False):
        s = torch.as_strided(src, tuple(dims), tuple(strides), src.storage_offset()).float()
        n = dims[0] * dims[1] * dims[2] * dims[3]
        d = _flat(dst, n)
        r = s.reshape(-1)
        if accumulate:
            r = r + d.float()
        d.copy_(r.to(dst.dtype))

    def nchw_to_nhwc_dual(self, src, dst_f32, dst_act, N, hw, C):
        for d in (dst_f32, dst_act):
            if d is not None:
                self.permute4(src, d, (N, hw, C, 1), (C * hw, 1, hw, 0))

    def add_indexed(self, dst, src, dst_idx, F_, n):
        d = _flat(dst, dst.numel()).reshape(-1, n) if dst.numel() % n == 0 else None
        s = _flat(src, F_ * n).reshape(F_, n).float()
        idx = dst_idx[:F_].long()
        d[idx] = (d[idx].float() + s).to(dst.dtype)

    def transpose_batched(self, src, dst, A, P, Q):
        m = _flat(src, A * P * Q).reshape(A, P, Q).float()
        _flat(dst, A * P * Q).copy_(m.transpose(1, 2).reshape(-1).to(dst.dtype))

    def blockdiag(self, src, dst, R, C, g):
        m = _flat(src, R * C).reshape(R, C).float()
        _flat(dst, g * R * g * C).copy_(torch.block_diag(*([m] * g)).reshape(-1).to(dst.dtype))

    def group_sum(self, inp, out, grp_src, G, F_, n):
        i = _flat(inp, G * n).reshape(G, n).float()
        o = torch.zeros(F_, n, device=inp.device)
        o.index_add_(0, grp_src[:G].long(), i)
        _flat(out, F_ * n).copy_(o.reshape(-1).to(out.dtype))

    # -- batch norm ----
    def bn_fwd_stats(self, x, G, R, C, gamma, beta, mean, invstd, var_unb, scale, shift, eps=1e-5):
        xv = _flat(x, G * R * C).reshape(G, R, C).double()
        m = xv.mean(1)
        var = (xv * xv).mean(1) - m * m
        var = var.clamp_min(0)
        istd = 1.0 / torch.sqrt(var + eps)
        _flat(mean, G * C).copy_(m.reshape(-1).float())
        _flat(invstd, G * C).copy_(istd.reshape(-1).float())
        vu = var * R / (R - 1) if R > 1 else var
        _flat(var_unb, G * C).copy_(vu.reshape(-1).float())
        sc = gamma[:C].float() * istd.float()
        _flat(scale, G * C).copy_(sc.reshape(-1))
        _flat(shift, G * C).copy_((beta[:C].float() - m.float() * sc).reshape(-1))

    def bn_act(self, x, y, scale, shift, G, R, C, act):
        xv = _flat(x, G * R * C).reshape(G, R, C).float()
        z = xv * _flat(scale, G * C).reshape(G, 1, C) + _flat(shift, G * C).reshape(G, 1, C)
        _flat(y, G * R * C).copy_(_act(z, act).reshape(-1).to(y.dtype))

    def bn_bwd(self, dy, x, y, mean, invstd, gamma, G, R, C, act, dx, sum_dz, sum_dzx):
        n = G * R * C
        dyv = _flat(dy, n).reshape(G, R, C).float()
        xv = _flat(x, n).reshape(G, R, C).float()
        yv = _flat(y, n).reshape(G, R, C).float()
        mu = _flat(mean, G * C).reshape(G, 1, C)
        istd = _flat(invstd, G * C).reshape(G, 1, C)
        dz = dyv * _act_grad(yv, act)
        xhat = (xv - mu) * istd
        s0 = dz.double().sum(1).float()
        s1 = (dz * xhat).double().sum(1).float()
        _flat(sum_dz, G * C).copy_(s0.reshape(-1))
        _flat(sum_dzx, G * C).copy_(s1.reshape(-1))
        r = gamma[:C].float() * istd * (dz - s0.reshape(G, 1, C) / R - xhat * s1.reshape(G, 1, C) / R)
        _flat(dx, n).copy_(r.reshape(-1).to(dx.dtype))

    def bn_param_grad(self, sum_dz, sum_dzx, G, C, dgamma, dbeta):
        dgamma[:C].copy_(_flat(sum_dzx, G * C).reshape(G, C).sum(0))
        dbeta[:C].copy_(_flat(sum_dz, G * C).reshape(G, C).sum(0))

    def bn_ema(self, rmean, rvar, mean, var_unb, order, ncalls, C, momentum=0.1):
        m = _flat(mean, mean.numel()).reshape(-1, C)
        v = _flat(var_unb, var_unb.numel()).reshape(-1, C)
        for k in range(ncalls):
            g = int(order[k])
            rmean[:C].mul_(1 - momentum).add_(m[g], alpha=momentum)
            rvar[:C].mul_(1 - momentum).add_(v[g], alpha=momentum)

    # -- recurrent ----
    def lstm_pointwise_fwd(self, gates, c_prev, c_out, h_out, B, R):
        g = _flat(gates, B * 4 * R).reshape(B, 4 * R)
        i, f, gg, o = torch.sigmoid(g[:, :R]), torch.sigmoid(g[:, R:2 * R]), torch.tanh(g[:, 2 * R:3 * R]), torch.sigmoid(g[:, 3 * R:])
        c = f * _flat(c_prev, B * R).reshape(B, R) + i * gg
        h = o * torch.tanh(c)
        g.copy_(torch.cat([i, f, gg, o], 1))
        _flat(c_out, B * R).copy_(c.reshape(-1))
        _flat(h_out, B * R).copy_(h.reshape(-1))

    def lstm_pointwise_bwd(self, dh, dc_next, gates, c_prev, c, dgates, dc_prev, B, R):
        g = _flat(gates, B * 4 * R).reshape(B, 4 * R)
        i, f, gg, o = g[:, :R], g[:, R:2 * R], g[:, 2 * R:3 * R], g[:, 3 * R:]
        cv = _flat(c, B * R).reshape(B, R)
        cp = _flat(c_prev, B * R).reshape(B, R)
        dhv = _flat(dh, B * R).reshape(B, R)
        tc = torch.tanh(cv)
        dc = dhv * o * (1 - tc * tc)
        if dc_next is not None:
            dc = dc + _flat(dc_next, B * R).reshape(B, R)
        dg = torch.cat([dc * gg * i * (1 - i), dc * cp * f * (1 - f), dc * i * (1 - gg * gg), dhv * tc * o * (1 - o)], 1)
        _flat(dgates, B * 4 * R).copy_(dg.reshape(-1))
        _flat(dc_prev, B * R).copy_((dc * f).reshape(-1))

    def reparam_kl_fwd(self, mu, lv, mu_p, lv_p, eps, eps_p, z, z_p, n, kl_sum):
        m1, l1, m2, l2 = (_flat(t, n) for t in (mu, lv, mu_p, lv_p))
        s1, s2 = (0.5 * l1).exp(), (0.5 * l2).exp()
        _flat(z, n).copy_(_flat(eps, n) * s1 + m1)
        _flat(z_p, n).copy_(_flat(eps_p, n) * s2 + m2)
        k = torch.log(s2 / s1) + (l1.exp() + (m1 - m2) ** 2) / (2 * l2.exp()) - 0.5
        kl_sum.reshape(-1)[0] = k.double().sum().float()

    def reparam_kl_bwd(self, mu, lv, mu_p, lv_p, eps, eps_p, dz, dz_p, kl_coef, dmu, dlv, dmu_p, dlv_p, n):
        m1, l1, m2, l2 = (_flat(t, n) for t in (mu, lv, mu_p, lv_p))
        e1, e2, d = l1.exp(), l2.exp(), m1 - m2
        gm1 = kl_coef * d / e2
        gl1 = kl_coef * (-0.5 + e1 / (2 * e2))
        gm2 = -gm1
        gl2 = kl_coef * (0.5 - (e1 + d * d) / (2 * e2))
        if dz is not None:
            v = _flat(dz, n)
            gm1 = gm1 + v
            gl1 = gl1 + v * _flat(eps, n) * 0.5 * (0.5 * l1).exp()
        if dz_p is not None:
            v = _flat(dz_p, n)
            gm2 = gm2 + v
            gl2 = gl2 + v * _flat(eps_p, n) * 0.5 * (0.5 * l2).exp()
        for t, s in ((dmu, gm1), (dlv, gl1), (dmu_p, gm2), (dlv_p, gl2)):
            _flat(t, n).copy_(s)

    def build_concat(self, dst, A, ia, ga, Bm, ib, gb, tuc, dt, S, B, ld=None):
        Av = _flat(A, A.numel()).reshape(-1, B, ga)
        Bv = _flat(Bm, Bm.numel()).reshape(-1, B, gb)
        a = Av[ia[:S].long()]
        b = Bv[ib[:S].long()]
        t1 = tuc[:S].reshape(S, 1, 1).expand(S, B, 1)
        t2 = dt[:S].reshape(S, 1, 1).expand(S, B, 1)
        W = ga + gb + 2
        ld = ld if ld is not None else W
        d = _flat(dst, S * B * ld).reshape(S * B, ld)
        d.zero_()
        d[:, :W] = torch.cat([a, b, t1, t2], 2).reshape(S * B, W)

    def gather_add_cols(self, dst, src, idx, S, T, B, g, W, col0, init=False):
        d = _flat(dst, T * B * g).reshape(T, B, g)
        if init:
            d.zero_()
        s = _flat(src, S * B * W).reshape(S, B, W)[:, :, col0:col0 + g]
        d.index_add_(0, idx[:S].long(), s)

    def align(self, H, in_idx, h_pred, P, B, g, coef, loss_partial, d_hpred, dH):
        Hv = _flat(H, H.numel()).reshape(-1, B, g)
        hp = _flat(h_pred, P * B * g).reshape(P, B, g)
        for s in range(P):
            t = int(in_idx[s])
            diff = Hv[t, 0].unsqueeze(0) - hp[s]
            loss_partial.reshape(-1)[s] = (diff.double() ** 2).mean().float()
            if d_hpred is not None:
                _flat(d_hpred, P * B * g).reshape(P, B, g)[s] += -coef * 2 * diff / (B * g)
            if dH is not None:
                _flat(dH, dH.numel()).reshape(-1, B, g)[t, 0] += coef * 2 * diff.sum(0) / (B * g)

    def colsum(self, x, rows, cols, ld, out, accumulate=False):
        s = _mat(x, rows, cols, ld).float().sum(0)
        if accumulate:
            out[:cols] += s
        else:
            out[:cols] = s

    def act_fwd(self, x, n, act):
        v = _flat(x, n)
        v.copy_(_act(v, act))

    def act_bwd(self, dy, y, dx, n, act):
        _flat(dx, n).copy_(_flat(dy, n) * _act_grad(_flat(y, n), act))

    # -- losses / optimiser ----
    def mse_chunks(self):
        return 32

    def sigmoid_mse(self, raw, x, tgt, coef, G, E, pred, d_raw, partial):
        r = _flat(raw, G * E).reshape(G, E).float()
        xt = _flat(x, x.numel()).reshape(-1, E)[tgt[:G].long()]
        s = torch.sigmoid(r)
        d = s - xt
        p = _flat(partial, G * 32).reshape(G, 32)
        p.zero_()
        p[:, 0] = (d.double() ** 2).sum(1).float()
        if pred is not None:
            _flat(pred, G * E).copy_(s.reshape(-1).to(pred.dtype))
        if d_raw is not None:
            _flat(d_raw, G * E).copy_((coef[:G].reshape(G, 1) * 2 * d * s * (1 - s)).reshape(-1).to(d_raw.dtype))

    def finalize_losses(self, mse_partial, n_recon, has_cpc, E, kl_sum, batch_size, align_partial, n_align, seq_len, out):
        p = _flat(mse_partial, (n_recon + int(has_cpc)) * 32).reshape(-1, 32).double().sum(1) / E
        o = out.reshape(-1)
        o[0] = float(p[:n_recon].sum()) / seq_len
        o[1] = float(kl_sum.reshape(-1)[0]) / batch_size / seq_len
        o[2] = float(p[n_recon]) / seq_len if has_cpc else 0.0
        o[3] = float(align_partial.reshape(-1)[:n_align].double().sum()) / seq_len

    def adam(self, p, g, m, v, n, lr, beta1, beta2, eps, step_t):
        import math
        t = int(step_t.reshape(-1)[0])
        pv, gv, mv, vv = (_flat(x, n) for x in (p, g, m, v))
        mv.mul_(beta1).add_(gv, alpha=1 - beta1)
        vv.mul_(beta2).addcmul_(gv, gv, value=1 - beta2)
        step_size = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
        pv.addcdiv_(mv, vv.sqrt().add_(eps), value=-step_size)

    def scale(self, x, n, a):
        _flat(x, n).mul_(a)
