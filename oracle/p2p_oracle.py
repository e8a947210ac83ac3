"""CPU oracle for the p2pvg training hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32 or fp64) *restatement* of the reference
algorithm.  It is the checker for the CUDA path, never the thing shipped or
measured: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.

Parity pin: the reference ships no golden vectors, so the oracle
is pinned against *outputs of the reference itself run in the build container*
(``tests/golden/make_golden.py`` imports ``/root/reference`` unmodified and
writes the fixtures ``tests/golden/*.pt``; ``tests/test_oracle_golden.py``
replays them through this file).

Every function cites the reference lines it restates (paths relative to the
reference checkout).  Parameters are carried in flat ``dict[str, Tensor]`` that
use exactly the reference's ``state_dict`` keys, so reference checkpoints can be
fed to the oracle directly.

Functions
---------
encoder_fwd / decoder_fwd      models/dcgan_64.py:28-88, models/dcgan_128.py:28-94
lstm_fwd / gaussian_lstm_fwd   models/lstm.py:37-44, 76-94
kl_criterion                   misc/criterion.py:10-15
legacy_adam_step               torch-1.0 ``optim.Adam.step`` (README.md:62 pins PyTorch 1.0)
train_step                     models/p2p_model.py:185-271 (+ :273-280)
skip_schedule                  models/p2p_model.py:209-229 (integer / rational control logic)
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LRELU = 0.2


# ----------------------------------------------------------------------------------------------
# backbone description
# ----------------------------------------------------------------------------------------------
def dcgan_stages(image_width: int) -> int:
    """Number of stride-2 stages: 4 for dcgan_64 (c1..c4), 5 for dcgan_128 (c1..c5)."""
    if image_width == 64:
        return 4
    if image_width == 128:
        return 5
    raise ValueError("dcgan backbone exists for 64 and 128 pixel frames only")


def dcgan_channels(image_width: int):
    """Output channels of the stride-2 stages (dcgan_64.py:34-40, dcgan_128.py:34-42)."""
    return [64, 128, 256, 512] if image_width == 64 else [64, 128, 256, 512, 512]


# ----------------------------------------------------------------------------------------------
# parameter construction — consumes the torch RNG in the same order as the reference
# ----------------------------------------------------------------------------------------------
def _init_weights_like_reference(mod: torch.nn.Module):
    """misc/utils.py:157-164 — class-name matching initialiser."""
    name = type(mod).__name__
    if "Conv" in name or "Linear" in name:
        mod.weight.data.normal_(0.0, 0.02)
        mod.bias.data.fill_(0)
    elif "BatchNorm" in name:
        mod.weight.data.normal_(1.0, 0.02)
        mod.bias.data.fill_(0)


def _lstm_container(in_dim, out_dim, hidden, layers, gaussian):
    nn = torch.nn
    m = nn.Module()
    m.embed = nn.Linear(in_dim, hidden)
    m.lstm = nn.ModuleList([nn.LSTMCell(hidden, hidden) for _ in range(layers)])
    if gaussian:
        m.mu_net = nn.Linear(hidden, out_dim)
        m.logvar_net = nn.Linear(hidden, out_dim)
    else:
        m.output = nn.Sequential(nn.Linear(hidden, out_dim), nn.Tanh())
    return m


def _dcgan_encoder_container(g_dim, nc, width):
    nn = torch.nn
    m = nn.Module()
    chans = dcgan_channels(width)
    cin = nc
    for idx, cout in enumerate(chans, start=1):
        blk = nn.Module()
        blk.main = nn.Sequential(nn.Conv2d(cin, cout, 4, 2, 1), nn.BatchNorm2d(cout), nn.LeakyReLU(LRELU))
        setattr(m, f"c{idx}", blk)
        cin = cout
    setattr(m, f"c{len(chans) + 1}", nn.Sequential(nn.Conv2d(cin, g_dim, 4, 1, 0), nn.BatchNorm2d(g_dim), nn.Tanh()))
    return m


def _dcgan_decoder_container(g_dim, nc, width):
    nn = torch.nn
    m = nn.Module()
    chans = dcgan_channels(width)
    top = chans[-1]
    m.upc1 = nn.Sequential(nn.ConvTranspose2d(g_dim, top, 4, 1, 0), nn.BatchNorm2d(top), nn.LeakyReLU(LRELU))
    cin = top
    outs = list(reversed(chans[:-1]))  # e.g. [256,128,64] / [512,256,128,64]
    if width == 128:
        outs = [512, 256, 128, 64]
    for k, cout in enumerate(outs, start=2):
        blk = nn.Module()
        blk.main = nn.Sequential(nn.ConvTranspose2d(cin * 2, cout, 4, 2, 1), nn.BatchNorm2d(cout), nn.LeakyReLU(LRELU))
        setattr(m, f"upc{k}", blk)
        cin = cout
    setattr(m, f"upc{len(outs) + 2}", nn.Sequential(nn.ConvTranspose2d(cin * 2, nc, 4, 2, 1), nn.Sigmoid()))
    return m


VGG_ENC = [[(None, 64), (64, 64)], [(64, 128), (128, 128)], [(128, 256), (256, 256), (256, 256)], [(256, 512), (512, 512), (512, 512)]]
VGG_DEC = [[(1024, 512), (512, 512), (512, 256)], [(512, 256), (256, 256), (256, 128)], [(256, 128), (128, 64)], [(128, 64)]]
# models/vgg_128.py:16-105: one more 512-channel stage on both sides
VGG_ENC_128 = VGG_ENC + [[(512, 512), (512, 512), (512, 512)]]
VGG_DEC_128 = [[(1024, 512), (512, 512), (512, 512)]] + VGG_DEC


def _vgg_lists(width):
    return (VGG_ENC_128, VGG_DEC_128) if width == 128 else (VGG_ENC, VGG_DEC)


def _vgg_width(p):
    """64 or 128, from the state_dict keys (vgg_128 has a c6 / upc6)."""
    return 128 if any(k.startswith(("c6.", "upc6.")) for k in p) else 64


def _vgg_layer(nin, nout):
    nn = torch.nn
    blk = nn.Module()
    blk.main = nn.Sequential(nn.Conv2d(nin, nout, 3, 1, 1), nn.BatchNorm2d(nout), nn.LeakyReLU(LRELU))
    return blk


def _vgg_encoder_container(g_dim, nc, width=64):
    """models/vgg_64.py:16-48, models/vgg_128.py:16-54."""
    nn = torch.nn
    m = nn.Module()
    enc = _vgg_lists(width)[0]
    for i, stage in enumerate(enc, 1):
        setattr(m, f"c{i}", nn.Sequential(*[_vgg_layer(nc if a is None else a, b) for a, b in stage]))
    setattr(m, f"c{len(enc) + 1}", nn.Sequential(nn.Conv2d(512, g_dim, 4, 1, 0), nn.BatchNorm2d(g_dim), nn.Tanh()))
    return m


def _vgg_decoder_container(g_dim, nc, width=64):
    """models/vgg_64.py:59-92, models/vgg_128.py:66-105."""
    nn = torch.nn
    m = nn.Module()
    dec = _vgg_lists(width)[1]
    m.upc1 = nn.Sequential(nn.ConvTranspose2d(g_dim, 512, 4, 1, 0), nn.BatchNorm2d(512), nn.LeakyReLU(LRELU))
    for i, stage in enumerate(dec[:-1], 2):
        setattr(m, f"upc{i}", nn.Sequential(*[_vgg_layer(a, b) for a, b in stage]))
    setattr(m, f"upc{len(dec) + 1}", nn.Sequential(_vgg_layer(128, 64), nn.ConvTranspose2d(64, nc, 3, 1, 1), nn.Sigmoid()))
    return m


def _vgg_block(p, pre, x, training=True):
    x = F.conv2d(x, p[pre + ".main.0.weight"], p[pre + ".main.0.bias"], stride=1, padding=1)
    return F.leaky_relu(_bn_train(x, p, pre + ".main.1", training), LRELU)


def vgg_encoder_fwd(p, x, training=True):
    """models/vgg_64.py:50-56, models/vgg_128.py:56-63."""
    skips, h = [], x
    enc = _vgg_lists(_vgg_width(p))[0]
    for i, stage in enumerate(enc, 1):
        if i > 1:
            h = F.max_pool2d(h, 2, 2)
        for j in range(len(stage)):
            h = _vgg_block(p, f"c{i}.{j}", h, training)
        skips.append(h)
    top = f"c{len(enc) + 1}"
    h = F.conv2d(F.max_pool2d(h, 2, 2), p[top + ".0.weight"], p[top + ".0.bias"])
    h = torch.tanh(_bn_train(h, p, top + ".1", training))
    return h.reshape(h.shape[0], -1), skips


def vgg_decoder_fwd(p, vec, skips, training=True):
    """models/vgg_64.py:94-105, models/vgg_128.py:107-120."""
    dec = _vgg_lists(_vgg_width(p))[1]
    n = len(dec)
    d = F.conv_transpose2d(vec.reshape(vec.shape[0], -1, 1, 1), p["upc1.0.weight"], p["upc1.0.bias"])
    d = F.leaky_relu(_bn_train(d, p, "upc1.1", training), LRELU)
    for i, stage in enumerate(dec[:-1], 2):
        d = torch.cat([F.interpolate(d, scale_factor=2, mode="nearest"), skips[n + 1 - i]], 1)
        for j in range(len(stage)):
            d = _vgg_block(p, f"upc{i}.{j}", d, training)
    d = torch.cat([F.interpolate(d, scale_factor=2, mode="nearest"), skips[0]], 1)
    last = f"upc{n + 1}"
    d = _vgg_block(p, last + ".0", d, training)
    return torch.sigmoid(F.conv_transpose2d(d, p[last + ".1.weight"], p[last + ".1.bias"], stride=1, padding=1))


def _residual_linear_container(nin, nout):
    """models/h36m_mlp.py:28-43 — shortcut Linear+ReLU, long path of three Linear+ReLU (hidden nin//2), LayerNorm."""
    nn = torch.nn
    m = nn.Module()
    m.shortcut = nn.Sequential(nn.Linear(nin, nout), nn.ReLU())
    m.long_path = nn.Sequential(nn.Linear(nin, nin // 2), nn.ReLU(), nn.Linear(nin // 2, nin // 2), nn.ReLU(),
                                nn.Linear(nin // 2, nout), nn.ReLU())
    m.norm = nn.LayerNorm(nout)
    return m


def _mlp_encoder_container(in_dim, out_dim, h_dim):
    nn = torch.nn
    m = nn.Module()
    m.fc1 = _residual_linear_container(in_dim, h_dim)
    m.fc2 = _residual_linear_container(h_dim, h_dim)
    m.fc3 = nn.Linear(h_dim, out_dim)
    return m


def _mlp_decoder_container(in_dim, out_dim, h_dim):
    nn = torch.nn
    m = nn.Module()
    m.fc1 = _residual_linear_container(in_dim, h_dim)
    m.fc2 = _residual_linear_container(h_dim * 2, h_dim)
    m.fc3 = nn.Linear(h_dim * 2, out_dim)
    return m


def residual_linear_fwd(p, pre, x):
    """models/h36m_mlp.py:45-46."""
    sc = F.relu(F.linear(x, p[pre + ".shortcut.0.weight"], p[pre + ".shortcut.0.bias"]))
    h = x
    for i in (0, 2, 4):
        h = F.relu(F.linear(h, p[pre + f".long_path.{i}.weight"], p[pre + f".long_path.{i}.bias"]))
    s = sc + h
    return F.layer_norm(s, (s.shape[-1],), p[pre + ".norm.weight"], p[pre + ".norm.bias"], 1e-5)


def mlp_encoder_fwd(p, x):
    """models/h36m_mlp.py:61-69 — x [B,17,3]; returns (latent [B,g], [h1, h2])."""
    h1 = residual_linear_fwd(p, "fc1", x.reshape(x.shape[0], -1))
    h2 = residual_linear_fwd(p, "fc2", h1)
    return torch.tanh(F.linear(h2, p["fc3.weight"], p["fc3.bias"])), [h1, h2]


def mlp_decoder_fwd(p, vec, skips):
    """models/h36m_mlp.py:85-95."""
    d1 = residual_linear_fwd(p, "fc1", vec)
    d2 = residual_linear_fwd(p, "fc2", torch.cat([d1, skips[1]], 1))
    out = F.linear(torch.cat([d2, skips[0]], 1), p["fc3.weight"], p["fc3.bias"])
    return out.reshape(out.shape[0], 17, 3)


def build_state(cfg: dict, seed: int | None = None, dtype=torch.float32) -> "OrderedDict[str, OrderedDict]":
    """Initial parameters + buffers of the five modules.

    Construction order and initialisation order restate models/p2p_model.py:28-38 and :64-69 so
    that ``torch.manual_seed(s)`` followed by this call yields the reference's initial weights.
    """
    if seed is not None:
        torch.manual_seed(seed)
    g, z, r = cfg["g_dim"], cfg["z_dim"], cfg["rnn_size"]
    fp = _lstm_container(g + z + 2, g, r, cfg.get("predictor_rnn_layers", 2), gaussian=False)
    post = _lstm_container(2 * g + 2, z, r, cfg.get("posterior_rnn_layers", 1), gaussian=True)
    prior = _lstm_container(2 * g + 2, z, r, cfg.get("prior_rnn_layers", 1), gaussian=True)
    if cfg.get("backbone", "dcgan") == "mlp":  # models/p2p_model.py:33-35 (h36m)
        enc = _mlp_encoder_container(51, g, g)
        dec = _mlp_decoder_container(g, 51, g)
    elif cfg.get("backbone") == "vgg":
        vw = cfg.get("vgg_width", 128 if cfg.get("image_width") == 128 else 64)
        enc = _vgg_encoder_container(g, cfg["channels"], vw)
        dec = _vgg_decoder_container(g, cfg["channels"], vw)
    else:
        nc, w = cfg["channels"], cfg["image_width"]
        enc = _dcgan_encoder_container(g, nc, w)
        dec = _dcgan_decoder_container(g, nc, w)
    mods = OrderedDict(frame_predictor=fp, posterior=post, prior=prior, encoder=enc, decoder=dec)
    for m in mods.values():
        m.apply(_init_weights_like_reference)
    out = OrderedDict()
    for name, m in mods.items():
        out[name] = OrderedDict((k, v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone())
                                for k, v in m.state_dict().items())
    return out


def is_param(key: str) -> bool:
    return not (key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked"))


# ----------------------------------------------------------------------------------------------
# modules (functional)
# ----------------------------------------------------------------------------------------------
def _bn_train(x, p, prefix, training=True):
    """nn.BatchNorm2d in training mode: batch statistics for normalisation, EMA of running stats
    with the unbiased variance, ``num_batches_tracked += 1``."""
    rm, rv = p[prefix + ".running_mean"], p[prefix + ".running_var"]
    y = F.batch_norm(x, rm, rv, p[prefix + ".weight"], p[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS)
    if training:
        p[prefix + ".num_batches_tracked"] += 1
    return y


def encoder_fwd(p: dict, x: torch.Tensor, width: int, training: bool = True):
    """models/dcgan_64.py:48-54 / models/dcgan_128.py:50-57.  Returns (latent [B,g], skips)."""
    n = dcgan_stages(width)
    h, skips = x, []
    for i in range(1, n + 1):
        pre = f"c{i}.main"
        h = F.conv2d(h, p[pre + ".0.weight"], p[pre + ".0.bias"], stride=2, padding=1)
        h = F.leaky_relu(_bn_train(h, p, pre + ".1", training), LRELU)
        skips.append(h)
    pre = f"c{n + 1}"
    h = F.conv2d(h, p[pre + ".0.weight"], p[pre + ".0.bias"], stride=1, padding=0)
    h = torch.tanh(_bn_train(h, p, pre + ".1", training))
    return h.reshape(h.shape[0], -1), skips


def decoder_fwd(p: dict, vec: torch.Tensor, skips, width: int, training: bool = True):
    """models/dcgan_64.py:81-88 / models/dcgan_128.py:86-94."""
    n = dcgan_stages(width)
    d = F.conv_transpose2d(vec.reshape(vec.shape[0], -1, 1, 1), p["upc1.0.weight"], p["upc1.0.bias"], stride=1, padding=0)
    d = F.leaky_relu(_bn_train(d, p, "upc1.1", training), LRELU)
    for k in range(2, n + 1):
        pre = f"upc{k}.main"
        d = F.conv_transpose2d(torch.cat([d, skips[n + 1 - k]], 1), p[pre + ".0.weight"], p[pre + ".0.bias"], stride=2, padding=1)
        d = F.leaky_relu(_bn_train(d, p, pre + ".1", training), LRELU)
    pre = f"upc{n + 1}"
    d = F.conv_transpose2d(torch.cat([d, skips[0]], 1), p[pre + ".0.weight"], p[pre + ".0.bias"], stride=2, padding=1)
    return torch.sigmoid(d)


def lstm_cell(p, prefix, x, hc):
    """nn.LSTMCell: gate order i,f,g,o along the 4R axis."""
    h, c = hc
    gates = F.linear(x, p[prefix + ".weight_ih"], p[prefix + ".bias_ih"]) + F.linear(h, p[prefix + ".weight_hh"], p[prefix + ".bias_hh"])
    i, f, g, o = gates.chunk(4, 1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def _n_layers(p):
    return len({k.split(".")[1] for k in p if k.startswith("lstm.")})


def init_hidden(p, batch, like):
    """models/lstm.py:21-27 / :60-66 — zero (h, c) per layer."""
    r = p["embed.weight"].shape[0]
    return [(like.new_zeros(batch, r), like.new_zeros(batch, r)) for _ in range(_n_layers(p))]


def lstm_fwd(p, hidden, inp):
    """models/lstm.py:37-44 — frame predictor.  Mutates ``hidden``."""
    h = F.linear(inp, p["embed.weight"], p["embed.bias"])
    for l in range(len(hidden)):
        hidden[l] = lstm_cell(p, f"lstm.{l}", h, hidden[l])
        h = hidden[l][0]
    return torch.tanh(F.linear(h, p["output.0.weight"], p["output.0.bias"]))


def gaussian_lstm_fwd(p, hidden, inp, eps):
    """models/lstm.py:83-94 with the N(0,1) draw of :78 supplied by the caller (``eps``)."""
    h = F.linear(inp, p["embed.weight"], p["embed.bias"])
    for l in range(len(hidden)):
        hidden[l] = lstm_cell(p, f"lstm.{l}", h, hidden[l])
        h = hidden[l][0]
    mu = F.linear(h, p["mu_net.weight"], p["mu_net.bias"])
    logvar = F.linear(h, p["logvar_net.weight"], p["logvar_net.bias"])
    sigma = (logvar * 0.5).exp()  # lstm.py:77
    z = eps * sigma + mu  # lstm.py:81
    return z, mu, logvar


def kl_criterion(mu1, lv1, mu2, lv2, batch_size):
    """misc/criterion.py:10-15 — note the division by the *configured* batch size."""
    s1 = (lv1 * 0.5).exp()
    s2 = (lv2 * 0.5).exp()
    kld = torch.log(s2 / s1) + (torch.exp(lv1) + (mu1 - mu2) ** 2) / (2 * torch.exp(lv2)) - 0.5
    return kld.sum() / batch_size


# ----------------------------------------------------------------------------------------------
# control logic (bit-exact part)
# ----------------------------------------------------------------------------------------------
def skip_schedule(seq_len: int, probs, skip_prob: float, n_past: int):
    """models/p2p_model.py:209-229.  Returns the executed timesteps with their time counters.

    Each entry: (i, time_until_cp, delta_time) with the counters computed as Python doubles exactly
    as the reference does before they are rounded to fp32 by ``fill_``.
    """
    cp_ix = seq_len - 1
    prev_i, skip_count = 0, 0
    max_skip = seq_len * skip_prob
    out = []
    for i in range(1, seq_len):
        if probs[i - 1] <= skip_prob and i >= n_past and skip_count < max_skip and i != 1 and i != cp_ix:
            skip_count += 1
            continue
        out.append((i, (cp_ix - i + 1) / cp_ix, (i - prev_i) / cp_ix))
        prev_i = i
    return out


# ----------------------------------------------------------------------------------------------
# optimiser
# ----------------------------------------------------------------------------------------------
def new_adam_state(params: dict):
    return {k: dict(step=0, m=torch.zeros_like(v), v=torch.zeros_like(v)) for k, v in params.items() if is_param(k)}


def legacy_adam_step(params: dict, grads: dict, state: dict, lr: float, beta1: float, beta2: float = 0.999, eps: float = 1e-8):
    """PyTorch-1.0 Adam (the version README.md:62 pins): ``denom = sqrt(v) + eps`` *before* the bias
    correction is folded into the step size.  Updates ``params`` in place through ``.data`` so that
    autograd version counters are not bumped (oracle Mode A)."""
    for k, g in grads.items():
        if g is None:
            continue
        st = state[k]
        st["step"] += 1
        st["m"].mul_(beta1).add_(g, alpha=1 - beta1)
        st["v"].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = st["v"].sqrt().add_(eps)
        bc1 = 1 - beta1 ** st["step"]
        bc2 = 1 - beta2 ** st["step"]
        step_size = lr * math.sqrt(bc2) / bc1
        params[k].data.addcdiv_(st["m"], denom, value=-step_size)


# ----------------------------------------------------------------------------------------------
# the train step
# ----------------------------------------------------------------------------------------------
MODULES = ("frame_predictor", "posterior", "prior", "encoder", "decoder")


def default_opt(**kw):
    o = dict(beta=1e-4, weight_cpc=100.0, weight_align=0.5, skip_prob=0.0, n_past=1, last_frame_skip=False,
             lr=1e-3, beta1=0.9, batch_size=None)
    o.update(kw)
    return o


def forward_losses(state, x, opt, width, eps, probs, tape=None):
    """models/p2p_model.py:195-257 — the per-timestep loop, in the reference's own call order
    (every frame encoded where the reference encodes it, so BN running statistics advance with the
    reference's multiplicity).  Returns (mse, kld, cpc, align) as graph-attached scalars.

    ``eps``: [S, 2, B, z] — posterior's draw then prior's draw for each executed step.
    """
    enc, dec = state["encoder"], state["decoder"]
    fp, post, prior = state["frame_predictor"], state["posterior"], state["prior"]
    if width == "mlp":  # h36m pose backbone: x [T,B,17,3]
        encoder_fwd_ = lambda p, xx, w: mlp_encoder_fwd(p, xx)
        decoder_fwd_ = lambda p, v, sk, w: mlp_decoder_fwd(p, v, sk)
    elif width == "vgg":
        encoder_fwd_ = lambda p, xx, w: vgg_encoder_fwd(p, xx)
        decoder_fwd_ = lambda p, v, sk, w: vgg_decoder_fwd(p, v, sk)
    else:
        encoder_fwd_, decoder_fwd_ = encoder_fwd, decoder_fwd
    seq_len, B = x.shape[0], x.shape[1]
    cp_ix = seq_len - 1
    hid_fp, hid_post, hid_prior = init_hidden(fp, B, x), init_hidden(post, B, x), init_hidden(prior, B, x)
    x_cp = x[cp_ix]
    global_z = encoder_fwd_(enc, x_cp, width)[0]  # p2p_model.py:71-78, not detached
    sched = skip_schedule(seq_len, probs, opt["skip_prob"], opt["n_past"])
    mse = kld = cpc = align = 0
    h = h_pred = skip = None
    for s, (i, tuc, dt) in enumerate(sched):
        if i > 1:
            # p2p_model.py:224-225 — `h` is already the [B,g] latent, so h[0] is batch row 0 broadcast
            align = align + F.mse_loss(h[0].expand_as(h_pred), h_pred)
        t_tuc = x.new_zeros(B, 1).fill_(tuc)
        t_dt = x.new_zeros(B, 1).fill_(dt)
        h_full = encoder_fwd_(enc, x[i - 1], width)
        h_target = encoder_fwd_(enc, x[i], width)[0]
        if opt["last_frame_skip"] or i <= opt["n_past"]:
            h, skip = h_full
        else:
            h = h_full[0]
        h_cpaw = torch.cat([h, global_z, t_tuc, t_dt], 1)
        h_target_cpaw = torch.cat([h_target, global_z, t_tuc, t_dt], 1)
        zt, mu, logvar = gaussian_lstm_fwd(post, hid_post, h_target_cpaw, eps[s, 0])
        zt_p, mu_p, logvar_p = gaussian_lstm_fwd(prior, hid_prior, h_cpaw, eps[s, 1])
        h_pred = lstm_fwd(fp, hid_fp, torch.cat([h, zt, t_tuc, t_dt], 1))
        x_pred = decoder_fwd_(dec, h_pred, skip, width)
        if i == cp_ix:
            h_pred_p = lstm_fwd(fp, hid_fp, torch.cat([h, zt_p, t_tuc, t_dt], 1))
            x_pred_p = decoder_fwd_(dec, h_pred_p, skip, width)
            cpc = F.mse_loss(x_pred_p, x_cp)
        mse = mse + F.mse_loss(x_pred, x[i])
        kld = kld + kl_criterion(mu, logvar, mu_p, logvar_p, opt["batch_size"])
        if tape is not None:
            tape.append(dict(i=i, tuc=tuc, dt=dt, h=h.detach(), h_pred=h_pred.detach(), mu=mu.detach(),
                             logvar=logvar.detach(), mu_p=mu_p.detach(), logvar_p=logvar_p.detach(),
                             zt=zt.detach(), x_pred=x_pred.detach()))
    return mse, kld, cpc, align



def p2p_generate(state, x, len_output, eval_cp_ix, opt, width, eps, probs, model_mode="full", skip_frame=False):
    """models/p2p_model.py:80-183 — autoregressive point-to-point generation, modules in eval mode (BatchNorm uses its
    running statistics: train.py:245, generate.py:82).  One sample per input sequence; skipped frames are zeros (:136);
    the posterior sees ground truth only while it exists (:167-171); model_mode picks the z that feeds the predictor
    (:159-162, :176-179).  ``eps``: [n_executed, 2, B, z] — posterior's draw then prior's draw per executed step
    (both gaussian LSTMs are called every executed step); ``probs``: the np.random.uniform(0,1,len_output-1) draw (:128).
    Returns the list of len_output frames."""
    enc, dec = state["encoder"], state["decoder"]
    fp, post, prior = state["frame_predictor"], state["posterior"], state["prior"]
    if width == "mlp":
        encoder_fwd_ = lambda p, xx: mlp_encoder_fwd(p, xx)
        decoder_fwd_ = lambda p, v, sk: mlp_decoder_fwd(p, v, sk)
    elif width == "vgg":
        encoder_fwd_ = lambda p, xx: vgg_encoder_fwd(p, xx, training=False)
        decoder_fwd_ = lambda p, v, sk: vgg_decoder_fwd(p, v, sk, training=False)
    else:
        encoder_fwd_ = lambda p, xx: encoder_fwd(p, xx, width, training=False)
        decoder_fwd_ = lambda p, v, sk: decoder_fwd(p, v, sk, width, training=False)
    with torch.no_grad():
        B = x[0].shape[0]
        gen_seq = [x[0]]
        x_in = x[0]
        hid_fp, hid_post, hid_prior = init_hidden(fp, B, x[0]), init_hidden(post, B, x[0]), init_hidden(prior, B, x[0])
        seq_len = len(x)
        x_cp = x[seq_len - 1]
        global_z = encoder_fwd_(enc, x_cp)[0]
        prev_i, skip_count, s = 0, 0, 0
        max_skip_count = seq_len * opt["skip_prob"]
        skip = None
        for i in range(1, len_output):
            if (probs[i - 1] <= opt["skip_prob"] and i >= opt["n_past"] and skip_count < max_skip_count and i != 1
                    and i != (len_output - 1) and skip_frame):
                skip_count += 1
                gen_seq.append(torch.zeros_like(x_in))
                continue
            tuc = x_cp.new_zeros(B, 1).fill_((eval_cp_ix - i + 1) / eval_cp_ix)
            dt = x_cp.new_zeros(B, 1).fill_((i - prev_i) / eval_cp_ix)
            prev_i = i
            h, sk = encoder_fwd_(enc, x_in)
            if opt["last_frame_skip"] or i == 1 or i < opt["n_past"]:
                skip = sk
            h_cpaw = torch.cat([h, global_z, tuc, dt], 1)
            if i < opt["n_past"]:
                h_target = encoder_fwd_(enc, x[i])[0]
                zt = gaussian_lstm_fwd(post, hid_post, torch.cat([h_target, global_z, tuc, dt], 1), eps[s, 0])[0]
                zt_p = gaussian_lstm_fwd(prior, hid_prior, h_cpaw, eps[s, 1])[0]
                lstm_fwd(fp, hid_fp, torch.cat([h, zt if model_mode in ("posterior", "full") else zt_p, tuc, dt], 1))
                x_in = x[i]
            else:
                if i < len(x):
                    h_target = encoder_fwd_(enc, x[i])[0]
                    h_target_cpaw = torch.cat([h_target, global_z, tuc, dt], 1)
                else:
                    h_target_cpaw = h_cpaw
                zt = gaussian_lstm_fwd(post, hid_post, h_target_cpaw, eps[s, 0])[0]
                zt_p = gaussian_lstm_fwd(prior, hid_prior, h_cpaw, eps[s, 1])[0]
                h_pred = lstm_fwd(fp, hid_fp, torch.cat([h, zt if model_mode == "posterior" else zt_p, tuc, dt], 1))
                x_in = decoder_fwd_(dec, h_pred, skip)
            s += 1
            gen_seq.append(x_in)
    return gen_seq


def train_step(state, adam, x, opt, width, eps, probs, mode="A", tape=None):
    """One call of P2PModel.forward (models/p2p_model.py:185-271).

    ``state``: module -> {key: tensor}; parameters are updated in place, BN buffers advanced.
    ``adam``:  module -> legacy Adam state (``new_adam_state``).
    mode "A": reference behaviour on its pinned torch (second backward sees post-step weights of the
              four already-updated modules together with pre-step saved activations).
    mode "B": all gradients evaluated at pre-step weights (clean variant).
    Returns dict(losses=(mse,kld,cpc,align)/seq_len as python floats, grads=module->{key: grad}).
    """
    opt = dict(opt)
    if opt.get("batch_size") is None:
        opt["batch_size"] = x.shape[1]
    for m in MODULES:
        for k, v in state[m].items():
            if is_param(k):
                v.requires_grad_(True)
                v.grad = None
    seq_len = x.shape[0]
    mse, kld, cpc, align = forward_losses(state, x, opt, width, eps, probs, tape)
    loss = mse + kld * opt["beta"] + align * opt["weight_align"]
    prior_loss = kld + cpc * opt["weight_cpc"]
    non_prior = [m for m in MODULES if m != "prior"]
    pkeys = {m: [k for k in state[m] if is_param(k)] for m in MODULES}

    loss.backward(retain_graph=True)  # p2p_model.py:262
    grads1 = {m: {k: (state[m][k].grad.detach().clone() if state[m][k].grad is not None else None) for k in pkeys[m]} for m in non_prior}
    if mode == "A":
        for m in ("frame_predictor", "posterior", "encoder", "decoder"):  # p2p_model.py:276-280
            legacy_adam_step(state[m], grads1[m], adam[m], opt["lr"], opt["beta1"])
    for k in pkeys["prior"]:  # p2p_model.py:266
        state["prior"][k].grad = None
    prior_loss.backward()  # p2p_model.py:268
    gprior = {k: state["prior"][k].grad.detach().clone() for k in pkeys["prior"]}
    if mode != "A":
        for m in ("frame_predictor", "posterior", "encoder", "decoder"):
            legacy_adam_step(state[m], grads1[m], adam[m], opt["lr"], opt["beta1"])
    legacy_adam_step(state["prior"], gprior, adam["prior"], opt["lr"], opt["beta1"])  # p2p_model.py:273-274
    grads = dict(grads1)
    grads["prior"] = gprior
    for m in MODULES:
        for k in pkeys[m]:
            state[m][k].requires_grad_(False)
            state[m][k].grad = None
    sc = lambda t: float(t.detach()) / seq_len if torch.is_tensor(t) else float(t) / seq_len
    return dict(losses=(sc(mse), sc(kld), sc(cpc), sc(align)), grads=grads)


def clone_state(state, dtype=None):
    out = OrderedDict()
    for m, d in state.items():
        out[m] = OrderedDict()
        for k, v in d.items():
            v = v.detach().clone()
            if dtype is not None and v.is_floating_point():
                v = v.to(dtype)
            out[m][k] = v
    return out


def draw_eps(n_steps, batch, z_dim, seed=None, dtype=torch.float32):
    """The reference draws posterior's then prior's N(0,1) [B,z] tensor per executed step from the
    torch global generator (models/lstm.py:78; p2p_model.py:244-245)."""
    if seed is not None:
        torch.manual_seed(seed)
    out = torch.empty(n_steps, 2, batch, z_dim, dtype=dtype)
    for s in range(n_steps):
        out[s, 0].normal_()
        out[s, 1].normal_()
    return out


def tensor_digest(t: torch.Tensor, n_samples: int = 32):
    """Compact signature of a tensor for the committed golden fixtures (full weight sets are 50 MB)."""
    f = t.detach().double().reshape(-1)
    g = torch.Generator().manual_seed(f.numel() % 9973 + 17)
    idx = torch.randint(0, f.numel(), (min(n_samples, f.numel()),), generator=g)
    return dict(numel=f.numel(), sum=float(f.sum()), l2=float(f.norm()), absmax=float(f.abs().max()),
                samples=f[idx].clone(), idx=idx)
