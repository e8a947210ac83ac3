#!/usr/bin/env python
"""More fixtures from the UNMODIFIED reference (build container only, same shims as make_golden.py):

  gen_<case>.pt        the reference's own ``P2PModel.p2p_generate`` (models/p2p_model.py:80-183) in eval mode
                       (BatchNorm on its running statistics), for model_mode in {full, posterior, prior} x skip_frame in
                       {False, True}: inputs, NumPy skip draws, the eps stream, BN buffers, every generated frame as a
                       digest and the last frame in full.
  ckpt_ref_small.pth   a checkpoint written by the reference's own ``P2PModel.save`` (models/p2p_model.py:289-308) after
                       one training step of a SMALL configuration (h36m_mlp backbone, g_dim 32, rnn_size 32: a few
                       hundred KB; the checkpoint format does not depend on the backbone)
  ckpt_ref_small_next.pt   the batch / eps / skip draws and the four losses of the reference's NEXT step from that
                       checkpoint (resume check), plus digests of the checkpointed tensors.

    python tests/golden/make_golden_extra.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import LegacyAdam, ROOT, import_reference, make_opt  # noqa: E402

sys.path.insert(0, ROOT)
from oracle.p2p_oracle import tensor_digest  # noqa: E402

GEN_CASES = {
    # n_past = 1 (the training default): ground truth is only the first frame
    "d64": dict(width=64, channels=1, T_in=5, len_output=7, B=2, opt=dict(skip_prob=0.5)),
    # n_past = 2 + last_frame_skip: the `i < n_past` branch (posterior / prior advance on ground truth) and fresh skips
    "d64_np2": dict(width=64, channels=1, T_in=4, len_output=6, B=2, opt=dict(skip_prob=0.5, n_past=2, last_frame_skip=True)),
}


def warm_bn(model, spec, gen):
    """Move the BatchNorm running statistics off their initial (0, 1) without touching any weight: a few train-mode
    forward calls of the reference's own encoder / decoder under no_grad."""
    model.train()
    with torch.no_grad():
        for _ in range(3):
            xb = torch.rand(4, spec["channels"], spec["width"], spec["width"], generator=gen)
            h, skips = model.encoder(xb)
            model.decoder([torch.tanh(torch.randn(4, 128, generator=gen)), skips])
    model.eval()


def run_gen_case(name, spec, p2p_model, backbones):
    torch.manual_seed(1)
    opt = make_opt(backbones[spec["width"]], batch_size=spec["B"], **spec["opt"])
    model = p2p_model.P2PModel(opt.batch_size, spec["channels"], 128, 10, 256, 1, 1, 2, opt=opt)
    gen = torch.Generator().manual_seed(4321)
    warm_bn(model, spec, gen)
    x = torch.rand(spec["T_in"], spec["B"], spec["channels"], spec["width"], spec["width"], generator=gen)
    mods = dict(encoder=model.encoder, decoder=model.decoder)
    fix = dict(case=name, init_seed=1, cfg=dict(g_dim=128, z_dim=10, rnn_size=256, channels=spec["channels"], image_width=spec["width"],
                                                 predictor_rnn_layers=2, posterior_rnn_layers=1, prior_rnn_layers=1),
               opt={k: getattr(opt, k) for k in ("beta", "weight_cpc", "weight_align", "skip_prob", "n_past", "last_frame_skip", "lr",
                                                 "beta1", "batch_size")},
               x=x, len_output=spec["len_output"], eval_cp_ix=spec["len_output"] - 1,
               bn_buffers={m: {k: v.detach().clone() for k, v in mods[m].state_dict().items() if "running_" in k or "num_batches" in k}
                           for m in mods},
               runs=[])
    n_calls = []
    hook = model.posterior.register_forward_hook(lambda *a: n_calls.append(1))
    for mode in ("full", "posterior", "prior"):
        for skip_frame in (False, True):
            seed = 100 + len(fix["runs"])
            np.random.seed(seed)
            probs = np.random.uniform(0, 1, spec["len_output"] - 1)
            np.random.seed(seed)
            torch.manual_seed(seed)
            n_calls.clear()
            with torch.no_grad():
                seq = model.p2p_generate(x, spec["len_output"], spec["len_output"] - 1, model_mode=mode, skip_frame=skip_frame)
            n_exec = len(n_calls)
            torch.manual_seed(seed)
            eps = torch.empty(n_exec, 2, spec["B"], 10)
            for s in range(n_exec):
                eps[s, 0].normal_()
                eps[s, 1].normal_()
            zeros = [bool((f == 0).all()) for f in seq]
            fix["runs"].append(dict(model_mode=mode, skip_frame=skip_frame, np_seed=seed, probs=torch.from_numpy(probs), eps=eps,
                                    n_exec=n_exec, zero_frames=zeros, digests=[tensor_digest(f) for f in seq],
                                    last=seq[-1].detach().clone(), mid=seq[len(seq) // 2].detach().clone()))
            print(f"[gen_{name}] mode={mode} skip_frame={skip_frame}: executed {n_exec}, zero frames {zeros}")
    hook.remove()
    path = os.path.join(HERE, f"gen_{name}.pt")
    torch.save(fix, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def run_ckpt_case(p2p_model, backbones):
    torch.manual_seed(1)
    g_dim, z_dim, rnn, B, T = 32, 4, 32, 4, 5
    opt = make_opt(backbones["mlp"], dataset="h36m", batch_size=B)
    model = p2p_model.P2PModel(B, 1, g_dim, z_dim, rnn, 1, 1, 2, opt=opt)
    model.opt.optimizer = LegacyAdam
    model.init_optimizer()
    model.train()
    gen = torch.Generator().manual_seed(99)

    def step(seed):
        x = 3 * torch.randn(T, B, 17, 3, generator=gen)
        np.random.seed(seed)
        probs = np.random.uniform(0, 1, T - 1)
        np.random.seed(seed)
        torch.manual_seed(seed)
        model.zero_grad()
        losses = model((None, x, None), 0, T - 1)
        torch.manual_seed(seed)
        eps = torch.empty(T - 1, 2, B, z_dim)
        for s in range(T - 1):
            eps[s, 0].normal_()
            eps[s, 1].normal_()
        return x, probs, eps, [float(v) for v in losses]

    step(7)
    path = os.path.join(HERE, "ckpt_ref_small.pth")
    # a real reference checkpoint pickles opt.optimizer = torch.optim.Adam (models/p2p_model.py:41); the Mode-A shim class
    # lives in this script and must not leak into the file
    model.opt.optimizer = torch.optim.Adam
    model.save(path, 3)   # the reference's own save(): epoch 3
    model.opt.optimizer = LegacyAdam
    mods = ("frame_predictor", "posterior", "prior", "encoder", "decoder")
    digests = {m: {k: tensor_digest(v) for k, v in getattr(model, m).state_dict().items() if v.is_floating_point()} for m in mods}
    x, probs, eps, losses = step(8)
    side = dict(cfg=dict(g_dim=g_dim, z_dim=z_dim, rnn_size=rnn, backbone="mlp", predictor_rnn_layers=2, posterior_rnn_layers=1,
                         prior_rnn_layers=1),
                B=B, T=T, epoch=3, digests=digests, next=dict(x=x, probs=torch.from_numpy(probs), eps=eps, losses=losses),
                post_digests={m: {k: tensor_digest(v) for k, v in getattr(model, m).state_dict().items() if v.is_floating_point()}
                              for m in mods})
    torch.save(side, os.path.join(HERE, "ckpt_ref_small_next.pt"))
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; next-step losses", losses)


def main():
    torch.set_num_threads(8)
    p2p_model, backbones = import_reference()
    for name, spec in GEN_CASES.items():
        run_gen_case(name, spec, p2p_model, backbones)
    run_ckpt_case(p2p_model, backbones)


if __name__ == "__main__":
    main()
