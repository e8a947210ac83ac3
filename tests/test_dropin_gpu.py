"""The drop-in Python API (same names / signatures as the reference's models.* and misc.criterion):
P2PModel training step through the public call, checkpoints, stand-alone module forwards, p2p_generate."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import p2p_oracle as O

pytestmark = pytest.mark.gpu
CFG = dict(g_dim=128, z_dim=10, rnn_size=256, channels=1, image_width=64, predictor_rnn_layers=2, posterior_rnn_layers=1,
           prior_rnn_layers=1)


def make_model(B, precision="fp32", graph="0", **optkw):
    from p2pvg_b200.models import dcgan_64
    from p2pvg_b200.models.p2p_model import P2PModel
    os.environ["P2PVG_PRECISION"] = precision
    os.environ["P2PVG_GRAPH"] = graph
    opt = types.SimpleNamespace(dataset="mnist", backbone_net=dcgan_64, lr=1e-3, beta1=0.9, beta=1e-4, weight_cpc=100.0,
                                weight_align=0.5, skip_prob=0.0, n_past=1, last_frame_skip=False, batch_size=B)
    for k, v in optkw.items():
        setattr(opt, k, v)
    torch.manual_seed(1)
    return P2PModel(B, 1, 128, 10, 256, 1, 1, 2, opt=opt).cuda()


def test_p2pmodel_forward_matches_oracle_and_exposes_reference_api(tmp_path):
    T, B = 5, 3
    model = make_model(B)
    model.train()
    state = O.build_state(CFG, seed=1)
    for m in O.MODULES:  # same seed -> same initial weights and the reference's state_dict keys
        sd = getattr(model, m).state_dict()
        assert list(sd.keys()) == list(state[m].keys())
        for k in sd:
            assert torch.equal(sd[k].cpu(), state[m][k]), (m, k)
    x = torch.rand(T, B, 1, 64, 64, generator=torch.Generator().manual_seed(5))
    np.random.seed(0)
    probs = np.random.uniform(0, 1, T - 1)
    np.random.seed(0)
    torch.manual_seed(123)
    eps = torch.randn(T - 1, 2, B, 10, device="cuda").cpu()
    torch.manual_seed(123)
    model.zero_grad()
    out = model(x.cuda(), 0, T - 1)
    assert len(out) == 4 and all(np.isfinite(float(v)) for v in out)
    adam = {m: O.new_adam_state(state[m]) for m in O.MODULES}
    ref = O.train_step(state, adam, x, O.default_opt(batch_size=B), 64, eps, probs, mode="A")
    np.testing.assert_allclose(np.array(out, dtype=np.float32), np.array(ref["losses"], dtype=np.float32), rtol=1e-4, atol=1e-7)
    for name, p in model.named_parameters():  # train.py:226-233 reads .data and .grad of every parameter
        assert p.grad is not None and p.grad.shape == p.shape and p.is_cuda
    g = dict(model.posterior.named_parameters())["embed.weight"].grad.cpu()
    cos = torch.nn.functional.cosine_similarity(g.flatten(), ref["grads"]["posterior"]["embed.weight"].flatten(), dim=0)
    assert cos > 0.9999
    # checkpoint round trip in the reference's dict layout (p2p_model.py:289-330)
    f = str(tmp_path / "model.pth")
    model.save(f, epoch=3)
    ck = torch.load(f, weights_only=False)
    assert set(ck.keys()) == {"encoder", "decoder", "frame_predictor", "posterior", "prior", "encoder_opt", "decoder_opt",
                              "frame_predictor_opt", "posterior_opt", "prior_opt", "epoch", "opt"}
    assert ck["encoder_opt"]["state"][0]["step"] == 1 and "exp_avg" in ck["encoder_opt"]["state"][0]
    w_before = model.encoder.c1.main[0].weight.detach().clone()
    model2 = make_model(B)
    assert model2.load(f) == 4
    assert torch.equal(model2.encoder.c1.main[0].weight, w_before)
    out2 = model2(x.cuda(), 0, T - 1)
    assert all(np.isfinite(float(v)) for v in out2)
    assert int(model2.encoder_optimizer.state_dict()["state"][0]["step"]) == 2


def test_graph_replay_equals_eager():
    T, B = 4, 4
    x = torch.rand(T, B, 1, 64, 64, generator=torch.Generator().manual_seed(7)).cuda()
    outs = {}
    for graph in ("0", "1"):
        model = make_model(B, precision="bf16", graph=graph)
        res = []
        for it in range(4):  # eager warm step, capture, replay, replay
            torch.manual_seed(50 + it)
            res.append(np.array(model(x, 0, T - 1), dtype=np.float64))
        outs[graph] = res
    for a, b in zip(outs["0"], outs["1"]):
        np.testing.assert_allclose(a, b, rtol=2e-3)


def test_early_loss_readback_equals_blocking_readback():
    """P2PModel.forward hands the four scalars over as soon as the forward half of the step has produced them (zero-copy store
    polled by the host, the backward passes / optimiser still running).  They must be the very numbers a blocking read-back
    after the whole step gives, for eager steps, the capture step and graph replays; and everything the caller does afterwards
    (here: reading the weights) must see the finished step (stream order)."""
    T, B = 4, 4
    x = torch.rand(T, B, 1, 64, 64, generator=torch.Generator().manual_seed(8)).cuda()
    runs = {}
    for early in ("1", "0"):
        os.environ["P2PVG_EARLY_LOSS"] = early
        try:
            model = make_model(B, precision="bf16", graph="1")
            res = []
            for it in range(5):
                torch.manual_seed(70 + it)
                out = np.array(model(x, 0, T - 1), dtype=np.float32)
                eng = model._engine
                assert eng.early_loss == (early == "1")
                torch.cuda.synchronize()
                assert np.array_equal(out, eng._bufs["loss_out"][:4].cpu().numpy()), (early, it)
                res.append((out, model.decoder.state_dict()["upc1.0.weight"].float().sum().item()))
            runs[early] = res
        finally:
            os.environ.pop("P2PVG_EARLY_LOSS", None)
    for (a, wa), (b, wb) in zip(runs["1"], runs["0"]):
        np.testing.assert_allclose(a, b, rtol=1e-5)
        assert abs(wa - wb) <= 1e-5 * abs(wb)


def test_module_forwards_match_oracle():
    os.environ["P2PVG_PRECISION"] = "fp32"
    model = make_model(2)
    state = O.clone_state(O.build_state(CFG, seed=1))
    x = torch.rand(3, 1, 64, 64, generator=torch.Generator().manual_seed(9))
    for training in (True, False):
        model.train(training)
        h, skips = model.encoder(x.cuda())
        h_ref, sk_ref = O.encoder_fwd(state["encoder"], x, 64, training=training)
        assert torch.allclose(h.cpu(), h_ref, rtol=1e-4, atol=1e-5)
        for a, b in zip(skips, sk_ref):
            assert a.shape == b.shape and torch.allclose(a.cpu(), b, rtol=1e-4, atol=1e-5)
        y = model.decoder([h, skips])
        y_ref = O.decoder_fwd(state["decoder"], h_ref, sk_ref, 64, training=training)
        assert y.shape == y_ref.shape and torch.allclose(y.cpu(), y_ref, rtol=1e-4, atol=1e-5)
    assert torch.allclose(model.encoder.c1.main[1].running_mean.cpu(), state["encoder"]["c1.main.1.running_mean"], rtol=1e-4, atol=1e-6)
    # recurrent modules keep a mutable .hidden list like the reference
    model.init_hidden(batch_size=3)
    hid_fp = O.init_hidden(state["frame_predictor"], 3, x)
    hid_post = O.init_hidden(state["posterior"], 3, x)
    for step in range(2):
        inp = torch.randn(3, 140, generator=torch.Generator().manual_seed(step))
        out = model.frame_predictor(inp.cuda())
        ref = O.lstm_fwd(state["frame_predictor"], hid_fp, inp)
        assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-5)
        inp2 = torch.randn(3, 258, generator=torch.Generator().manual_seed(10 + step))
        z, mu, lv = model.posterior(inp2.cuda())
        _, mu_r, lv_r = O.gaussian_lstm_fwd(state["posterior"], hid_post, inp2, torch.zeros(3, 10))
        assert torch.allclose(mu.cpu(), mu_r, rtol=1e-4, atol=1e-5) and torch.allclose(lv.cpu(), lv_r, rtol=1e-4, atol=1e-5)
        assert z.shape == mu.shape and not torch.equal(z, mu)
    assert torch.allclose(model.frame_predictor.hidden[1][0].cpu(), hid_fp[1][0], rtol=1e-4, atol=1e-5)


def test_kl_criterion_and_generate():
    from p2pvg_b200.misc.criterion import KLCriterion
    opt = types.SimpleNamespace(batch_size=5)
    mu1, lv1, mu2, lv2 = (torch.randn(4, 10, generator=torch.Generator().manual_seed(i)) * 0.3 for i in range(4))
    got = KLCriterion(opt)(mu1.cuda(), lv1.cuda(), mu2.cuda(), lv2.cuda())
    ref = O.kl_criterion(mu1, lv1, mu2, lv2, 5)
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref)) + 1e-7
    model = make_model(2, precision="bf16")
    model.eval()
    x = [t for t in torch.rand(6, 2, 1, 64, 64, generator=torch.Generator().manual_seed(3)).cuda()]
    np.random.seed(1)
    seq = model.p2p_generate(x, len_output=8, eval_cp_ix=7, model_mode="full", skip_frame=False)
    assert len(seq) == 8 and torch.equal(seq[0], x[0])
    for fr in seq:
        assert fr.shape == x[0].shape and torch.isfinite(fr).all() and fr.min() >= 0 and fr.max() <= 1
