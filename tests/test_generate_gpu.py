"""CUDA p2p_generate (p2pvg_b200/infer.py, through the drop-in P2PModel.p2p_generate) against the frames the reference's
own P2PModel.p2p_generate produced (tests/golden/gen_*.pt): eval-mode BatchNorm, model_mode in {full, posterior, prior},
skip_frame in {False, True}, the reference's NumPy skip draws and eps stream replayed.  Also: resume from a checkpoint
written by the reference's own save().  Tolerances: fp32 mode 2e-4 absolute on frames in [0,1]; bf16 mode 4e-2 worst
pixel / 6e-3 mean absolute (5 autoregressive steps through bf16 conv stacks)."""
import glob
import os
import types

import numpy as np
import pytest
import torch

from oracle import p2p_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
GEN = sorted(glob.glob(os.path.join(GOLD, "gen_*.pt")))


def build_model(fix):
    from p2pvg_b200.models import dcgan_64
    from p2pvg_b200.models.p2p_model import P2PModel
    o = fix["opt"]
    opt = types.SimpleNamespace(dataset="mnist", backbone_net=dcgan_64, **o)
    torch.manual_seed(fix["init_seed"])
    cfg = fix["cfg"]
    model = P2PModel(o["batch_size"], cfg["channels"], cfg["g_dim"], cfg["z_dim"], cfg["rnn_size"], 1, 1, 2, opt=opt)
    ref = O.build_state(cfg, seed=fix["init_seed"])
    for m in ("encoder", "decoder", "posterior"):   # same seed -> bit-identical initial weights as the reference
        sd = getattr(model, m).state_dict()
        for k, v in ref[m].items():
            assert torch.equal(sd[k], v), f"{m}.{k}"
    for m, bufs in fix["bn_buffers"].items():
        mod = getattr(model, m)
        for k, v in bufs.items():
            owner, _, leaf = k.rpartition(".")
            getattr(mod.get_submodule(owner), leaf).copy_(v)
    return model.cuda().eval()


@pytest.mark.parametrize("precision,tol_max,tol_mean", [("fp32", 2e-4, 2e-5), ("bf16", 4e-2, 6e-3)])
@pytest.mark.parametrize("path", GEN, ids=lambda p: os.path.basename(p)[4:-3])
def test_p2p_generate_matches_reference(path, precision, tol_max, tol_mean):
    from p2pvg_b200.infer import eps_stream
    fix = torch.load(path, weights_only=False)
    prev = os.environ.get("P2PVG_PRECISION")
    os.environ["P2PVG_PRECISION"] = precision
    try:
        model = build_model(fix)
        x = fix["x"].cuda()
        for run in fix["runs"]:
            np.random.seed(run["np_seed"])   # the drop-in draws its skip vector from NumPy's global RNG like the reference
            draws = [run["eps"][s, j] for s in range(run["n_exec"]) for j in (0, 1)]
            with eps_stream(draws) as es:
                seq = model.p2p_generate(x, fix["len_output"], fix["eval_cp_ix"], model_mode=run["model_mode"], skip_frame=run["skip_frame"])
                assert len(es.draws) == 0, "the drop-in executed fewer gaussian-LSTM calls than the reference"
            what = f"{precision} {run['model_mode']}/skip_frame={run['skip_frame']}"
            assert len(seq) == fix["len_output"]
            assert [bool((f == 0).all()) for f in seq] == run["zero_frames"], what
            for i, (f, d) in enumerate(zip(seq, run["digests"])):
                v = f.detach().double().reshape(-1).cpu()
                assert v.numel() == d["numel"]
                err = (v[d["idx"]] - d["samples"]).abs().max().item()
                assert err <= tol_max, f"{what} frame {i}: {err:.3e}"
            for f, r in ((seq[-1], run["last"]), (seq[len(seq) // 2], run["mid"])):
                e = (f.float().cpu() - r).abs()
                assert e.max().item() <= tol_max and e.mean().item() <= tol_mean, f"{what}: max {e.max().item():.3e} mean {e.mean().item():.3e}"
    finally:
        if prev is None:
            del os.environ["P2PVG_PRECISION"]
        else:
            os.environ["P2PVG_PRECISION"] = prev


def test_resume_from_reference_checkpoint():
    """reference save() -> drop-in load() -> the next training step reproduces the reference's next-step losses and
    post-step weights (models/p2p_model.py:289-330)."""
    from tests.test_generate_ckpt_golden import small_model
    from tests.test_oracle_golden import check_digest
    prev = os.environ.get("P2PVG_PRECISION")
    os.environ["P2PVG_PRECISION"] = "fp32"
    try:
        model, side = small_model()
        model.load(os.path.join(GOLD, "ckpt_ref_small.pth"))
        model = model.cuda().train()
        nx = side["next"]
        eng = model.engine(0)
        eng.opt = model._opt_dict()
        got = eng.step(nx["x"].cuda(), probs=nx["probs"].numpy(), eps=nx["eps"].cuda())
        np.testing.assert_allclose(got, np.array(nx["losses"], dtype=np.float32), rtol=2e-4, atol=1e-7)
        assert int(eng.arena["encoder"].step_t.item()) == 2, "Adam step counter must continue from the checkpoint"
        for m, digs in side["post_digests"].items():
            sd = getattr(model, m).state_dict()
            for k, d in digs.items():
                if k.endswith("num_batches_tracked"):
                    continue
                check_digest(sd[k].cpu(), d, 2e-3, 2e-5, f"post-step {m}.{k}")
    finally:
        if prev is None:
            del os.environ["P2PVG_PRECISION"]
        else:
            os.environ["P2PVG_PRECISION"] = prev


def test_batched_samples_equal_looped_calls():
    """p2p_generate_samples (nsample tiled along the batch, one pass) against nsample separate p2p_generate calls fed the
    same noise draws (fp32 mode: rows are independent in eval mode)."""
    from p2pvg_b200.infer import eps_stream
    fix = torch.load(GEN[0], weights_only=False)
    prev = os.environ.get("P2PVG_PRECISION")
    os.environ["P2PVG_PRECISION"] = "fp32"
    try:
        model = build_model(fix)
        x = fix["x"].cuda()
        B, z, ns, L = x.shape[1], 10, 3, fix["len_output"]
        n_exec = L - 1
        g = torch.Generator().manual_seed(9)
        draws = torch.randn(ns, n_exec, 2, B, z, generator=g)
        looped = []
        for s in range(ns):
            with eps_stream([draws[s, i, j] for i in range(n_exec) for j in (0, 1)]):
                looped.append(model.p2p_generate(x, L, L - 1, model_mode="full", skip_frame=False))
        with eps_stream([draws[:, i, j].reshape(ns * B, z) for i in range(n_exec) for j in (0, 1)]):
            batched = model.p2p_generate_samples(x, ns, L, L - 1, model_mode="full", skip_frame=False)
        assert len(batched) == ns and len(batched[0]) == L
        for s in range(ns):
            for a, b in zip(batched[s], looped[s]):
                assert a.shape == b.shape and torch.allclose(a.float(), b.float(), rtol=1e-4, atol=2e-5)
    finally:
        if prev is None:
            del os.environ["P2PVG_PRECISION"]
        else:
            os.environ["P2PVG_PRECISION"] = prev
