"""human3.6m pose backbone on the GPU: kernels vs emulation, train step vs oracle / reference fixture, drop-in API."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import p2p_oracle as O
from p2pvg_b200.engine import StepPlan

pytestmark = pytest.mark.gpu
CFG = dict(g_dim=128, z_dim=10, rnn_size=256, backbone="mlp", predictor_rnn_layers=2, posterior_rnn_layers=1, prior_rnn_layers=1)
GOLD = os.path.join(os.path.dirname(__file__), "golden", "step_h36m_mlp.pt")


def test_layernorm_relu_mse_kernels():
    from p2pvg_b200._lib import CudaKernels
    from tests.emu_mlp import EmuKernelsMLP
    Kc, Ke = CudaKernels("cuda"), EmuKernelsMLP("cuda")
    torch.manual_seed(0)
    rows, C = 1000, 128
    x = torch.randn(rows, C, device="cuda") * 2 + 0.5
    gamma, beta = torch.randn(C, device="cuda") * 0.1 + 1, torch.randn(C, device="cuda") * 0.1
    dy = torch.randn(rows, C, device="cuda")
    res = []
    for K in (Kc, Ke):
        y, mean, rstd = torch.empty_like(x), torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
        K.layernorm_fwd(x, gamma, beta, y, mean, rstd, rows, C)
        dx, dg, db = torch.empty_like(x), torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
        K.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dg, db, rows, C)
        a = x.clone()
        K.act_fwd(a, a.numel(), 4)
        da = torch.empty_like(a)
        K.act_bwd(dy, a, da, a.numel(), 4)
        G, E = 3, 5 * 51
        pred, tg = torch.randn(G, E, device="cuda"), torch.randn(6, E, device="cuda")
        tgt = torch.tensor([1, 4, 5], dtype=torch.int32, device="cuda")
        coef = torch.tensor([0.1, 0.2, 3.0], device="cuda")
        torch.manual_seed(1)
        dp, part = torch.empty(G, E, device="cuda"), torch.zeros(G * 32, device="cuda")
        K.mse_plain(pred, tg, tgt, coef, G, E, dp, part)
        res.append((y, mean, rstd, dx, dg, db, a, da, part.reshape(G, 32).sum(1)))
        torch.manual_seed(0)
        torch.randn(1)
    for i, (a, b) in enumerate(zip(*res[:2])):
        if i == 8:
            continue
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-4), i


def run_step(optkw, T, B, precision, np_seed=0):
    from p2pvg_b200._lib import CudaKernels
    from p2pvg_b200.engine_mlp import TrainEngineMLP
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    state = O.build_state(CFG, seed=1)
    opt = O.default_opt(**optkw)
    opt["batch_size"] = opt["batch_size"] or B
    adt = torch.float32 if precision == "fp32" else torch.bfloat16
    eng = TrainEngineMLP(O.clone_state(state), CFG, opt, CudaKernels("cuda"), act_dtype=adt)
    adam = {m: O.new_adam_state(state[m]) for m in O.MODULES}
    x = torch.randn(T, B, 17, 3, generator=torch.Generator().manual_seed(5))
    np.random.seed(np_seed)
    probs = np.random.uniform(0, 1, T - 1)
    eps = O.draw_eps(StepPlan(T, probs, opt).S, B, 10, seed=11)
    ref = O.train_step(state, adam, x, opt, "mlp", eps, probs, mode="A")
    got = eng.step(x.cuda(), probs=probs, eps=eps.cuda())
    return ref, got, eng


@pytest.mark.parametrize("precision,rtol,mincos", [("fp32", 1e-4, 1 - 1e-5), ("bf16", 1e-2, 0.98)])
@pytest.mark.parametrize("optkw,T,B,seed", [({}, 6, 5, 0), (dict(skip_prob=0.5, n_past=2, last_frame_skip=True), 9, 3, 3)])
def test_mlp_step_vs_oracle(precision, rtol, mincos, optkw, T, B, seed):
    ref, got, eng = run_step(optkw, T, B, precision, seed)
    np.testing.assert_allclose(got, np.array(ref["losses"], dtype=np.float32), rtol=rtol, atol=1e-6)
    for m in O.MODULES:
        for k, gref in ref["grads"][m].items():
            g = eng.arena[m].g[k].cpu()
            cos = torch.nn.functional.cosine_similarity(g.flatten().double(), gref.flatten().double(), dim=0).item()
            assert cos >= mincos, f"{precision} grad {m}.{k}: cosine {cos:.6f}"


def test_mlp_step_vs_reference_golden():
    from p2pvg_b200._lib import CudaKernels
    from p2pvg_b200.engine_mlp import TrainEngineMLP
    fix = torch.load(GOLD, weights_only=False)
    state = O.build_state(fix["cfg"], seed=fix["init_seed"])
    eng = TrainEngineMLP(state, fix["cfg"], dict(fix["opt"]), CudaKernels("cuda"), act_dtype=torch.float32)
    rec = fix["steps"][0]
    got = eng.step(rec["x"].cuda(), probs=rec["probs"].numpy(), eps=rec["eps"].cuda())
    np.testing.assert_allclose(got, np.array(rec["losses"], dtype=np.float32), rtol=1e-4, atol=1e-7)
    for m, digs in rec["grad_digest"].items():
        for k, d in digs.items():
            f = eng.arena[m].g[k].double().reshape(-1).cpu()
            err = (f[d["idx"]] - d["samples"]).abs().max().item()
            assert err <= 2e-3 * max(d["absmax"], 1e-30), f"grad {m}.{k}: {err}"


def test_p2pmodel_pose_dropin():
    from p2pvg_b200.models import h36m_mlp
    from p2pvg_b200.models.p2p_model import P2PModel
    os.environ["P2PVG_PRECISION"], os.environ["P2PVG_GRAPH"] = "bf16", "1"
    T, B = 8, 16
    opt = types.SimpleNamespace(dataset="h36m", backbone_net=h36m_mlp, lr=1e-3, beta1=0.9, beta=1e-4, weight_cpc=100.0,
                                weight_align=0.5, skip_prob=0.0, n_past=1, last_frame_skip=False, batch_size=B)
    torch.manual_seed(1)
    model = P2PModel(B, 1, 128, 10, 512, 1, 1, 2, opt=opt).cuda()   # rnn_size 512 (BASELINE config 5)
    x = torch.randn(T, B, 17, 3, generator=torch.Generator().manual_seed(2)).cuda() * 3
    outs = [model((None, x, None), 0, T - 1) for _ in range(4)]      # eager, capture, replay, replay
    for o in outs:
        assert len(o) == 4 and all(np.isfinite(float(v)) for v in o)
    assert float(outs[-1][0]) < float(outs[0][0])                    # the reconstruction loss goes down
    model.eval()
    seq = model.p2p_generate((None, [t for t in x], None), len_output=8, eval_cp_ix=7)
    assert len(seq) == 8 and seq[3].shape == (B, 17, 3) and torch.isfinite(seq[3]).all()
