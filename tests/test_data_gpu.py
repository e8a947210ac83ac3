"""DevicePrefetcher (p2pvg_b200/data.py): batches arrive in order, intact, from pinned and pageable host memory."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prefetcher_order_and_values():
    from p2pvg_b200.data import DevicePrefetcher
    host = [torch.full((3, 4, 1, 8, 8), float(i)) + torch.arange(8.0) for i in range(7)]
    host[2] = host[2].pin_memory()
    seen = []
    pf = DevicePrefetcher(iter(host), "cuda")
    for x in pf:
        assert x.is_cuda
        y = x * 2          # consumer work on the current stream
        pf.release()
        seen.append((x.clone(), y))
    assert len(seen) == len(host)
    for i, (x, y) in enumerate(seen):
        assert torch.equal(x.cpu(), host[i]) and torch.equal(y.cpu(), host[i] * 2)


def test_prefetcher_rejects_cpu_target():
    from p2pvg_b200.data import DevicePrefetcher
    with pytest.raises(RuntimeError):
        DevicePrefetcher(iter([]), "cpu")


def test_prefetcher_early_release_through_the_graph_step():
    """early_release=True: the CUDA-graph train step reports the moment it has copied the batch into its static input buffer and
    the slot is refilled while the step still runs.  Every step must still train on ITS batch: the losses of a run fed by the
    prefetcher equal, step for step, those of a run fed with the batches already resident."""
    import os
    import types
    import numpy as np
    from p2pvg_b200.data import DevicePrefetcher
    from p2pvg_b200.models import dcgan_64
    from p2pvg_b200.models.p2p_model import P2PModel
    T, B, n = 4, 4, 7
    g = torch.Generator().manual_seed(11)
    host = [torch.rand(T, B, 1, 64, 64, generator=g).pin_memory() for _ in range(n)]

    def run(feed):
        os.environ["P2PVG_PRECISION"], os.environ["P2PVG_GRAPH"] = "bf16", "1"
        opt = types.SimpleNamespace(dataset="mnist", backbone_net=dcgan_64, lr=1e-3, beta1=0.9, beta=1e-4, weight_cpc=100.0,
                                    weight_align=0.5, skip_prob=0.0, n_past=1, last_frame_skip=False, batch_size=B)
        torch.manual_seed(1)
        model = P2PModel(B, 1, 128, 10, 256, 1, 1, 2, opt=opt).cuda()
        out = []
        for i, x in enumerate(feed()):
            torch.manual_seed(200 + i)
            out.append(np.array(model(x, 0, T - 1), dtype=np.float64))
        return out

    def prefetched():
        pf = DevicePrefetcher(iter(host), "cuda", early_release=True)
        for x in pf:
            assert hasattr(x, "_p2pvg_on_consumed")
            yield x

    a = run(prefetched)
    b = run(lambda: (h.cuda() for h in host))
    assert len(a) == len(b) == n
    for i, (u, v) in enumerate(zip(a, b)):
        np.testing.assert_allclose(u, v, rtol=1e-5, err_msg=f"step {i}")
    assert len({tuple(u) for u in a}) == n      # the batches differ, so do the losses


@pytest.mark.parametrize("streams", [1, 3])
def test_prefetcher_chunked_copies(streams):
    """copy_streams > 1: every batch is split into pieces copied on separate streams; the consumer must see whole batches"""
    from p2pvg_b200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(5)
    host = [torch.rand(3, 5, 1, 300, 300, generator=g).pin_memory() for _ in range(5)]   # 1.35 M elements each
    got = []
    pf = DevicePrefetcher(iter(host), "cuda", copy_streams=streams)
    for x in pf:
        got.append(x.sum(dtype=torch.float64) + 0)     # consumer work on the current stream, before the slot is refilled
        pf.release()
    for a, h in zip(got, host):
        assert abs(a.item() - h.sum(dtype=torch.float64).item()) < 1e-6 * h.numel()
