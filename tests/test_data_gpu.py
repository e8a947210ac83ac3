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
