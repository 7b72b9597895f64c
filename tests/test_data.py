"""SURVEY.md §8(f) row f3 — DataLoader / samplers (reference pydynet/data.py:4-123): batch order and contents against fixtures the
unmodified reference produced (tests/golden/make_golden_data.py), and the device input pipeline (pinned double-buffered H2D on a copy
stream) on cuda:0."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_loader.json")


def _idx_dataset(n):
    from pydynet_b200.data import Dataset

    class Idx(Dataset):

        def __getitem__(self, index):
            return list(index)

        def __len__(self):
            return n

    return Idx()


def test_sampler_batches_match_reference():
    from pydynet_b200.data import DataLoader
    g = json.load(open(GOLD))
    assert len(g["cases"]) == 16
    for c in g["cases"]:
        np.random.seed(c["seed"])
        dl = DataLoader(_idx_dataset(c["n"]), c["batch_size"], c["shuffle"], c["drop_last"])
        assert [[b for b in dl] for _ in range(2)] == c["epochs"], c
        assert len(dl) == c["len"] == len(dl.batch_sampler)


def test_data_loader_helper_matches_reference():
    from pydynet_b200.data import data_loader
    import pydynet_b200 as pdn
    g = json.load(open(GOLD))
    X, y = np.arange(40).reshape(10, 4), np.arange(10)
    np.random.seed(g["xy_seed"])
    got = [[bx.tolist(), by.tolist()] for bx, by in data_loader(X, y, 4, True)]
    assert got == g["xy"]
    # Tensor-backed sets (the examples' usage, mnist.py:143-152) yield Tensors in the same order
    np.random.seed(g["xy_seed"])
    got_t = [[bx.numpy().tolist(), by.numpy().tolist()] for bx, by in data_loader(pdn.Tensor(X), pdn.Tensor(y), 4, True)]
    assert got_t == g["xy"]


@pytest.mark.gpu
def test_device_prefetch_pipeline_matches_host_loader():
    import pydynet_b200 as pdn
    from pydynet_b200.data import data_loader
    rng = np.random.default_rng(0)
    X = rng.standard_normal((203, 3, 9, 5)).astype(np.float32)
    y = rng.integers(0, 10, 203)
    for shuffle in (False, True):
        np.random.seed(11)
        host = [(bx.copy(), by.copy()) for bx, by in data_loader(X, y, 32, shuffle)]
        np.random.seed(11)
        got, sums = [], []
        for bx, by in data_loader(X, y, 32, shuffle, device="cuda:0"):
            assert bx.device.is_cuda and by.device.is_cuda and bx.dtype == np.float32 and by.dtype == y.dtype
            sums.append((bx * 2.0).sum())  # a kernel consumes the batch while the next one is in flight
            got.append((bx, by))
        assert len(got) == len(host) == 7
        for (hx, hy), (dx, dy), s in zip(host, got, sums):
            np.testing.assert_array_equal(dx.numpy(), hx)
            np.testing.assert_array_equal(dy.numpy(), hy)
            np.testing.assert_allclose(s.item(), 2.0 * hx.astype(np.float64).sum(), rtol=1e-4, atol=1e-3)
    # cpu Tensors in the set (the examples' usage) and a host-side dtype cast
    np.random.seed(3)
    ref = [bx.copy() for bx, _ in data_loader(X.astype(np.float64), y, 50, True)]
    np.random.seed(3)
    out = [bx for bx, _ in data_loader(pdn.Tensor(X.astype(np.float64)), pdn.Tensor(y), 50, True, device="cuda:0", dtype=np.float32)]
    for r, o in zip(ref, out):
        assert o.dtype == np.float32
        np.testing.assert_array_equal(o.numpy(), r.astype(np.float32))
