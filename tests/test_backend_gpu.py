"""GPU parity of the array seam (pydynet_b200.backend.ndarray — the object that replaces cupy.ndarray behind
Device.xp, reference pydynet/cuda.py:90-91) against NumPy, which is the arithmetic the reference's CPU path
bottoms out in.  Mirrors reference tests/test_tensor_basic.py:87-117 and
tests/test_ops_extended.py:15-91 at the array level."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xp():
    from pydynet_b200 import backend
    from pydynet_b200.backend import lib
    lib.call("pdn_init", 0)
    return backend


def _rand(rng, shape, dtype):
    return rng.standard_normal(shape).astype(dtype)


SHAPE_PAIRS = [((3, 4), (3, 4)), ((5, 1, 7), (1, 6, 7)), ((2, 3, 4, 5), (5, )), ((1, ), (4, 3)), ((8, 1), (1, 9)),
               ((257, 513), (513, )), ((64, 1, 33), (64, 17, 1)), ((), (3, 2))]


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.float16])
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "pow", "maximum", "minimum"])
def test_binary_broadcast(xp, op, dtype):
    rng = np.random.default_rng(0)
    for sa, sb in SHAPE_PAIRS:
        a, b = _rand(rng, sa, dtype), _rand(rng, sb, dtype)
        if op == "pow":
            a = np.abs(a) + dtype(0.5)
        da, db = xp.array(a), xp.array(b)
        f = {"add": lambda x, y: x + y, "sub": lambda x, y: x - y, "mul": lambda x, y: x * y, "div": lambda x, y: x / y,
             "pow": lambda x, y: x**y}.get(op)
        if f is None:
            ref, got = getattr(np, op)(a, b), getattr(xp, op)(da, db)
        else:
            ref, got = f(a, b), f(da, db)
        assert got.shape == ref.shape and got.dtype == ref.dtype
        tol = {np.float16: 2e-3, np.float32: 2e-6, np.float64: 1e-12}[dtype]
        np.testing.assert_allclose(got.get(), ref, rtol=tol, atol=tol)


def test_mixed_dtype_promotion(xp):
    rng = np.random.default_rng(1)
    for d1 in (np.float16, np.float32, np.float64):
        for d2 in (np.float16, np.float32, np.float64):
            a, b = _rand(rng, (4, 5), d1), _rand(rng, (5, ), d2)
            got = xp.array(a) * xp.array(b)
            ref = a * b
            assert got.dtype == ref.dtype
            np.testing.assert_allclose(got.get(), ref, rtol=2e-3, atol=2e-3)


def test_scalar_ops_and_inplace(xp):
    rng = np.random.default_rng(2)
    a = _rand(rng, (33, 65), np.float32)
    d = xp.array(a)
    np.testing.assert_allclose((2.0 - d).get(), 2.0 - a, rtol=1e-6)
    np.testing.assert_allclose((d / 3).get(), a / 3, rtol=1e-6)
    np.testing.assert_allclose((1 / (1 + xp.exp(-d))).get(), 1 / (1 + np.exp(-a)), rtol=1e-5)
    d += 1.5
    a += 1.5
    d *= xp.array(a[0])
    a *= a[0].copy()
    np.testing.assert_allclose(d.get(), a, rtol=1e-6)
    assert (d > 0).dtype == np.bool_
    np.testing.assert_array_equal((d > 0).get(), a > 0)
    np.testing.assert_array_equal((d == d).get(), np.ones_like(a, bool))


@pytest.mark.parametrize("name", ["exp", "log", "abs", "sign", "sqrt", "square"])
def test_unary(xp, name):
    rng = np.random.default_rng(3)
    a = _rand(rng, (7, 129), np.float64)
    if name in ("log", "sqrt"):
        a = np.abs(a) + 0.1
    got = getattr(xp, name)(xp.array(a)).get()
    np.testing.assert_allclose(got, getattr(np, name)(a), rtol=1e-12, atol=1e-12)
    a32 = a.astype(np.float32)
    got = getattr(xp, name)(xp.array(a32)[:, ::2]).get()  # strided view
    np.testing.assert_allclose(got, getattr(np, name)(a32[:, ::2]), rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_reductions(xp, dtype):
    rng = np.random.default_rng(4)
    a = _rand(rng, (6, 37, 129), dtype)
    d = xp.array(a)
    tol = 1e-5 if dtype == np.float32 else 1e-12
    for axis, keep in [(None, False), (0, False), (1, True), (2, False), ((0, 2), True), ((1, 2), False), (-1, True)]:
        for name in ("sum", "mean", "max", "min"):
            ref = getattr(a, name)(axis=axis, keepdims=keep)
            got = getattr(d, name)(axis=axis, keepdims=keep)
            assert got.shape == np.shape(ref), (name, axis, keep)
            np.testing.assert_allclose(got.get(), ref, rtol=tol, atol=tol * 10)
    for axis in (None, 0, 1, 2):
        for name in ("argmax", "argmin"):
            ref = getattr(a, name)(axis=axis)
            got = getattr(d, name)(axis=axis)
            assert got.dtype == np.int64
            np.testing.assert_array_equal(got.get(), ref)
    # first-occurrence tie rule, large row
    t = np.zeros((3, 40000), dtype)
    t[:, [5, 70, 39999]] = 7
    np.testing.assert_array_equal(xp.array(t).argmax(axis=1).get(), t.argmax(axis=1))
    # transposed (non-contiguous) input
    np.testing.assert_allclose(d.transpose(2, 0, 1).sum(axis=1).get(), a.transpose(2, 0, 1).sum(axis=1), rtol=tol, atol=tol * 10)
    big = _rand(rng, (1 << 20, ), dtype)
    np.testing.assert_allclose(xp.array(big).sum().get(), big.sum(dtype=np.float64), rtol=1e-5, atol=1e-2)


def test_views_and_indexing(xp):
    rng = np.random.default_rng(5)
    a = _rand(rng, (4, 6, 8), np.float32)
    d = xp.array(a)
    np.testing.assert_array_equal(d.reshape(24, 8).get(), a.reshape(24, 8))
    np.testing.assert_array_equal(d.transpose(2, 0, 1).reshape(8, 24).get(), a.transpose(2, 0, 1).reshape(8, 24))
    np.testing.assert_array_equal(d[1:3, ::2, -1].get(), a[1:3, ::2, -1])
    np.testing.assert_array_equal(d[..., 0].get(), a[..., 0])
    np.testing.assert_array_equal(d[:, None, 2].get(), a[:, None, 2])
    idx = np.array([3, 0, 3, 1])
    np.testing.assert_array_equal(d[idx].get(), a[idx])
    np.testing.assert_array_equal(d[:, [0, 5, 5]].get(), a[:, [0, 5, 5]])
    np.testing.assert_array_equal(d[[0, 1, 2], [1, 2, 3]].get(), a[[0, 1, 2], [1, 2, 3]])
    m = a > 0.3
    np.testing.assert_array_equal(d[xp.array(m)].get(), a[m])
    # assignment: basic, fancy (last write wins), mask
    d[1, 2:4] = 5.0
    a[1, 2:4] = 5.0
    v = _rand(rng, (4, 6, 8), np.float32)
    d[idx] = xp.array(v)
    a[idx] = v
    np.testing.assert_array_equal(d.get(), a)
    d[:, 0, :] = xp.array(v[:, 1, :])
    a[:, 0, :] = v[:, 1, :]
    np.testing.assert_array_equal(d.get(), a)
    # view aliasing (KV-cache style in-place write, reference llm/llama/model.py:106-107)
    w = d[2]
    w[...] = 1.25
    a[2] = 1.25
    np.testing.assert_array_equal(d.get(), a)
    # add.at semantics
    z = np.zeros((5, 3), np.float32)
    dz = xp.array(z)
    ii = np.array([1, 1, 4, 1])
    vals = _rand(rng, (4, 3), np.float32)
    xp.add_at(dz, ii, xp.array(vals))
    np.add.at(z, ii, vals)
    np.testing.assert_allclose(dz.get(), z, rtol=1e-6)
    c = xp.concatenate([d, d[:, :2]], axis=1)
    np.testing.assert_array_equal(c.get(), np.concatenate([a, a[:, :2]], axis=1))
    p = xp.pad(d, ((0, 0), (1, 2), (3, 0)))
    np.testing.assert_array_equal(p.get(), np.pad(a, ((0, 0), (1, 2), (3, 0))))
    np.testing.assert_array_equal(d.astype(np.float64).get(), a.astype(np.float64))
    np.testing.assert_array_equal(xp.array(np.arange(10)).astype(np.float32).get(), np.arange(10, dtype=np.float32))


def _relerr(got, ref):
    return float(np.linalg.norm(got.astype(np.float64) - ref) / max(np.linalg.norm(ref), 1e-30))


MM_SHAPES = [((512, 512), (512, 512)), ((256, 2450), (2450, 500)), ((50176, 180), (180, 50)), ((3, 5, 64, 96), (96, 80)),
             ((4, 1, 130, 70), (1, 3, 70, 65)), ((2, 8, 512, 64), (2, 8, 64, 512)), ((1, 288), (288, 32000)), ((7, 9), (9, 3)),
             ((16, 288), (288, 768)), ((1000, 33), (33, 17)), ((300, 1000), (1000, 260))]


@pytest.mark.parametrize("sa,sb", MM_SHAPES)
def test_matmul_f32(xp, sa, sb):
    """x @ y (reference tensor.py:657-659). Tolerance: normwise 1e-4 vs fp64 NumPy (BASELINE north_star);
    the tcgen05 BF16x3 path is expected at ~5e-6."""
    rng = np.random.default_rng(6)
    a, b = _rand(rng, sa, np.float32), _rand(rng, sb, np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    got = (xp.array(a) @ xp.array(b))
    assert got.shape == ref.shape and got.dtype == np.float32
    err = _relerr(got.get(), ref)
    assert err < 2e-5, err


def test_matmul_paths_and_views(xp):
    from pydynet_b200.backend import lib
    rng = np.random.default_rng(7)
    a, b = _rand(rng, (512, 384), np.float32), _rand(rng, (384, 640), np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    da, db = xp.array(a), xp.array(b)
    for prec, path in ((1, 0), (2, 1)):
        out = xp.gemm_into(None, da, db, prec=prec)
        assert lib.load().pdn_gemm_last_path() == path
        assert _relerr(out.get(), ref) < 2e-5
    # grads use swapaxes views: g @ Bᵀ and Aᵀ @ g (reference tensor.py:669-676)
    g = _rand(rng, (512, 640), np.float32)
    dg = xp.array(g)
    assert _relerr((dg @ db.swapaxes(-1, -2)).get(), g.astype(np.float64) @ b.T.astype(np.float64)) < 2e-5
    assert _relerr((da.swapaxes(-1, -2) @ dg).get(), a.T.astype(np.float64) @ g.astype(np.float64)) < 2e-5
    # bias epilogue + accumulate
    bias = _rand(rng, (640, ), np.float32)
    out = xp.gemm_into(None, da, db, bias=xp.array(bias))
    assert _relerr(out.get(), ref + bias) < 2e-5
    xp.gemm_into(out, da, db, accumulate=True)
    assert _relerr(out.get(), 2 * ref + bias) < 2e-5
    # fp64 / fp16 / 1-D operands
    a64, b64 = a[:40, :30].astype(np.float64), b[:30, :20].astype(np.float64)
    np.testing.assert_allclose((xp.array(a64) @ xp.array(b64)).get(), a64 @ b64, rtol=1e-12)
    v = _rand(rng, (384, ), np.float32)
    np.testing.assert_allclose((da @ xp.array(v)).get(), a @ v, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose((xp.array(v) @ db).get(), v @ b, rtol=1e-4, atol=1e-4)


def test_operand_plane_cache_tracks_writes(xp):
    """pdn_gemm_cached keeps the tensor-core operand planes of a buffer until its write counter changes or it is freed: every
    way the Python layer writes device memory in place must invalidate them (stale planes would silently reuse old values)."""
    import ctypes as C
    from pydynet_b200.backend import lib
    rng = np.random.default_rng(11)
    a, b = _rand(rng, (256, 192), np.float32), _rand(rng, (192, 320), np.float32)
    da, db = xp.array(a), xp.array(b)

    def check(tag):
        got = (da @ db).get()
        assert _relerr(got, da.get().astype(np.float64) @ db.get().astype(np.float64)) < 2e-5, tag

    def stats():
        v = [C.c_uint64() for _ in range(4)]
        lib.call("pdn_plane_cache_stats", *[C.byref(x) for x in v])
        return [x.value for x in v]

    check("first")
    h0 = stats()[0]
    check("second (cache hit)")
    assert stats()[0] >= h0 + 2  # both operands served from the cache
    da[3:7, :] = 5.0  # __setitem__
    check("after setitem")
    db += 1.5  # in-place arithmetic
    check("after iadd")
    da[...] = xp.array(_rand(rng, (256, 192), np.float32))  # copy-into
    check("after copy")
    # transposed re-use of the same planes (MN-major) and a view of the same buffer
    gt = (da.swapaxes(0, 1) @ xp.array(_rand(rng, (256, 64), np.float32))).get()
    assert gt.shape == (192, 64)
    # freed and re-allocated memory must not inherit planes: allocate/free in a loop with fresh values
    for i in range(4):
        t = xp.array(_rand(rng, (256, 192), np.float32))
        ref = t.get().astype(np.float64) @ db.get().astype(np.float64)
        assert _relerr((t @ db).get(), ref) < 2e-5, f"realloc {i}"
        del t
