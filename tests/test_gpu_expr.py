"""Device-side expressions and filter compaction (csrc/expr.cu, vaex_b200/expression.py): bit-exact
against numpy's evaluation of the same expression — what the reference does per chunk (vaex/scopes.py:108-128) — and, for the
filtered / virtual-column frames, against the oracle fed with the numpy-evaluated columns."""
import numpy as np
import pytest

from helpers import to_device

pytestmark = pytest.mark.gpu


def _columns(n, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n).astype("f4")
    y = rng.standard_normal(n) * 1e3
    x[::97] = np.nan
    y[::89] = 0.0
    return {"x": x, "y": y, "i": rng.integers(-2 ** 31, 2 ** 31 - 1, n).astype("i4"), "j": rng.integers(-10, 10, n).astype("i8"),
            "u": rng.integers(0, 255, n).astype("u1"), "h": rng.integers(-32768, 32767, n).astype("i2")}


EXPRESSIONS = ["x + y", "x * 2.5", "i + 1", "i * 2.5", "i / j", "x / 3", "(x - 1) * (y + 2) > 0.5", "(i > 3) & (x < 0.5)", "~(x > 0)", "-x", "abs(i)", "sqrt(j)",
               "sqrt(x)", "h + h", "u * u", "x.astype('float64') + 1", "i + j", "u + i", "h * 100", "(x > 0) | (y != y)", "x - x * y / (y + 1.5)", "i * i",
               "y / 0.1", "x == x", "(y >= 0) & ~(x != x)", "j / j", "-i", "abs(x) - 0.25"]


@pytest.mark.parametrize("device", [False, True])
def test_expressions_are_bit_identical_to_numpy(device):
    from vaex_b200.frame import Frame
    n = 50_003
    cols = _columns(n)
    ns = dict(cols, abs=np.abs, sqrt=np.sqrt)
    df = Frame({k: (to_device(v) if device else v) for k, v in cols.items()})
    for e in EXPRESSIONS:
        with np.errstate(all="ignore"):
            want = eval(e, dict(ns))
        got = df.evaluate(e)
        assert got.dtype == want.dtype, e
        # bit for bit; the one freedom is the payload of a NaN (the GPU produces the canonical quiet NaN, x86 propagates an operand's)
        if want.dtype.kind == "f":
            nan = np.isnan(want)
            assert np.array_equal(np.isnan(got), nan), e
            assert np.array_equal(got[~nan].view("u1"), np.ascontiguousarray(want[~nan]).view("u1")), e
        else:
            assert np.array_equal(got, want), e


def test_ordinal_values_and_compaction():
    from vaex_b200 import expression, hash as vhash
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(3)
    n = 20_011
    k = rng.integers(0, 500, n).astype("i8") * 7 + 3
    f = rng.integers(0, 40, n).astype("f8") * 0.5
    f[::41] = np.nan
    for keys in (k, f):
        hm = vhash.HashMapUnique(keys.dtype, 3)
        hm.add(keys[: n // 2])  # the second half holds unknown keys -> -1
        df = Frame({"k": keys}, variables={"hm": hm})
        got = df.evaluate("_ordinal_values(k, hm)")
        want = hm.map(keys)
        assert got.dtype == want.dtype and np.array_equal(got, want)
    cols = _columns(n, seed=9)
    keep = (cols["x"] > 0) & (cols["j"] != 0)
    for device in (False, True):
        inp = [to_device(v) if device else v for v in cols.values()]
        count, out = expression.compact(0, to_device(keep.view("u1")) if device else keep, inp)
        assert count == int(keep.sum())
        for name, o in zip(cols, out):
            assert np.array_equal(o.to_numpy().view("u1"), cols[name][keep].view("u1")), name


@pytest.mark.parametrize("device", [False, True])
def test_filtered_and_virtual_frames_match_the_oracle(device, oracle):
    """df[df.x > 0.1], a virtual column as binby axis, an expression as selection: the reference evaluates these with numpy per
    chunk and compresses the columns (vaex/execution.py:516-551); here everything stays on the device"""
    from vaex_b200.frame import Frame
    from vaex_b200 import execution
    n = 300_007
    cols = _columns(n, seed=5)
    cols["y"] = cols["y"] / 1e3
    src = {k: (to_device(v) if device else v) for k, v in cols.items()}
    df = Frame(src, executor=execution.Executor(nthreads=3, chunk_size=70_001))
    with np.errstate(all="ignore"):
        keep = (cols["x"] > 0.1) & (cols["y"] < 1.5)
        r = cols["x"] * 2.5 + cols["y"]
        sel = (cols["j"] >= 0)
    fdf = df.filter("(x > 0.1) & (y < 1.5)")
    fdf.add_virtual_column("r", "x * 2.5 + y")
    lim = [[-3, 3], [-4, 4]]
    # count + sum over the filtered frame, binned by a real and a virtual column
    got = fdf.count(binby=["x", "r"], limits=lim, shape=[32, 16], edges=True)
    b = [oracle.scalar(cols["x"][keep], -3, 3, 32), oracle.scalar(r[keep], -4, 4, 16)]
    assert np.array_equal(got, oracle.binby(b, [oracle.agg("count")])[0])
    got = fdf.sum("y", binby=["x", "r"], limits=lim, shape=[32, 16], edges=True)
    want = oracle.binby(b, [oracle.agg("sum", cols["y"][keep])])[0]
    assert np.allclose(got, want, rtol=1e-6, atol=1e-9 * np.abs(want).max())
    # an expression as selection, on top of the filter
    got = fdf.count("y", binby=["x"], limits=[[-3, 3]], shape=32, selection="j >= 0", edges=True)
    want = oracle.binby([oracle.scalar(cols["x"][keep], -3, 3, 32)], [oracle.agg("count", cols["y"][keep], mask=sel[keep].astype("u1"))])[0]
    assert np.array_equal(got, want)
    # the limits pre-pass sees the filter and the virtual column too
    assert np.array_equal(fdf.minmax("r", raw=True), oracle.minmax(r[keep], raw=True), equal_nan=True)
    assert np.array_equal(fdf.minmax("y", raw=True), oracle.minmax(cols["y"][keep], raw=True))
    # groupby on a filtered frame
    out = fdf.groupby("j").agg({"y": ["sum", "count"]})
    order = np.argsort(out["j"])
    ks = np.unique(cols["j"][keep])
    assert np.array_equal(np.asarray(out["j"])[order], ks)
    assert np.array_equal(np.asarray(out["count"])[order], [int((keep & (cols["j"] == v)).sum()) for v in ks])


def test_cancel_through_progress_and_unsupported_syntax():
    from vaex_b200.frame import Frame
    from vaex_b200 import execution
    n = 100_000
    cols = _columns(n)
    df = Frame(cols, executor=execution.Executor(nthreads=1, chunk_size=10_000))
    with pytest.raises(execution.UserAbort):
        df.count(binby=["x"], limits=[[-3, 3]], shape=8, progress=lambda f: f < 0.25)
    with pytest.raises(NotImplementedError):
        df.filter("x ** 2 > 1")
