"""GPU parity against the committed golden vectors of the compiled reference (no oracle in the loop)."""
import numpy as np
import pytest

import golden_util
from helpers import b200_binby, same

pytestmark = pytest.mark.gpu

GOLD = golden_util.load()
BINBY = sorted(k for k in GOLD if not k.startswith(("set_", "hash64")))
SETS = sorted(k for k in GOLD if k.startswith("set_"))


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("name", BINBY)
def test_binby_matches_golden(name, device):
    binners, aggs, n, expected = golden_util.binby_case(GOLD[name])
    got = b200_binby(binners, aggs, n, device=device)
    for a, w, g in zip(aggs, expected, got):
        if a["op"] in ("sum", "sum_moment") and np.asarray(a["data"]).dtype.kind == "f":
            assert w.shape == g.shape and np.allclose(g, w, rtol=1e-6, atol=1e-9 * max(1.0, float(np.abs(w).max()))), (name, a["op"])
        else:
            assert same(w, g), (name, a["op"])


@pytest.mark.parametrize("name", SETS)
def test_sets_match_golden(name):
    from vaex_b200 import superutils
    c = GOLD[name]
    dtype, nmaps = name.split("_")[1], int(name.split("_")[2])
    s = getattr(superutils, "ordered_set_" + dtype)(nmaps)
    vals, mi = s.update(c["keys"], c["mask"], 0, return_values=True)
    assert np.array_equal(vals, c["values"]) and np.array_equal(mi, c["map_index"])
    assert np.array_equal(s.key_array(), c["key_array"], equal_nan=True)
    assert s.offsets() == c["offsets"].tolist()
    mo = s.map_ordinal(c["keys"])
    assert mo.dtype == c["map_ordinal"].dtype and np.array_equal(mo, c["map_ordinal"])
    assert [s.null_index, s.nan_index, s.null_count, s.nan_count] == c["null_nan"].tolist()


def test_hash64_golden():
    from vaex_b200 import superutils
    for i, o in zip(GOLD["hash64"]["in"], GOLD["hash64"]["out"]):
        assert superutils.hash(int(i)) == int(o)
