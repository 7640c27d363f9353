"""Generate tests/golden/strings_golden.npz from the COMPILED reference's ordered_set<> over StringList64 (src/hash_string.hpp,
src/superstring.hpp through oracle/ref_strset_shim.cpp).  Run where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden_strings.py

Per case: the update calls (arrow large_string buffers: offsets / bytes / null mask), what every call returned with
return_values=True (local ordinals + shard), the final key array, shard offsets, map_ordinal of a probe list with unknown strings
and nulls, null bookkeeping; plus known answers of std::hash<string_view>."""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_driver as R  # noqa: E402


def cases():
    rnd = random.Random(20260923)
    out = {}
    alphabet = "abcdefghijklmnopqrstuvwxyz0123456789 _-äß€"
    words = ["".join(rnd.choice(alphabet) for _ in range(rnd.choice([0, 1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 31, 40]))) for _ in range(400)]
    for nmaps in (1, 3, 7):
        for variant in ("plain", "nulls"):
            name = f"strset_{nmaps}_{variant}"
            s = R.RefStringSet(nmaps)
            ncalls = 3
            out[f"{name}/ncalls"] = np.array(ncalls)
            for c in range(ncalls):
                n = rnd.choice([1, 50, 1200])
                strs = [rnd.choice(words) for _ in range(n)]
                if variant == "nulls" and c >= 1:  # the first null arrives in the second call
                    strs = [None if rnd.random() < 0.1 else w for w in strs]
                off, by, mask = R.pack_strings(strs)
                vals, mi = s.update(strs, 0, True)
                out[f"{name}/c{c}_offsets"], out[f"{name}/c{c}_bytes"], out[f"{name}/c{c}_mask"] = off, by, mask
                out[f"{name}/c{c}_values"], out[f"{name}/c{c}_map_index"] = np.asarray(vals), np.asarray(mi)
            koff, kby, knull = R.pack_strings(s.keys())
            out[f"{name}/key_offsets"], out[f"{name}/key_bytes"], out[f"{name}/key_nulls"] = koff, kby, knull
            out[f"{name}/shard_offsets"] = np.array(s.offsets())
            probe = [rnd.choice(words + ["not a member", "zzz"]) for _ in range(300)] + [None, ""]
            poff, pby, pmask = R.pack_strings(probe)
            out[f"{name}/probe_offsets"], out[f"{name}/probe_bytes"], out[f"{name}/probe_mask"] = poff, pby, pmask
            out[f"{name}/probe_ordinals"] = s.map_ordinal(probe)
            out[f"{name}/info"] = np.array([len(s), s.null_count, s.null_index])
    mod = R.strset_module()
    kat = [b"", b"a", b"abcdefg", b"abcdefgh", b"abcdefghi", b"hello world, this is longer than sixteen bytes", "äß€".encode("utf8")]
    off, by, _ = R.pack_strings([k.decode("utf8") for k in kat])
    out["strhash/offsets"], out["strhash/bytes"] = off, by
    out["strhash/hash"] = np.array([mod.hash(k) for k in kat], dtype=np.uint64)
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    data = cases()
    np.savez_compressed(os.path.join(here, "strings_golden.npz"), **data)
    print("wrote", len(data), "arrays")
