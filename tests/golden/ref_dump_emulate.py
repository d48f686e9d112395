"""Writes the JSON that tools/ref_dump (the cargo project) would print, but computed by the ORACLE -- only to exercise
tests/test_oracle.py::test_reference_dump_if_present in an image without cargo:

    python tests/golden/ref_dump_emulate.py /tmp/emulated.json && BPHF_REF_DUMP=/tmp/emulated.json python -m pytest tests/test_oracle.py -k reference_dump

Never commit its output as tests/golden/ref_dump.json: that file must come from the real crates."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import boomphf_oracle as o
from test_oracle import _fnv1a64_levels, _keys_for

out = {"source": "EMULATED from the oracle (not a reference run)", "hash_with_seed": [], "mphf": []}
for k in [0, 0xFFFFFFFFFFFFFFFF, 1, 2, 0x0123456789ABCDEF, 1999998]:
    for it in [0, 1, 5, 31, 32, 33, 63, 64, 100]:
        out["hash_with_seed"].append({"key": "%016x" % k, "iter": it, "h": "%016x" % o.hash_with_seed(it, k)})
spec = [("even_1000_g1.7", 1.7, None, False), ("doctest_9", 1.7, None, False), ("one", 1.7, None, False),
        ("rand_5000_g1.7", 1.7, None, False), ("rand_100k_g1.7_parallel", 1.7, None, True), ("rand_100k_g1.7_seed5", 1.7, 5, True)]
for name, gamma, seed, par in spec:
    keys = _keys_for(name[:-9] if name.endswith("_parallel") else name, o)
    m = o.OracleMphf.new_parallel(gamma, keys, seed) if par else o.OracleMphf.new(gamma, keys)
    lv = m.levels()
    e = {"name": name, "gamma": gamma, "starting_seed": seed, "n": len(keys), "constructor": "new_parallel" if par else "new",
         "bits": [int(l[0]) for l in lv], "fnv1a64": _fnv1a64_levels(lv)}
    if seed is None:
        e["hash_first"] = [int(m.hash(int(k))) for k in keys[:6]]
    if len(keys) <= 5000:
        e["levels"] = [{"bits": int(b), "words": ["%016x" % int(w) for w in ws], "ranks": [int(r) for r in rs]} for b, ws, rs in lv]
    out["mphf"].append(e)
json.dump(out, open(sys.argv[1], "w"))
print("wrote", sys.argv[1])
