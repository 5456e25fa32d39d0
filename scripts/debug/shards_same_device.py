import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from claude_semantic_search_b200 import _native
from oracle import search_oracle as so
rng = np.random.default_rng(22)
d, n = 768, 38097
x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
q = so.normalize_rows(rng.standard_normal((40, d), dtype=np.float32))
for mode in sys.argv[1:] or ["default"]:
    for kv in mode.split(","):
        if "=" in kv:
            k_, v_ = kv.split("=")
            _native.set_option(k_, int(v_))
    idx = _native.Index(d, devices=[0, 0])
    idx.add(x)
    Dr, Ir = so.flat_search_c(x, q, 10)
    for nq in (1, 1, 1, 9, 40):
        t0 = time.perf_counter()
        try:
            D, I = idx.search(q[:nq], 10)
            ok, why = so.compare_topk(Dr[:nq], Ir[:nq], D, I)
        except Exception as e:
            ok, why = False, repr(e)
        print(mode, "nq", nq, "ok", ok, "%.3f s" % (time.perf_counter() - t0), why[:100] if not ok else "", flush=True)
    idx.close()
