"""TEST INFRASTRUCTURE ONLY -- tests/golden/combine_*.json from the UNMODIFIED reference
``hicpeaks/utilities.py`` (``combine_annotations``, ``_parse_peakfile``), run in the build container only:

    python oracle/make_golden_combine.py

``utilities.py`` imports h5py and cooler at module level (neither is installed here and neither is touched by the two
functions): empty stand-in modules are registered under those names for the import; the reference source is not modified.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/hicpeaks/utilities.py"


def load_reference_utilities():
    for name, attrs in {"h5py": [], "cooler": ["ice", "create_cooler"], "cooler.util": ["binnify", "parse_cooler_uri"],
                        "cooler.reduce": ["CoolerMerger"], "cooler.api": ["Cooler"]}.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, None)
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("ref_utilities", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synth_calls(rng, resolutions, chroms, n_loops, jitter_bins, drop):
    """Loops of a hidden truth seen at every resolution with bin-level jitter, some dropped, plus private false calls."""
    byres = {r: {} for r in resolutions}
    for c in chroms:
        x = rng.integers(0, 2000, n_loops) * 5000
        span = rng.integers(4, 400, n_loops) * 5000
        for r in resolutions:
            calls = []
            for k in range(n_loops):
                if rng.random() < drop:
                    continue
                jx, jy = rng.integers(-jitter_bins, jitter_bins + 1, 2)
                a = int((x[k] // r + jx) * r)
                b = int(((x[k] + span[k]) // r + jy) * r)
                if a < 0 or b <= a:
                    continue
                calls.append((a, a + r, b, b + r))
            for _ in range(n_loops // 5):
                a = int(rng.integers(0, 2000) * 5000 // r * r)
                b = a + int(rng.integers(2, 300)) * r
                calls.append((a, a + r, b, b + r))
            if calls and rng.random() > 0.1:           # now and then a resolution has no call on a chromosome
                byres[r][c] = calls
    return byres


def main():
    ref = load_reference_utilities()
    rng = np.random.default_rng(20181105)
    cases = {
        "three": dict(resolutions=[5000, 10000, 20000], chroms=["1", "2", "X"], n_loops=120, jitter_bins=2, drop=0.3,
                      kw=dict(good_res=20000, mindis=200000, max_res=10000)),          # script defaults
        "defaults": dict(resolutions=[5000, 10000], chroms=["1", "7"], n_loops=150, jitter_bins=1, drop=0.25,
                         kw=dict(good_res=10000, mindis=100000, max_res=10000)),       # function defaults
        "coarse_out": dict(resolutions=[5000, 10000, 25000, 40000], chroms=["3"], n_loops=200, jitter_bins=3, drop=0.4,
                           kw=dict(good_res=10000, mindis=150000, max_res=25000)),
        "single": dict(resolutions=[10000], chroms=["2", "1"], n_loops=40, jitter_bins=0, drop=0.0,
                       kw=dict(good_res=10000, mindis=100000, max_res=10000)),
    }
    for name, cfg in cases.items():
        byres = synth_calls(rng, cfg["resolutions"], cfg["chroms"], cfg["n_loops"], cfg["jitter_bins"], cfg["drop"])
        if name == "three":                            # the same call present at two resolutions, and an exact-radius pair
            byres[10000].setdefault("1", []).append(byres[5000]["1"][0])
            a = byres[5000]["1"][1]
            byres[20000].setdefault("1", []).append((a[0] + 12000, a[0] + 32000, a[2] + 16000, a[2] + 36000))   # distance 20 000
        out = ref.combine_annotations({r: {c: list(v) for c, v in d.items()} for r, d in byres.items()}, **cfg["kw"])
        doc = dict(kw=cfg["kw"], byres={str(r): d for r, d in byres.items()}, out=[list(t) for t in out])
        with open(os.path.join(GOLD, "combine_%s.json" % name), "w") as f:
            json.dump(doc, f)
        print(name, {r: sum(len(v) for v in d.values()) for r, d in byres.items()}, "->", len(out))
    # the peak-file reader: header rows, 'chr' prefixes
    path = os.path.join(GOLD, "combine_peakfile.txt")
    with open(path, "w") as f:
        f.write("#chrom1\tx1\tx2\tchrom2\ty1\ty2\n")
        for c, p in [("chr1", (10000, 20000, 300000, 310000)), ("chr1", (50000, 60000, 900000, 910000)),
                     ("chrX", (0, 10000, 250000, 260000)), ("chr10", (70000, 80000, 170000, 180000))]:
            f.write("\t".join([c, str(p[0]), str(p[1]), c, str(p[2]), str(p[3]), "extra"]) + "\n")
    parsed = ref._parse_peakfile(path, 1)
    with open(os.path.join(GOLD, "combine_peakfile.json"), "w") as f:
        json.dump({k: [list(t) for t in v] for k, v in parsed.items()}, f)
    print("peakfile", parsed)


if __name__ == "__main__":
    main()
