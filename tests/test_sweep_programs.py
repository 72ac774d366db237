"""The sweep programs compiled into the score / level kernels (``SProg<...>`` in csrc/hp_score_spec.cuh: step order, the
rings every step adds, the rings that feed ``Reads``) against the oracle's restatement of callers.py:15-23 and :132-198.
The program is evaluated on the host: a few-line ``main`` that includes the kernel header is compiled with nvcc (no GPU is
needed to run it) and prints the compile-time tables."""
import os
import shutil
import subprocess

import pytest

from oracle import hiccups_oracle as ho

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the programs instantiated in csrc/hp_api.cu (g_specs)
PROGRAMS = {"p2w5": ([2], [5]), "p1w3": ([1], [3]), "p4w7": ([4], [7]), "p124w357": ([1, 2, 4], [3, 5, 7])}
MAXWW = 10

SRC = r'''
#include <cstdio>
#include "hp_score_spec.cuh"
using namespace hp;
template <class PG> void dump(const char* name) {
    printf("%s %d\n", name, PG::nsteps());
    for (int s = 0; s < PG::nsteps(); ++s)
        printf("%d %d %d %u %u %d\n", s, PG::step_p(s), PG::step_w(s), PG::mask(s), PG::rmask(s), PG::prev_same_pair(s));
}
int main() {
    dump<SProg<10, 1, 2, 5>>("p2w5");
    dump<SProg<10, 1, 1, 3>>("p1w3");
    dump<SProg<10, 1, 4, 7>>("p4w7");
    dump<SProg<10, 3, 1, 3, 2, 5, 4, 7>>("p124w357");
    return 0;
}
'''


@pytest.fixture(scope="module")
def compiled(tmp_path_factory):
    nvcc = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc) and not shutil.which(nvcc):
        pytest.skip("nvcc not available")
    d = tmp_path_factory.mktemp("sprog")
    (d / "dump.cu").write_text(SRC)
    subprocess.run([nvcc, "-std=c++17", "-O0", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-I", os.path.join(ROOT, "hicpeaks_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                    str(d / "dump.cu"), "-o", str(d / "dump")], check=True, capture_output=True, timeout=600)
    out = subprocess.run([str(d / "dump")], check=True, capture_output=True, text=True).stdout.split("\n")
    progs, i = {}, 0
    while i < len(out) and out[i].strip():
        name, n = out[i].split()
        rows = [tuple(int(x) for x in out[i + 1 + k].split()) for k in range(int(n))]
        progs[name] = rows
        i += 1 + int(n)
    return progs


@pytest.mark.parametrize("name", sorted(PROGRAMS))
def test_compiled_program_equals_the_reference_sweep(compiled, name):
    pw, ww = PROGRAMS[name]
    steps = ho.step_program(pw, ww, MAXWW)
    rows = compiled[name]
    assert len(rows) == len(steps) == sum(MAXWW - w + 1 for w in ww)
    last_of_pair = {}
    for s, ((p, w, ops), row) in enumerate(zip(steps, rows)):
        assert row[0] == s and (row[1], row[2]) == (p, w)                    # callers.py:15-23 order
        k_rings = {max(abs(a), abs(b)) for a, b, _, _ in ops}
        r_rings = {max(abs(a), abs(b)) for a, b, _, is_r in ops if is_r}
        assert row[3] == sum(1 << g for g in k_rings), (name, s)             # rings whose off-cross cells join K
        assert row[4] == sum(1 << g for g in r_rings), (name, s)             # ... and Reads (callers.py:197-198)
        # a ring is added whole: every off-cross cell of it, rows outer, columns inner
        exp = [(a, b) for a in range(-w, w + 1) for b in range(-w, w + 1)
               if a and b and max(abs(a), abs(b)) in k_rings]
        assert [(a, b) for a, b, _, _ in ops] == exp
        assert all(is_y == (a > 0 and b < 0) for a, b, is_y, _ in ops)
        assert row[5] == last_of_pair.get(p, -1)
        last_of_pair[p] = s


def _random_programs(n, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    out = [([2], [5], 10), ([1, 2, 4], [3, 5, 7], 10), ([2], [5], 20), ([0], [1], 3), ([3, 1], [4, 6], 9), ([5], [5], 5)]
    while len(out) < n:
        npw = int(rng.integers(1, 5))
        maxww = int(rng.integers(1, 21))
        pw = [int(x) for x in rng.choice(8, npw, replace=False)]
        ww = [int(rng.integers(1, maxww + 1)) for _ in pw]
        if sum(maxww - w + 1 for w in ww) <= 160:
            out.append((pw, ww, maxww))
    return out


def test_runtime_program_builder_equals_the_reference_sweep():
    """hp_program_dump (host only): the op list the table-driven kernels walk -- any (pw, ww, maxww), including peak widths
    at or above the window, unsorted pairs and the reference's re-added rings in union mode -- against the oracle."""
    from hicpeaks_b200 import _capi
    for pw, ww, maxww in _random_programs(60, 1):
        try:
            got = _capi.program_dump(pw, ww, maxww)
        except _capi.EngineError as e:                      # longer than this build's op table: a declared limit, not a mismatch
            assert e.code == _capi.HP_ERR_INVALID and "too long" in str(e), (pw, ww, maxww, str(e))
            continue
        exp = ho.step_program(pw, ww, maxww)
        assert [(p, w) for p, w, _ in got] == [(p, w) for p, w, _ in exp], (pw, ww, maxww)
        for (p, w, ops), (_, _, eops) in zip(got, exp):
            assert ops == [(a, b, bool(y), bool(r)) for a, b, y, r in eops], (pw, ww, maxww, p, w)
