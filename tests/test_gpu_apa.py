"""GPU: the APA path (hicpeaks_b200.apa through the C ABI) against the outputs of the unmodified reference
``hicpeaks/apa.py`` (tests/golden/apa_*.npz) and against the oracle on seeded inputs."""
import numpy as np
import pytest

import golden_util as gu
from hicpeaks_b200 import apa
from oracle import apa_oracle as ao

pytestmark = pytest.mark.gpu


class BandMatrix:
    """Minimal stand-in for the scipy CSR the reference passes: ``shape`` and ``diagonal(k)``."""

    def __init__(self, diags, n):
        self.diags, self.shape = diags, (n, n)

    def diagonal(self, k):
        return self.diags[k] if k < len(self.diags) else np.zeros(self.shape[0] - k)


@pytest.mark.parametrize("name", gu.names("apa"))
def test_apa_matches_reference(name):
    z, n, Diags, weights = gu.load_apa(name)
    w, cw = int(z["w"]), int(z["cw"])
    M = BandMatrix(ao.balanced_diags(Diags, weights), n)
    wins = apa.apa_submatrix(M, [tuple(p) for p in z["pos"]], w=w)
    assert len(wins) == int(z["n_windows"])
    assert np.array_equal(wins.mean_arr, z["mean_arr"])                 # bit-exact window means
    assert np.array_equal(wins[0], z["win_first"]) and np.array_equal(wins[len(wins) - 1], z["win_last"])
    avg, score, zz, p, maxi = apa.apa_analysis(wins, w=w, cw=cw)
    assert np.array_equal(avg, z["avg"])                                # bit-exact pileup
    assert np.array_equal(np.array([score, zz, p, maxi]), z["stats"])
    # the plain-array entry (what scripts/apa-analysis passes after np.r_[apa]) gives the same answer
    avg2, score2, z2, p2, maxi2 = apa.apa_analysis(np.asarray(wins), w=w, cw=cw)
    assert np.array_equal(avg2, avg) and score2 == score and z2 == zz
    # several chromosomes: handles in a list == one concatenated array
    avg3 = apa.apa_analysis([wins, wins], w=w, cw=cw)[0]
    both = np.concatenate([np.asarray(wins), np.asarray(wins)])
    assert np.array_equal(avg3, ao.apa_analysis(both, w=w, cw=cw)[0])


def test_apa_rejections_and_edges():
    from hicpeaks_b200.synth import synth_chromosome
    n, w = 400, 4
    inp = synth_chromosome(n, 80, 1, maxww=0, seed=2, scale=5.0, nan_frac=0.05)
    diags = ao.balanced_diags(inp["Diags"], inp["weights"])
    rng = np.random.default_rng(0)
    pos = [(int(i), int(min(i + d, n - 1))) for i, d in zip(rng.integers(0, n, 600), rng.integers(0, 70, 600))]
    pos += [(0, 10), (n - 1, n - 1), (w, w), (n - w - 1, n - w - 1), (50, 45)]      # edges, (j < i) mirror
    exp, valid = ao.apa_submatrix(diags, n, pos, w=w)
    got = apa.apa_submatrix(BandMatrix(diags, n), pos, w=w)
    assert len(got) == len(exp) and len(exp) > 50 and (~valid).sum() > 20
    assert np.array_equal(np.asarray(got), np.asarray(exp))
