"""Front ends: flags / chromosome selection / worker input preparation on CPU, the whole pyHICCUPS / pyBHFDR run
(fake cooler in, BEDPE-like text out) on the GPU against the reference's golden peak tables."""
import os

import numpy as np
import pytest

import golden_util as gu
from fake_cooler import FakeCooler
from hicpeaks_b200 import cli


def test_flags_and_defaults_match_reference():
    a = cli.hiccups_parser().parse_args(['-O', 'x', '-p', 'y', '--pw', '1', '2', '--ww', '3', '5'])
    assert (a.maxww, a.siglevel, a.sumq, a.double_fold, a.single_fold) == (10, 0.05, 0.01, 1.75, 2)
    assert (a.clr_weight_name, a.use_raw, a.min_marginal_peaks, a.min_local_reads, a.only_anchors) == ('weight', False, 2, 16, False)
    assert (a.maxapart, a.nproc, a.chroms, a.logFile) == (10000000, 1, ['#', 'X'], 'pyHICCUPS.log')
    b = cli.bhfdr_parser().parse_args(['-O', 'x', '-p', 'y'])
    assert (b.pw, b.ww, b.maxww, b.siglevel, b.maxapart, b.logFile) == (2, 5, 10, 0.05, 2000000, 'pyBHFDR.log')


def test_chromosome_selection():
    names = ['chr1', 'chr2', 'chrX', 'chrY', 'chrM', 'chr10_random']
    assert cli.select_chromosomes(names, ['#', 'X']) == ['chr1', 'chr2', 'chrX']
    assert cli.select_chromosomes(names, []) == names
    assert cli.select_chromosomes(names, ['Y']) == ['chrY']


def test_prepare_chromosome_matches_worker_containers():
    z, inp, kw, res = gu.load("synth_p2w5")
    Lib = FakeCooler(res, {"chr7": (inp["Diags"], inp["weights"])})
    b = cli.prepare_chromosome(Lib, "chr7", "weight", kw["maxapart"], kw["maxww"], min(kw["ww"]), res)
    assert b["n"] == inp["n"] and b["num"] == inp["num"]
    for d in range(inp["num"]):
        assert np.array_equal(b["Diags"][d], inp["Diags"][d])
    for i in range(inp["num"] - inp["min_ww"]):
        assert np.array_equal(b["cDiags"][i], inp["cDiags"][i])
    assert all(b["IR"][k] == inp["IR"][k] or (np.isnan(b["IR"][k]) and np.isnan(inp["IR"][k])) for k in inp["IR"])
    assert np.array_equal(b["biases"], inp["biases"])


def _expected_lines(table_rows, fmt, chrom, res, ncols):
    lines = set()
    for row in table_rows:
        x, y = int(row[0]), int(row[1])
        vals = tuple(row[5:5 + ncols])
        c = 'chr' + chrom
        lines.add(fmt.format(c, x, x + res, c, y, y + res, '.', vals[0], '.', '.', *vals[1:]))
    return lines


@pytest.mark.gpu
def test_pyhiccups_end_to_end(tmp_path):
    z, inp, kw, res = gu.load("chr21_25k_p1w3")
    Lib = FakeCooler(res, {"chr21": (inp["Diags"], inp["weights"]), "chrM": (inp["Diags"], inp["weights"])})
    out = tmp_path / "loops.txt"
    cli.run_hiccups(['-O', str(out), '-p', 'fake.cool', '--logFile', str(tmp_path / 'log.txt'), '--pw', '1', '--ww', '3',
                     '--maxapart', str(kw["maxapart"]), '-C', '21'], Lib=Lib)
    got = set(open(out).read().splitlines(True))
    exp = _expected_lines(z["table"], cli.HICCUPS_LINE, "21", res, 7)
    # O is exact; p / q agree with the reference within 1e-6, so '.3g' strings can differ only in the last digit
    assert len(got) == len(exp) == z["table"].shape[0]
    key = lambda ln: tuple(ln.split('\t')[:8])
    assert {key(l) for l in got} == {key(l) for l in exp}
    same = len(got & exp)
    assert same >= 0.98 * len(exp), (same, len(exp))
    assert "Done!" in open(tmp_path / 'log.txt').read()


@pytest.mark.gpu
def test_pybhfdr_end_to_end(tmp_path):
    z, inp, kw, res = gu.load("bh_synth_p2w5")
    Lib = FakeCooler(res, {"chr3": (inp["Diags"], inp["weights"])})
    out = tmp_path / "bh.txt"
    cli.run_bhfdr(['-O', str(out), '-p', 'fake.cool', '--logFile', str(tmp_path / 'log.txt'), '--pw', '2', '--ww', '5',
                   '--maxapart', str(kw["maxapart"]), '-C', '3'], Lib=Lib)
    got = set(open(out).read().splitlines(True))
    exp = _expected_lines(z["table"], cli.BHFDR_LINE, "3", res, 4)
    key = lambda ln: tuple(ln.split('\t')[:8])
    assert len(got) == len(exp) and {key(l) for l in got} == {key(l) for l in exp}


def test_peakfile_parser(tmp_path):
    f = tmp_path / "loops.bedpe"
    f.write_text("#header\nchr1\t100\t200\tchr1\t900\t1000\t.\t5\nchr2\t10\t20\tchr2\t70\t80\t.\t3\nchr1\t300\t400\tchr1\t500\t600\n")
    D = cli.parse_peakfile(str(f), skip=1)
    assert D == {"1": [(100, 200, 900, 1000), (300, 400, 500, 600)], "2": [(10, 20, 70, 80)]}


@pytest.mark.gpu
def test_apa_analysis_end_to_end(tmp_path):
    """Loop file + (fake) cooler in, averaged window out: the front end against the oracle on the same anchors."""
    from oracle import apa_oracle as ao
    z, n, Diags, weights = gu.load_apa("apa_w5")
    res, w = 10000, 5
    Lib = FakeCooler(res, {"chr1": (Diags, weights)})
    rng = np.random.default_rng(3)
    loops = tmp_path / "loops.bedpe"
    with open(loops, "w") as f:
        for _ in range(400):
            i = int(rng.integers(10, n - 130)); j = i + int(rng.integers(12, 100))
            f.write("chr1\t%d\t%d\tchr1\t%d\t%d\t.\t1\n" % (i * res, (i + 2) * res, j * res, (j + 1) * res))
    out = tmp_path / "apa.txt"
    avg, score, zz, p, maxi, nwin = cli.run_apa(['-O', str(out), '-p', 'fake.cool', '-I', str(loops), '-W', str(w), '-M', '10'], Lib=Lib)
    M = Lib.matrix(balance="weight", sparse=True).fetch("chr1").tocsr()
    pos = cli.locate_anchors(M, cli.parse_peakfile(str(loops), 0)["1"], res, 10)
    exp, valid = ao.apa_submatrix(ao.balanced_diags(Diags, weights), n, pos, w=w)
    e_avg, e_score, e_z, e_p, e_maxi, _, _ = ao.apa_analysis(exp, w=w, cw=3)
    assert nwin == len(exp) > 100
    assert np.array_equal(avg, e_avg) and score == e_score and zz == e_z and maxi == e_maxi
    assert os.path.exists(out)


def test_worker_threads_spread_over_gpus():
    """--nproc beyond the GPU count = several chromosomes in flight per GPU (thread t on GPU t % ngpu); every key is
    computed exactly once and an exception in a worker propagates like Pool.map's (scripts/pyHICCUPS:192-198)."""
    import threading
    from types import SimpleNamespace
    from hicpeaks_b200 import cli
    assert cli._n_workers(SimpleNamespace(nproc=1), 2) == 2
    assert cli._n_workers(SimpleNamespace(nproc=6), 2) == 6
    assert cli._n_workers(SimpleNamespace(nproc=64), 2) == 16
    seen, lock = [], threading.Lock()

    def fn(key, gpu):
        with lock:
            seen.append((key, gpu, threading.get_ident()))
        return key * 2

    keys = list(range(40))
    out = cli._map_over_gpus(keys, {k: k for k in keys}, 2, fn, nworkers=6)
    assert out == {k: 2 * k for k in keys}
    assert sorted(k for k, _, _ in seen) == keys and {g for _, g, _ in seen} <= {0, 1}
    by_thread = {}
    for _, g, t in seen:
        by_thread.setdefault(t, set()).add(g)
    assert all(len(g) == 1 for g in by_thread.values())          # a thread stays on its GPU

    def boom(key, gpu):
        if key == 7:
            raise KeyError("chromosome 7")
        return key

    with pytest.raises(KeyError):
        cli._map_over_gpus(keys, {k: 1 for k in keys}, 2, boom, nworkers=4)
