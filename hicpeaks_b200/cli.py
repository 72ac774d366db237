"""Command-line front ends with the reference's surface: ``pyHICCUPS`` (/root/reference/scripts/pyHICCUPS) and
``pyBHFDR`` (/root/reference/scripts/pyBHFDR) -- same flags and defaults (:19-73 / :19-50), same log format
(:89-105), same output lines (:200-207 / :169-176) -- on top of the CUDA engine.

What replaces what:

* ``prepare_chromosome``  <- the worker's input preparation, pyHICCUPS:142-166 (cooler fetch, band diagonals,
                             per-distance expected ``IR``, biases).  Host side, numpy, same arithmetic.
* ``run``                 <- ``Pool(nproc).map(worker, Params)`` (pyHICCUPS:184-198): one host thread per GPU
                             (``--gpus``), chromosomes handed out longest first; ``--nproc`` = host worker threads
                             (several chromosomes in flight per GPU when it exceeds the GPU count).
* extra flags: ``--gpus N`` and, for pyHICCUPS, ``--fdr-scope {chrom,genome}`` (``chrom`` = the reference).

``cooler`` is imported lazily; anything with ``binsize``, ``chromnames``, ``matrix(balance=, sparse=True).fetch(c)``
and ``bins().fetch(c)[name].values`` works (tests use an in-memory stand-in, the image has no cooler / h5py).
"""
from __future__ import annotations

import argparse
import logging
import logging.handlers
import sys
import threading

import numpy as np

__version__ = "0.1.0"

log = logging.getLogger("hicpeaks_b200.cli")


# ---------------------------------------------------------------------------------------------------------
def hiccups_parser(prog="pyHICCUPS"):
    p = argparse.ArgumentParser(prog=prog, usage='%(prog)s <-O output> [options]',
                                description='A B200 (CUDA) implementation of the HiCCUPS algorithm.',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('-v', '--version', action='version', version=' '.join(['%(prog)s', __version__]),
                   help='Print version number and exit.')
    p.add_argument('-O', '--output', help='Output file name.')
    p.add_argument('--logFile', default='pyHICCUPS.log', help='Logging file name.')
    g1 = p.add_argument_group(title='Relate to Hi-C data:')
    g1.add_argument('-p', '--path', help='Cooler URI.')
    g1.add_argument('-C', '--chroms', nargs='*', default=['#', 'X'],
                    help='List of chromosome labels. "#" stands for chromosomes with numerical labels. '
                         '"--chroms" with zero argument will include all chromosome data.')
    g2 = p.add_argument_group(title='Algorithm Parameters:')
    g2.add_argument('--pw', type=int, nargs='+', help='List of the peak widths.')
    g2.add_argument('--ww', type=int, nargs='+', help='List of the donut widths.')
    g2.add_argument('--maxww', type=int, default=10, help='Maximum donut width.')
    g2.add_argument('--siglevel', type=float, default=0.05, help='Significant Level.')
    g2.add_argument('--sumq', type=float, default=0.01,
                    help='Isolated peak pixels are dropped when the sum of their 2 q-values exceeds this threshold.')
    g2.add_argument('--double-fold', type=float, default=1.75, help='Minimum fold enrichment for both backgrounds.')
    g2.add_argument('--single-fold', type=float, default=2, help='Minimum fold enrichment for either background.')
    g2.add_argument('--clr-weight-name', default='weight', help='Name of the weight column in the Cooler URI.')
    g2.add_argument('--use-raw', action='store_true', help='Sort peak pixels by raw signal during local clustering.')
    g2.add_argument('--min-marginal-peaks', type=int, default=2,
                    help='Minimum marginal number of peaks when detecting peak anchors.')
    g2.add_argument('--min-local-reads', type=int, default=16,
                    help='Minimum sum of contacts in the vicinity of a valid loop.')
    g2.add_argument('--only-anchors', action='store_true', help='Either of the peak loci must be an anchor.')
    g2.add_argument('--maxapart', type=int, default=10000000, help='Maximum genomic distance between two loci.')
    g2.add_argument('--nproc', type=int, default=1, help='Host worker threads (chromosomes in flight); spread over the GPUs in use, at least one per GPU.')
    g3 = p.add_argument_group(title='GPU engine:')
    g3.add_argument('--gpus', type=int, default=0, help='Number of GPUs (0 = all visible).')
    g3.add_argument('--fdr-scope', choices=['chrom', 'genome'], default='chrom',
                    help='"chrom": BH within lambda-chunks per chromosome (the reference). "genome": histograms merged '
                         'over all chromosomes and GPUs before BH.')
    return p


def bhfdr_parser(prog="pyBHFDR"):
    p = argparse.ArgumentParser(prog=prog, usage='%(prog)s <-O output> [options]',
                                description='A B200 (CUDA) implementation of the BH-FDR algorithm.',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('-v', '--version', action='version', version=' '.join(['%(prog)s', __version__]),
                   help='Print version number and exit.')
    p.add_argument('-O', '--output', help='Output file name.')
    p.add_argument('--logFile', default='pyBHFDR.log', help='Logging file name.')
    g1 = p.add_argument_group(title='Relate to Hi-C data:')
    g1.add_argument('-p', '--path', help='Cooler URI.')
    g1.add_argument('-C', '--chroms', nargs='*', default=['#', 'X'], help='List of chromosome labels.')
    g2 = p.add_argument_group(title='Algorithm Parameters:')
    g2.add_argument('--pw', type=int, default=2, help='Width of the interaction region surrounding the peak.')
    g2.add_argument('--ww', type=int, default=5, help='Width of the donut sampled.')
    g2.add_argument('--maxww', type=int, default=10, help='Maximum donut width.')
    g2.add_argument('--siglevel', type=float, default=0.05, help='Significant Level.')
    g2.add_argument('--maxapart', type=int, default=2000000, help='Maximum genomic distance between two loci.')
    g2.add_argument('--clr-weight-name', default='weight', help='Name of the weight column in the Cooler URI.')
    g2.add_argument('--nproc', type=int, default=1, help='Host worker threads (chromosomes in flight); spread over the GPUs in use, at least one per GPU.')
    g3 = p.add_argument_group(title='GPU engine:')
    g3.add_argument('--gpus', type=int, default=0, help='Number of GPUs (0 = all visible).')
    return p


def setup_logging(logfile, rotating=False):
    """Root logger as the reference sets it up (pyHICCUPS:89-105; pyBHFDR uses a rotating file, :70-72)."""
    root = logging.getLogger()
    root.setLevel(10)
    console = logging.StreamHandler()
    if rotating:
        filehandler = logging.handlers.RotatingFileHandler(logfile, maxBytes=200000, backupCount=5)
    else:
        filehandler = logging.FileHandler(logfile)
    console.setLevel('INFO')
    filehandler.setLevel('INFO')
    formatter = logging.Formatter(fmt='%(name)-21s %(levelname)-7s @ %(asctime)s: %(message)s', datefmt='%m/%d/%y %H:%M:%S')
    console.setFormatter(formatter)
    filehandler.setFormatter(formatter)
    root.addHandler(console)
    root.addHandler(filehandler)
    return console, filehandler


# ---------------------------------------------------------------------------------------------------------
def select_chromosomes(chromnames, chroms):
    """pyHICCUPS:184-187."""
    out = []
    for key in chromnames:
        label = key.lstrip('chr')
        if (not chroms) or (label.isdigit() and '#' in chroms) or (label in chroms):
            out.append(key)
    return out


def prepare_chromosome(Lib, key, weight_name, maxapart, maxww, min_ww, res):
    """The containers the reference's worker hands to the caller (pyHICCUPS:142-166), without the two scipy
    band matrices ``M`` / ``cM`` (the engine reads ``Diags`` / ``cDiags``)."""
    H = Lib.matrix(balance=False, sparse=True).fetch(key)
    cHeatMap = Lib.matrix(balance=weight_name, sparse=True).fetch(key)
    chromLen = H.shape[0]
    num = maxapart // res + maxww + 1
    Diags = [np.asarray(H.diagonal(i)) for i in np.arange(num)]
    IR = {}
    cDiags = []
    for i in np.arange(min_ww, num):
        diag = np.array(cHeatMap.diagonal(i), dtype=np.float64)
        mask = np.isnan(diag)
        notnan = diag[np.logical_not(mask)]
        with np.errstate(invalid="ignore", divide="ignore"):
            IR[int(i)] = notnan.mean()
        diag[mask] = 0
        cDiags.append(diag)
    tmp = np.asarray(Lib.bins().fetch(key)[weight_name].values, dtype=np.float64)
    mask = np.logical_not((tmp == 0) | np.isnan(tmp))
    biases = np.zeros_like(tmp)
    biases[mask] = 1 / tmp[mask]
    return dict(n=int(chromLen), num=int(num), min_ww=int(min_ww), Diags=Diags, cDiags=cDiags, IR=IR, biases=biases)


# weight columns cooler treats as divisive (cooler/api.py: `divisive_weights` defaults to True for these names): the
# balanced value is count / (w_i w_j), while the reference's biases stay 1 / column (scripts/pyHICCUPS:163-166)
DIVISIVE_WEIGHT_NAMES = ("KR", "VC", "VC_SQRT")


def worker_input(Lib, key, weight_name, maxapart, maxww, min_ww, res):
    """Input of one chromosome for the engine.  The fast boundary -- raw int32 count diagonals + the weight column, the
    balanced band / IR / biases derived on the GPU -- applies when that derivation is what cooler + the reference's worker
    do: multiplicative weights (balanced = bias1[row] * bias2[col] * data) and integer counts.  Otherwise (a divisive
    weight column, float or out-of-range counts) the containers are built on the host exactly as the worker builds
    them, through ``Lib.matrix(balance=name)`` (``prepare_chromosome``), and the operator-level entry is used."""
    if weight_name not in DIVISIVE_WEIGHT_NAMES:
        b = prepare_counts(Lib, key, weight_name, maxapart, maxww, res)
        ok = True
        for d in b["Diags"]:
            a = np.asarray(d)
            if not (np.issubdtype(a.dtype, np.integer) and a.dtype.itemsize <= 4 and a.dtype != np.uint32):
                if not (np.issubdtype(a.dtype, np.integer) and a.size and np.abs(a).max() < 2 ** 31) and a.size:
                    ok = False
                    break
        if ok:
            return b
    return prepare_chromosome(Lib, key, weight_name, maxapart, maxww, min_ww, res)


def chrom_sizes(Lib, keys, res, num):
    """{key: (bins, stored diagonals)} for the longest-first order / the LPT partition."""
    sizes = {}
    cs = getattr(Lib, "chromsizes", None)
    for k in keys:
        try:
            n = int(-(-int(cs[k]) // res))
        except Exception:
            n = 1
        sizes[k] = (n, num)
    return sizes


def prepare_counts(Lib, key, weight_name, maxapart, maxww, res):
    """Worker-level input for the engine: the raw count diagonals and the bin weights only; the balanced band, ``IR``
    and the biases of pyHICCUPS:149-166 are derived on the GPU (``hicpeaks_b200.callers.hiccups_from_counts``)."""
    H = Lib.matrix(balance=False, sparse=True).fetch(key)
    num = maxapart // res + maxww + 1
    Diags = [np.asarray(H.diagonal(i)) for i in np.arange(num)]
    weights = np.asarray(Lib.bins().fetch(key)[weight_name].values, dtype=np.float64)
    return dict(n=int(H.shape[0]), num=int(num), Diags=Diags, weights=weights)


def _open_cooler(path):
    try:
        import cooler
    except ImportError as e:                                    # pragma: no cover (no cooler in this image)
        raise SystemExit("the 'cooler' package is needed to read %s (%s)" % (path, e))
    return cooler.Cooler(path)


def _n_gpus(args):
    from . import _capi
    have = _capi.device_count()
    if have == 0:
        raise _capi.EngineError(_capi.HP_ERR_NO_DEVICE, "no CUDA device (this engine has no CPU fallback)")
    want = args.gpus or have                 # --gpus 0: all visible; --nproc only sets the number of host threads
    return max(1, min(want, have))


def _n_workers(args, ngpu):
    """Host threads pulling chromosomes: at least one per GPU; ``--nproc`` beyond the GPU count puts several chromosomes
    in flight per GPU (thread t drives GPU t % ngpu with its own context and stream), so that the host side of one
    chromosome -- cooler read, diagonal extraction, upload packing, clustering -- overlaps the kernels of another, the
    way ``Pool(nproc)`` overlaps whole chromosomes in the reference (scripts/pyHICCUPS:192-198)."""
    return max(ngpu, min(int(getattr(args, "nproc", 1) or 1), 8 * ngpu))


def _map_over_gpus(keys, sizes, ngpu, fn, nworkers=None):
    """Longest chromosome first, host threads pulling from a shared list (thread t on GPU t % ngpu); returns
    {key: fn(key, gpu)}."""
    order = sorted(keys, key=lambda k: -sizes[k])
    lock = threading.Lock()
    out, errors = {}, []

    def loop(gpu):
        while True:
            with lock:
                if not order or errors:
                    return
                key = order.pop(0)
            try:
                out[key] = fn(key, gpu)
            except BaseException as e:                          # propagate like Pool.map does
                errors.append(e)
                return

    threads = [threading.Thread(target=loop, args=(t % ngpu,)) for t in range(max(ngpu, nworkers or ngpu))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out


HICCUPS_LINE = '{0}\t{1}\t{2}\t{3}\t{4}\t{5}\t{6}\t{7:.3g}\t{8}\t{9}\t{10:.3g}\t{11:.3g}\t{12:.3g}\t{13:.3g}\t{14:.3g}\t{15:.3g}\n'
BHFDR_LINE = '{0}\t{1}\t{2}\t{3}\t{4}\t{5}\t{6}\t{7:.3g}\t{8}\t{9}\t{10:.3g}\t{11:.3g}\t{12:.3g}\n'


def write_table(OF, fmt, key, pixel_table, res):
    """pyHICCUPS:200-207 / pyBHFDR:169-176."""
    for pixel in pixel_table:
        tmp = pixel_table[pixel]
        c = 'chr' + key.lstrip('chr')
        content = (c, pixel[0], pixel[0] + res, c, pixel[1], pixel[1] + res, '.', tmp[3], '.', '.') + tuple(tmp[4:])
        OF.write(fmt.format(*content))


def run_hiccups(argv=None, Lib=None):
    args = hiccups_parser().parse_args(argv if argv else ['-h'])
    handlers = setup_logging(args.logFile)
    try:
        from . import callers, dispatch
        for k in ('output', 'logFile', 'chroms', 'path', 'pw', 'ww', 'maxww', 'maxapart', 'siglevel', 'sumq', 'double_fold',
                  'single_fold', 'clr_weight_name', 'min_local_reads', 'gpus', 'fdr_scope'):
            log.info('# %s = %s', k, getattr(args, k))
        log.info('Loading Hi-C data ...')
        Lib = Lib if Lib is not None else _open_cooler(args.path)
        res = Lib.binsize
        keys = select_chromosomes(Lib.chromnames, args.chroms)
        log.info('Calling Peaks ...')
        kw = dict(pw=list(args.pw), ww=list(args.ww), sig=args.siglevel, sumq=args.sumq, maxww=args.maxww,
                  maxapart=args.maxapart, double_fold=args.double_fold, single_fold=args.single_fold, res=res,
                  use_raw=args.use_raw, min_marginal_peaks=args.min_marginal_peaks, onlyanchor=args.only_anchors,
                  min_local_reads=args.min_local_reads)

        ngpu = _n_gpus(args)
        num = args.maxapart // res + args.maxww + 1
        sizes = chrom_sizes(Lib, keys, res, num)

        def load_any(key):
            return worker_input(Lib, key, args.clr_weight_name, args.maxapart, args.maxww, min(args.ww), res)

        if args.fdr_scope == 'genome':
            # one host thread (= rank) per GPU; chromosomes LPT-partitioned over them; the lambda-chunk histograms are
            # merged on the devices with one NCCL all-reduce (hp_allreduce_hist) before BH
            comms = dispatch.ThreadComm.group(ngpu) if ngpu > 1 else [dispatch.LocalComm()]
            results, errors = [None] * ngpu, []

            def rank_main(g):
                eng = dispatch.CudaEngine(g)
                try:
                    runner = dispatch.GenomeRunner(comm=comms[g], engine=eng, fdr_scope='genome')
                    results[g] = runner.run({k: (lambda k=k: load_any(k)) for k in keys}, sizes, **kw)
                    if g == 0 and runner.merge_ms is not None:
                        log.info('Genome-wide lambda-chunk histograms merged over %d GPU(s) in %.3f ms (device)', ngpu, runner.merge_ms)
                except BaseException as e:
                    errors.append(e)
                    if ngpu > 1:
                        comms[g].shared.barrier.abort()
                finally:
                    eng.close()

            threads = [threading.Thread(target=rank_main, args=(g,)) for g in range(ngpu)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            if errors:
                raise errors[0]
            tables = results[0]
        else:
            def one(key, gpu):
                b = load_any(key)
                if "cDiags" in b:
                    return callers.hiccups(None, None, b["biases"], b["biases"], b["IR"], b["n"], b["Diags"], b["cDiags"], b["num"],
                                           key.lstrip('chr'), device=gpu, **kw)
                return callers.hiccups_from_counts(b["weights"], b["n"], b["Diags"], b["num"], key.lstrip('chr'), device=gpu, **kw)
            tables = _map_over_gpus(keys, {k: sizes[k][0] * sizes[k][1] for k in keys}, ngpu, one, _n_workers(args, ngpu))
        with open(args.output, 'w') as OF:
            for key in keys:
                write_table(OF, HICCUPS_LINE, key.lstrip('chr'), tables[key], res)
        log.info('Done!')
    finally:
        for h in handlers:
            logging.getLogger().removeHandler(h)
            h.close()


def run_bhfdr(argv=None, Lib=None):
    args = bhfdr_parser().parse_args(argv if argv else ['-h'])
    handlers = setup_logging(args.logFile, rotating=True)
    try:
        from . import callers
        log.info('Loading Hi-C data ...')
        Lib = Lib if Lib is not None else _open_cooler(args.path)
        res = Lib.binsize
        keys = select_chromosomes(Lib.chromnames, args.chroms)
        log.info('Calling Peaks ...')

        def one(key, gpu):
            b = worker_input(Lib, key, args.clr_weight_name, args.maxapart, args.maxww, args.ww, res)
            kwb = dict(pw=args.pw, ww=args.ww, sig=args.siglevel, maxww=args.maxww, maxapart=args.maxapart, res=res, device=gpu)
            if "cDiags" in b:
                return callers.bhfdr(None, None, b["biases"], b["biases"], b["IR"], b["n"], b["Diags"], b["cDiags"], b["num"],
                                     key.lstrip('chr'), **kwb)
            return callers.bhfdr_from_counts(b["weights"], b["n"], b["Diags"], b["num"], key.lstrip('chr'), **kwb)

        ngpu = _n_gpus(args)
        sizes = chrom_sizes(Lib, keys, res, args.maxapart // res + args.maxww + 1)
        tables = _map_over_gpus(keys, {k: sizes[k][0] * sizes[k][1] for k in keys}, ngpu, one, _n_workers(args, ngpu))
        with open(args.output, 'w') as OF:
            for key in keys:
                write_table(OF, BHFDR_LINE, key.lstrip('chr'), tables[key], res)
        log.info('Done!')
    finally:
        for h in handlers:
            logging.getLogger().removeHandler(h)
            h.close()


# ---------------------------------------------------------------------------------------------------------
# apa-analysis (/root/reference/scripts/apa-analysis)
def apa_parser(prog="apa-analysis"):
    p = argparse.ArgumentParser(prog=prog, description='Perform Aggregate Peak Analysis (APA).',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('-v', '--version', action='version', version=' '.join(['%(prog)s', __version__]))
    p.add_argument('-O', '--output', help='Output file name.')
    p.add_argument('--dpi', default=200, type=int, help='The resolution in dots per inch of the output figure.')
    p.add_argument('-p', '--path', help='Cooler URI.')
    p.add_argument('-I', '--loop-file', help='Loop file in bedpe format.')
    p.add_argument('-S', '--skip-rows', default=0, type=int, help='Number of leading lines in the loop file to skip.')
    p.add_argument('-M', '--min-dis', default=10, type=int,
                   help='Only peak calls whose loci are separated by at least this number of bins are examined.')
    p.add_argument('-W', '--window', default=5, type=int, help='Width of the window in APA analysis.')
    p.add_argument('-C', '--corner-size', default=3, type=int, help='Lower-/upper-corner size of the resulted APA matrix.')
    p.add_argument('--clr-weight-name', default='weight', help='Weight column; "raw" uses the raw signals.')
    p.add_argument('--colormap-name', default='traditional', help='Name of the colormap in matplotlib.')
    p.add_argument('--vmax', type=float, help='The maximum value that the colorbar covers.')
    return p


def find_chrom_pre(chromlabels):
    """utilities.py:433-441."""
    return 'chr' if chromlabels[0].startswith('chr') else ''


def parse_peakfile(filpath, skip=1):
    """utilities.py:443-467: {chrom label without prefix: [(x1, x2, y1, y2)]} from BEDPE columns 0, 1, 2, 4, 5."""
    D = {}
    with open(filpath, 'r') as source:
        for i, line in enumerate(source):
            if i < skip:
                continue
            parse = line.rstrip().split()
            D.setdefault(parse[0], []).append((int(parse[1]), int(parse[2]), int(parse[4]), int(parse[5])))
    pre = find_chrom_pre(list(D.keys())) if D else ''
    return {chrom.lstrip(pre): v for chrom, v in D.items()}


def locate_anchors(M, loops, res, min_dis):
    """apa-analysis:98-119: the strongest pixel inside the bin ranges of every loop, upper-triangle order."""
    pos = []
    n = M.shape[0]
    for p in loops:
        x, y = p[0], p[2]
        if abs(y - x) < min_dis * res:
            continue
        s_l = range(p[0] // res, int(np.ceil(p[1] / float(res))))
        e_l = range(p[2] // res, int(np.ceil(p[3] / float(res))))
        si, ei = None, None
        for st in s_l:
            for et in e_l:
                if (st < n) and (et < n):
                    if si is None:
                        si, ei = st, et
                    elif M[st, et] > M[si, ei]:
                        si, ei = st, et
        if si is not None:
            pos.append((si, ei) if si < ei else (ei, si))
    return pos


def run_apa(argv=None, Lib=None):
    """Returns (avg, score, z, p, maxi, n_windows); writes the figure when matplotlib is available, else the averaged
    window as text (``numpy.savetxt``) to ``--output``."""
    args = apa_parser().parse_args(argv if argv else ['-h'])
    from . import apa as apa_mod
    correct = False if args.clr_weight_name.lower() == 'raw' else args.clr_weight_name
    Lib = Lib if Lib is not None else _open_cooler(args.path)
    res = Lib.binsize
    pre = find_chrom_pre(Lib.chromnames)
    peaks = parse_peakfile(args.loop_file, args.skip_rows)
    parts = []
    for c in peaks:
        chrom = pre + c
        if chrom not in Lib.chromnames:
            continue
        M = Lib.matrix(balance=correct, sparse=True).fetch(chrom).tocsr()
        pos = locate_anchors(M, peaks[c], res, args.min_dis)
        if pos:
            parts.append(apa_mod.apa_submatrix(M, pos, w=args.window))
    n_windows = sum(len(p) for p in parts)
    print(n_windows)
    avg, score, z, p, maxi = apa_mod.apa_analysis(parts, w=args.window, cw=args.corner_size)
    for part in parts:                     # one engine context per chromosome: give the GPU memory back
        part.close()
    vmax = maxi if args.vmax is None else args.vmax
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
        from matplotlib.colors import LinearSegmentedColormap
        cmap = LinearSegmentedColormap.from_list('interaction', ['#FFFFFF', '#ff9292', '#ff6767', '#F70000'])
        plt.imshow(avg, cmap=cmap if args.colormap_name == 'traditional' else args.colormap_name, vmax=vmax, interpolation='none')
        plt.tick_params(axis='both', bottom=False, top=False, left=False, right=False, labelbottom=False, labeltop=False,
                        labelleft=False, labelright=False)
        plt.colorbar()
        plt.savefig(args.output, dpi=args.dpi, bbox_inches='tight')
        plt.close()
    except ImportError:
        np.savetxt(args.output, avg, header='APA score = {0:.6g}, z = {1:.6g}, p-value = {2:.6g}, vmax = {3:.6g}'.format(score, z, p, vmax))
    return avg, score, z, p, maxi, n_windows


# ---------------------------------------------------------------------------------------------------------
# combine-resolutions (/root/reference/scripts/combine-resolutions)
def combine_parser(prog="combine-resolutions"):
    p = argparse.ArgumentParser(prog=prog, usage='%(prog)s <-O output> [options]',
                                description='Combine loop calls from different resolutions.',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('-v', '--version', action='version', version=' '.join(['%(prog)s', __version__]),
                   help='Print version number and exit.')
    p.add_argument('-O', '--output', help='Output peak file name.')
    p.add_argument('-p', '--paths', nargs='+', help='List of peak file paths at different resolutions.')
    p.add_argument('-R', '--resolutions', type=int, nargs='+',
                   help='List of resolutions corresponding to the input peak files.')
    p.add_argument('-S', '--skip-rows', type=int, default=0, help='Number of leading lines to skip.')
    p.add_argument('-G', '--good-res', type=int, default=20000,
                   help='Peaks detected at finer resolutions (less than this value) are likely to be false positives if '
                        'there are no peak annotations at coarser resolutions in the neighborhood. We keep these peaks only '
                        'if the two loci are <mindis apart.')
    p.add_argument('-M', '--min-dis', type=int, default=200000, help='See --good-res.')
    p.add_argument('--max-res', type=int, default=10000,
                   help='Allowed largest resolution for output, i.e., only peaks originally at this or less than this '
                        'resolution will be outputed.')
    return p


def run_combine(argv=None):
    """combine-resolutions:52-73: one peak file per resolution in, merged BEDPE-style coordinates out."""
    from .combine import combine_annotations
    args = combine_parser().parse_args(argv if argv else ['-h'])
    byres = {res: parse_peakfile(path, args.skip_rows) for res, path in zip(args.resolutions, args.paths)}
    peak_list = combine_annotations(byres, good_res=args.good_res, mindis=args.min_dis, max_res=args.max_res)
    with open(args.output, 'w') as out:
        for t in peak_list:
            out.write('\t'.join(('chr' + t[0], str(t[1]), str(t[2]), 'chr' + t[3], str(t[4]), str(t[5]))) + '\n')
    return peak_list


if __name__ == "__main__":
    (run_bhfdr if (len(sys.argv) > 1 and sys.argv[1] == "bhfdr") else run_hiccups)(sys.argv[2:] if len(sys.argv) > 1 and
                                                                                   sys.argv[1] in ("bhfdr", "hiccups") else sys.argv[1:])
