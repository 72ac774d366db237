"""TEST INFRASTRUCTURE ONLY -- stand-in for ``statsmodels.sandbox.stats.multicomp``.

statsmodels is not installed in this image and the reference imports
``multipletests`` at module level (/root/reference/hicpeaks/callers.py:11; call
sites :273 and :545, always ``method='fdr_bh'``).  This is a restatement of the
published Benjamini-Hochberg branch of ``statsmodels.stats.multitest
.multipletests`` (version unpinned by the reference: README.rst:21 names the
package only): sort, ecdf = arange(1, n+1)/float(n), step-up reject, reversed
cumulative minimum of p/ecdf, clip at 1, unsort.  Only the first two return
values are consumed by the reference.
"""
import numpy as np


def multipletests(pvals, alpha=0.05, method='fdr_bh', is_sorted=False, returnsorted=False):
    if method not in ('fdr_bh', 'fdr_i', 'fdr_p', 'indep', 'p', 'poscorr'):
        raise NotImplementedError('shim implements fdr_bh only')
    pvals = np.asarray(pvals, dtype=float)
    n = pvals.size
    order = np.argsort(pvals)
    ps = np.take(pvals, order)
    ecdf = np.arange(1, n + 1) / float(n)
    rej = ps <= ecdf * alpha
    if rej.any():
        last = np.nonzero(rej)[0].max()
        rej[:last] = True
    raw = ps / ecdf
    q = np.minimum.accumulate(raw[::-1])[::-1]
    q[q > 1] = 1
    q_out = np.empty_like(q)
    r_out = np.empty_like(rej)
    q_out[order] = q
    r_out[order] = rej
    return r_out, q_out, None, None
