"""In-memory stand-in for the slice of the cooler API the front ends use (the image has no cooler / h5py):
``binsize``, ``chromnames``, ``matrix(balance=, sparse=True).fetch(chrom)``, ``bins().fetch(chrom)[name].values``.
Balanced values follow cooler (api.py, sparse branch of ``matrix``): ``bias1[row] * bias2[col] * data`` evaluated left to
right, i.e. ``(w[i] * w[j]) * count`` for stored pixels, NaN where a weight is NaN."""
import numpy as np
from scipy import sparse


class _Matrix:
    def __init__(self, owner, balance):
        self.owner, self.balance = owner, balance

    def fetch(self, chrom):
        Diags, w = self.owner.data[chrom]
        n = len(Diags[0])
        rows, cols, vals = [], [], []
        for d, v in enumerate(Diags):
            nz = np.nonzero(v)[0]
            x = v[nz].astype(np.float64 if self.balance else v.dtype)
            if self.balance:
                with np.errstate(invalid="ignore"):
                    x = w[nz] * w[nz + d] * x            # cooler/api.py: mat.data = bias1[mat.row] * bias2[mat.col] * mat.data
            rows.append(nz); cols.append(nz + d); vals.append(x)
            if d:
                rows.append(nz + d); cols.append(nz); vals.append(x)
        return sparse.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()


class _Col:
    def __init__(self, values):
        self.values = values


class _Bins:
    def __init__(self, owner):
        self.owner = owner

    def fetch(self, chrom):
        return {self.owner.weight_name: _Col(self.owner.data[chrom][1].copy())}


class FakeCooler:
    def __init__(self, binsize, data, weight_name="weight"):
        """data: {chrom: (Diags list of int arrays for d = 0 .. num-1, weights)}"""
        self.binsize, self.data, self.weight_name = binsize, data, weight_name
        self.chromnames = list(data)
        self.chromsizes = {k: len(v[0][0]) * binsize for k, v in data.items()}

    def matrix(self, balance=False, sparse=True):
        return _Matrix(self, balance)

    def bins(self):
        return _Bins(self)
