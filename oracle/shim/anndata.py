"""anndata-surface shim -- ORACLE / TEST INFRASTRUCTURE ONLY.

The reference imports `anndata` at module import time in baf/io.py, rdr/io.py, utils/csp_io.py and baf/rpc.py;
on the two counting paths an AnnData is only built when `cellsnp_dir` is given (local phasing:
xcltk/utils/csp_io.py:16-63 builds it, xcltk/baf/fc/main.py:112-149,419-454 slices it).  This stand-in keeps
obs / var as pandas DataFrames and the layers as dense arrays, and supports exactly those uses: construction,
uns / obsm / layers dictionaries, transpose(), shape, copy() and [rows, cols] selection by boolean masks, by
pandas Index labels or by `:`.
"""

import numpy as np
import pandas as pd


class AnnData(object):
    def __init__(self, X=None, obs=None, var=None):
        self.X = X
        self.obs = obs if obs is not None else pd.DataFrame()
        self.var = var if var is not None else pd.DataFrame()
        self.uns, self.obsm, self.varm, self.layers = {}, {}, {}, {}

    @property
    def shape(self):
        return (len(self.obs), len(self.var))

    def transpose(self):
        t = AnnData(None, self.var, self.obs)
        t.uns = dict(self.uns)
        t.obsm, t.varm = dict(self.varm), dict(self.obsm)
        t.layers = {k: np.asarray(v).T for k, v in self.layers.items()}
        return t

    def copy(self):
        c = AnnData(None, self.obs.copy(), self.var.copy())
        c.uns = dict(self.uns)
        c.obsm, c.varm = dict(self.obsm), dict(self.varm)
        c.layers = {k: np.array(v) for k, v in self.layers.items()}
        return c

    @staticmethod
    def _positions(sel, frame):
        if isinstance(sel, slice):
            return np.arange(len(frame))[sel]
        if isinstance(sel, pd.Index):
            pos = frame.index.get_indexer(sel)
            assert (pos >= 0).all()
            return pos
        arr = np.asarray(sel)
        if arr.dtype == bool:
            assert len(arr) == len(frame)
            return np.nonzero(arr)[0]
        return arr.astype(np.int64)

    def __getitem__(self, key):
        rows, cols = key if isinstance(key, tuple) else (key, slice(None))
        r, c = self._positions(rows, self.obs), self._positions(cols, self.var)
        out = AnnData(None, self.obs.iloc[r], self.var.iloc[c])
        out.uns = dict(self.uns)
        out.layers = {k: np.asarray(v)[np.ix_(r, c)] for k, v in self.layers.items()}
        return out


def read_h5ad(*a, **k):
    raise NotImplementedError("shim: anndata is not available in this image")
