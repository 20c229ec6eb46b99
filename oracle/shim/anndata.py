"""anndata-surface shim -- ORACLE / TEST INFRASTRUCTURE ONLY.

The reference imports `anndata` at module import time in baf/io.py, rdr/io.py,
utils/csp_io.py and baf/rpc.py; nothing on the two counting paths constructs an
AnnData unless `cellsnp_dir` is given (local phasing, out of scope in round 1).
"""


class AnnData(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: anndata is not available in this image")


def read_h5ad(*a, **k):
    raise NotImplementedError("shim: anndata is not available in this image")
