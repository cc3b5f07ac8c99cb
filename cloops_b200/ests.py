"""Distance cut-off estimation (cLoops/ests.py) -- O(N) numpy on the host; it is the one cross-
chromosome synchronisation point of a clustering round (cLoops/pipe.py:259)."""
from collections import Counter

import numpy as np
import pandas as pd


def estFragSize(ds, top=500):
    """cLoops/ests.py:23-33: median of the 500 most frequent opposite-strand PET distances."""
    s = pd.Series(Counter(ds))
    s.sort_values(inplace=True, ascending=False)
    return int(np.median(s[:top].index))


def estIntSelCutFrag(di, ds, log=1):
    """cLoops/ests.py:36-61 -> (distance cut-off, fragment size); both integers."""
    di = np.abs(np.array(di))
    ds = np.abs(np.array(ds))
    di = di[~np.isnan(di)]
    ds = ds[~np.isnan(ds)]
    di = di[di > 0]
    ds = ds[ds > 0]
    if log:
        di = np.log2(di)
        ds = np.log2(ds)
    cut1 = np.median(ds) + 3 * ds.std()
    cut2 = (ds.mean() * ds.std() + di.mean() * di.std()) / (ds.std() + di.std())
    rcut = int(2 ** min([cut1, cut2]))
    rfrags = int(2 ** np.median(ds))
    return rcut, rfrags
