"""Row filters of the reference's consumers over the rows of one alignment job, same function names:

    filterData, filterDataIdentical   repeatPhaserLib/abunHouseKeeper.py:146-160 (headTailMatch :177-202, identicalItem :162-175)
    checkSatisfy rows                 repeatPhaserLib/associatedReadFinder.py:16-43
    completelyEmbedSR2TR rows         misassemblyFixerLib/merger.py:101-124 (LR mode)

The reference walks a Python list of 9-field rows and tests every row with dictionary look-ups; here the predicates of all
rows are one mfsc_row_flags call over the job's SoA row buffer (alignerRobot.RowSet) and the functions return the rows
the reference would keep, in the reference's order (the show-coords -r order extractMumData yields).
"""
import numpy as np

from . import _lib, alignerRobot


def row_flags(rowset, E=None):
    """MFSC_RF_* bits of every row of the RowSet, in its own (delta) order."""
    E = E or alignerRobot.get_engine()
    gid = {}
    rg = np.array([gid.setdefault(n, len(gid)) for n in rowset.ref_names], dtype=np.int32)
    qg = np.array([gid.setdefault(n, len(gid)) for n in rowset.qry_names], dtype=np.int32)
    return E.row_flags(rowset.rows, rowset.ref_lens, rowset.qry_lens, rg, qg)


def _keep(rowset, mask_fn, E=None):
    f = row_flags(rowset, E)
    o = rowset.order()
    dl = rowset.datalist()
    return [dl[k] for k in range(len(o)) if mask_fn(int(f[o[k]]))]


def filterData(rowset, E=None):
    """abunHouseKeeper.filterData: rows with headTailMatch."""
    return _keep(rowset, lambda f: f & _lib.RF_HEADTAIL, E)


def filterDataIdentical(rowset, E=None):
    """abunHouseKeeper.filterDataIdentical: rows that are not identicalItem."""
    return _keep(rowset, lambda f: not (f & _lib.RF_IDENTICAL), E)


def associatedRows(rowset, E=None):
    """associatedReadFinder.getAllAssociatedReads keeps the rows with checkSatisfy."""
    return _keep(rowset, lambda f: f & _lib.RF_ASSOCIATED, E)


def embeddedRows(rowset, E=None):
    """merger.alignSR2LC keeps the rows with completelyEmbedSR2TR (LR mode)."""
    return _keep(rowset, lambda f: f & _lib.RF_EMBEDDED_LR, E)
