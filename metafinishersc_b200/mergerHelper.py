"""The consumer of the coverage profile on the hot path's output side, same names as the reference:

    formatIntervalAndFilter  misassemblyFixerLib/mergerHelper.py:137-163
    groupCTestData           the data preparation of mergerHelper.groupCTest (:80-95), up to the statistics

The reference reloads coveragePerContigs.json (one Python int per contig base) and calls np.sum(coverageData[start:end])
for every gap between repeat intervals.  Here coverageData is a ContigCoverage view of the profile that stayed in HBM
(engine.CoverageProfile): all sums of a contig are one mfsc_cov_interval_sums call, two prefix look-ups each.  The
statistics that follow (groupctest.predictDecisionBoundary) are host logic on a handful of numbers and out of scope.
"""
import numpy as np


class ContigCoverage:
    """coverageData of one contig, backed by the resident profile: len(), interval sums, and (on request) the list."""

    def __init__(self, profile, contig):
        self.profile, self.contig = profile, contig

    def __len__(self):
        return self.profile.contig_len(self.contig)

    def sums(self, starts, ends):
        c = np.full(len(starts), self.contig, dtype=np.int32)
        return self.profile.sums(c, starts, ends)

    def tolist(self):
        return self.profile.array(self.contig).tolist()


def _intervals(repeatInterval, n):
    """[start, end) of the stretches before, between and after the repeat intervals that the reference keeps
    (mergerHelper.py:144-161: initThres = 2, lenThres = 1)."""
    initThres, lenThres = 2, 1
    cand = [[initThres, repeatInterval[0][0]]]
    for i in range(0, len(repeatInterval) - 1):
        cand.append([repeatInterval[i][1], repeatInterval[i + 1][0]])
    cand.append([repeatInterval[-1][1], n - initThres])
    return [c for c in cand if c[1] - c[0] > lenThres]


def formatIntervalAndFilter(repeatInterval, coverageData):
    """mergerHelper.py:137-163.  coverageData: ContigCoverage.  Returns (dataList, locList) with
    dataList[k] = [sum of the coverage over at most largestRange bases of the stretch, number of bases summed]."""
    largestRange = 50000
    locList = _intervals(repeatInterval, len(coverageData))
    starts = [s for s, _ in locList]
    ends = [min(s + largestRange, e) for s, e in locList]
    sums = coverageData.sums(starts, ends) if locList else []
    dataList = [[int(sums[k]), ends[k] - starts[k]] for k in range(len(locList))]
    return dataList, locList


def groupCTestData(repeatIntervalDic, profile, names):
    """xDic, locDic of mergerHelper.groupCTest (:88-91) for every contig that has repeat intervals."""
    cid = {n: i for i, n in enumerate(names)}
    xDic, locDic = {}, {}
    for eachitem in repeatIntervalDic:
        xDic[eachitem], locDic[eachitem] = formatIntervalAndFilter(repeatIntervalDic[eachitem], ContigCoverage(profile, cid[eachitem]))
    return xDic, locDic
