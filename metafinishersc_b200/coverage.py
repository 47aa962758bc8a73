"""The two consumers of reads-vs-contigs rows on the hot path, same names and files as the reference:

    alignSR2LC        misassemblyFixerLib/merger.py:142-190  -> coveragePerContigs.json  (M-Fixer)
    evaluateCoverage  repeatPhaserLib/abunSplitter.py:260-351 -> contig -> abundance      (A-Splitter)

Row selection stays on the host because it is order-sensitive and seeded (stable sort by read name,
groupby, one Mersenne-Twister draw per read that has a candidate); the per-base accumulation
`for i in range(S1-1, E1): cov[contig][i] += 1` (merger.py:139-140) runs on the GPU as a difference
array + prefix sum (mfsc_coverage).
"""
import json
import random
from itertools import groupby
from operator import itemgetter

import numpy as np

from . import IORobot, alignerRobot, houseKeeper

mergerGlobalLCReads = "LR"      # merger.py:9; mFixer.py:27-28 turns a missing -t option into "LR"


def completelyEmbedSR2TR(eachsub, lenDic):
    """merger.py:101-124."""
    if mergerGlobalLCReads == "SR":
        return abs(lenDic[eachsub[-1]] - eachsub[5]) < 7
    thres = 20
    rdStart, rdEnd = min(eachsub[2], eachsub[3]), max(eachsub[2], eachsub[3])
    return rdStart < thres and rdEnd >= lenDic[eachsub[-1]] - thres


def py2_randint0(n):
    """random.randint(0, n-1) as Python 2 draws it: int(random() * n), exactly one random() per call
    (Python 3's randint uses getrandbits rejection sampling and would pick other rows)."""
    return int(random.random() * n)


def select_rows(dataList, SRLenDic):
    """merger.py:164-187 without the accumulation: the one row per read whose interval gets counted."""
    random.seed(0)
    dataList.sort(key=itemgetter(-1))
    chosen = []
    for key, items in groupby(dataList, itemgetter(-1)):
        itemList = [eachsub for eachsub in items if completelyEmbedSR2TR(eachsub, SRLenDic)]
        if len(itemList) > 0:
            chosen.append(itemList[py2_randint0(len(itemList))])
    return chosen


def coverage_profile_from_rows(chosen, LCLenDic, engine=None):
    """assignCoverage (merger.py:126-140) over all chosen rows at once, on the GPU, left resident in HBM.
    Returns (engine.CoverageProfile, contig names in LCLenDic's order)."""
    E = engine or alignerRobot.get_engine()
    names = list(LCLenDic.keys())
    cid = {n: i for i, n in enumerate(names)}
    off = np.zeros(len(names) + 1, dtype=np.int64)
    np.cumsum([LCLenDic[n] for n in names], out=off[1:])
    rows = [r for r in chosen if r[-2] in cid]
    for r in chosen:
        if r[-2] not in cid:
            print("contig not found" + r[-2])
    ref = np.array([cid[r[-2]] for r in rows], dtype=np.int32)
    s1 = np.array([r[0] for r in rows], dtype=np.int32)
    e1 = np.array([r[1] for r in rows], dtype=np.int32)
    return E.coverage_profile(ref, s1, e1, off), names


def coverage_from_rows(chosen, LCLenDic, engine=None):
    """Same, brought back to the host: {contigName: int32 array}; contigs keep LCLenDic's order."""
    prof, names = coverage_profile_from_rows(chosen, LCLenDic, engine)
    cov, off = prof.array(), prof.off
    prof.free()
    return {n: cov[off[i]:off[i + 1]] for i, n in enumerate(names)}


def alignSR2LC(folderName, mummerLink, incontigName, keep_profile=False):
    """merger.py:142-190: SR.fasta reads vs <incontigName>.fasta contigs -> coveragePerContigs.json."""
    print("alignSR2LC")
    LCLenDic = IORobot.obtainLength(folderName, incontigName + ".fasta")
    outputName, referenceName, queryName = "SR2" + incontigName + "Align", incontigName, "SR"
    dataList = alignerRobot.largeRvsQAlign(folderName, 20, mummerLink, referenceName, queryName, outputName)
    SRLenDic = IORobot.obtainLength(folderName, "SR.fasta")
    chosen = select_rows(dataList, SRLenDic)
    prof, names = coverage_profile_from_rows(chosen, LCLenDic)
    full, off = prof.array(), prof.off
    cov = {n: full[off[i]:off[i + 1]] for i, n in enumerate(names)}
    with open(folderName + "coveragePerContigs.json", "w") as outfile:
        json.dump({k: v.tolist() for k, v in cov.items()}, outfile)
    if keep_profile:     # the group C-test's interval sums then run on the device (mergerHelper.groupCTestData)
        return cov, prof, names
    prof.free()
    return cov


def evaluateCoverage(dataList, lenDic, readLenDic, folderName, mummerLink, continueFilter, contigFilename):
    """abunSplitter.py:260-351: per read the first row maximising IDY/100*LEN1; count[contig] += readLen;
    abundance = count / contig length."""
    myCountDic = {eachitem: 0 for eachitem in lenDic}
    dataList.sort(key=itemgetter(-1))
    ctkk, ctbase = 0, 0
    toAddBackDic = dict(readLenDic)
    for key, items in groupby(dataList, itemgetter(-1)):
        maxMatch, bestname = -1, ""
        for eachitem in items:
            ct = eachitem[6] / 100.0 * eachitem[4]
            if ct > maxMatch:
                maxMatch, bestname = ct, eachitem[-2]
        myCountDic[bestname] += readLenDic[key]
        ctkk += 1
        ctbase += readLenDic[key]
        toAddBackDic[key] = -1
    cttot = sum(readLenDic.values())
    print("Missed coverage  ", (cttot - ctbase) / (4.7 * pow(10, 6)))
    print("percentage miss read", (len(readLenDic) - ctkk) / (1.0 * len(readLenDic)))
    toAddReadList = [k for k in toAddBackDic if toAddBackDic[k] >= 0]
    if continueFilter:
        # abunSplitter.py:318-342: the unmapped reads are re-aligned with -l 10 in 20 parts
        from . import fasta
        numberOfFiles = houseKeeper.globalParallelFileNum
        names, seqs = fasta.read_fasta(folderName + "raw_reads.fasta")
        want = set(toAddReadList)
        keep = [(n, s) for n, s in zip(names, seqs) if n in want]
        fasta.write_fasta(folderName + "selected_raw.fasta", [n for n, _ in keep], [s for _, s in keep])
        fasta.split_fasta(folderName + "selected_raw.fasta", numberOfFiles, out_dir=folderName)
        workerList = []
        for i in range(1, numberOfFiles + 1):
            idx = alignerRobot.zeropadding(i)
            workerList.append(["outAbunRefine" + idx, contigFilename + ".fasta", folderName + "selected_raw.part-" + idx + ".fasta",
                               "abunMissOut" + idx])
        alignerRobot.useMummerAlignBatch(mummerLink, folderName, workerList, houseKeeper.globalParallel, specialForRaw=True,
                                         refinedVersion=True)
        alignerRobot.combineMultipleCoorMum(True, mummerLink, folderName, "outAbunRefine", "abunMissOut", numberOfFiles)
    for eachitem in lenDic:
        print(eachitem, myCountDic[eachitem] / (1.0 * lenDic[eachitem]))
        myCountDic[eachitem] = myCountDic[eachitem] / (1.0 * lenDic[eachitem])
    return myCountDic


def generateAbundanceCounts(folderName, mummerLink, contigFilename="improved3", readsetFilename="raw_reads",
                            continueFilter=False):
    """The alignment + counting part of abunSplitter.generateAbundanceGraph (abunSplitter.py:353-432):
    20 jobs <contigFilename>.fasta x raw_reads.part-NN.fasta -> outAbunNNOut, then evaluateCoverage and
    myCountDic.json."""
    from . import fasta
    numberOfFiles = houseKeeper.globalParallelFileNum
    fasta.split_fasta(folderName + readsetFilename + ".fasta", numberOfFiles, out_dir=folderName)
    workerList = []
    for i in range(1, numberOfFiles + 1):
        idx = alignerRobot.zeropadding(i)
        workerList.append(["outAbun" + idx, contigFilename + ".fasta", readsetFilename + ".part-" + idx + ".fasta", "outAbun" + idx])
    alignerRobot.useMummerAlignBatch(mummerLink, folderName, workerList, houseKeeper.globalParallel, False)
    dataList = []
    for i in range(1, numberOfFiles + 1):
        dataList = dataList + alignerRobot.extractMumData(folderName, "outAbun" + alignerRobot.zeropadding(i) + "Out")
    lenDic = IORobot.obtainLength(folderName, contigFilename + ".fasta")
    readLenDic = IORobot.obtainLength(folderName, readsetFilename + ".fasta")
    myCountDic = evaluateCoverage(dataList, lenDic, readLenDic, folderName, mummerLink, continueFilter, contigFilename)
    with open(folderName + "myCountDic.json", "w") as f:
        json.dump(myCountDic, f)
    return myCountDic
