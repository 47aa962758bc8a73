"""The two IORobot functions on the alignment path (reference
srcRefactor/repeatPhaserLib/finisherSCCoreLib/IORobot.py): obtainLength :8-30 and the small pairwise
overlap aligner align / alignWithName :495-545 (`nucmer -l 10 --maxmatch` on <= 50 kb windows)."""
from . import alignerRobot


def obtainLength(folderName, fileName):
    """name -> summed sequence-line length (IORobot.py:8-30)."""
    lenDic = {}
    tmplen, tmpName = 0, ""
    with open(folderName + fileName, "r") as f:
        for tmp in f:
            tmp = tmp.rstrip()
            if len(tmp) == 0:
                break
            if tmp[0] == ">":
                if tmplen != 0:
                    lenDic[tmpName] = tmplen
                    tmplen = 0
                tmpName = tmp[1:]
            else:
                tmplen += len(tmp)
    lenDic[tmpName] = tmplen
    return lenDic


def align(leftSeg, rightSeg, folderName, mummerLink):
    return alignWithName(leftSeg, rightSeg, folderName, mummerLink, "overlap")


def alignWithName(leftSeg, rightSeg, folderName, mummerLink, nameOfOut):
    """[LEN1, LEN2] of the best end-to-start overlap between the last 50 kb of leftSeg and the first
    50 kb of rightSeg (IORobot.py:498-545); [0, 0] when there is none."""
    left = leftSeg if len(leftSeg) < 50000 else leftSeg[-50000:]
    right = rightSeg if len(rightSeg) < 50000 else rightSeg[0:50000]
    lLen = len(left)
    with open(folderName + nameOfOut + "leftSeg.fasta", "w") as f:
        f.write(">SegL\n")
        f.write(left)
    with open(folderName + nameOfOut + "rightSeg.fasta", "w") as f:
        f.write(">SegR\n")
        f.write(right)
    alignerRobot.useMummerAlign(mummerLink, folderName, nameOfOut, nameOfOut + "leftSeg.fasta", nameOfOut + "rightSeg.fasta",
                                specialForRaw=False, specialName="", refinedVersion=True)
    dataList = alignerRobot.extractMumData(folderName, nameOfOut + "Out")
    thres = 10
    myMax = [0, 0]
    for eachitem in dataList:
        if eachitem[1] > lLen - thres and eachitem[2] < thres:
            if eachitem[5] > myMax[1]:
                myMax[0] = eachitem[4]
                myMax[1] = eachitem[5]
    return myMax
