"""The two IORobot functions on the alignment path (reference
srcRefactor/repeatPhaserLib/finisherSCCoreLib/IORobot.py): obtainLength :8-30 and the small pairwise
overlap aligner align / alignWithName :495-545 (`nucmer -l 10 --maxmatch` on <= 50 kb windows)."""
import numpy as np

from . import alignerRobot, fasta

_RC_TABLE = np.zeros(256, dtype=np.uint8)      # houseKeeper.reverseComplement: ACGT (either case) -> TGCA, N/n -> N, anything else dropped
for _a, _b in zip(b"ACGTNacgtn", b"TGCANTGCAN"):
    _RC_TABLE[_a] = _b


def reverseComplement(myStr):
    """houseKeeper.reverseComplement (houseKeeper.py:106-126): upper-case complement of the reversed string; N stays N;
    any other character is dropped from the result (the reference prints it and appends nothing)."""
    b = np.frombuffer(myStr.encode() if isinstance(myStr, str) else bytes(myStr), dtype=np.uint8)
    out = _RC_TABLE[b[::-1]]
    out = out[out != 0].tobytes()
    return out.decode() if isinstance(myStr, str) else out


def writeToFile_Double1(folderName, fileName1, fileName2, option="contig"):
    """IORobot.writeToFile_Double1 (IORobot.py:107-140): every record of <fileName1> written twice to <fileName2>, as
    `>Contig<i>_p` / `>Read<i>_p` (sequence on one line, as read) and `>..<i>_d` (its reverse complement).  These
    `*_Double.fasta` files are the inputs of the CC / RR / CR alignments (readContigGraphFormer.py:99-131).  The reference
    stops reading at the first empty line and always emits at least one (possibly empty) record; both are kept.  One
    vectorised pass over the parsed arrays (table look-up on the reversed bytes) instead of a per-character loop."""
    with open(folderName + fileName1, "rb") as f:
        data = f.read()
    reads, cur = [], []
    for line in data.split(b"\n"):
        line = line.rstrip()
        if not line:
            break
        if line[:1] == b">":
            if cur:
                reads.append(b"".join(cur))
                cur = []
        else:
            cur.append(line)
    reads.append(b"".join(cur))
    print("len(readSet)", len(reads))
    header = b">Contig" if option == "contig" else b">Read"
    with open(folderName + fileName2, "wb") as f2:
        for i, seq in enumerate(reads):
            f2.write(header + str(i).encode() + b"_p\n" + seq + b"\n")
            f2.write(header + str(i).encode() + b"_d\n" + reverseComplement(seq) + b"\n")


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
