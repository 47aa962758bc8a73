"""Host-side mirror of the reference's aligner wrapper and coverage consumers, driven end to end on a
scaled README-shape data set (two genomes sharing a segment, two chimeric contigs, noisy reads).

On CPU the wrapper is bound to the host-emulation build; the GPU variant of the same test lives in
test_gpu_pipeline.py.  Expected values come from the oracle rows pushed through a line-by-line
Python 3 transliteration of the reference loops (merger.py:126-187, abunSplitter.py:283-349)."""
import json
import os
import random
from itertools import groupby
from operator import itemgetter

import numpy as np

from metafinishersc_b200 import IORobot, alignerRobot, coverage, datagen, fasta, houseKeeper


def make_folder(tmp_path, n_reads=240):
    d = datagen.abun_gap(G=60_000, repeat_len=1200, L=3000, seed=4)
    folder = str(tmp_path) + "/"
    reads = d["reads"][:n_reads // 2] + d["reads"][-n_reads // 2:]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = [r.translate(comp)[::-1] if i % 3 == 0 else r for i, r in enumerate(reads)]
    fasta.write_fasta(folder + "LC_filtered.fasta", datagen.names("Segkk", 2), d["contigs"])
    fasta.write_fasta(folder + "SR.fasta", datagen.names("Segkk", len(reads)), reads)
    fasta.write_fasta(folder + "reference.fasta", datagen.names("Segkk", 2), d["reference"])
    return folder, d, reads


def reference_rows(O, contigs, cnames, parts, folder_reads_names):
    """Rows as the reference would collect them: per part, show-coords -r order, parts concatenated."""
    out = []
    for names, seqs in parts:
        o = O.align(contigs, seqs, band_cap=248)
        out += O.coords_rows(o["rows"], cnames, names)
    return out


def transliterated_alignSR2LC(dataList, LCLenDic, SRLenDic):
    """merger.py:159-187, Python 3 spelling, py2 randint."""
    cov = {k: [0] * LCLenDic[k] for k in LCLenDic}
    random.seed(0)
    dataList.sort(key=itemgetter(-1))
    for key, items in groupby(dataList, itemgetter(-1)):
        itemList = []
        for eachsub in items:
            rdStart, rdEnd = min(eachsub[2], eachsub[3]), max(eachsub[2], eachsub[3])
            if rdStart < 20 and rdEnd >= SRLenDic[eachsub[-1]] - 20:
                itemList.append(eachsub)
        if len(itemList) > 0:
            i = int(random.random() * len(itemList))
            for j in range(itemList[i][0] - 1, itemList[i][1]):
                cov[itemList[i][-2]][j] += 1
    return cov


def run_pipeline(engine, oracle, tmp_path):
    alignerRobot.set_engine(engine)
    folder, d, reads = make_folder(tmp_path)
    cwd = os.getcwd()
    try:
        cov = coverage.alignSR2LC(folder, "/unused/mummer/", "LC_filtered")
    finally:
        os.chdir(cwd)
    # files the rest of the pipeline re-reads
    for i in range(1, 21):
        idx = alignerRobot.zeropadding(i)
        assert os.path.exists(folder + "SR2LC_filteredAlign" + idx + "Out")
        assert os.path.exists(folder + "SR2LC_filteredAlign" + idx + ".delta")
        assert os.path.exists(folder + "SR.part-" + idx + ".fasta")
    # expected: oracle rows per part through the transliterated loops
    parts = []
    for i in range(1, 21):
        parts.append(fasta.read_fasta(folder + "SR.part-" + alignerRobot.zeropadding(i) + ".fasta"))
    cnames = ["Segkk0", "Segkk1"]
    want_rows = reference_rows(oracle, d["contigs"], cnames, parts, None)
    got_rows = []
    for i in range(1, 21):
        got_rows += alignerRobot.extractMumData(folder, "SR2LC_filteredAlign" + alignerRobot.zeropadding(i) + "Out")
    assert got_rows == want_rows
    LCLenDic = IORobot.obtainLength(folder, "LC_filtered.fasta")
    SRLenDic = IORobot.obtainLength(folder, "SR.fasta")
    want_cov = transliterated_alignSR2LC(want_rows, LCLenDic, SRLenDic)
    on_disk = json.load(open(folder + "coveragePerContigs.json"))
    assert on_disk == want_cov
    assert all(np.array_equal(cov[k], np.array(want_cov[k])) for k in want_cov)
    # text identical to what the oracle formats
    o1 = oracle.align(d["contigs"], parts[0][1], band_cap=248)
    text = oracle.coords_text(o1["rows"], cnames, parts[0][0], folder + "LC_filtered.fasta", folder + "SR.part-01.fasta")
    assert open(folder + "SR2LC_filteredAlign01Out").read().split("\n")[1:] == text.split("\n")[1:]
    # the delta file is a real MUMmer delta: header, per alignment the oracle's indel offsets, terminating 0
    dl = open(folder + "SR2LC_filteredAlign01.delta").read().split("\n")
    assert dl[1] == "NUCMER" and dl[2].startswith(">Segkk")
    k = 3
    for row in o1["rows"]:
        if dl[k].startswith(">"):
            k += 1
        hdr = dl[k].split()
        assert [int(x) for x in hdr[:5]] == [int(row[3]), int(row[4]), int(row[5]), int(row[6]), int(row[7])] and hdr[6] == "0"
        want_d = [str(int(x)) for x in o1["deltas"][row[9]:row[10]]] + ["0"]
        assert dl[k + 1:k + 1 + len(want_d)] == want_d
        k += 1 + len(want_d)
    # resume path: rows reloaded from the delta file give the same coords text (both delta forms)
    alignerRobot.clear_caches()
    alignerRobot.showCoorMummer(False, "", folder, "SR2LC_filteredAlign01", "")
    assert open(folder + "SR2LC_filteredAlign01Out").read().split("\n")[5:] == text.split("\n")[5:]
    alignerRobot.fullDelta = False
    alignerRobot.useMummerAlign("", folder, "hdronly", "LC_filtered.fasta", "SR.part-01.fasta")
    alignerRobot.fullDelta = True
    alignerRobot.clear_caches()
    alignerRobot.showCoorMummer(False, "", folder, "hdronly", "")
    assert open(folder + "hdronlyOut").read().split("\n")[5:] == text.split("\n")[5:]
    # A-Splitter abundance on the same rows
    fasta.write_fasta(folder + "improved3.fasta", cnames, d["contigs"])
    fasta.write_fasta(folder + "raw_reads.fasta", datagen.names("Segkk", len(reads)), reads)
    counts = coverage.generateAbundanceCounts(folder, "/unused/", "improved3", "raw_reads")
    want = {k: 0 for k in cnames}
    rows2 = sorted(want_rows, key=itemgetter(-1))
    for key, items in groupby(rows2, itemgetter(-1)):
        best, name = -1, ""
        for it in items:
            ct = it[6] / 100.0 * it[4]
            if ct > best:
                best, name = ct, it[-2]
        want[name] += SRLenDic[key]
    assert counts == {k: want[k] / (1.0 * LCLenDic[k]) for k in cnames}
    # small pairwise overlap (IORobot.align, nucmer -l 10)
    c0 = d["contigs"][0].decode()
    ov = IORobot.align(c0[1000:9000], c0[7000:15000], folder, "/unused/")
    assert ov == [2000, 2000]
    assert IORobot.align(c0[0:3000], c0[20000:23000], folder, "/unused/") == [0, 0]
    # default-mode self alignment (adaptorFix.py:21) and transformCoor
    rs = alignerRobot.nucmerDefault("", folder + "adaptorCk", folder + "LC_filtered.fasta", folder + "LC_filtered.fasta")
    alignerRobot.showCoorMummer(False, "", folder, "adaptorCk", "")
    dl = alignerRobot.extractMumData(folder, "adaptorCkOut")
    od = oracle.align(d["contigs"], d["contigs"], band_cap=248, maxmatch=False)
    assert dl == oracle.coords_rows(od["rows"], cnames, cnames)
    assert all(r[2] <= r[3] for r in alignerRobot.transformCoor([list(r) for r in got_rows]))
    # a missing input leaves an empty coords file and an empty row list, like the ignored nucmer failure
    alignerRobot.useMummerAlign("", folder, "nothing", "absent.fasta", "SR.fasta")
    assert alignerRobot.extractMumData(folder, "nothingOut") == []
    alignerRobot.set_engine(None)


def test_host_pipeline_emu(emu_engine, oracle, tmp_path):
    run_pipeline(emu_engine, oracle, tmp_path)


def test_large_mode_order(emu_engine, oracle, tmp_path):
    alignerRobot.set_engine(emu_engine)
    folder, d, reads = make_folder(tmp_path, n_reads=60)
    houseKeeper.globalLarge = True
    try:
        alignerRobot.useMummerAlignBatch("", folder, [["big", "LC_filtered.fasta", "SR.fasta", "big"]], 4)
    finally:
        houseKeeper.globalLarge = False
    rows = alignerRobot.extractMumData(folder, "bigOut")
    o = oracle.align(d["contigs"], reads, band_cap=248)
    want = oracle.coords_rows(o["rows"], ["Segkk0", "Segkk1"], datagen.names("Segkk", len(reads)))
    assert sorted(map(tuple, rows)) == sorted(map(tuple, want))
    alignerRobot.set_engine(None)


def test_batch_over_two_devices_emu(emu_engine, oracle, tmp_path):
    """The in-process fan-out over GPUs (one host thread per device) writes the same files as the serial loop."""
    alignerRobot.set_engine(emu_engine)
    folder, d, reads = make_folder(tmp_path, n_reads=80)
    serial = alignerRobot.largeRvsQAlign(folder, 20, "", "LC_filtered", "SR", "serial")
    houseKeeper.globalGPUs = 2
    try:
        par = alignerRobot.largeRvsQAlign(folder, 20, "", "LC_filtered", "SR", "par")
    finally:
        houseKeeper.globalGPUs = 1
    assert par == serial and len(par) > 0
    for i in range(1, 21):
        idx = alignerRobot.zeropadding(i)
        assert open(folder + "par" + idx + "Out").read().split("\n")[1:] == open(folder + "serial" + idx + "Out").read().split("\n")[1:]
        assert open(folder + "par" + idx + ".delta").read().split("\n")[1:] == open(folder + "serial" + idx + ".delta").read().split("\n")[1:]
    alignerRobot.set_engine(None)
