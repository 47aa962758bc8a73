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
    alignerRobot.fullDelta = True           # this test also checks the MUMmer delta records
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
    assert open(folder + "hdronly.delta").read().split("\n")[2].startswith("#MFSC rows-only")
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
    alignerRobot.fullDelta = False
    alignerRobot.set_engine(None)


def test_host_pipeline_emu(emu_engine, oracle, tmp_path):
    run_pipeline(emu_engine, oracle, tmp_path)


def test_large_mode_order(emu_engine, oracle, tmp_path):
    alignerRobot.set_engine(emu_engine)
    folder, d, reads = make_folder(tmp_path, n_reads=60)
    houseKeeper.globalLarge = True
    cwd = os.getcwd()
    os.chdir(folder)
    try:
        alignerRobot.useMummerAlignBatch("", folder, [["big", "LC_filtered.fasta", "SR.fasta", "big"]], 4)
    finally:
        houseKeeper.globalLarge = False
        os.chdir(cwd)
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


def test_large_mode_tile_order(emu_engine, oracle, tmp_path):
    """globalLarge (alignerRobot.py:168-236): the coords file is the concatenation over reference part i = 1..10, query part
    j = 1..10 of `show-coords -r` of tile (i, j) -- the ORDER is asserted, tile by tile, against the oracle run on the part
    files the Perl splitter rule produces."""
    alignerRobot.set_engine(emu_engine)
    d = datagen.abun_gap(G=30_000, repeat_len=900, L=2500, seed=9)
    folder = str(tmp_path) + "/"
    # twelve reference records, so that the ten reference parts are not trivial
    contigs = []
    for c in d["contigs"]:
        contigs += [c[i:i + 5000] for i in range(0, len(c), 5000)]
    cnames = ["ctg%d" % i for i in range(len(contigs))]
    reads = d["reads"][:70]
    rnames = datagen.names("Segkk", len(reads))
    fasta.write_fasta(folder + "ref.fasta", cnames, contigs)
    fasta.write_fasta(folder + "qry.fasta", rnames, reads)
    houseKeeper.globalLarge = True
    cwd = os.getcwd()
    os.chdir(folder)                       # fasta-splitter.pl writes the parts into the CWD and the tiles are read from there
    try:
        alignerRobot.useMummerAlignBatch("", folder, [["big", "ref.fasta", "qry.fasta", "big"]], 4)
    finally:
        houseKeeper.globalLarge = False
        os.chdir(cwd)
    rows = alignerRobot.extractMumData(folder, "bigOut")
    assert os.path.exists(folder + "big0101.delta") and os.path.exists(folder + "ref.part-01.fasta")
    want = []
    rb = fasta.part_boundaries(cnames, contigs, 10)
    qb = fasta.part_boundaries(rnames, reads, 10)
    assert len(rb) > 3 and len(qb) > 3
    for (r0, r1) in rb:
        for (q0, q1) in qb:
            o = oracle.align(contigs[r0:r1], reads[q0:q1], band_cap=248)
            want += oracle.coords_rows(o["rows"], cnames[r0:r1], rnames[q0:q1])
    assert len(want) > 20 and rows == want
    alignerRobot.set_engine(None)


def test_combine_and_special_for_raw(emu_engine, oracle, tmp_path):
    """combineMultipleCoorMum (alignerRobot.py:74-96) and the specialForRaw=True call shape of gapFiller.formRelatedReadsFile
    (gapFiller.py:116-182: reference = 4 x 400-base contig ends, query = read parts addressed by BARE path, coords written to
    <folder><specialName>NN and then joined), followed by gapFiller's own inline parser and its `S1 < 10 and E1 > 390` rule."""
    alignerRobot.set_engine(emu_engine)
    folder, d, reads = make_folder(tmp_path, n_reads=90)
    trunc, tnames = [], []
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for i, c in enumerate(d["contigs"]):
        trunc += [c[:400], c[-400:], c[:400].translate(comp)[::-1], c[-400:].translate(comp)[::-1]]
        tnames += ["Segkk%d_%s" % (i, t) for t in ("ph", "pt", "dh", "dt")]
    fasta.write_fasta(folder + "improved3Trunc.fasta", tnames, trunc)
    # reads that run over the contig ends (noisy copies of the first / last 2.5 kb at several offsets), both strands
    rng = np.random.default_rng(17)
    ends = []
    for c in d["contigs"]:
        g = np.frombuffer(c, dtype=np.uint8)
        for off in (0, 3, 9, 12, 150, 395):
            for piece in (g[off:off + 2500], g[len(g) - 2500 - off:len(g) - off]):
                r, _ = datagen.noisy_reads(piece, 1, 2500, 0.01, 0.01, rng)
                ends.append(r[0] if len(ends) % 2 else r[0].translate(comp)[::-1])
    reads = reads[:40] + ends
    fasta.write_fasta(folder + "SR.fasta", datagen.names("Segkk", len(reads)), reads)
    n_parts = 5
    paths = fasta.split_fasta(folder + "SR.fasta", n_parts, out_dir=folder)
    raw_dir = str(tmp_path) + "/"                     # gapFiller passes the raw read parts with their own (bare) path
    workerList = []
    for i in range(1, n_parts + 1):
        idx = alignerRobot.zeropadding(i)
        workerList.append(["outGapFillRaw" + idx, "improved3Trunc.fasta", raw_dir + "SR.part-" + str(i) + ".fasta", "fromMumRaw" + idx])
    alignerRobot.useMummerAlignBatch("", folder, workerList, 2, specialForRaw=True)
    for i in range(1, n_parts + 1):
        assert os.path.exists(folder + "fromMumRaw" + alignerRobot.zeropadding(i))
        assert not os.path.exists(folder + "outGapFillRaw" + alignerRobot.zeropadding(i) + "Out")
    alignerRobot.combineMultipleCoorMum(True, "", folder, "outGapFillRaw", "fromMumRaw", n_parts)
    joined = open(folder + "fromMumRaw").read().split("\n")
    want_rows, texts = [], []
    for i, pth in enumerate(paths, start=1):
        names, seqs = fasta.read_fasta(pth)
        o = oracle.align(trunc, seqs, band_cap=248)
        want_rows += oracle.coords_rows(o["rows"], tnames, names)
        texts.append(oracle.coords_text(o["rows"], tnames, names, folder + "improved3Trunc.fasta", pth))
    # head -5 of part 01, then tail -n+6 of every part
    assert joined[1:5] == texts[0].split("\n")[1:5]
    body = []
    for t in texts:
        body += t.split("\n")[5:-1]
    assert joined[5:-1] == body and len(body) > 10
    assert alignerRobot.extractMumData(folder, "fromMumRaw") == want_rows
    # gapFiller.py:151-182: its own parser over every per-part file keeps the reads spanning a whole 400-base end
    kept = []
    endThres, thres = 10, 400
    for i in range(1, n_parts + 1):
        with open(folder + "fromMumRaw" + alignerRobot.zeropadding(i)) as f:
            for _ in range(6):
                tmp = f.readline()
            while len(tmp) > 0:
                infoArr = tmp.split("|")
                rdGpArr = infoArr[-1].split("\t")
                contigName, readName = rdGpArr[0].rstrip().lstrip(), rdGpArr[1].rstrip().lstrip()
                pos = [int(x) for x in infoArr[0].split(" ") if len(x) > 0]
                if pos[0] < endThres and pos[1] > thres - endThres:
                    kept.append((contigName, readName))
                tmp = f.readline()
    want_kept = [(r[7], r[8]) for r in want_rows if r[0] < 10 and r[1] > 390]
    assert kept == want_kept and len(kept) > 5
    alignerRobot.set_engine(None)


def test_refined_special_for_raw(emu_engine, oracle, tmp_path):
    """The `-l 10` refine shape of abunSplitter.evaluateCoverage (abunSplitter.py:318-342): refinedVersion=True together with
    specialForRaw=True, query parts by bare path."""
    alignerRobot.set_engine(emu_engine)
    folder, d, reads = make_folder(tmp_path, n_reads=30)
    paths = fasta.split_fasta(folder + "SR.fasta", 3, out_dir=folder)
    workerList = [["refine" + alignerRobot.zeropadding(i), "LC_filtered.fasta", p, "refineOut" + alignerRobot.zeropadding(i)] for i, p in enumerate(paths, start=1)]
    alignerRobot.useMummerAlignBatch("", folder, workerList, 2, specialForRaw=True, refinedVersion=True)
    got, want = [], []
    for i, p in enumerate(paths, start=1):
        got += alignerRobot.extractMumData(folder, "refineOut" + alignerRobot.zeropadding(i))
        names, seqs = fasta.read_fasta(p)
        o = oracle.align(d["contigs"], seqs, band_cap=248, minmatch=10)
        want += oracle.coords_rows(o["rows"], ["Segkk0", "Segkk1"], names)
    assert got == want and len(got) > 0
    alignerRobot.set_engine(None)


def test_index_cache_is_keyed_on_content(emu_engine, tmp_path):
    """IORobot.alignWithName rewrites <name>leftSeg.fasta in place with the same size; with a coarse mtime the second call must
    not be served the first call's index (regression: cache keyed on path/mtime/size)."""
    alignerRobot.set_engine(emu_engine)
    folder = str(tmp_path) + "/"
    rng = np.random.default_rng(3)
    a = datagen.random_genome(3000, rng).tobytes().decode()
    b = datagen.random_genome(3000, rng).tobytes().decode()
    real_stat = os.stat

    def run(left, right):
        out = IORobot.alignWithName(left, right, folder, "", "ov")
        for f in ("ovleftSeg.fasta", "ovrightSeg.fasta"):
            os.utime(folder + f, ns=(10**18, 10**18))          # pin the time stamps: same path, same size, same mtime
        return out
    assert run(a, a[2500:] + b[:2500]) == [500, 500]
    assert run(b, b[2500:] + a[:2500]) == [500, 500]
    assert run(a, b) == [0, 0]
    assert real_stat is os.stat
    alignerRobot.set_engine(None)


def test_write_double_and_reverse_complement(tmp_path):
    """IORobot.writeToFile_Double1 (IORobot.py:107-140) and houseKeeper.reverseComplement (houseKeeper.py:106-126) against a
    character-by-character restatement."""
    folder = str(tmp_path) + "/"
    rng = np.random.default_rng(2)
    seqs = [datagen.random_genome(n, rng).tobytes().decode() for n in (1, 59, 60, 61, 500)]
    seqs[1] = seqs[1][:10] + "NnacgtX" + seqs[1][10:]
    with open(folder + "in.fasta", "w") as f:
        for i, s in enumerate(seqs):
            f.write(">r%d extra words\n" % i)
            for k in range(0, len(s), 60):
                f.write(s[k:k + 60] + "\n")
        f.write("\n>after_blank\nACGT\n")

    def rc_ref(s):
        m = {"A": "T", "a": "T", "T": "A", "t": "A", "C": "G", "c": "G", "G": "C", "g": "C", "N": "N", "n": "N"}
        return "".join(m.get(ch, "") for ch in s[::-1])
    for option, hdr in (("contig", ">Contig"), ("read", ">Read")):
        IORobot.writeToFile_Double1(folder, "in.fasta", "out.fasta", option)
        want = "".join("%s%d_p\n%s\n%s%d_d\n%s\n" % (hdr, i, s, hdr, i, rc_ref(s)) for i, s in enumerate(seqs))
        assert open(folder + "out.fasta").read() == want
    assert IORobot.reverseComplement("ACGTNacgtnX") == rc_ref("ACGTNacgtnX")


def test_pipelined_align_equals_one_call(emu_engine):
    """Engine.align splits a large batch into pieces (upload of piece k+1 on a second host thread while piece k is aligned):
    the rows, the delta lists and the statistics are those of one call."""
    from metafinishersc_b200 import engine
    d = datagen.abun_gap(G=40_000, repeat_len=900, L=2500, seed=11)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = [r.translate(comp)[::-1] if i % 2 else r for i, r in enumerate(d["reads"][:50])] + [b"", b"ACGT"]
    P = engine.make_params()
    ix = emu_engine.index(d["contigs"], P)
    for flags in (0, engine.WANT_DELTAS):
        one = emu_engine.align(ix, reads, P, flags=flags, pieces=1)
        many = emu_engine.align(ix, reads, P, flags=flags, pieces=4)
        for f in ("ref_id", "qry_id", "s1", "e1", "s2", "e2", "errors", "cols"):
            assert np.array_equal(getattr(one, f), getattr(many, f)), f
        assert one.stats["cells"] == many.stats["cells"] and one.stats["max_band"] == many.stats["max_band"] and len(one) > 40
        if flags:
            for i in range(len(one)):
                assert np.array_equal(one.deltas[one.delta_begin[i]:one.delta_end[i]], many.deltas[many.delta_begin[i]:many.delta_end[i]])
    # pieces=None: the first call on an index goes up whole, later ones in pieces when the device rate of the previous call was high
    assert getattr(ix, "_rate", 0) > 0
    one = emu_engine.align(ix, reads, P, pieces=1)
    old = engine.Engine.PIECE_BASES, engine.Engine.PIECE_MIN_RATE
    try:
        engine.Engine.PIECE_BASES, engine.Engine.PIECE_MIN_RATE = 20_000, 0.0
        auto = emu_engine.align(ix, reads, P)
        assert auto.stats["launches"] > one.stats["launches"] and np.array_equal(auto.s1, one.s1) and np.array_equal(auto.qry_id, one.qry_id)
        engine.Engine.PIECE_MIN_RATE = 1e30
        assert emu_engine.align(ix, reads, P).stats["launches"] == one.stats["launches"]
    finally:
        engine.Engine.PIECE_BASES, engine.Engine.PIECE_MIN_RATE = old
    ix.free()


def test_host_buffer_pool(emu_engine):
    """fasta.host_buffer: blocks come from mfsc_host_alloc (page-locked on a GPU box; malloc in the emulation build used here), stay
    alive as long as any view of them does, and are reused by the next request of a similar size."""
    import gc
    from metafinishersc_b200 import fasta
    L = emu_engine.L
    with fasta._PIN_LOCK:
        del fasta._PIN_FREE[:]
    a = fasta.host_buffer(5 << 20, L)
    assert a.size == 5 << 20 and a.flags["WRITEABLE"] and not a.flags["OWNDATA"]
    addr = a.ctypes.data
    a[:] = 3
    v = a[1000:2000]
    del a
    gc.collect()
    assert fasta._PIN_FREE == [] and int(v.sum()) == 3000          # a view keeps the block out of the pool
    del v
    gc.collect()
    assert [b for b, _ in fasta._PIN_FREE] == [16 << 20]
    b = fasta.host_buffer(6 << 20, L)                               # similar size: the same block again
    assert b.ctypes.data == addr and fasta._PIN_FREE == []
    c = fasta.host_buffer(40 << 20, L)                              # a much larger request gets its own block
    assert c.ctypes.data != addr and c.size == 40 << 20
    small = fasta.host_buffer(1000, L)
    assert small.flags["OWNDATA"]                                   # small buffers stay pageable
    del b, c
    gc.collect()
    assert len(fasta._PIN_FREE) == 2


def test_fasta_arrays_threaded_paths(tmp_path):
    """mfsc_fasta_count / mfsc_fasta_parse above 8 MB run on several host threads (chunks cut at line starts / header lines): same
    names and sequences as the line-by-line Python reader on a file with wrapped, unwrapped, CRLF and blank lines."""
    from metafinishersc_b200 import fasta
    rng = np.random.default_rng(3)
    recs, p = [], str(tmp_path / "x.fasta")
    with open(p, "wb") as f:
        f.write(b"ACGT stray line before any header\n")
        for i in range(2500):
            L = int(rng.integers(1, 9000))
            s = bytes(rng.choice(list(b"ACGTN"), size=L).astype(np.uint8))
            w = int(rng.choice([0, 60, 70, 1]))
            f.write(b">r%d some text\r\n" % i if i % 5 == 0 else b">r%d\n" % i)
            if w == 0:
                f.write(s + b"\n")
            else:
                for k in range(0, L, w):
                    f.write(s[k:k + w] + (b"\r\n" if i % 7 == 0 else b"\n"))
            if i % 11 == 0:
                f.write(b"\n\n")
            recs.append(s)
    assert os.path.getsize(p) > (8 << 20)
    names, bases, off = fasta.read_fasta_arrays(p)
    n2, s2 = fasta.read_fasta(p)
    assert names == n2 and len(names) == 2500
    assert all(bases[off[i]:off[i + 1]].tobytes() == recs[i] == s2[i] for i in range(2500))


def test_merged_calls_of_any_size_write_the_same_files(emu_engine, tmp_path):
    """The 20 jobs of largeRvsQAlign are merged into GPU calls of up to alignerRobot._SUPER_BASES query bases (uploads of the next
    call on a second thread): the rows and every per-part file are the same whether that is one call, a few, or one per part."""
    alignerRobot.set_engine(emu_engine)
    folder, d, reads = make_folder(tmp_path, n_reads=120)
    old = alignerRobot._SUPER_BASES
    outs = {}
    try:
        for tag, sb in (("one", 1 << 40), ("few", 60_000), ("each", 1)):
            alignerRobot._SUPER_BASES = sb
            outs[tag] = alignerRobot.largeRvsQAlign(folder, 20, "", "LC_filtered", "SR", tag)
    finally:
        alignerRobot._SUPER_BASES = old
        alignerRobot.set_engine(None)
    assert outs["one"] == outs["few"] == outs["each"] and len(outs["one"]) > 100
    for i in range(1, 21):
        idx = alignerRobot.zeropadding(i)
        for tag in ("few", "each"):
            assert open(folder + tag + idx + "Out").read().split("\n")[1:] == open(folder + "one" + idx + "Out").read().split("\n")[1:]
            assert open(folder + tag + idx + ".delta").read().split("\n")[1:] == open(folder + "one" + idx + ".delta").read().split("\n")[1:]
