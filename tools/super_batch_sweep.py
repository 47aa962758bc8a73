import os, sys, tempfile, time
sys.path.insert(0, os.getcwd())
from metafinishersc_b200 import alignerRobot, datagen, fasta
d = datagen.sc_lr(G=5_000_000, n_reads=25000, seed=100, read_seed=2100)
with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
    folder = tmp + "/"
    fasta.write_fasta(folder + "LC.fasta", ["Contig0"], d["contigs"])
    fasta.write_fasta(folder + "LR.fasta", datagen.names("Segkk", len(d["reads"])), d["reads"])
    os.chdir(tmp)
    alignerRobot.largeRvsQAlign(folder, 20, "", "LC", "LR", "warm")
    for sb in (32_000_000, 64_000_000, 96_000_000, 128_000_000, 260_000_000):
        alignerRobot._SUPER_BASES = sb
        best = 1e9
        for it in range(3):
            t0 = time.perf_counter()
            alignerRobot.largeRvsQAlign(folder, 20, "", "LC", "LR", "h%d_%d" % (sb, it))
            best = min(best, time.perf_counter() - t0)
        print("super %d Mbp: %.3f s" % (sb // 1000000, best))
