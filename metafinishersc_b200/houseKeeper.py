"""Process-wide switches of the reference's houseKeeper module that steer the aligner wrapper
(reference srcRefactor/repeatPhaserLib/finisherSCCoreLib/houseKeeper.py:4-9): same names, same
defaults, read by alignerRobot at call time."""
globalFast = False            # nucmer -b 50
globalParallel = 1            # advisory here: jobs run in-process on the GPU
globalLarge = False           # 10x10 reference x query tiling in the reference; row order is reproduced
globalRelaxThres = False
globalRunMPI = False          # the MPI fan-out is replaced by in-process dispatch
globalParallelFileNum = 20    # read parts per reads-vs-contigs alignment
globalGPUs = 1                # GPUs of this box the batch calls may use (one host thread per GPU); not in the reference
