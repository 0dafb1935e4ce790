#!/usr/bin/env python3
"""Randomised parity sweep, CPU only: the kernel sources compiled for the host (tests/cudasim) against the oracle over
random read shapes (20-301 bases, single / paired, error, N, repeat and chimeric rates), random option values and flags
of the Java API, and random subsets of the forced fall-back paths (B200_TEST_* hooks), on three small genomes.
usage: python tests/fuzz_sim.py [seed] [seconds] [sim|engine] [fraction of long-query calls] [fraction of ragged batches]   -- prints every mismatch and a summary; exit status 1 on any.
`engine` runs the same sweep through the CUDA library on a B200 (tests/test_fuzz_sim.py::test_fuzz_slice_gpu).
(tests/test_fuzz_sim.py runs a short fixed-seed slice of it in the CPU suite.)"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ['B200_TEST_HOOKS'] = '1'
import common
from common import simdata, run_blob, sim_lib, engine_lib, oracle_lib, first_diff, bwa
seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tmax = float(sys.argv[2]) if len(sys.argv) > 2 else 300
which = sys.argv[3] if len(sys.argv) > 3 else 'sim'
LONG_FRAC = float(sys.argv[4]) if len(sys.argv) > 4 else 0.03
RAGGED_FRAC = float(sys.argv[5]) if len(sys.argv) > 5 else 0.15
rng = random.Random(seed0)
genomes = [common.synth_image("g200k", 7, [120000, 60000, 20000], repeat_frac=0.08, tandem_frac=0.01, n_segdup=3),
           common.synth_image("g30k_rep", 11, [20000, 8000, 2000], repeat_frac=0.3, tandem_frac=0.05, n_segdup=4),
           common.synth_image("g3k", 13, [3000], repeat_frac=0.1, tandem_frac=0.02, n_segdup=1),
           common.alt_contig_image()]
t0 = time.time(); it = 0; bad = 0
while time.time() - t0 < tmax:
    it += 1
    g, img = rng.choice(genomes)
    paired = rng.random() < 0.6
    rl = rng.choice([20, 25, 36, 50, 76, 100, 150, 151, 200, 250, 301])
    n = rng.choice([1, 2, 7, 40, 150])
    if rng.random() < LONG_FRAC:  # assembled-contig-like queries (the long-read mode above 1000 bases), a few per call, from the genomes with long contigs
        rl = rng.choice([600, 1100, 2500, 6000]); n = rng.choice([1, 2, 3]); paired = paired and rl <= 1100
        g, img = rng.choice([genomes[0], genomes[3]])
    kw = dict(err_rate=rng.choice([0, 0.01, 0.05, 0.12]), n_rate=rng.choice([0, 0, 0.01, 0.1]), repeat_read_frac=rng.choice([0, 0.2, 0.8]),
              chimeric_frac=rng.choice([0, 0, 0.2]), mut_rate=rng.choice([0.001, 0.02]), indel_frac=rng.choice([0.15, 0.5]))
    if paired: kw.update(ins_mean=rng.choice([200, 400, 800]) + (rl if rl > 301 else 0), ins_std=rng.choice([10, 50, 150]))
    blob, _ = simdata.make_reads(g, rng.randrange(1 << 30), n, rl, paired, **kw)
    if rng.random() < RAGGED_FRAC and rl <= 301:  # a ragged batch: units of a second length (down to the empty read) spliced in
        import struct
        rl2 = rng.choice([0, 1, 18, 19, 33, 120, 251])
        units = [simdata.split_blob(bytes(blob))[k:k + (2 if paired else 1)] for k in range(0, n * (2 if paired else 1), 2 if paired else 1)]
        if rl2 >= 18:
            blob2, _ = simdata.make_reads(g, rng.randrange(1 << 30), max(1, n // 3), rl2, paired, **kw)
            r2 = simdata.split_blob(bytes(blob2))
            units += [r2[k:k + (2 if paired else 1)] for k in range(0, len(r2), 2 if paired else 1)]
        else:
            units += [[b'ACGTN'[:rl2]] * (2 if paired else 1) for _ in range(3)]
        rng.shuffle(units)
        reads = [r for u in units for r in u]
        blob = bytearray(struct.pack('<I', len(reads)) + b''.join(r + b'\0' for r in reads))
        kw = dict(kw, ragged=rl2)
    opts = {}
    if rng.random() < 0.7:
        for name, choices in [("MinSeedLengthOption", [10, 15, 19, 25]), ("BandwidthOption", [5, 30, 100, 200]), ("ZDropOption", [10, 100, 300]),
                              ("MaxSeedOccurencesOption", [3, 50, 500]), ("MatchScoreOption", [1, 2, 3]), ("MismatchPenaltyOption", [2, 4, 9]),
                              ("DGapOpenPenaltyOption", [2, 6, 16]), ("IGapOpenPenaltyOption", [3, 6, 16]), ("DGapExtendPenaltyOption", [1, 2]), ("IGapExtendPenaltyOption", [1, 3]),
                              ("Clip5PenaltyOption", [0, 5, 20]), ("Clip3PenaltyOption", [0, 5, 20]), ("UnpairedPenaltyOption", [0, 17, 40]),
                              ("OutputScoreThresholdOption", [0, 20, 30, 60]), ("MaxMemIntvOption", [0, 5, 20]), ("SplitWidthOption", [2, 10, 50]),
                              ("MaxChainGapOption", [100, 10000]), ("MaxMateSWOption", [1, 50]), ("MaxXAHitsOption", [0, 2, 5, 30]), ("MaxXAHitsAltOption", [5, 200]),
                              ("MinChainWeightOption", [0, 0, 20, 40]), ("MaxChainExtendOption", [1, 3, 1 << 30]), ("MaxInsOption", [500, 10000])]:
            if rng.random() < 0.25: opts[name] = rng.choice(choices)
        for name, choices in [("SplitFactorOption", [1.0, 1.5, 2.5]), ("MaxLevelOption", [0.3, 0.5, 0.9]), ("DropRatioOption", [0.3, 0.5, 0.8]), ("XADropRatio", [0.5, 0.8, 0.95]), ("MaskLevelRedunOption", [0.8, 0.95])]:
            if rng.random() < 0.2: opts[name] = rng.choice(choices)
    flag = 2 if paired else 0
    for f, p in [(0x4, 0.1), (0x8, 0.15), (0x10, 0.1), (0x20, 0.15), (0x200, 0.15), (0x800, 0.1), (0x400, 0.05)]:
        if rng.random() < p: flag |= f
    HOOKS = ["B200_TEST_RESCUE_FALLBACK", "B200_TEST_TINY_SCRATCH", "B200_TEST_ALL_LIGHT", "B200_TEST_ALL_HEAVY", "B200_WARP_EXTEND", "B200_NO_EXT_TASKS", "B200_TEST_CHAIN_ALL_WARP",
             "B200_TEST_CHAIN_ALL_THREAD", "B200_FINAL_ONE_PASS", "B200_SEED_ONE_LAUNCH", "B200_SEED_FUSED_ROUNDS", "B200_TEST_TINY_SEED_POOL", "B200_TX_CHUNK=1", "B200_TX_CHUNK=0", "B200_EXT_WIDE_CELLS",
             "B200_TEST_FINAL_RING", "B200_TEST_TINY_GA_CAP", "B200_TEST_TINY_SEED_TASKS", "B200_EXT_LEGACY", "B200_EXT_NO_SORT", "B200_TEST_LONG_MODE", "B200_RESCUE_WARP_SW", "B200_EXT_TASK_MACHINE",
             "B200_GA_NO_SORT", "BWA_B200_KMER=3", "B200_TEST_OCC_SB=2", "B200_TEST_SEED_SLICES=3", "B200_TEST_TINY_INTV_CAP", "B200_TEST_TINY_WARP_SCRATCH", "BWA_B200_SA_INTV=1", "BWA_B200_SA_INTV=32"]
    for h in HOOKS: os.environ.pop(h.split('=')[0], None)
    hooks = [h for h in HOOKS if rng.random() < 0.06]
    for h in hooks:
        k, _, v = h.partition('='); os.environ[k] = v or '1'
    opts = dict(opts); opts_dbg = dict(opts, hooks=hooks)
    def setup(al, opts=opts):
        for k, v in opts.items():
            getattr(al, "set" + k)(v)
    if os.environ.get('FUZZ_VERBOSE'): print('CFG', it, dict(rl=rl, n=n, paired=paired, kw=kw, opts=opts_dbg, flag=hex(flag), img=os.path.basename(img)), flush=True)
    stats = None  # pestat: inferred per call, declared failed, or given (BwaMemAligner.java:178-190)
    if paired and rng.random() < 0.25:
        stats = bwa.BwaMemPairEndStats.DO_NOT_INFER if rng.random() < 0.3 else bwa.BwaMemPairEndStats.of(rng.choice([200.0, 400.0, 800.0]), rng.choice([10.0, 50.0, 150.0]), rng.choice([1, 100]), rng.choice([600, 1200, 5000]))
        opts_dbg['stats'] = str(stats)
    try:
        a = run_blob(oracle_lib(), img, blob, flag, stats, setup)
        b = run_blob(engine_lib() if which == 'engine' else sim_lib(), img, blob, flag, stats, setup)
    except Exception as e:
        print("EXC", it, e, dict(rl=rl, n=n, paired=paired, kw=kw, opts=opts_dbg, flag=hex(flag), img=os.path.basename(img))); bad += 1; continue
    if a != b:
        bad += 1
        print("DIFF", it, dict(rl=rl, n=n, paired=paired, kw=kw, opts=opts_dbg, flag=hex(flag), img=os.path.basename(img)), first_diff(a, b, blob, paired))
print("iterations", it, "bad", bad, "seed", seed0, which)
sys.exit(1 if bad else 0)
