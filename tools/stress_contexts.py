#!/usr/bin/env python3
"""Stress of concurrent createAlignments calls on one shared handle: CONTEXTS contexts (BWA_B200_CONTEXTS), THREADS caller
threads, batches of different sizes per thread so that workspaces are re-grown while other contexts are running.  Every
result must equal the one the same batch gives on an idle handle.  usage: stress_contexts.py [contexts] [threads] [reps]"""
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
contexts = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n_threads = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
os.environ["BWA_B200_CONTEXTS"] = str(contexts)
import common  # noqa: E402
from common import bwa, simdata, engine_lib  # noqa: E402

g, img = common.synth_image("g200k", 7, [120000, 60000, 20000], repeat_frac=0.08, tandem_frac=0.01, n_segdup=3)
sizes = [300, 2500, 900, 4000, 150, 1800, 3200, 600]
blobs = [simdata.make_reads(g, 400 + i, sizes[i % len(sizes)], 150, True, repeat_read_frac=0.2)[0] for i in range(n_threads)]
idx = bwa.BwaMemIndex(img, lib=engine_lib())
al0 = bwa.BwaMemAligner(idx); al0.alignPairs()
want = [idx.doAlignment(bytearray(b), al0._opts_addr, None) for b in blobs]
bad = []


def worker(t):
    al = bwa.BwaMemAligner(idx); al.alignPairs()
    for r in range(reps):
        k = (t + r) % n_threads  # a different batch size on every call of a thread
        out = idx.doAlignment(bytearray(blobs[k]), al._opts_addr, None)
        if out != want[k]:
            bad.append((t, r, k))
    al.close()


ths = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
for t in ths:
    t.start()
for t in ths:
    t.join()
idx.close()
print(f"contexts {contexts} threads {n_threads} reps {reps}: {'OK' if not bad else 'MISMATCH ' + str(bad[:8])}")
sys.exit(1 if bad else 0)
