"""Multi-GPU host logic on CPU: world-size-2 gloo.  Each rank runs its share of a list of createAlignments calls
(shard-by-call) and rank 0 gathers the record streams in call order; the gathered list must equal what one process
produces for the same calls -- "identical output regardless of GPU count".  No GPU exists here, so the per-call
engine is the kernel-debug harness (tests/cudasim, the CUDA kernel source compiled for the host: test tooling), and
the expected streams come from the oracle."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from common import oracle_lib, pkg, run_blob, sim_lib, simdata, synth_image

N_CALLS = 5


def _calls():
    g, img = synth_image("g200k", 7, [120000, 60000, 20000], repeat_frac=0.08, tandem_frac=0.01, n_segdup=3)
    blobs = [simdata.make_reads(g, 300 + i, 60 + 10 * i, 150, True)[0] for i in range(N_CALLS)]
    return img, blobs


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img, blobs = _calls()
        mine = pkg.shard.calls_of_rank(len(blobs), rank, world)
        local = {i: run_blob(sim_lib(), img, blobs[i], 2) for i in mine}
        got = pkg.shard.ordered_gather(local, len(blobs))
        if rank == 0:
            with open(out_path, "wb") as f:
                for s in got:
                    f.write(len(s).to_bytes(8, "little") + s)
    finally:
        dist.destroy_process_group()


def test_calls_of_rank_partitions_the_calls():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in pkg.shard.calls_of_rank(11, r, world))
        assert seen == list(range(11))
    with pytest.raises(ValueError):
        pkg.shard.calls_of_rank(4, 2, 2)


def test_two_ranks_gather_in_call_order(tmp_path):
    img, blobs = _calls()
    sim_lib()  # build the harness once, before forking workers
    want = [run_blob(oracle_lib(), img, b, 2) for b in blobs]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.bin")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    raw = open(out, "rb").read()
    got, k = [], 0
    while k < len(raw):
        n = int.from_bytes(raw[k:k + 8], "little")
        got.append(raw[k + 8:k + 8 + n])
        k += 8 + n
    assert got == want
