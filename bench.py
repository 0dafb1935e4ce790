#!/usr/bin/env python
"""bench.py -- throughput of the hot path (jnibwa_createAlignments: a batch of reads in -> BwaMemAlignment records out).

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c2|c3|c4|c5]` prints ONE JSON line
(rank 0).  A "step" is one createAlignments call over one batch of synthetic reads (a batch = one pestat domain,
exactly like one BwaMemAligner.alignSeqs call).  Under torchrun every rank owns one GPU with its own index replica
and its own stream of batches (shard-by-call, SURVEY.md 8e: no data-path collective); `value` is the whole-job
aggregate.  The default workload is BASELINE.json configs[2] (= C3: 3 Gbp genome, 2x150 bp pairs, 500 000 pairs per
call); --config selects the other BASELINE configurations.

  value   device time of the pipeline with the batch already resident in HBM (CUDA events on the context's stream)
  e2e     the same batches through the reference-facing C-ABI with HOST buffers (blob in, malloc'd record stream
          out), wall clock, host<->device copies inside the timed region, from 1 and from --contexts (3) caller threads (the
          reference's usage: one shared BwaMemIndex, one BwaMemAligner per thread); under torchrun it also includes
          the host-side ordered gather of every rank's record stream into one shared host buffer
  parity  every distinct batch that was timed, re-run (untimed) by the CPU oracle and byte-compared; a mismatch
          makes the run exit non-zero
  roofline / stages / kernels   per-stage device times of the last timed step against the bound of each stage
  cpu_baseline   the CPU oracle (restatement of bwa-mem @cb95061 semantics; the real bwa is not vendored in the
                 reference and cannot be built offline) on one full batch, all host cores

`--impl reference` times that CPU restatement alone (all host threads) on the same config.
`--single-process` (with --gpus N, no torchrun) drives N GPUs from ONE process through the library's device pool
(BWA_B200_DEVICES), 2N caller threads.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

# BASELINE.json configs[1..4] as SURVEY.md 8d defines them
CONFIGS = {
    "c2": dict(genome_bp=4_600_000, gseed=1, repeats=False, paired=False, read_len=150, units=1_000_000, reads={},
               what="4.6 Mbp single-contig genome (index L2-resident), 150 bp single-end reads (0.1% mutations, 1% base error)"),
    "c3": dict(genome_bp=3_000_000_000, gseed=2, repeats=True, paired=True, read_len=150, units=500_000, reads={},
               what="3 Gbp genome (24 contigs, 5% interspersed repeats, 1% tandem, segdups), 2x150 bp wgsim-style pairs (insert N(400,50), 0.1% mutations, 1% base error)"),
    "c4": dict(genome_bp=3_000_000_000, gseed=2, repeats=True, paired=True, read_len=250, units=500_000,
               reads=dict(ins_mean=550.0, ins_std=80.0, err_rate=0.05, mut_rate=0.02, indel_frac=0.5),
               what="3 Gbp genome, 2x250 bp pairs at 5% base error + 1% indels (insert N(550,80)): stresses the ksw_extend2 band and z-drop"),
    "c5": dict(genome_bp=3_000_000_000, gseed=2, repeats=True, paired=True, read_len=150, units=500_000,
               reads=dict(chimeric_frac=0.10, repeat_read_frac=0.20),
               what="3 Gbp genome, 2x150 bp pairs with 10% chimeric reads and 20% of reads from repeat families (supplementary + XA)"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=os.environ.get("B200_BENCH_CONFIG", "c3"), choices=sorted(CONFIGS))
    ap.add_argument("--genome-mbp", type=float, default=None, help="override the config's genome size (development)")
    ap.add_argument("--units", type=int, default=None, help="reads (SE) or pairs (PE) per createAlignments call per GPU (default: the config's)")
    ap.add_argument("--ref-sample-units", type=int, default=100_000, help="--impl reference: units per step (a bounded sample of the call)")
    ap.add_argument("--parity-seconds", type=float, default=150.0, help="oracle wall-time budget of the parity check (at least one full batch is always compared)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle (no cpu_baseline, no parity): development only")
    ap.add_argument("--single-process", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="HBM-resident leg only (no e2e legs, no oracle): what the ncu captures under profiles/ run")
    ap.add_argument("--contexts", type=int, default=3, help="batches in flight per device (BWA_B200_CONTEXTS) = caller threads of the multi-caller e2e leg")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu: int):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0])); self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def contig_lengths(total: int):
    """24 contigs of 50-250 Mbp at human scale (deterministic spread); 4 equal contigs below 1 Gbp; one below 8 Mbp."""
    if total < 1_000_000_000:
        n_ctg = 4 if total >= 8_000_000 else 1
        lens = [total // n_ctg] * n_ctg
        lens[-1] += total - sum(lens)
        return lens
    w = [50 + (i * 37) % 201 for i in range(24)]
    lens = [int(total * x / sum(w)) for x in w]
    lens[-1] += total - sum(lens)
    return lens


def genome_of(cfg):
    import common
    lens = contig_lengths(cfg["genome_bp"])
    if cfg["repeats"]:
        return common.simdata.make_genome(cfg["gseed"], lens, repeat_frac=0.05, tandem_frac=0.01, n_segdup=50 if cfg["genome_bp"] >= 50_000_000 else 5)
    return common.simdata.make_genome(cfg["gseed"], lens)


def image_path(cfg):
    import common
    return os.path.join(common.workdir(), f"bench_g{cfg['genome_bp']}_s{cfg['gseed']}{'r' if cfg['repeats'] else 'u'}.img")


def build_image_here(cfg, g, use_gpu: bool):
    """Builds the index image of the genome into tests/_work (not timed).  use_gpu: csrc/sa_build.cu through the product
    library (3 Gbp in ~15 s); else the host suffix sorter."""
    import common
    img = image_path(cfg)
    if os.path.exists(img):
        return img
    t = time.time()
    dll = common.engine_lib().dll if use_gpu else None
    common.simdata.build_image(g, img + ".tmp", dll=dll)
    os.replace(img + ".tmp", img)
    sys.stderr.write(f"[bench] built the index image for {cfg['genome_bp']} bp in {time.time() - t:.1f}s ({'device' if dll else 'host'} suffix sorter)\n")
    return img


def workload(cfg, rank: int, in_subprocess_on_gpu: bool = False):
    """Seeded synthetic genome + index image.  The image is built once by rank 0 and shared through tests/_work; none of
    this is timed.  in_subprocess_on_gpu: build the image in a CHILD process (the reference arm must not map the
    product library into its own process, yet a 3 Gbp index needs the GPU sorter to be ready within the run)."""
    t = time.time()
    g = genome_of(cfg)
    if rank == 0:
        sys.stderr.write(f"[bench] generated the {cfg['genome_bp']} bp genome in {time.time() - t:.1f}s\n")
    img = image_path(cfg)
    if rank == 0 and not os.path.exists(img):
        big = cfg["genome_bp"] >= 8_000_000
        if in_subprocess_on_gpu and big:
            code = ("import sys, json; sys.path.insert(0, %r); import bench, torch; cfg = json.loads(%r); "
                    "g = bench.genome_of(cfg); bench.build_image_here(cfg, g, torch.cuda.is_available())") % (ROOT, json.dumps(cfg))
            subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
        else:
            use_gpu = False
            if big:
                try:
                    import torch
                    use_gpu = torch.cuda.is_available()
                except Exception:
                    use_gpu = False
            build_image_here(cfg, g, use_gpu)
    return g, img


def make_batch(g, cfg, seed: int, units: int) -> bytes:
    import common
    return bytes(common.simdata.make_reads(g, seed, units, cfg["read_len"], cfg["paired"], **cfg["reads"])[0])


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def allreduce(x: float, world, op: str) -> float:
    if world <= 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


ORC_FIELDS = ("n_occ", "n_sa", "n_extend", "n_sa_lookup", "cells_ext", "cells_msw", "cells_glb", "n_ext_calls", "n_msw_calls",
              "n_glb_calls", "n_seeds", "n_chains", "n_regs", "n_occ_sa", "n_ext_retry", "n_ext_zdrop")


def oracle_counters(lib):
    class C(ctypes.Structure):
        _fields_ = [(n, ctypes.c_uint64) for n in ORC_FIELDS]
    c = C()
    lib.dll.orc_get_counters(ctypes.byref(c), 1)
    return {n: int(getattr(c, n)) for n in ORC_FIELDS}


def run_oracle(img, blob, flag, threads):
    """One createAlignments call of the CPU oracle with `threads` host threads; returns (seconds, counters, bytes)."""
    import common
    lib = common.oracle_lib()
    idx = common.bwa.BwaMemIndex(img, lib=lib)
    al = common.bwa.BwaMemAligner(idx)
    al.setFlagOption(flag)
    al.setNThreadsOption(threads)
    oracle_counters(lib)
    t = time.time()
    out = idx.doAlignment(bytearray(blob), al._opts_addr, None)
    dt = time.time() - t
    cnt = oracle_counters(lib)
    al.close(); idx.close()
    return dt, cnt, out


def record_fractions(out: bytes, n_reads: int, sample_reads: int = 100_000):
    """fraction of records carrying XA alternatives / flagged supplementary, over the first reads."""
    import common
    import numpy as np
    n = min(n_reads, sample_reads)
    # the stream has no index: decode a prefix that certainly holds n reads (<= 64 int32 per read is far above any mean)
    ints = np.frombuffer(out, dtype="<i4")
    recs = None
    for cap in (len(ints),):
        try:
            recs = common.bwa.decode_alignments(out, n_reads) if n == n_reads else None
        except Exception:
            recs = None
    if recs is None:
        recs = decode_prefix(out, n)
    n_rec = sum(len(a) for a in recs)
    n_xa = sum(1 for a in recs for r in a if r.xaTag)
    n_sup = sum(1 for a in recs for r in a if r.samFlag & 0x800)
    n_unm = sum(1 for a in recs for r in a if r.samFlag & 0x4)
    return {"reads_sampled": n, "records": n_rec, "frac_records_with_XA": n_xa / max(1, n_rec), "frac_records_supplementary": n_sup / max(1, n_rec),
            "frac_records_unmapped": n_unm / max(1, n_rec)}


def decode_prefix(out: bytes, n: int):
    """Records of the first n reads of a stream (walks the int32 layout of ...BwaMemIndex.c:115-141 to find the prefix)."""
    import common
    import struct
    p, k = 0, 0
    ints = memoryview(out).cast("i")
    while k < n:
        na = ints[p]; p += 1
        for _ in range(na):
            fm = ints[p]; p += 1
            flag = (fm >> 16) & 0xffff
            if not (flag & 0x4):
                p += 5
                nc = ints[p]; p += 1 + nc
                md = ints[p]; p += 1 + ((md + 3) >> 2)
                xa = ints[p]; p += 1 + ((xa + 3) >> 2)
            if (flag & 0x9) == 1:
                p += 3
        k += 1
    return common.bwa.decode_alignments(bytes(out[:4 * p]), n)


def main():
    args = parse_args()
    rank, world, local = dist_env()
    cores = os.cpu_count() or 1
    cfg = dict(CONFIGS[args.config])
    if args.genome_mbp is not None:
        cfg["genome_bp"] = int(args.genome_mbp * 1e6)
    units = args.units or int(os.environ.get("B200_BENCH_UNITS", "0")) or cfg["units"]
    paired = cfg["paired"]
    flag = 2 if paired else 0
    unit_name = "pairs" if paired else "reads"
    metric = "aligned read-pairs/sec" if paired else "aligned reads/sec"
    n_ctg = len(contig_lengths(cfg["genome_bp"]))
    config = {"workload": f"{args.config.upper()}: {cfg['what']}; {units} {unit_name} per createAlignments call per GPU, default mem_opt_t{' + MEM_F_PE, pestat inferred per call' if paired else ''}",
              "config": args.config, "genome_bp": cfg["genome_bp"], "contigs": n_ctg, f"{unit_name}_per_step_per_gpu": units, "read_len": cfg["read_len"],
              "parallelism": f"index replica per GPU, shard-by-call x{world}", "cache": "index + batch exceed L2; a different batch every step" if cfg["genome_bp"] > 100_000_000 else "index fits L2 (by design of this config); a different batch (150 MB) every step",
              "index": "BwaMemIndex image of the genome built before the timed region (device suffix sorter csrc/sa_build.cu for >= 8 Mbp)"}

    # ---------------- reference arm: the CPU restatement on all host cores -----------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        import common
        g, img = workload(cfg, 0, in_subprocess_on_gpu=True)
        n = min(units, args.ref_sample_units)
        config[f"{unit_name}_per_step"] = n
        config["sample"] = f"each step is one oracle call over {n} {unit_name} of the workload (the GPU arm's calls hold {units})"
        times = []
        for s in range(args.warmup + args.steps):
            blob = make_batch(g, cfg, 1000 + s, n)
            dt, _, _ = run_oracle(img, blob, flag, cores)
            if s >= args.warmup:
                times.append(dt)
        tot = sum(times)
        val = n * len(times) / tot
        print(json.dumps({"metric": metric, "value": val, "unit": f"{unit_name}/s", "impl": "reference", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": f"{unit_name}/s", "cores": cores, "kind": "port",
                                           "sample": f"{n} {unit_name} per step of the same workload (CPU restatement of bwa-mem @cb95061 semantics -- not bwa itself -- gcc -O2, SSE2 striped ksw_align2 like bwa's, scalar ksw_extend2 / ksw_global2, {cores} threads with work stealing over reads)"},
                          "e2e": {"value": val, "unit": f"{unit_name}/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---------------- B200 arm ------------------------------------------------------------------------------
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    single = args.single_process and world == 1 and args.gpus > 1
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["BWA_B200_CONTEXTS"] = str(args.contexts)
    if single:
        os.environ["BWA_B200_DEVICES"] = ",".join(str(d) for d in range(args.gpus))
        config["parallelism"] = f"ONE process, in-library device pool of {args.gpus} GPUs (BWA_B200_DEVICES), {args.contexts * args.gpus} caller threads"
    import common
    eng = common.engine_lib()
    assert eng.dll.b200_setDevice(local) == 0
    g, img = workload(cfg, rank)
    barrier(world)
    idx = common.bwa.BwaMemIndex(img, lib=eng)
    al = common.bwa.BwaMemAligner(idx)
    if paired:
        al.alignPairs()
    dll = eng.dll
    dll.b200_alignStaged.restype = ctypes.c_longlong
    dll.b200_alignStaged.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    dll.b200_stageBatch.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    dll.b200_lastBatchMs.restype = ctypes.c_float
    dll.b200_lastBatchMs.argtypes = [ctypes.c_void_p]
    dll.b200_lastStageTimes.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    dll.b200_lastCounters.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    dll.jnibwa_createAlignments.restype = ctypes.c_void_p
    dll.jnibwa_createAlignments.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
    dll.jnibwa_free.argtypes = [ctypes.c_void_p]
    h = ctypes.c_void_p(idx._addr)

    n_steps = args.warmup + args.steps
    n_callers = args.contexts * args.gpus if single else args.contexts
    n_distinct = min(n_steps, 4) if not single else min(max(n_steps, n_callers), 2 * n_callers)
    blobs = [make_batch(g, cfg, 1000 + rank * 100 + s, units) for s in range(n_distinct)]
    h2d = len(blobs[0])
    host_blobs = [ctypes.create_string_buffer(b, len(b)) for b in blobs]

    # ---- HBM-resident leg: stage (untimed), run (timed by CUDA events on the context's stream)
    sampler = ClockSampler(local) if rank == 0 else None
    dev_ms, out_bytes = [], 0
    stages, cnt = [], (ctypes.c_uint64 * 8)()
    if not single:
        for s in range(n_steps):
            if s == args.warmup:
                barrier(world); torch.cuda.synchronize()
                if sampler:
                    sampler.start()
            assert dll.b200_stageBatch(h, blobs[s % n_distinct]) == 0
            sz = dll.b200_alignStaged(h, al._opts_addr, None)
            assert sz > 0, "alignStaged failed"
            if s >= args.warmup:
                dev_ms.append(float(dll.b200_lastBatchMs(h)))
                out_bytes = int(sz)
        torch.cuda.synchronize(); barrier(world)
        ms = (ctypes.c_float * 64)(); names = (ctypes.c_char_p * 64)()
        ns = dll.b200_lastStageTimes(h, ms, names, 64)
        stages = [(names[i].decode(), float(ms[i])) for i in range(ns)]
        dll.b200_lastCounters(h, cnt)
        t_dev = allreduce(sum(dev_ms) / 1e3, world, "max")
        value = world * units * args.steps / t_dev
    elif sampler:
        sampler.start()

    # ---- end-to-end leg: host blob in -> host record stream out through jnibwa_createAlignments
    # (the C-ABI call itself, as the JNI shim makes it: the read blob sits in a host buffer owned by the caller -- the
    # direct ByteBuffer of BwaMemAligner.java:203-210 -- and the result is a malloc'd host buffer the caller frees).
    # Under torchrun the record stream of every call is also copied into this rank's slot of ONE shared host buffer, in
    # (call, rank) order: the host-side ordered gather of north_star, inside the timed region.
    gather = None
    if world > 1:
        from multiprocessing import shared_memory
        slot = int(out_bytes * 1.25) + (1 << 20)
        name = f"b200_bench_gather_{os.environ.get('MASTER_PORT', '0')}"
        if rank == 0:
            try:
                shared_memory.SharedMemory(name=name).unlink()
            except Exception:
                pass
            shm = shared_memory.SharedMemory(name=name, create=True, size=slot * world * 2)
        barrier(world)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=name)
        gather = (shm, slot, ctypes.addressof(ctypes.c_char.from_buffer(shm.buf)))
    lock = threading.Lock()

    def e2e_leg(n_threads: int, steps: int, warm_per_thread: int):
        """`steps` timed calls handed out to n_threads caller threads (after warm_per_thread untimed calls each); returns seconds."""
        cursor = {"next": 0, "sum": 0, "bytes": 0, "fail": None}
        start_evt = threading.Event()
        ready = threading.Barrier(n_threads + 1)

        def call(i):
            size = ctypes.c_size_t(0)
            res = dll.jnibwa_createAlignments(h, al._opts_addr, None, ctypes.cast(host_blobs[i % n_distinct], ctypes.c_void_p), ctypes.byref(size))
            if not res or size.value == 0:
                cursor["fail"] = "createAlignments failed"
                return None, 0
            return res, size.value

        def worker(t):
            for k in range(warm_per_thread):
                res, sz = call(t + k)
                if res:
                    dll.jnibwa_free(res)
            ready.wait()
            start_evt.wait()
            while True:
                with lock:
                    i = cursor["next"]
                    if i >= steps:
                        break
                    cursor["next"] = i + 1
                res, sz = call(i)
                if not res:
                    break
                last = ctypes.c_int32.from_address(res + sz - 4).value  # the result is host memory: touch its last word
                if gather is not None:  # ordered gather: slot of (call parity, rank) in the shared host buffer
                    shm, slot, base = gather
                    ctypes.memmove(base + ((i & 1) * world + rank) * slot, res, min(sz, slot))
                dll.jnibwa_free(res)
                with lock:
                    cursor["sum"] ^= last; cursor["bytes"] = sz
        ths = [threading.Thread(target=worker, args=(t,), daemon=True) for t in range(n_threads)]  # (daemon: a failure in the main thread must end the process)
        for t in ths:
            t.start()
        ready.wait()
        barrier(world); torch.cuda.synchronize()
        t0 = time.perf_counter()
        start_evt.set()
        for t in ths:
            t.join()
        for d in range(args.gpus if single else 1):
            torch.cuda.synchronize(d if single else local)
        dt = time.perf_counter() - t0
        assert cursor["fail"] is None, cursor["fail"]
        return allreduce(dt, world, "max"), cursor["bytes"]

    if args.device_only:
        e2e1 = e2e2 = 0.0
        args.no_cpu_baseline = True
    elif single:
        k_steps = max(args.steps, n_callers) * 2
        t_pool, ob = e2e_leg(n_callers, k_steps, max(2, args.warmup))
        out_bytes = ob
        e2e1 = e2e2 = units * k_steps / t_pool
        value = e2e2
        t_dev = t_pool * args.steps / k_steps
    else:
        t1, _ = e2e_leg(1, args.steps, args.warmup)
        t2, _ = e2e_leg(n_callers, n_callers * args.steps, max(2, args.warmup))
        e2e1 = world * units * args.steps / t1
        e2e2 = world * units * n_callers * args.steps / t2
    clocks = sampler.stop() if sampler else None
    if gather is not None:
        barrier(world)
        gather[0].close()
        if rank == 0:
            gather[0].unlink()

    if rank != 0:
        import torch.distributed as dist
        dist.destroy_process_group()
        return

    # ---- parity: every distinct timed batch against the CPU oracle (untimed), cpu_baseline from the first one
    cpu = None
    work = None
    parity = None
    extras = {}
    if not args.no_cpu_baseline:
        spent, n_ok, n_cmp, first_cnt = 0.0, 0, 0, None
        for b in range(n_distinct):
            if b > 0 and spent + (spent / b) > args.parity_seconds:
                break
            dt, oc, ref_out = run_oracle(img, blobs[b], flag, cores)
            spent += dt
            # the engine's record stream for this batch through the same C-ABI call the e2e legs timed (re-run untimed: copying
            # 65 MB results out inside the timed loop would hold the interpreter lock against the other caller threads)
            got = idx.doAlignment(bytearray(blobs[b]), al._opts_addr, None)
            n_cmp += 1
            n_ok += int(got == ref_out)
            if b == 0:
                first_cnt = oc
                cpu = {"value": units / dt, "unit": f"{unit_name}/s", "cores": cores, "kind": "port",
                       "sample": f"one full batch ({units} {unit_name}) of the timed workload (CPU restatement of bwa-mem @cb95061 semantics -- not bwa itself -- gcc -O2, SSE2 striped ksw_align2 like bwa's, scalar ksw_extend2 / ksw_global2, {cores} threads with work stealing over reads)"}
                if args.config == "c5" or args.config == "c3":
                    extras["records"] = record_fractions(got, units * (2 if paired else 1))
        parity = {"batches_compared": n_cmp, "batches_timed_distinct": n_distinct, f"{unit_name}_compared": n_cmp * units, "identical": n_ok == n_cmp,
                  "oracle_seconds": round(spent, 1)}
        work = first_cnt
        if first_cnt and first_cnt["n_ext_calls"]:
            extras["extend"] = {"ksw_extend2_calls_per_read": first_cnt["n_ext_calls"] / (units * (2 if paired else 1)),
                                "band_retry_rate": first_cnt["n_ext_retry"] / first_cnt["n_ext_calls"],
                                "zdrop_exit_rate": first_cnt["n_ext_zdrop"] / first_cnt["n_ext_calls"]}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6545.9))
    dpx = {}
    try:
        dpx = json.load(open(os.path.join(ROOT, "profiles", "dpx_peak.json")))
    except Exception:
        pass
    # Per-kernel table.  Algorithmic work comes from the oracle's counters on the same batch: FM-index
    # kernels in bytes (64 B per occurrence-block touch, 8 B per suffix-array sample read), DP kernels in cell updates.
    # Kernels that split one stage (thread-level / task / warp variants) share its work.
    groups = {"seed": ["body_seed_q", "body_seed12_q", "body_seed3_q", "body_seedF1_q", "body_seedF2_q", "body_seedB_q", "body_seedB2_q", "wbody_seedB_heavy", "body_seed12_fb_q"],
              "sal": ["body_sal"], "chain": ["body_chain", "wbody_chain_heavy", "body_chain_classify"],
              "extend": ["body_extend_q", "body_extend_q8", "wbody_heavy_prep", "body_extend_task_q", "body_extend_task_q8", "wbody_extend_heavy", "wbody_extend",
                         "body_ext_prep", "body_ext_left_q", "body_ext_right_q", "body_ext_left_q8", "body_ext_right_q8", "body_ext_replay"],
              "rescue_sw": ["wbody_rescue_sw", "body_tsw_fwd_q", "body_tsw_rev_q", "wbody_rescue_sw_fb"],
              "finalize": ["body_final_pe_t", "wbody_final_pe_redo", "body_final_se_t", "wbody_final_se_redo", "body_final_pe_collect", "body_final_se_collect",
                           "body_galign_task_q", "wbody_galign_fb", "body_final_pe_consume", "body_final_se_consume"]}
    # kernels launched on the context's auxiliary stream run beside a kernel of the same stage on the main stream: their entry
    # in the stage table is only the part their neighbour did not cover (possibly 0), but their DRAM traffic counts in full
    aux_kernels = {"seed": ["wbody_seedB_heavy"], "chain": ["wbody_chain_heavy"]}
    ms_of = {gr: sum(t for nm, t in stages if nm in names_) for gr, names_ in groups.items()}
    kern = []
    tot_ms = sum(t for _, t in stages) or 1.0
    for nm, t in stages:
        kern.append({"kernel": nm, "ms": round(t, 3), "share": round(t / tot_ms, 4)})
    stage_work = None
    if work and stages:
        stage_work = {
            "seed": {"algorithmic_bytes": 64.0 * (work["n_occ"] - work["n_occ_sa"]), "ms": round(ms_of["seed"], 3)},
            "sal": {"algorithmic_bytes": 64.0 * work["n_occ_sa"] + 8.0 * work["n_sa"], "ms": round(ms_of["sal"], 3)},
            "extend": {"cells": work["cells_ext"], "ms": round(ms_of["extend"], 3)},
            "rescue_sw": {"cells": work["cells_msw"], "ms": round(ms_of["rescue_sw"], 3)},
            "finalize": {"cells": work["cells_glb"], "ms": round(ms_of["finalize"], 3)},
        }
        # DP roof (tools/dpx_peak.cu -> profiles/dpx_peak.json): the min/max pipe issues R three-input min/max
        # instructions per second; a cell of the affine-gap recurrence needs DPX_PER_CELL of them in the scalar (one
        # cell per instruction) formulation these kernels use: peak cells/s = R / DPX_PER_CELL.
        # (rescue_sw runs the packed s16x2 forms, two cells per instruction: 4 DPX + 1 PRMT-class instructions per PAIR of cells)
        dpx_per_cell = {"extend": 4.0, "rescue_sw": 2.0, "finalize": 5.0}
        r32 = min(dpx.get("vimax3_s32", 0.0), dpx.get("viaddmax_s32_relu", 0.0), dpx.get("vimax3_s16x2", 1e30), dpx.get("viaddmax_s16x2_relu", 1e30)) if dpx else 0.0
        for k_, v in stage_work.items():
            if v["ms"] > 0:
                if "cells" in v:
                    v["gcups"] = round(v["cells"] / (v["ms"] * 1e6), 2)
                    if r32 > 0:
                        v["dpx_peak_gcups"] = round(r32 / dpx_per_cell[k_] / 1e9, 1)
                        v["frac_dpx"] = round(v["gcups"] / v["dpx_peak_gcups"], 4)
                else:
                    v["gbs"] = round(v["algorithmic_bytes"] / (v["ms"] * 1e6), 1)
                    v["frac_hbm"] = round(v["gbs"] / hbm_peak, 4)
                    if k_ == "sal":  # the reference walks to a sample every 32 rows, the device keeps one every 4 (BWA_B200_SA_INTV):
                        v["frac_hbm"] = None  # ~8x fewer steps than the algorithmic count, so this rate is not a bandwidth
                        v["note"] = "algorithmic bytes are the reference's (SA sample every 32 rows); the device keeps a sample every 4 rows and walks ~8x fewer LF steps"
    # The roofline object describes the dominant STAGE: a stage is one algorithmic step of the path and may be split over
    # several launches (SMEM seeding = forward scans, backward sweeps, round 3: profiles/*_launch_summary.txt).
    roof = None
    if stages:
        dom_group = max(ms_of, key=lambda gr: ms_of[gr])
        dom_ms = ms_of[dom_group]
        dom_kernels = [nm for nm, t in stages if nm in groups[dom_group] and t > 0]
        roof = {"bound": "hbm", "kernel": "+".join(dom_kernels), "stage": dom_group, "ms": round(dom_ms, 3), "achieved": None, "peak": hbm_peak, "unit": "GB/s",
                "frac": None, "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback"}
        if stage_work and dom_group in ("seed", "sal"):
            # the FM-index walks: random 64-byte occurrence-block reads of an index far larger than L2
            roof["achieved"] = stage_work[dom_group]["algorithmic_bytes"] / (dom_ms * 1e6)
            roof["algorithmic_bytes_per_step"] = stage_work[dom_group]["algorithmic_bytes"]
            roof["note"] = "achieved = 64 B x occurrence-block touches of the SMEM search (oracle counters on the same batch; the reference's 64-byte block per touch -- the device reads a 32-byte sector per touch and a prefix table for the first 11 extensions of every forward scan) / summed kernel time of the stage's launches (CUDA events on the context's stream)"
        elif stage_work and dom_group in ("extend", "rescue_sw", "finalize"):
            cells = stage_work[dom_group]["cells"]
            roof["achieved"] = cells * 8.0 / (dom_ms * 1e6)
            roof["gcups"] = cells / (dom_ms * 1e6)
            roof["frac_dpx"] = stage_work[dom_group].get("frac_dpx")
            roof["note"] = "dominant stage is integer DP out of shared memory / registers (no HBM or tensor roof applies; see stages.*.frac_dpx for the min/max-pipe roof): achieved = cell updates x 8 B of H/E state / time"
        if roof["achieved"] is not None:
            roof["frac"] = roof["achieved"] / hbm_peak
            if dom_group in ("seed", "sal"):
                try:  # the roof that applies to 64-byte random reads, measured with tools/hbm_random.cu (profiles/hbm_random.json)
                    rr = json.load(open(os.path.join(ROOT, "profiles", "hbm_random.json")))["roof_for_the_fm_index_stage_gbs"]
                    roof["peak_random64"] = rr
                    roof["frac_random64"] = roof["achieved"] / rr
                except Exception:
                    pass
        try:  # DRAM bytes of the stage's launches from the committed ncu --set full captures of this very workload, if they exist
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ents = [prof.get(k) for k in dom_kernels + [k for k in aux_kernels.get(dom_group, []) if k not in dom_kernels]]
            if all(e and e.get("units_per_step", e.get("pairs_per_step")) == units and e.get("genome_bp") == cfg["genome_bp"] and e.get("config", "c3") == args.config for e in ents):
                roof["traffic"] = sum(e["dram_bytes_per_launch"] for e in ents)
                roof["traffic_source"] = sorted(set(e.get("source") for e in ents))
        except Exception:
            pass
    e2e_obj = {"value": max(e2e1, e2e2), "unit": f"{unit_name}/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_bytes,
               "caller_threads": n_callers if (single or e2e2 >= e2e1) else 1, "value_1_thread": e2e1, f"value_{n_callers}_threads": e2e2,
               "note": "concurrent createAlignments calls on one shared handle (BwaMemAligner.java:18) overlap upload, kernels and result copy of consecutive batches"
                       + ("; includes the host-side ordered gather into one shared host buffer" if world > 1 else "")}
    line = {"metric": metric, "value": value, "unit": f"{unit_name}/s", "n_gpus": args.gpus if single else world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e_obj,
            "gpu_launches": int(cnt[7]) * args.steps if not single else None, "roofline": roof, "stages": stage_work, "kernels": kern, "cpu_baseline": cpu,
            "parity": parity, "extras": extras}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if parity is not None and not parity["identical"]:
        sys.stderr.write("[bench] PARITY MISMATCH against the CPU oracle\n")
        sys.exit(3)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:  # noqa: BLE001 -- a failed run must END (helper threads may be blocked inside the library), never hang the box
        import traceback
        traceback.print_exc()
        sys.stderr.flush(); sys.stdout.flush()
        os._exit(1)
