#!/usr/bin/env python
"""bench.py -- throughput of the hot path (jnibwa_createAlignments: a batch of reads in -> BwaMemAlignment records out).

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line (rank 0).
A "step" is one createAlignments call over one batch of synthetic 2x150 bp read pairs (batch = one pestat domain,
exactly like one BwaMemAligner.alignSeqs call).  Under torchrun every rank owns one GPU with its own index replica
and its own stream of batches (shard-by-call, SURVEY.md 8e: no data-path collective); `value` is the whole-job
aggregate.

  value   device time of the pipeline with the batch already resident in HBM (CUDA events on the engine's stream)
  e2e     the same batches through the reference-facing C-ABI with HOST buffers (blob in, malloc'd record stream
          out), wall clock, host<->device copies inside the timed region
  roofline / kernels   per-stage device times of the last timed step, the dominant kernel against its bound
  cpu_baseline         the CPU oracle (restatement of bwa-mem @cb95061 semantics; the real bwa is not vendored in the
                       reference and cannot be built offline) on a bounded sample, all host cores

`--impl reference` times that CPU restatement alone (all host threads) on the same config.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome-mbp", type=float, default=float(os.environ.get("B200_BENCH_GENOME_MBP", "3000")))
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("B200_BENCH_PAIRS", "250000")),
                    help="read pairs per step (per GPU)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--cpu-sample-pairs", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    """24 contigs of 50-250 Mbp at human scale (deterministic spread); 4 equal contigs below 1 Gbp."""
    if total < 1_000_000_000:
        n_ctg = 4 if total >= 4_000_000 else 1
        return [total // n_ctg] * n_ctg
    w = [50 + (i * 37) % 201 for i in range(24)]
    lens = [int(total * x / sum(w)) for x in w]
    lens[-1] += total - sum(lens)
    return lens


def workload(args, rank: int, world: int):
    """Seeded synthetic genome (contigs, injected repeat families / tandem repeats / segmental duplications)
    + wgsim-style 2x150 bp pairs.  The index image is built once by rank 0 (on the GPU when one is present:
    csrc/sa_build.cu; the host sorter otherwise) and shared through tests/_work; none of this is timed."""
    import common
    total = int(args.genome_mbp * 1e6)
    lens = contig_lengths(total)
    tag = f"bench_g{total}"
    img = os.path.join(common.workdir(), tag + ".img")
    t = time.time()
    g = common.simdata.make_genome(2, lens, repeat_frac=0.05, tandem_frac=0.01, n_segdup=50 if total >= 50_000_000 else 5)
    if rank == 0:
        sys.stderr.write(f"[bench] generated the {total} bp genome in {time.time() - t:.1f}s\n")
    if rank == 0 and not os.path.exists(img):
        t = time.time()
        dll = None
        try:
            import torch
            if torch.cuda.is_available():
                dll = common.engine_lib().dll
        except Exception:
            dll = None
        common.simdata.build_image(g, img + ".tmp", dll=dll)
        os.replace(img + ".tmp", img)
        sys.stderr.write(f"[bench] built the index image for {total} bp in {time.time() - t:.1f}s ({'device' if dll else 'host'} suffix sorter)\n")
    return g, img


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def allreduce_max(x: float, world) -> float:
    if world <= 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_sum(x: float, world) -> float:
    if world <= 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def oracle_counters(lib):
    class C(ctypes.Structure):
        _fields_ = [(n, ctypes.c_uint64) for n in ("n_occ", "n_sa", "n_extend", "n_sa_lookup", "cells_ext", "cells_msw",
                                                    "cells_glb", "n_ext_calls", "n_msw_calls", "n_glb_calls", "n_seeds",
                                                    "n_chains", "n_regs", "n_occ_sa")]
    c = C()
    lib.dll.orc_get_counters(ctypes.byref(c), 1)
    return {n: int(getattr(c, n)) for n, _ in C._fields_}


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


def main():
    args = parse_args()
    rank, world, local = dist_env()
    cores = os.cpu_count() or 1
    cfg = {"workload": f"synthetic {args.genome_mbp:g} Mbp genome ({len(contig_lengths(int(args.genome_mbp * 1e6)))} contigs, 5% interspersed repeats, 1% tandem, segdups), "
                       f"2x{args.read_len} bp wgsim-style pairs (insert N(400,50), 0.1% mutations, 1% base error), "
                       f"{args.pairs} pairs per createAlignments call per GPU, default mem_opt_t + MEM_F_PE, pestat inferred per call",
           "genome_bp": int(args.genome_mbp * 1e6), "pairs_per_step_per_gpu": args.pairs, "read_len": args.read_len,
           "parallelism": f"index replica per GPU, shard-by-call x{world}", "cache": "index + batch exceed L2; a different batch every step",
           "index": "BwaMemIndex image of the genome built on the GPU before the timed region (csrc/sa_build.cu)"}

    # ---------------- reference arm: the CPU restatement on all host cores -----------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        import common
        g, img = workload(args, 0, 1)
        n = min(args.pairs, args.cpu_sample_pairs)
        times = []
        for s in range(args.warmup + args.steps):
            blob, _ = common.simdata.make_reads(g, 1000 + s, n, args.read_len, True)
            dt, _, _ = run_oracle(img, blob, 2, cores)
            if s >= args.warmup:
                times.append(dt)
        tot = sum(times)
        val = n * len(times) / tot
        print(json.dumps({"metric": "aligned read-pairs/sec", "value": val, "unit": "pairs/s", "impl": "reference", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port",
                                           "sample": f"{n} pairs per step of the same workload (CPU restatement of bwa-mem, gcc -O2, {cores} threads)"},
                          "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---------------- B200 arm ------------------------------------------------------------------------------
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import common
    eng = common.engine_lib()
    assert eng.dll.b200_setDevice(local) == 0
    g, img = workload(args, rank, world)
    barrier(world)
    idx = common.bwa.BwaMemIndex(img, lib=eng)
    al = common.bwa.BwaMemAligner(idx)
    al.alignPairs()
    dll = eng.dll
    dll.b200_alignStaged.restype = ctypes.c_longlong
    dll.b200_alignStaged.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    dll.b200_stageBatch.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    dll.b200_lastBatchMs.restype = ctypes.c_float
    dll.b200_lastBatchMs.argtypes = [ctypes.c_void_p]
    dll.b200_lastStageTimes.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    dll.b200_lastCounters.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    h = ctypes.c_void_p(idx._addr)

    n_steps = args.warmup + args.steps
    n_distinct = min(n_steps, 4)
    blobs = [bytes(common.simdata.make_reads(g, 1000 + rank * 100 + s, args.pairs, args.read_len, True)[0]) for s in range(n_distinct)]
    h2d = len(blobs[0])

    # ---- HBM-resident leg: stage (untimed), run (timed by CUDA events on the engine's stream)
    sampler = ClockSampler(local) if rank == 0 else None
    dev_ms, out_bytes = [], 0
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
    clocks = sampler.stop() if sampler else None
    ms = (ctypes.c_float * 32)(); names = (ctypes.c_char_p * 32)()
    ns = dll.b200_lastStageTimes(h, ms, names, 32)
    stages = [(names[i].decode(), float(ms[i])) for i in range(ns)]
    cnt = (ctypes.c_uint64 * 8)()
    dll.b200_lastCounters(h, cnt)
    t_dev = allreduce_max(sum(dev_ms) / 1e3, world)
    value = world * args.pairs * args.steps / t_dev

    # ---- end-to-end leg: host blob in -> host record stream out through jnibwa_createAlignments
    # (the C-ABI call itself, as the JNI shim makes it: the read blob sits in a host buffer owned by the caller -- the
    # direct ByteBuffer of BwaMemAligner.java:203-210 -- and the result is a malloc'd host buffer the caller frees)
    dll.jnibwa_createAlignments.restype = ctypes.c_void_p
    dll.jnibwa_createAlignments.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
    dll.jnibwa_free.argtypes = [ctypes.c_void_p]
    host_blobs = [ctypes.create_string_buffer(b, len(b)) for b in blobs]
    checksum = 0
    for s in range(args.warmup + args.steps):
        if s == args.warmup:
            barrier(world); torch.cuda.synchronize()
            t0 = time.perf_counter()
        size = ctypes.c_size_t(0)
        res = dll.jnibwa_createAlignments(h, al._opts_addr, None, ctypes.cast(host_blobs[s % n_distinct], ctypes.c_void_p), ctypes.byref(size))
        assert res and size.value > 0, "createAlignments failed"
        checksum ^= ctypes.c_int32.from_address(res + size.value - 4).value  # the result is host memory: touch its last word
        dll.jnibwa_free(res)
    torch.cuda.synchronize()
    t_e2e = allreduce_max(time.perf_counter() - t0, world)
    e2e = world * args.pairs * args.steps / t_e2e

    if rank != 0:
        import torch.distributed as dist
        dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel + per-kernel table (algorithmic work from the oracle's counters on a sample)
    n_s = min(args.pairs, args.cpu_sample_pairs)
    sample_blob = blobs[(n_steps - 1) % n_distinct]
    reads = sample_blob[4:].split(b"\0")[:2 * n_s]
    sblob = common.bwa.pack_sequences(reads)
    cpu = None
    work = None
    parity = None
    if not args.no_cpu_baseline:
        dt, oc, ref_out = run_oracle(img, sblob, 2, cores)
        # the oracle's stream for the sample doubles as a full-scale parity check of the engine (same batch, untimed)
        eng_out = idx.doAlignment(bytearray(sblob), al._opts_addr, None)
        parity = {"sample_pairs": n_s, "identical": bool(eng_out == ref_out), "bytes": len(ref_out)}
        scale = args.pairs / n_s
        work = {k: v * scale for k, v in oc.items()}
        cpu = {"value": n_s / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": f"first {n_s} pairs of the last timed batch (CPU restatement of bwa-mem @cb95061 semantics, gcc -O2, {cores} threads)"}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6545.9))
    # Per-kernel table.  Algorithmic work comes from the oracle's counters on the sample, scaled to the batch
    # (DESIGN.md 5): FM-index kernels in bytes (64 B per occurrence-block touch, 8 B per suffix-array sample read),
    # DP kernels in cell updates.  Kernels that split one stage (thread-level / task / warp variants) share its work.
    groups = {"seed": ["body_seed_q", "body_seed12_q", "body_seed3_q", "body_seedF1_q", "body_seedF2_q", "body_seedB_q", "body_seedB2_q", "wbody_seedB_heavy", "body_seed12_fb_q"],
              "sal": ["body_sal"], "chain": ["body_chain", "wbody_chain_heavy"],
              "extend": ["body_extend_q", "body_extend_q8", "wbody_heavy_prep", "body_extend_task_q", "body_extend_task_q8", "wbody_extend_heavy", "wbody_extend"],
              "rescue_sw": ["wbody_rescue_sw"],
              "finalize": ["body_final_pe_t", "wbody_final_pe_redo", "body_final_se_t", "wbody_final_se_redo", "body_final_pe_collect", "body_final_se_collect",
                           "body_galign_task_q", "body_final_pe_consume", "body_final_se_consume"]}
    # kernels launched on the engine's auxiliary stream run beside a kernel of the same stage on the main stream: their entry
    # in the stage table is only the part their neighbour did not cover (possibly 0), but their DRAM traffic counts in full
    aux_kernels = {"seed": ["wbody_seedB_heavy"], "chain": ["wbody_chain_heavy"]}
    ms_of = {g: sum(t for nm, t in stages if nm in names) for g, names in groups.items()}
    kern = []
    tot_ms = sum(t for _, t in stages) or 1.0
    for nm, t in stages:
        kern.append({"kernel": nm, "ms": round(t, 3), "share": round(t / tot_ms, 4)})
    stage_work = None
    if work:
        stage_work = {
            "seed": {"algorithmic_bytes": 64.0 * (work["n_occ"] - work["n_occ_sa"]), "ms": round(ms_of["seed"], 3)},
            "sal": {"algorithmic_bytes": 64.0 * work["n_occ_sa"] + 8.0 * work["n_sa"], "ms": round(ms_of["sal"], 3)},
            "extend": {"cells": work["cells_ext"], "ms": round(ms_of["extend"], 3)},
            "rescue_sw": {"cells": work["cells_msw"], "ms": round(ms_of["rescue_sw"], 3)},
            "finalize": {"cells": work["cells_glb"], "ms": round(ms_of["finalize"], 3)},
        }
        for v in stage_work.values():
            if v["ms"] > 0:
                if "cells" in v:
                    v["gcups"] = round(v["cells"] / (v["ms"] * 1e6), 2)
                else:
                    v["gbs"] = round(v["algorithmic_bytes"] / (v["ms"] * 1e6), 1)
    # The roofline object describes the dominant STAGE: a stage is one algorithmic step of the path and may be split over
    # several launches (SMEM seeding = forward scans, backward sweeps, round 3: profiles/r1z_launch_summary.txt).
    dom_group = max(ms_of, key=lambda g: ms_of[g])
    dom_ms = ms_of[dom_group]
    dom_kernels = [nm for nm, t in stages if nm in groups[dom_group] and t > 0]
    roof = {"bound": "hbm", "kernel": "+".join(dom_kernels), "stage": dom_group, "ms": round(dom_ms, 3), "achieved": None, "peak": hbm_peak, "unit": "GB/s",
            "frac": None, "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback"}
    if stage_work and dom_group in ("seed", "sal"):
        # the FM-index walks: random 64-byte occurrence-block reads of an index far larger than L2
        roof["achieved"] = stage_work[dom_group]["algorithmic_bytes"] / (dom_ms * 1e6)
        roof["algorithmic_bytes_per_step"] = stage_work[dom_group]["algorithmic_bytes"]
        roof["note"] = "achieved = 64 B x occurrence-block touches of the SMEM search (oracle counters on the same reads) / summed kernel time of the stage's launches (CUDA events on the engine's stream)"
    elif stage_work and dom_group in ("extend", "rescue_sw", "finalize"):
        cells = stage_work[dom_group]["cells"]
        roof["achieved"] = cells * 8.0 / (dom_ms * 1e6)
        roof["gcups"] = cells / (dom_ms * 1e6)
        roof["note"] = "dominant stage is integer DP out of shared memory / registers (no HBM or tensor roof applies): achieved = cell updates x 8 B of H/E state / time"
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
        if all(e and e.get("pairs_per_step") == args.pairs and e.get("genome_bp") == int(args.genome_mbp * 1e6) for e in ents):
            roof["traffic"] = sum(e["dram_bytes_per_launch"] for e in ents)
            roof["traffic_source"] = sorted(set(e.get("source") for e in ents))
    except Exception:
        pass
    line = {"metric": "aligned read-pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_bytes},
            "gpu_launches": int(cnt[7]) * args.steps, "roofline": roof, "stages": stage_work, "kernels": kern, "cpu_baseline": cpu, "parity": parity}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
