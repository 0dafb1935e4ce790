"""Synthetic genomes / reads and index images for tests and bench.py (ctypes over tools/libbwasim.so and the
product library's index builder).  Everything is seeded and generated offline; nothing here is on the timed path."""
from __future__ import annotations

import ctypes
import os
import struct
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gatk-bwamem-jni_b200")
_SIM_SO = os.path.join(ROOT, "tools", "libbwasim.so")
_IDX_SO = os.path.join(PKG, "lib", "libbwaidx_b200.so")


def build_tools(force: bool = False) -> None:
    """g++ builds of the simulator and of the host-only index builder (no CUDA needed for either)."""
    os.makedirs(os.path.dirname(_IDX_SO), exist_ok=True)
    src = os.path.join(ROOT, "tools", "simulate.cpp")
    if force or not os.path.exists(_SIM_SO) or os.path.getmtime(_SIM_SO) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _SIM_SO, src])
    src = os.path.join(PKG, "csrc", "index_build.cpp")
    hdr = os.path.join(PKG, "csrc", "bwa_image.h")
    newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
    if force or not os.path.exists(_IDX_SO) or os.path.getmtime(_IDX_SO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _IDX_SO, src, "-lz", "-lpthread"])


_sim = None
_idx = None


def _simlib():
    global _sim
    if _sim is None:
        build_tools()
        _sim = ctypes.CDLL(_SIM_SO)
        _sim.sim_genome.restype = ctypes.c_int64
        _sim.sim_genome.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _sim.sim_reads.restype = ctypes.c_int64
        _sim.sim_reads.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_void_p]
    return _sim


def idxlib():
    """The host-only index builder (same sources as the builder inside libbwa_b200.so)."""
    global _idx
    if _idx is None:
        build_tools()
        _idx = ctypes.CDLL(_IDX_SO)
        _idx.b200_buildImageFromBases.restype = ctypes.c_int
        _idx.b200_buildImageFromBases.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_char_p, ctypes.c_char_p]
        _idx.jnibwa_createIndexFile.restype = ctypes.c_int
        _idx.jnibwa_createIndexFile.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        _idx.jnibwa_createReferenceIndex.restype = ctypes.c_int
        _idx.jnibwa_createReferenceIndex.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
    return _idx


class Genome:
    def __init__(self, bases: np.ndarray, lens: Sequence[int], names: Sequence[str], repeats: np.ndarray):
        self.bases = bases            # uint8 ASCII, contigs concatenated
        self.lens = np.asarray(lens, dtype=np.int64)
        self.names = list(names)
        self.repeats = repeats        # int64 [n,2] global intervals of injected repeat copies

    @property
    def total(self) -> int:
        return int(self.lens.sum())


def make_genome(seed: int, lens: Sequence[int], repeat_frac: float = 0.0, tandem_frac: float = 0.0,
                n_segdup: int = 0, names: Optional[Sequence[str]] = None) -> Genome:
    lens_a = np.asarray(lens, dtype=np.int64)
    total = int(lens_a.sum())
    out = np.empty(total, dtype=np.uint8)
    cap = 1 << 20
    rep = np.zeros((cap, 2), dtype=np.int64)
    n_rep = ctypes.c_int64(0)
    _simlib().sim_genome(seed, len(lens_a), lens_a.ctypes.data, repeat_frac, tandem_frac, n_segdup,
                         out.ctypes.data, rep.ctypes.data, cap, ctypes.byref(n_rep))
    n = min(int(n_rep.value), cap)
    names = list(names) if names else [f"chr{i + 1}" for i in range(len(lens_a))]
    return Genome(out, lens_a, names, rep[:n].copy())


def make_reads(genome: Genome, seed: int, n_units: int, read_len: int, paired: bool, ins_mean: float = 400.0,
               ins_std: float = 50.0, mut_rate: float = 0.001, indel_frac: float = 0.15, indel_ext: float = 0.3,
               err_rate: float = 0.01, chimeric_frac: float = 0.0, repeat_read_frac: float = 0.0,
               n_rate: float = 0.0, want_truth: bool = False) -> Tuple[bytearray, Optional[np.ndarray]]:
    """Returns the wire blob (BwaMemAligner.java:198-209 format) and optionally the truth positions."""
    n_reads = n_units * (2 if paired else 1)
    cap = 4 + n_reads * (read_len + 1)
    out = np.empty(cap, dtype=np.uint8)
    truth = np.zeros(n_reads, dtype=np.int64) if want_truth else None
    rep = genome.repeats if len(genome.repeats) else None
    sz = _simlib().sim_reads(seed, genome.bases.ctypes.data, len(genome.lens), genome.lens.ctypes.data, n_units,
                             read_len, 1 if paired else 0, ins_mean, ins_std, mut_rate, indel_frac, indel_ext,
                             err_rate, chimeric_frac, repeat_read_frac,
                             rep.ctypes.data if rep is not None else None, len(rep) if rep is not None else 0,
                             n_rate, out.ctypes.data, cap, truth.ctypes.data if truth is not None else None)
    if sz < 0:
        raise RuntimeError("read blob capacity too small")
    return bytearray(out[:sz].tobytes()), truth


def split_blob(blob: bytes) -> List[bytes]:
    n = struct.unpack_from("<I", blob, 0)[0]
    parts = bytes(blob[4:]).split(b"\0")
    return parts[:n]


def build_image(genome: Genome, img_path: str, prefix: Optional[str] = None, dll=None) -> None:
    """dll: a loaded library exporting b200_buildImageFromBases (default: the host-only builder; pass the engine's
    library to use the device suffix sorter, csrc/sa_build.cu)."""
    names = b"".join(n.encode() + b"\0" for n in genome.names)
    if dll is not None:
        dll.b200_buildImageFromBases.restype = ctypes.c_int
        dll.b200_buildImageFromBases.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_char_p, ctypes.c_char_p]
    rc = (dll or idxlib()).b200_buildImageFromBases(len(genome.lens), names, genome.lens.ctypes.data, genome.bases.ctypes.data,
                                           img_path.encode(), prefix.encode() if prefix else None)
    if rc != 0:
        raise RuntimeError(f"image build failed rc={rc}")


def write_fasta(genome: Genome, path: str, width: int = 60) -> None:
    with open(path, "wb") as f:
        off = 0
        for name, ln in zip(genome.names, genome.lens):
            f.write(b">" + name.encode() + b"\n")
            seq = genome.bases[off:off + int(ln)].tobytes()
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + b"\n")
            off += int(ln)
