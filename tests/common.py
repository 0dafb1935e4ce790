"""Shared helpers of the test-suite: library handles (oracle = checker, engine = product, sim = kernel-debug
harness), the rotavirus fixture of the reference (tests/golden/rotavirus, copied from
/root/reference/src/test/resources by tests/golden/make_golden.py) and seeded synthetic datasets."""
from __future__ import annotations

import ctypes
import importlib.util
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "gatk-bwamem-jni_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
WORK = os.path.join(ROOT, "tests", "_work")
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _load_pkg():
    """The package directory name contains '-', so it is loaded by path under the name bwamem_b200."""
    name = "bwamem_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = _load_pkg()
bwa = pkg.bwa
import simdata  # noqa: E402  (tools/)

ORACLE_SO = os.path.join(ROOT, "oracle", "liborc.so")
SIM_SO = os.path.join(ROOT, "tests", "cudasim", "libsim.so")
ENGINE_SO = os.environ.get("LIBBWA_PATH") or os.path.join(PKG_DIR, "lib", "libbwa_b200.so")


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def build_sim():
    src = os.path.join(ROOT, "tests", "cudasim", "sim_engine.cpp")
    csrc = os.path.join(PKG_DIR, "csrc")
    newest = max([os.path.getmtime(src)] + [os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc)])
    if os.path.exists(SIM_SO) and os.path.getmtime(SIM_SO) >= newest:
        return
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-Wno-unknown-pragmas", "-I", csrc,
                           "-o", SIM_SO, src])


_libs = {}


def oracle_lib():
    if "orc" not in _libs:
        build_oracle()
        _libs["orc"] = bwa.NativeLib(ORACLE_SO, prefix="orc_", opt_init="orc_opt_init", free="orc_free")
    return _libs["orc"]


def sim_lib():
    if "sim" not in _libs:
        build_sim()
        _libs["sim"] = bwa.NativeLib(SIM_SO, prefix="sim_", opt_init="sim_opt_init", free="sim_free")
    return _libs["sim"]


def engine_lib():
    if "eng" not in _libs:
        _libs["eng"] = bwa.NativeLib(ENGINE_SO)
    return _libs["eng"]


def workdir():
    os.makedirs(WORK, exist_ok=True)
    return WORK


def fixture_image() -> str:
    """tests/_work/rotavirus.img built from the checked-in bwa index files of the reference fixture."""
    path = os.path.join(workdir(), "rotavirus.img")
    if not os.path.exists(path):
        rc = simdata.idxlib().jnibwa_createIndexFile(os.path.join(GOLDEN, "rotavirus", "ref.fa").encode(), path.encode())
        assert rc == 0
    return path


def synth_image(tag: str, seed: int, lens, **kw):
    """(genome, image path) of a seeded synthetic genome, cached in tests/_work."""
    g = simdata.make_genome(seed, lens, **kw)
    path = os.path.join(workdir(), f"{tag}.img")
    if not os.path.exists(path):
        simdata.build_image(g, path)
    return g, path


def run_blob(lib, img_path: str, blob: bytearray, flag: int = 0, stats=None, setup=None) -> bytes:
    """One createAlignments call through the Java-API mirror; returns the raw int32 stream."""
    idx = bwa.BwaMemIndex(img_path, lib=lib)
    try:
        al = bwa.BwaMemAligner(idx)
        if flag:
            al.setFlagOption(al.getFlagOption() | flag)
        if setup:
            setup(al)
        al.pairEndStats = stats
        out = idx.doAlignment(bytearray(blob), al._opts_addr, al.pairEndStats)
        al.close()
        return out
    finally:
        idx.close()


def first_diff(a: bytes, b: bytes, blob: bytes, paired: bool):
    """Human-readable description of the first read whose records differ."""
    n = int.from_bytes(blob[:4], "little")
    if paired:
        n -= n & 1
    ra = bwa.decode_alignments(a, n) if a is not None else None
    rb = bwa.decode_alignments(b, n)
    reads = simdata.split_blob(bytes(blob))
    for i in range(n):
        if ra[i] != rb[i]:
            return f"read {i} ({reads[i].decode('latin-1')}):\n  A={ra[i]}\n  B={rb[i]}"
    return "streams differ but decoded records are equal"


def repeat_family_image():
    """(genome, image) of a 1.5 Mbp genome carrying 1500 copies (3 % divergence) of one 320-bp element: reads drawn from
    the copies have hundreds of hits, which drives the warp-level chaining, the per-seed extension tasks, the
    warp-cooperative region de-duplication and mate-rescue chains of max_matesw hits with incremental list updates."""
    import numpy as np
    rng = np.random.default_rng(5)
    L = 1_500_000
    codes = rng.integers(0, 4, L).astype(np.uint8)
    cons = rng.integers(0, 4, 320).astype(np.uint8)
    iv = []
    for _ in range(1500):
        p = int(rng.integers(1000, L - 2000))
        c = cons.copy()
        m = rng.random(320) < 0.03
        c[m] = rng.integers(0, 4, int(m.sum()))
        codes[p:p + 320] = c
        iv.append((p, p + 320))
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].copy()
    g = simdata.Genome(bases, [L], ["chrR"], np.array(iv, dtype=np.int64))
    img = os.path.join(workdir(), "repeat_family.img")
    if not os.path.exists(img):
        simdata.build_image(g, img)
    return g, img


def alt_contig_image():
    """(genome, image) of a genome with two primary contigs and three ALT contigs (diverged copies of primary regions,
    named in <prefix>.alt, which jnibwa_createIndexFile reads like bwa_idx_load does: jnibwa.c:127).  Reads drawn from it
    hit both the primary assembly and the ALT contigs: exercises is_alt in chain filtering, the two-round primary
    marking, XA with max_XA_hits_alt and the extra supplementary record of paired output."""
    import numpy as np
    rng = np.random.default_rng(17)
    base = simdata.make_genome(41, [90000, 50000], repeat_frac=0.04, tandem_frac=0.005, n_segdup=2)
    prim = base.bases
    alts, names = [], ["chr1", "chr2"]
    for k, (off, ln, div) in enumerate([(20000, 6000, 0.02), (100000, 4000, 0.03), (60000, 5000, 0.01)]):
        c = prim[off:off + ln].copy()
        m = rng.random(ln) < div
        c[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
        if k == 1:  # one ALT with a 30-base deletion and a 12-base insertion
            c = np.concatenate([c[:1500], c[1530:3000], np.frombuffer(b"ACGTTGCAAGCT", dtype=np.uint8), c[3000:]])
        alts.append(c); names.append(f"alt{k + 1}")
    bases = np.concatenate([prim] + alts)
    lens = [90000, 50000] + [len(a) for a in alts]
    g = simdata.Genome(bases, lens, names, base.repeats)
    img = os.path.join(workdir(), "alt_contigs.img")
    if not os.path.exists(img):
        prefix = os.path.join(workdir(), "alt_contigs.fa")
        simdata.write_fasta(g, prefix)
        assert simdata.idxlib().jnibwa_createReferenceIndex(prefix.encode(), prefix.encode(), b"auto") == 0
        with open(prefix + ".alt", "w") as f:
            f.write("@SQ\tSN:chr1\n")
            for nm in names[2:]:
                f.write(f"{nm}\t0\tchr1\t1\t0\t100M\t*\t0\t0\t*\t*\n")
        assert simdata.idxlib().jnibwa_createIndexFile(prefix.encode(), img.encode()) == 0
    return g, img
