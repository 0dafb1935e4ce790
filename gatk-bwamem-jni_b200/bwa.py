"""Host-side mirror of the reference's Java API over the C-ABI.

The reference's host language is Java (no JDK in this image), so the classes the
reference exposes are mirrored here 1:1 in Python over ctypes, with the same
names, argument meaning and error behaviour:

  BwaMemIndex          /root/reference/src/main/java/org/broadinstitute/hellbender/utils/bwa/BwaMemIndex.java:1-488
  BwaMemAligner        .../BwaMemAligner.java:20-327   (option offsets :46-137, alignSeqs :192-311)
  BwaMemAlignment      .../BwaMemAlignment.java:8-65
  BwaMemPairEndStats   .../BwaMemPairEndStats.java:15-205

The native side is whatever library object is passed in (`NativeLib`): the CUDA
engine `libbwa_b200.so` (the product), or -- in tests only -- the CPU oracle,
which exports the same entry points under an `orc_` prefix.  No algorithmic work
happens in this file: it packs the byte blob Java packs, calls the C-ABI, and
decodes the int32 record stream exactly like BwaMemAligner.java:215-308.
"""
from __future__ import annotations

import ctypes
import math
import os
import struct
import threading
from dataclasses import dataclass
from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np

__all__ = [
    "NativeLib", "BwaMemIndex", "BwaMemAligner", "BwaMemAlignment", "BwaMemPairEndStats",
    "CouldNotReadImageException", "CouldNotCreateIndexImageException", "CouldNotCreateIndexException",
    "CouldNotReadIndexException", "CouldNotReadReferenceException", "InvalidFileFormatException",
    "InvalidInputException", "decode_alignments", "pack_sequences",
]


# --- exception types (the seven of the reference, .../bwa/*Exception.java) ---------------------
class BwaMemException(RuntimeError):
    pass


class CouldNotReadImageException(BwaMemException):
    pass


class CouldNotCreateIndexImageException(BwaMemException):
    pass


class CouldNotCreateIndexException(BwaMemException):
    pass


class CouldNotReadIndexException(BwaMemException):
    pass


class CouldNotReadReferenceException(BwaMemException):
    pass


class InvalidFileFormatException(BwaMemException):
    pass


class InvalidInputException(BwaMemException):
    pass


class _PeStat(ctypes.Structure):  # mem_pestat_t, 32 bytes
    _fields_ = [("low", ctypes.c_int32), ("high", ctypes.c_int32), ("failed", ctypes.c_int32),
                ("pad", ctypes.c_int32), ("avg", ctypes.c_double), ("std", ctypes.c_double)]


class NativeLib:
    """ctypes binding of the six C entry points of jnibwa.h:11-16 (+ mem_opt_init, free, version)."""

    def __init__(self, path: str, prefix: str = "jnibwa_", opt_init: str = "mem_opt_init",
                 free: str = "jnibwa_free"):
        self.path = path
        self.dll = ctypes.CDLL(path)  # RTLD_LOCAL: several native libraries may coexist in one process
        d = self.dll
        self.openIndex = getattr(d, prefix + "openIndex")
        self.openIndex.restype = ctypes.c_void_p
        self.openIndex.argtypes = [ctypes.c_int]
        self.destroyIndex = getattr(d, prefix + "destroyIndex")
        self.destroyIndex.restype = ctypes.c_int
        self.destroyIndex.argtypes = [ctypes.c_void_p]
        self.getRefContigNames = getattr(d, prefix + "getRefContigNames")
        self.getRefContigNames.restype = ctypes.c_void_p
        self.getRefContigNames.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
        self.createAlignments = getattr(d, prefix + "createAlignments")
        self.createAlignments.restype = ctypes.c_void_p
        self.createAlignments.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.POINTER(ctypes.c_size_t)]
        self.opt_init = getattr(d, opt_init)
        self.opt_init.restype = ctypes.c_void_p
        self.opt_init.argtypes = []
        self.free = getattr(d, free)
        self.free.restype = None
        self.free.argtypes = [ctypes.c_void_p]
        # index construction lives only in the product library
        self.createIndexFile = getattr(d, "jnibwa_createIndexFile", None)
        if self.createIndexFile is not None:
            self.createIndexFile.restype = ctypes.c_int
            self.createIndexFile.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        self.createReferenceIndex = getattr(d, "jnibwa_createReferenceIndex", None)
        if self.createReferenceIndex is not None:
            self.createReferenceIndex.restype = ctypes.c_int
            self.createReferenceIndex.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
        self.getVersion = getattr(d, "jnibwa_getVersion", None)
        if self.getVersion is not None:
            self.getVersion.restype = ctypes.c_char_p
            self.getVersion.argtypes = []


_default_lib: Optional[NativeLib] = None
_default_lib_lock = threading.Lock()


def default_lib() -> NativeLib:
    """Load the CUDA engine; fails loudly when it has not been built (there is no CPU fallback).

    Mirrors BwaMemIndex.loadNativeLibrary (BwaMemIndex.java:434-477): LIBBWA_PATH wins, else the
    library that ships inside the package directory."""
    global _default_lib
    with _default_lib_lock:
        if _default_lib is None:
            path = os.environ.get("LIBBWA_PATH")
            if not path:
                path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libbwa_b200.so")
            if not os.path.exists(path):
                raise CouldNotReadIndexException(
                    f"native library {path} not found: build it with __graft_entry__.build() "
                    f"(the B200 engine has no CPU fallback)")
            _default_lib = NativeLib(path)
        return _default_lib


# --- BwaMemPairEndStats ------------------------------------------------------------------------
@dataclass(frozen=True)
class BwaMemPairEndStats:
    """Mirror of BwaMemPairEndStats.java:15-167 (insert-size distribution of proper FR pairs)."""
    average: float = float("nan")
    std: float = float("nan")
    low: int = 0
    high: int = 0
    failed: bool = False

    DEFAULT_STD_RANGE = 4.0

    @staticmethod
    def of(average: float, std: float, low: Optional[int] = None, high: Optional[int] = None) -> "BwaMemPairEndStats":
        if not (average == average) or average < 0 or math.isinf(average):
            raise ValueError("the average must be a finite positive number")
        if not (std == std) or std < 0 or math.isinf(std):
            raise ValueError("the std must be a finite positive number")
        if low is None:
            low = max(1, int(round(average - BwaMemPairEndStats.DEFAULT_STD_RANGE * std)))
        if high is None:
            high = max(1, int(round(average + BwaMemPairEndStats.DEFAULT_STD_RANGE * std)))
        if low > high:
            raise ValueError("low cannot be larger than high")
        return BwaMemPairEndStats(float(average), float(std), int(low), int(high), False)


BwaMemPairEndStats.FAILED = BwaMemPairEndStats(failed=True)  # type: ignore[attr-defined]
BwaMemPairEndStats.DO_NOT_INFER = BwaMemPairEndStats.FAILED  # type: ignore[attr-defined]


def _pestat_array(stats: Optional[BwaMemPairEndStats]):
    """jobject_to_mem_pestat_t (…BwaMemIndex.c:21-40): slot 1 (FR) from the object, slots 0/2/3 failed."""
    if stats is None:
        return None
    arr = (_PeStat * 4)()
    for i in range(4):
        if i == 1:
            arr[i].failed = 1 if stats.failed else 0
            if not stats.failed:
                arr[i].low, arr[i].high, arr[i].avg, arr[i].std = stats.low, stats.high, stats.average, stats.std
        else:
            arr[i].failed = 1
    return arr


# --- BwaMemAlignment ---------------------------------------------------------------------------
@dataclass(frozen=True)
class BwaMemAlignment:
    """Mirror of BwaMemAlignment.java:8-65."""
    samFlag: int
    refId: int
    refStart: int
    refEnd: int
    seqStart: int
    seqEnd: int
    mapQual: int
    nMismatches: int
    alignerScore: int
    suboptimalScore: int
    cigar: str
    mdTag: Optional[str]
    xaTag: Optional[str]
    mateRefId: int
    mateRefStart: int
    templateLen: int

    # Java-style getters so tests read like BwaMemIndexTest.java
    def getSamFlag(self): return self.samFlag
    def getRefId(self): return self.refId
    def getRefStart(self): return self.refStart
    def getRefEnd(self): return self.refEnd
    def getSeqStart(self): return self.seqStart
    def getSeqEnd(self): return self.seqEnd
    def getMapQual(self): return self.mapQual
    def getNMismatches(self): return self.nMismatches
    def getAlignerScore(self): return self.alignerScore
    def getSuboptimalScore(self): return self.suboptimalScore
    def getCigar(self): return self.cigar
    def getMDTag(self): return self.mdTag
    def getXATag(self): return self.xaTag
    def getMateRefId(self): return self.mateRefId
    def getMateRefStart(self): return self.mateRefStart
    def getTemplateLen(self): return self.templateLen


_CIGAR_OPS = "MID?S???????????"


def decode_alignments(buf: bytes, n_sequences: int) -> List[List[BwaMemAlignment]]:
    """The decode loop of BwaMemAligner.java:215-308 over the native int32 stream."""
    ints = np.frombuffer(buf, dtype="<i4")
    raw = memoryview(buf)
    p = 0
    out: List[List[BwaMemAlignment]] = []

    def get_tag():
        nonlocal p
        n = int(ints[p]); p += 1
        if n == 0:
            return None
        s = bytes(raw[p * 4:p * 4 + n]).decode("latin-1")
        p += (n + 3) >> 2
        return s

    for _ in range(n_sequences):
        n_aligns = int(ints[p]); p += 1
        alns: List[BwaMemAlignment] = []
        for _ in range(n_aligns):
            flag_mapq = int(ints[p]) & 0xFFFFFFFF; p += 1
            flags = flag_mapq >> 16
            mapq = flag_mapq & 0xFF
            if flags & 0x4:
                ref_id = ref_start = ref_end = seq_start = seq_end = -1
                nm = a_s = xs = 0
                cigar = ""
                md = xa = None
            else:
                ref_id = int(ints[p]); ref_start = int(ints[p + 1]); nm = int(ints[p + 2])
                a_s = int(ints[p + 3]); xs = int(ints[p + 4]); n_cig = int(ints[p + 5]); p += 6
                parts = []
                if n_cig <= 0:
                    seq_start = seq_end = 0
                    ref_end = ref_start
                else:
                    ref_len = seq_len = 0
                    seq_start = 0
                    for k in range(n_cig):
                        len_op = int(ints[p]) & 0xFFFFFFFF; p += 1
                        ln = len_op >> 4
                        op = _CIGAR_OPS[len_op & 0xF]
                        parts.append(f"{ln}{op}")
                        if k == 0 and op == "S":
                            seq_start = ln
                        if op in "MD":
                            ref_len += ln
                        if op in "MI":
                            seq_len += ln
                    ref_end = ref_start + ref_len
                    seq_end = seq_start + seq_len
                cigar = "".join(parts)
                md = get_tag()
                xa = get_tag()
            if (flags & 0x1) == 0 or (flags & 0x8) != 0:
                mate_ref_id = mate_start = -1
                tlen = 0
            else:
                mate_ref_id = int(ints[p]); mate_start = int(ints[p + 1]); tlen = int(ints[p + 2]); p += 3
            alns.append(BwaMemAlignment(flags, ref_id, ref_start, ref_end, seq_start, seq_end, mapq, nm, a_s, xs,
                                        cigar, md, xa, mate_ref_id, mate_start, tlen))
        out.append(alns)
    if p != len(ints):
        raise InvalidInputException(f"alignment stream not consumed exactly: {p} of {len(ints)} ints")
    return out


def pack_sequences(seqs: Sequence[bytes]) -> bytearray:
    """BwaMemAligner.java:198-209: i32 count, then each sequence NUL-terminated."""
    total = 4 + sum(len(s) + 1 for s in seqs)
    if total >= 2 ** 31:
        raise InvalidInputException("too many bases for one call (Java int capacity, hazard H10)")
    buf = bytearray(total)
    struct.pack_into("<i", buf, 0, len(seqs))
    p = 4
    for s in seqs:
        buf[p:p + len(s)] = s
        p += len(s) + 1
    return buf


# --- BwaMemIndex -------------------------------------------------------------------------------
class BwaMemIndex:
    """Mirror of BwaMemIndex.java: an open ".img" index image, shareable between aligners/threads."""

    class Algorithm:
        AUTO = "auto"
        RB2 = "rb2"
        IS = "is"
        BWTSW = "bwtsw"  # listed by the Java enum but rejected by the native side, as in the reference

    INDEX_FILE_EXTENSIONS = (".amb", ".ann", ".bwt", ".pac", ".sa")
    FASTA_FILE_EXTENSIONS = (".fasta", ".fa")
    IMAGE_FILE_EXTENSION = ".img"

    def __init__(self, indexImageFile: str, lib: Optional[NativeLib] = None):
        self._lib = lib or default_lib()
        self._lock = threading.Lock()
        self._refcount = 0
        self.indexImageFile = indexImageFile
        try:
            fd = os.open(indexImageFile, os.O_RDONLY)
        except OSError:
            fd = -1
        addr = self._lib.openIndex(fd) if fd != -1 else None  # the native side owns and closes fd
        if not addr:
            raise CouldNotReadImageException("Could not read bwa index image: " + indexImageFile)
        self._addr = addr
        size = ctypes.c_size_t(0)
        p = self._lib.getRefContigNames(addr, ctypes.byref(size))
        if not p:
            raise IllegalStateException("Could not get reference contig names")  # noqa: F821
        blob = ctypes.string_at(p, size.value)
        self._lib.free(p)
        n = struct.unpack_from("<i", blob, 0)[0]
        names, off = [], 4
        for _ in range(n):
            ln = struct.unpack_from("<i", blob, off)[0]; off += 4
            names.append(blob[off:off + ln].decode()); off += ln
        self._names = names

    # -- index creation ------------------------------------------------------------------------
    @staticmethod
    def createIndexImageFromIndexFiles(referenceFasta: str, imageFile: str, lib: Optional[NativeLib] = None) -> None:
        """BwaMemIndex.java:114-123 -> jnibwa_createIndexFile (jnibwa.c:126-152)."""
        lib = lib or default_lib()
        for ext in BwaMemIndex.INDEX_FILE_EXTENSIONS:
            if not os.path.isfile(referenceFasta + ext):
                raise CouldNotReadIndexException("Missing bwa index file: " + referenceFasta + ext)
        if lib.createIndexFile(referenceFasta.encode(), imageFile.encode()) != 0:
            raise CouldNotCreateIndexImageException("Failed to create index image " + imageFile)

    @staticmethod
    def createIndexImageFromFastaFile(fasta: str, imageFile: str, algo: str = "auto", lib: Optional[NativeLib] = None) -> None:
        """BwaMemIndex.java:218-230 -> createReferenceIndex + createIndexImageFile."""
        lib = lib or default_lib()
        if not os.path.isfile(fasta):
            raise CouldNotReadReferenceException("cannot read reference " + fasta)
        with open(fasta, "rb") as f:
            first = f.read(1)
        if first != b">":
            raise InvalidFileFormatException("not a FASTA file: " + fasta)
        prefix = imageFile + ".tmpidx"
        rc = lib.createReferenceIndex(fasta.encode(), prefix.encode(), algo.encode())
        if rc == 3:
            raise ValueError(f"wrong algorithm name '{algo}'")
        try:
            if rc != 0:
                raise CouldNotCreateIndexException("Failed to index " + fasta)
            if lib.createIndexFile(prefix.encode(), imageFile.encode()) != 0:
                raise CouldNotCreateIndexImageException("Failed to create index image " + imageFile)
        finally:
            for ext in BwaMemIndex.INDEX_FILE_EXTENSIONS:
                try:
                    os.unlink(prefix + ext)
                except OSError:
                    pass

    # -- lifecycle (BwaMemIndex.java:353-397) --------------------------------------------------
    def isOpen(self) -> bool:
        return self._addr is not None

    def _addr_as_void(self):
        """The native handle as a ctypes pointer (for the b200_* instrumentation calls, which take void*)."""
        return ctypes.c_void_p(self._addr)

    def getReferenceContigNames(self) -> List[str]:
        return list(self._names)

    def refIndex(self) -> int:
        with self._lock:
            if self._addr is None:
                raise RuntimeError("Index image has been closed")
            self._refcount += 1
            return self._addr

    def deRefIndex(self) -> None:
        with self._lock:
            self._refcount -= 1

    def close(self) -> None:
        with self._lock:
            if self._addr is not None:
                if self._refcount != 0:
                    raise RuntimeError("Index image is still in use")
                self._lib.destroyIndex(self._addr)
                self._addr = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @staticmethod
    def getBWAVersion(lib: Optional[NativeLib] = None) -> str:
        lib = lib or default_lib()
        return lib.getVersion().decode() if lib.getVersion else "oracle"

    def doAlignment(self, seqs_buf, opts_addr: int, stats: Optional[BwaMemPairEndStats]) -> bytes:
        """BwaMemIndex.java:410-416 -> native createAlignments (…BwaMemIndex.c:142-155)."""
        size = ctypes.c_size_t(0)
        pes = _pestat_array(stats)
        cbuf = (ctypes.c_char * len(seqs_buf)).from_buffer(seqs_buf)
        p = self._lib.createAlignments(self._addr, opts_addr, ctypes.byref(pes) if pes is not None else None,
                                       ctypes.addressof(cbuf), ctypes.byref(size))
        if not p:
            raise RuntimeError("Unable to get alignments from bwa-mem index " + self.indexImageFile)
        out = ctypes.string_at(p, size.value)
        self._lib.free(p)  # destroyByteBuffer
        return out


# --- BwaMemAligner -----------------------------------------------------------------------------
class BwaMemAligner:
    """Mirror of BwaMemAligner.java: owns a native 168-byte mem_opt_t and pokes it by raw offset."""

    MEM_F_PE = 0x2
    MEM_F_NOPAIRING = 0x4
    MEM_F_ALL = 0x8
    MEM_F_NO_MULTI = 0x10
    MEM_F_NO_RESCUE = 0x20
    MEM_F_REF_HDR = 0x100
    MEM_F_SOFTCLIP = 0x200
    MEM_F_SMARTPE = 0x400
    MEM_F_PRIMARY5 = 0x800

    def __init__(self, index: BwaMemIndex):
        if not index.isOpen():
            raise RuntimeError("Can't create aligner: bwa-mem index has been closed")
        self.index = index
        self._lib = index._lib
        self._opts_addr = self._lib.opt_init()
        self._opts = (ctypes.c_uint8 * 168).from_address(self._opts_addr)
        self.pairEndStats: Optional[BwaMemPairEndStats] = None

    def isOpen(self) -> bool:
        return self._opts_addr is not None

    def close(self) -> None:
        if self._opts_addr is not None:
            self._lib.free(self._opts_addr)
            self._opts_addr = None
            self._opts = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _o(self):
        if self._opts is None:
            raise RuntimeError("The aligner has been closed.")
        return self._opts

    def _geti(self, off): return struct.unpack_from("<i", self._o(), off)[0]
    def _seti(self, off, v): struct.pack_into("<i", self._o(), off, int(v))
    def _getf(self, off): return struct.unpack_from("<f", self._o(), off)[0]
    def _setf(self, off, v): struct.pack_into("<f", self._o(), off, float(v))

    def getExpectedOptsSize(self): return 168
    def getOptsSize(self): return ctypes.sizeof(self._o())

    def setIntraCtgOptions(self):
        self.setDGapOpenPenaltyOption(16); self.setIGapOpenPenaltyOption(16); self.setMismatchPenaltyOption(9)
        self.setClip5PenaltyOption(5); self.setClip3PenaltyOption(5)

    def alignPairs(self): self.setFlagOption(self.MEM_F_PE | self.getFlagOption())
    def inferPairEndStats(self): self.pairEndStats = None
    def dontInferPairEndStats(self): self.pairEndStats = BwaMemPairEndStats.DO_NOT_INFER
    def setProperPairEndStats(self, stats: BwaMemPairEndStats): self.pairEndStats = stats
    def getIndex(self): return self.index

    def getMaxMemIntvOption(self): return struct.unpack_from("<q", self._o(), 48)[0]
    def setMaxMemIntvOption(self, v): struct.pack_into("<q", self._o(), 48, int(v))
    def getScoringMatrixOption(self): return bytes(self._o()[140:165])

    def setScoringMatrixOption(self, mat: bytes):
        o = self._o()
        for i, b in enumerate(bytes(mat)[:25]):
            o[140 + i] = b

    def alignSeqs(self, sequences: Iterable, func: Callable = lambda s: s) -> List[List[BwaMemAlignment]]:
        """BwaMemAligner.java:192-311."""
        opts_addr = self._opts_addr
        if opts_addr is None:
            raise RuntimeError("The aligner has been closed.")
        self.index.refIndex()
        try:
            seqs = [bytes(func(e)) for e in sequences]
            blob = pack_sequences(seqs)
            out = self.index.doAlignment(blob, opts_addr, self.pairEndStats)
        finally:
            self.index.deRefIndex()
        n = len(seqs)
        if (self.getFlagOption() & self.MEM_F_PE) and (n & 1):
            n -= 1  # hazard H11: the odd trailing read yields no records
        return decode_alignments(out, n)


def _int_opt(name: str, off: int):
    setattr(BwaMemAligner, "get" + name, lambda self, _o=off: self._geti(_o))
    setattr(BwaMemAligner, "set" + name, lambda self, v, _o=off: self._seti(_o, v))


def _float_opt(name: str, off: int, set_name: Optional[str] = None):
    setattr(BwaMemAligner, "get" + name, lambda self, _o=off: self._getf(_o))
    setattr(BwaMemAligner, "set" + (set_name or name), lambda self, v, _o=off: self._setf(_o, v))


# option accessors by raw byte offset, BwaMemAligner.java:46-124
for _n, _off in [("MatchScoreOption", 0), ("MismatchPenaltyOption", 4), ("DGapOpenPenaltyOption", 8),
                 ("DGapExtendPenaltyOption", 12), ("IGapOpenPenaltyOption", 16), ("IGapExtendPenaltyOption", 20),
                 ("UnpairedPenaltyOption", 24), ("Clip5PenaltyOption", 28), ("Clip3PenaltyOption", 32),
                 ("BandwidthOption", 36), ("ZDropOption", 40), ("OutputScoreThresholdOption", 56),
                 ("FlagOption", 60), ("MinSeedLengthOption", 64), ("MinChainWeightOption", 68),
                 ("MaxChainExtendOption", 72), ("SplitWidthOption", 80), ("MaxSeedOccurencesOption", 84),
                 ("MaxChainGapOption", 88), ("NThreadsOption", 92), ("ChunkSizeOption", 96),
                 ("MapQCoefFacOption", 120), ("MaxInsOption", 124), ("MaxMateSWOption", 128),
                 ("MaxXAHitsOption", 132), ("MaxXAHitsAltOption", 136)]:
    _int_opt(_n, _off)
for _n, _off, _s in [("SplitFactorOption", 76, None), ("MaskLevelOption", 100, "MaxLevelOption"),
                     ("DropRatioOption", 104, None), ("XADropRatio", 108, None),
                     ("MaskLevelRedunOption", 112, None), ("MapQCoefLenOption", 116, None)]:
    _float_opt(_n, _off, _s)


class IllegalStateException(RuntimeError):
    pass
