"""CPU: the host-side scan that finds the read boundaries of a request (csrc/host_scan.cpp; the reference walks the blob
with strlen, jnibwa.c:199-212).  The AVX2 path must agree with the plain one for every alignment of the buffer, for empty
reads and ragged lengths, and must not touch memory behind the last NUL (the buffer ends exactly at a page boundary that
is followed by an inaccessible page)."""
import ctypes
import mmap
import random

import pytest

from common import ENGINE_SO


@pytest.fixture(scope="module")
def scan():
    dll = ctypes.CDLL(ENGINE_SO)
    dll.b200_scanReads.restype = ctypes.c_uint64
    dll.b200_scanReads.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    return dll.b200_scanReads


def _expect(reads):
    off, p = [], 0
    for r in reads:
        off.append(p)
        p += len(r) + 1
    return off + [p], max([len(r) for r in reads], default=0)


def _run(scan, addr, n, plain):
    off = (ctypes.c_uint64 * (n + 1))()
    mx = ctypes.c_int(0)
    total = scan(addr, n, off, ctypes.byref(mx), plain)
    return list(off), mx.value, total


@pytest.mark.parametrize("plain", [0, 1], ids=["simd", "plain"])
def test_ragged_reads_at_every_alignment(scan, plain):
    rnd = random.Random(3)
    reads = [bytes(rnd.choice(b"ACGTN") for _ in range(rnd.choice([0, 1, 2, 31, 32, 33, 63, 64, 65, 150, 151, 250]))) for _ in range(400)]
    body = b"\0".join(reads) + b"\0"
    want_off, want_max = _expect(reads)
    for shift in range(0, 40):
        buf = ctypes.create_string_buffer(b"\x01" * shift + body + b"\x01" * 64)  # non-zero guard bytes on both sides
        off, mx, total = _run(scan, ctypes.addressof(buf) + shift, len(reads), plain)
        assert off == want_off and mx == want_max and total == len(body)


def test_zero_reads(scan):
    buf = ctypes.create_string_buffer(b"ACGT")
    off, mx, total = _run(scan, ctypes.addressof(buf), 0, 0)
    assert off == [0] and mx == 0 and total == 0


def test_never_reads_past_the_last_nul(scan):
    """The body ends on the last byte of a page; the next page is PROT_NONE, so any read behind the final NUL faults."""
    page = mmap.PAGESIZE
    libc = ctypes.CDLL(None, use_errno=True)
    libc.mmap.restype = ctypes.c_void_p
    libc.mmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long]
    libc.mprotect.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    base = libc.mmap(None, 3 * page, mmap.PROT_READ | mmap.PROT_WRITE, mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS, -1, 0)
    assert base not in (None, ctypes.c_void_p(-1).value)
    assert libc.mprotect(base + 2 * page, page, 0) == 0
    reads = [b"ACGT" * 37, b"", b"GATTACA" * 9, b"T" * 150]
    body = b"\0".join(reads) + b"\0"
    want_off, want_max = _expect(reads)
    for slack in (0, 1, 31, 32, 33):  # distance of the final NUL from the end of the readable area
        start = base + 2 * page - len(body) - slack
        ctypes.memset(base, 0x41, 2 * page)
        ctypes.memmove(start, body, len(body))
        off, mx, total = _run(scan, start, len(reads), 0)
        assert off == want_off and mx == want_max and total == len(body)
