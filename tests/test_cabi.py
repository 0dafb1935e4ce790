"""CPU: the product library loads and exports every symbol include/bwamem_b200.h declares; without a CUDA device
the engine refuses to open an index (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from common import ENGINE_SO, ROOT, bwa, fixture_image


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "bwamem_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b((?:jnibwa_|b200_|mem_opt_)\w+)\s*\(", hdr)))


def test_library_exports_declared_symbols():
    assert os.path.exists(ENGINE_SO), "build with __graft_entry__.build()"
    dll = ctypes.CDLL(ENGINE_SO)
    syms = declared_symbols()
    assert {"jnibwa_openIndex", "jnibwa_destroyIndex", "jnibwa_getRefContigNames", "jnibwa_createAlignments",
            "jnibwa_createIndexFile", "mem_opt_init"} <= set(syms)
    for s in syms:
        assert hasattr(dll, s), s


def test_default_options_block():
    lib = bwa.NativeLib(ENGINE_SO)
    p = lib.opt_init()
    raw = ctypes.string_at(p, 168)
    lib.free(p)
    import struct
    assert struct.unpack_from("<11i", raw, 0) == (1, 4, 6, 1, 6, 1, 17, 5, 5, 100, 100)
    assert struct.unpack_from("<q", raw, 48)[0] == 20 and struct.unpack_from("<i", raw, 64)[0] == 19
    assert raw[140] == 1 and raw[141] == 0xFC


def test_no_device_means_no_index(have_gpu):
    if have_gpu:
        pytest.skip("CUDA device present")
    with pytest.raises(bwa.CouldNotReadImageException):
        bwa.BwaMemIndex(fixture_image(), lib=bwa.NativeLib(ENGINE_SO))
