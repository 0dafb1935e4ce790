"""The JNI layer under a mock JNIEnv, in two builds: our shim (jni/jni_shim.c) and the reference's OWN glue files compiled
unmodified against include/compat and linked with libbwa_b200.so (INTEGRATION.md option A; skipped where
/root/reference does not exist, e.g. on the GPU box).  Checked: symbol surface, JNI_OnLoad field-ID cache, option block, error
behaviour (CPU), and the full createAlignments path against the oracle (GPU).  Mirrors what BwaMemIndexTest.java
exercises through a real JVM (/root/reference/src/test/java/.../BwaMemIndexTest.java:27-43,84-127)."""
import ctypes
import os
import subprocess

import pytest

from common import GOLDEN, bwa, fixture_image, oracle_lib, run_blob, workdir
import jni_harness

NATIVES = ["createReferenceIndex", "createIndexImageFile", "openIndex", "destroyIndex", "createDefaultOptions", "getRefContigNames",
           "createAlignments", "destroyByteBuffer", "getVersion"]


@pytest.fixture(scope="module", params=["shim", "reference_glue"])
def jvm(request):
    if request.param == "reference_glue" and not jni_harness.have_reference_glue():
        pytest.skip("/root/reference is not on this machine")
    return jni_harness.MockJvm(request.param)


@pytest.mark.parametrize("variant", ["shim", "reference_glue"])
def test_exports_the_nine_natives_and_onload(variant):
    if variant == "reference_glue" and not jni_harness.have_reference_glue():
        pytest.skip("/root/reference is not on this machine")
    out = subprocess.check_output(["nm", "-D", "--defined-only", jni_harness.build(variant)], text=True)
    syms = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert "JNI_OnLoad" in syms
    for n in NATIVES:
        assert jni_harness.P + n in syms, n


def test_onload_caches_field_ids(jvm):
    assert jvm.onload == 0x00010008  # JNI_VERSION_1_8, init.c:28


def test_default_options_block(jvm):
    """testOptsSize (BwaMemIndexTest.java:38-43): 168 bytes; defaults readable at the offsets the Java setters use."""
    buf = jvm.call("createDefaultOptions")
    assert jvm.dll.mock_buffer_cap(buf) == 168
    raw = jvm.buffer_bytes(buf)
    i32 = lambda off: int.from_bytes(raw[off:off + 4], "little", signed=True)
    assert (i32(0), i32(4), i32(36), i32(56), i32(64)) == (1, 4, 100, 30, 19)  # a, b, w, T, min_seed_len
    jvm.release(buf)


def test_version_string(jvm):
    s = jvm.call("getVersion")
    assert b"bwamem-b200" in ctypes.string_at(s)


def test_unknown_algorithm_raises_illegal_argument(jvm, tmp_path):
    jvm.dll.mock_clear_exception()
    ok = jvm.call("createReferenceIndex", b"/nonexistent.fa", str(tmp_path / "x").encode(), b"bwtsw")
    assert ok == 0
    assert jvm.dll.mock_exception_class() == b"java/lang/IllegalArgumentException"
    assert jvm.dll.mock_exception_message() == b"wrong algorithm name 'bwtsw'"
    assert jvm.dll.mock_live_strings() == 0  # every GetStringUTFChars was released


def test_image_creation_and_open_failure_modes(jvm):
    img = os.path.join(workdir(), "jni_rotavirus.img")
    if os.path.exists(img):
        os.remove(img)
    assert jvm.call("createIndexImageFile", os.path.join(GOLDEN, "rotavirus", "ref.fa").encode(), img.encode()) == 1
    assert open(img, "rb").read() == open(fixture_image(), "rb").read()
    assert jvm.call("openIndex", b"/no/such/image") == 0  # -> CouldNotReadImageException in Java
    assert jvm.call("destroyIndex", 0) == 0
    assert not jvm.call("getRefContigNames", 0)


@pytest.mark.parametrize("algo", [b"auto", b"is", b"rb2"])
def test_reference_index_through_the_natives(jvm, tmp_path, algo):
    """createReferenceIndex + createIndexImageFile (BwaMemIndex.java:218-230) on the reference's own FASTA reproduce the
    reference's checked-in index files and the image built from them, byte for byte -- through bwa_idx_build when the
    glue is the reference's own (...BwaMemIndex.c:59)."""
    fa = os.path.join(GOLDEN, "rotavirus", "ref.fa")
    prefix = str(tmp_path / "idx")
    assert jvm.call("createReferenceIndex", fa.encode(), prefix.encode(), algo) == 1
    for ext in (".amb", ".ann", ".bwt", ".pac", ".sa"):
        assert open(prefix + ext, "rb").read() == open(fa + ext, "rb").read(), ext
    img = str(tmp_path / "idx.img")
    assert jvm.call("createIndexImageFile", prefix.encode(), img.encode()) == 1
    assert open(img, "rb").read() == open(fixture_image(), "rb").read()
    assert jvm.dll.mock_live_strings() == 0


@pytest.mark.gpu
def test_alignments_through_jni_match_oracle(jvm):
    """testSimple / testMulti / testPair through the JNI natives, byte-compared with the oracle's stream."""
    img = fixture_image()
    addr = jvm.call("openIndex", img.encode())
    assert addr != 0
    names = jvm.call("getRefContigNames", addr)
    raw_names = jvm.buffer_bytes(names)
    assert int.from_bytes(raw_names[:4], "little") == 1  # one contig in the rotavirus fixture
    jvm.release(names)
    opts = jvm.call("createDefaultOptions")
    import json
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    reads = [r.encode() for r in kat["testSimple"]["reads"] + kat["testMulti"]["reads"]]
    got = jvm.align(addr, opts, reads)
    want = run_blob(oracle_lib(), img, bwa.pack_sequences(reads), 0)
    assert got == want
    # paired mode with supplied FR statistics (testPair, BwaMemIndexTest.java:84-127): flag bit MEM_F_PE at offset 60
    raw = (ctypes.c_int32 * 42).from_address(jvm.dll.mock_buffer_addr(opts))
    raw[15] |= 2
    stats = bwa.BwaMemPairEndStats.of(200.0, 10.0, 1, 600)
    pair = [r.encode() for r in kat["testPair"]["reads"]]
    got = jvm.align(addr, opts, pair, stats)
    want = run_blob(oracle_lib(), img, bwa.pack_sequences(pair), 2, stats)
    assert got == want
    jvm.release(opts)
    assert jvm.call("destroyIndex", addr) == 0
