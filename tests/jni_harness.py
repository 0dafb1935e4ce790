"""TEST TOOLING: builds jni/jni_shim.c + tests/jni_mock/mock_env.c into one library and wraps the Java_* natives the
way BwaMemIndex.java / BwaMemAligner.java call them (BwaMemIndex.java:479-487), with a mock JNIEnv instead of a JVM."""
import ctypes
import os
import subprocess

from common import ROOT, PKG_DIR, bwa

SO = os.path.join(ROOT, "tests", "jni_mock", "libjnimock.so")
SO_REFGLUE = os.path.join(ROOT, "tests", "jni_mock", "libjnimock_refglue.so")
REF_C = "/root/reference/src/main/c"
P = "Java_org_broadinstitute_hellbender_utils_bwa_BwaMemIndex_"


def have_reference_glue() -> bool:
    return os.path.isfile(os.path.join(REF_C, "org_broadinstitute_hellbender_utils_bwa_BwaMemIndex.c"))


def build(variant: str = "shim"):
    """variant "shim": jni/jni_shim.c (ours).  variant "reference_glue": the reference's OWN JNI files, unmodified and
    compiled where they lie (/root/reference/src/main/c/org_..._BwaMemIndex.c + init.c), against include/compat
    (jni.h, bwa/bwamem.h, bwa/bwa_commit.h) and linked with -lbwa_b200 -- INTEGRATION.md option A."""
    lib = os.path.join(PKG_DIR, "lib")
    mock = os.path.join(ROOT, "tests", "jni_mock", "mock_env.c")
    if variant == "reference_glue":
        srcs = [os.path.join(REF_C, "org_broadinstitute_hellbender_utils_bwa_BwaMemIndex.c"), os.path.join(REF_C, "init.c"), mock]
        compat = os.path.join(ROOT, "include", "compat")
        deps = srcs + [os.path.join(compat, "jni.h"), os.path.join(compat, "bwa", "bwamem.h"), os.path.join(compat, "bwa", "bwa_commit.h"), os.path.join(ROOT, "jni", "jni_min.h")]
        if not os.path.exists(SO_REFGLUE) or os.path.getmtime(SO_REFGLUE) < max(os.path.getmtime(f) for f in deps):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-w", "-I", compat, "-I", os.path.join(ROOT, "jni"), "-I", os.path.join(ROOT, "include"),
                                   "-o", SO_REFGLUE] + srcs + ["-L", lib, "-lbwa_b200", "-Wl,-rpath," + lib])
        return SO_REFGLUE
    srcs = [os.path.join(ROOT, "jni", "jni_shim.c"), mock]
    deps = srcs + [os.path.join(ROOT, "jni", "jni_min.h"), os.path.join(ROOT, "include", "bwamem_b200.h")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in deps):
        subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "jni"),
                               "-o", SO] + srcs + ["-L", lib, "-lbwa_b200", "-Wl,-rpath," + lib])
    return SO


class MockJvm:
    """Loads the shim like System.load + JNI_OnLoad, and exposes the natives with Java-like signatures."""

    def __init__(self, variant: str = "shim"):
        self.dll = d = ctypes.CDLL(build(variant))
        vp, ll, i32 = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
        for name, res, args in [("mock_env", vp, []), ("mock_vm", vp, []), ("mock_new_buffer", vp, [vp, ll]), ("mock_buffer_addr", vp, [vp]),
                                ("mock_buffer_cap", ll, [vp]), ("mock_free_buffer_object", None, [vp]),
                                ("mock_new_pestats", vp, [i32, i32, i32, ctypes.c_double, ctypes.c_double]),
                                ("mock_exception_class", ctypes.c_char_p, []), ("mock_exception_message", ctypes.c_char_p, []),
                                ("mock_clear_exception", None, []), ("mock_live_strings", i32, []), ("JNI_OnLoad", i32, [vp, vp]),
                                (P + "createReferenceIndex", ctypes.c_ubyte, [vp, vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]),
                                (P + "createIndexImageFile", ctypes.c_ubyte, [vp, vp, ctypes.c_char_p, ctypes.c_char_p]),
                                (P + "openIndex", ll, [vp, vp, ctypes.c_char_p]), (P + "destroyIndex", i32, [vp, vp, ll]),
                                (P + "createDefaultOptions", vp, [vp, vp]), (P + "getRefContigNames", vp, [vp, vp, ll]),
                                (P + "createAlignments", vp, [vp, vp, vp, ll, vp, vp]), (P + "destroyByteBuffer", None, [vp, vp, vp]),
                                (P + "getVersion", vp, [vp, vp])]:
            f = getattr(d, name)
            f.restype, f.argtypes = res, args
        self.vm = d.mock_vm()
        self.env = d.mock_env()
        self.onload = d.JNI_OnLoad(self.vm, None)

    def call(self, name, *args):
        return getattr(self.dll, P + name)(self.env, None, *args)

    def buffer_bytes(self, buf) -> bytes:
        return ctypes.string_at(self.dll.mock_buffer_addr(buf), self.dll.mock_buffer_cap(buf))

    def release(self, buf):
        """BwaMemIndex.destroyByteBuffer + dropping the Java object."""
        self.call("destroyByteBuffer", buf)
        self.dll.mock_free_buffer_object(buf)

    def align(self, idx_addr: int, opts_buf, reads, stats=None) -> bytes:
        """BwaMemAligner.alignSeqs up to the decode step (BwaMemAligner.java:192-214): pack, call, copy out, release."""
        blob = ctypes.create_string_buffer(bytes(bwa.pack_sequences(reads)))
        seqs = self.dll.mock_new_buffer(ctypes.cast(blob, ctypes.c_void_p), len(blob))
        st = None
        if stats is not None:
            st = self.dll.mock_new_pestats(int(stats.failed), stats.low, stats.high, stats.average, stats.std)
        out = self.call("createAlignments", seqs, idx_addr, opts_buf, st)
        self.dll.mock_free_buffer_object(seqs)
        if not out:
            return None
        data = self.buffer_bytes(out)
        self.release(out)
        return data
