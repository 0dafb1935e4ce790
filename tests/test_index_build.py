"""CPU: the from-scratch index builder against the reference's checked-in index files (byte-exact) and the
testIndexReference scenario of BwaMemIndexTest.java:142-177."""
import os
import random

import pytest

from common import GOLDEN, bwa, oracle_lib, simdata, workdir

FIX = os.path.join(GOLDEN, "rotavirus", "ref.fa")


def test_builder_reproduces_reference_index_files():
    prefix = os.path.join(workdir(), "rebuilt_ref")
    rc = simdata.idxlib().jnibwa_createReferenceIndex(FIX.encode(), prefix.encode(), b"auto")
    assert rc == 0
    for ext in (".amb", ".ann", ".bwt", ".pac", ".sa"):
        assert open(prefix + ext, "rb").read() == open(FIX + ext, "rb").read(), ext


def test_image_from_files_equals_image_from_fasta():
    a = os.path.join(workdir(), "a.img")
    b = os.path.join(workdir(), "b.img")
    assert simdata.idxlib().jnibwa_createIndexFile(FIX.encode(), a.encode()) == 0
    prefix = os.path.join(workdir(), "rebuilt_ref2")
    assert simdata.idxlib().jnibwa_createReferenceIndex(FIX.encode(), prefix.encode(), b"is") == 0
    assert simdata.idxlib().jnibwa_createIndexFile(prefix.encode(), b.encode()) == 0
    assert open(a, "rb").read() == open(b, "rb").read()


def test_unknown_algorithm_rejected():
    assert simdata.idxlib().jnibwa_createReferenceIndex(FIX.encode(), b"/tmp/x", b"bwtsw") == 3


@pytest.mark.parametrize("algo", ["auto", "rb2", "is"])
def test_index_reference_two_contigs(algo):
    rnd = random.Random(13)
    fa = os.path.join(workdir(), f"two_{algo}.fasta")
    with open(fa, "w") as f:
        for name, n in (("seq1", 45212), ("seq2", 13415)):
            s = "".join(rnd.choice("ACGT") for _ in range(n))
            f.write(f">{name}\n")
            for i in range(0, n, 60):
                f.write(s[i:i + 60] + "\n")
    img = fa + ".idx"
    bwa.BwaMemIndex.createIndexImageFromFastaFile(fa, img, algo, lib=_IdxLib())
    idx = bwa.BwaMemIndex(img, lib=oracle_lib())
    assert idx.getReferenceContigNames() == ["seq1", "seq2"]
    idx.close()


class _IdxLib:
    """index-construction entry points of the product library, host-only build"""
    def __init__(self):
        d = simdata.idxlib()
        self.createIndexFile = d.jnibwa_createIndexFile
        self.createReferenceIndex = d.jnibwa_createReferenceIndex
