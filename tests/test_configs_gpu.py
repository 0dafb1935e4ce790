"""GPU: the BASELINE.json configurations as parity cases at a scale the oracle finishes in seconds (the full sizes are
what bench.py runs, with its own full-scale parity check).  Genome: 64 Mbp, 8 contigs, repeat families / tandem
repeats / segmental duplications, index built by the device suffix sorter.  Bar: byte-identical record streams."""
import os

import pytest

from common import bwa, engine_lib, first_diff, oracle_lib, run_blob, simdata, workdir

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def genome64():
    lens = [12_000_000, 10_000_000, 9_000_000, 8_000_000, 8_000_000, 7_000_000, 6_000_000, 4_000_000]
    g = simdata.make_genome(33, lens, repeat_frac=0.05, tandem_frac=0.01, n_segdup=20)
    img = os.path.join(workdir(), "g64m.img")
    if not os.path.exists(img):
        os.environ["BWA_B200_BUILD"] = "device"
        try:
            simdata.build_image(g, img + ".tmp", dll=engine_lib().dll)
        finally:
            del os.environ["BWA_B200_BUILD"]
        os.replace(img + ".tmp", img)
    return g, img


def _threads(al):
    al.setNThreadsOption(os.cpu_count() or 8)


CONFIGS = [
    # C2: single-end 150 bp
    ("c2_se150", dict(seed=101, n_units=100000, read_len=150, paired=False), 0),
    # C3: 2x150 bp paired-end with mate rescue + pestat (the bench workload)
    ("c3_pe150", dict(seed=102, n_units=40000, read_len=150, paired=True), 2),
    # C4: 2x250 bp at high error / indel rate (band retries, z-drop)
    ("c4_pe250_noisy", dict(seed=103, n_units=15000, read_len=250, paired=True, ins_mean=550, ins_std=80, err_rate=0.05, mut_rate=0.01,
                            indel_frac=0.5), 2),
    # C5: chimeric / split reads with supplementary + XA alternate hits
    ("c5_chimeric_xa", dict(seed=104, n_units=30000, read_len=150, paired=True, chimeric_frac=0.1, repeat_read_frac=0.2), 2),
]


@pytest.mark.parametrize("name,kw,flag", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_baseline_config(genome64, name, kw, flag):
    g, img = genome64
    kw = dict(kw)
    blob, _ = simdata.make_reads(g, kw.pop("seed"), kw.pop("n_units"), kw.pop("read_len"), kw.pop("paired"), **kw)
    a = run_blob(oracle_lib(), img, blob, flag, None, _threads)
    b = run_blob(engine_lib(), img, blob, flag, None, _threads)
    assert a == b, first_diff(a, b, blob, bool(flag & 2))


def test_size_independent_properties(genome64):
    """Properties that hold at any size: every read gets >= 1 record; a mapped record's CIGAR consumes the whole read
    and NM equals what its MD implies; paired flags are mutually consistent; a second call returns the same bytes."""
    g, img = genome64
    n = 50000
    blob, _ = simdata.make_reads(g, 105, n, 150, True)
    out = run_blob(engine_lib(), img, blob, 2)
    assert out == run_blob(engine_lib(), img, blob, 2)
    recs = bwa.decode_alignments(out, 2 * n)
    assert len(recs) == 2 * n
    n_mapped = 0
    for i, alns in enumerate(recs):
        assert len(alns) >= 1
        for al in alns:
            assert al.samFlag & 0x1 and bool(al.samFlag & 0x40) != bool(al.samFlag & 0x80)
            if al.samFlag & 0x4:
                continue
            n_mapped += 1
            import re
            ops = re.findall(r"(\d+)([MIDSH])", al.cigar)
            assert sum(int(l) for l, op in ops if op in "MIS") == 150, al.cigar
            ref_len = sum(int(l) for l, op in ops if op in "MD")
            assert al.refEnd - al.refStart == ref_len
            mism = len(re.findall(r"[ACGTN]", re.sub(r"\^[ACGTN]+", "", al.mdTag)))
            gaps = sum(int(l) for l, op in ops if op in "ID")
            assert al.nMismatches == mism + gaps, (al.cigar, al.mdTag, al.nMismatches)
    assert n_mapped > 1.9 * n  # wgsim-style reads from the genome itself: nearly all map
