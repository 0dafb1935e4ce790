"""CPU: the oracle against every known-answer vector the reference's own tests hold for this path
(BwaMemIndexTest.java:38-127, restated in tests/golden/kat.json) and the Java-API mirror's plumbing."""
import json
import os

import pytest

from common import GOLDEN, bwa, fixture_image, oracle_lib

KAT = json.load(open(os.path.join(GOLDEN, "kat.json")))


def check(aln, exp):
    ref_start, ref_end, seq_start, seq_end, cigar, nm, flag = exp[:7]
    assert aln.getRefStart() == ref_start and aln.getRefEnd() == ref_end
    assert aln.getSeqStart() == seq_start and aln.getSeqEnd() == seq_end
    assert aln.getCigar() == cigar and aln.getNMismatches() == nm
    assert aln.getRefId() == 0 and aln.getSamFlag() == flag
    if len(exp) > 7:
        assert aln.getMateRefStart() == exp[7] and aln.getTemplateLen() == exp[8]


def run_kat(lib):
    idx = bwa.BwaMemIndex(fixture_image(), lib=lib)
    try:
        with bwa.BwaMemAligner(idx) as al:  # testOptsSize
            assert al.getOptsSize() == al.getExpectedOptsSize() == 168
        for name in ("testSimple", "testMulti"):
            al = bwa.BwaMemAligner(idx)
            res = al.alignSeqs(KAT[name]["reads"], lambda s: s.encode())
            assert len(res) == len(KAT[name]["reads"])
            for alns, exp in zip(res, KAT[name]["expect"]):
                assert len(alns) == 1
                check(alns[0], exp)
            al.close()
        for mode in (0, 1, 2):
            al = bwa.BwaMemAligner(idx)
            al.alignPairs()
            if mode == 1:
                al.setProperPairEndStats(bwa.BwaMemPairEndStats.of(*KAT["testPair"]["stats_mode_1"]))
            elif mode == 2:
                al.dontInferPairEndStats()
            else:
                al.inferPairEndStats()
            res = al.alignSeqs(KAT["testPair"]["reads"], lambda s: s.encode())
            assert len(res) == 2
            for alns, exp in zip(res, KAT["testPair"]["expect"][str(mode)]):
                assert len(alns) == 1
                check(alns[0], exp)
            al.close()
        assert idx.getReferenceContigNames() == ["gi|9626372|ref|NC_001422.1|"] or len(idx.getReferenceContigNames()) == 1
    finally:
        idx.close()


def test_oracle_known_answers():
    run_kat(oracle_lib())


def test_option_layout_and_defaults():
    """BwaMemAligner.java:46-137 offsets against mem_opt_init defaults."""
    idx = bwa.BwaMemIndex(fixture_image(), lib=oracle_lib())
    al = bwa.BwaMemAligner(idx)
    assert (al.getMatchScoreOption(), al.getMismatchPenaltyOption()) == (1, 4)
    assert (al.getDGapOpenPenaltyOption(), al.getDGapExtendPenaltyOption()) == (6, 1)
    assert (al.getIGapOpenPenaltyOption(), al.getIGapExtendPenaltyOption()) == (6, 1)
    assert al.getUnpairedPenaltyOption() == 17 and al.getClip5PenaltyOption() == 5 and al.getClip3PenaltyOption() == 5
    assert al.getBandwidthOption() == 100 and al.getZDropOption() == 100 and al.getMaxMemIntvOption() == 20
    assert al.getOutputScoreThresholdOption() == 30 and al.getFlagOption() == 0 and al.getMinSeedLengthOption() == 19
    assert al.getSplitWidthOption() == 10 and al.getMaxSeedOccurencesOption() == 500 and al.getMaxChainGapOption() == 10000
    assert abs(al.getSplitFactorOption() - 1.5) < 1e-7 and abs(al.getXADropRatio() - 0.8) < 1e-7
    assert al.getMaxInsOption() == 10000 and al.getMaxMateSWOption() == 50 and al.getMaxXAHitsOption() == 5
    assert al.getMaxXAHitsAltOption() == 200 and al.getMapQCoefFacOption() == 3
    mat = al.getScoringMatrixOption()
    assert mat[0] == 1 and mat[1] == 0xFC and mat[4] == 0xFF and mat[24] == 0xFF
    al.setIntraCtgOptions()
    assert al.getMismatchPenaltyOption() == 9 and al.getScoringMatrixOption() == mat  # hazard H1: matrix not refreshed
    al.close()
    with pytest.raises(RuntimeError):
        al.getFlagOption()
    idx.close()
    with pytest.raises(RuntimeError):
        bwa.BwaMemAligner(idx)


def test_missing_image_raises():
    with pytest.raises(bwa.CouldNotReadImageException):
        bwa.BwaMemIndex("/nonexistent/none.img", lib=oracle_lib())


def test_sse2_and_scalar_striped_kernels_agree():
    """ksw_align2's striped kernel exists twice in the oracle: SSE2 intrinsics (oracle/ksw_sse2.c, what the timed CPU arm runs
    -- bwa itself runs these SIMD kernels) and a scalar emulation (oracle/ksw.c: the readable statement of the semantics).
    Identical results on random related / unrelated sequence pairs, byte and short kernels, all xtra modes."""
    import ctypes
    import random
    dll = oracle_lib().dll
    if not dll.orc_have_sse2():
        pytest.skip("no SSE2 on this machine")

    class R(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int) for n in ("score", "te", "qe", "score2", "te2", "tb", "qb")]
    dll.orc_ksw_align2.restype = R
    dll.orc_ksw_align2.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 5
    mat = bytes((v & 0xff) for v in ([1, -4, -4, -4, -1, -4, 1, -4, -4, -1, -4, -4, 1, -4, -1, -4, -4, -4, 1, -1, -1, -1, -1, -1, -1]))
    rng = random.Random(5)
    XBYTE, XSTOP, XSUBO, XSTART = 0x10000, 0x20000, 0x40000, 0x80000
    n = 0
    for trial in range(400):
        qlen = rng.choice([1, 7, 16, 17, 31, 100, 150, 151, 250])
        tlen = rng.choice([qlen, qlen + 37, 3 * qlen + 11, 550])
        t = [rng.randrange(4) for _ in range(tlen)]
        if trial % 3:  # a mutated copy of a target window, with an indel now and then
            off = rng.randrange(max(1, tlen - qlen + 1))
            q = list(t[off:off + qlen])
            for k in range(len(q)):
                if rng.random() < 0.06:
                    q[k] = rng.randrange(5)
            if len(q) > 20 and rng.random() < 0.5:
                cut = rng.randrange(5, len(q) - 5)
                q = q[:cut] + ([rng.randrange(4)] * rng.randrange(1, 6) if rng.random() < 0.5 else []) + q[cut + rng.randrange(0, 4):]
            q = (q + [rng.randrange(4) for _ in range(qlen)])[:qlen]
        else:
            q = [rng.randrange(5) for _ in range(qlen)]
        gaps = rng.choice([(6, 1, 6, 1), (16, 1, 16, 1), (4, 2, 5, 1)])
        for xtra in (XSTART | XSUBO | 19, XSTART | XSUBO | XBYTE | 19, XSTART, XBYTE | XSTOP | 40, XSTOP | 25):
            if (xtra & XBYTE) and qlen >= 250:
                continue
            out = []
            for simd in (0, 1):
                dll.orc_set_simd(simd)
                r = dll.orc_ksw_align2(qlen, bytes(q), tlen, bytes(t), mat, *gaps, xtra)
                out.append((r.score, r.te, r.qe, r.score2, r.te2, r.tb, r.qb))
            assert out[0] == out[1], (trial, qlen, tlen, hex(xtra), gaps, out)
            n += 1
    dll.orc_set_simd(1)
    assert n > 1500


def test_oracle_pipeline_is_the_same_with_either_kernel():
    import ctypes
    from common import run_blob, simdata, synth_image
    dll = oracle_lib().dll
    if not dll.orc_have_sse2():
        pytest.skip("no SSE2 on this machine")
    g, img = synth_image("g200k", 7, [120000, 60000, 20000], repeat_frac=0.08, tandem_frac=0.01, n_segdup=3)
    for read_len, kw in ((150, dict(repeat_read_frac=0.3, chimeric_frac=0.05)), (250, dict(ins_mean=550, ins_std=80, err_rate=0.05, mut_rate=0.01, indel_frac=0.5))):
        blob, _ = simdata.make_reads(g, 61, 400, read_len, True, **kw)
        try:
            dll.orc_set_simd(0)
            a = run_blob(oracle_lib(), img, blob, 2)
        finally:
            dll.orc_set_simd(1)
        b = run_blob(oracle_lib(), img, blob, 2)
        assert a == b
