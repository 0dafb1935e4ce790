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
