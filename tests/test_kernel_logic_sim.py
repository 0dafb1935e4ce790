"""CPU: the engine's kernel source (csrc/*.cuh) compiled by the kernel-debug harness (tests/cudasim, -DB200_SIM:
warp variables become 32-lane arrays, collectives become loops) against the oracle.  This exercises the host
orchestration (engine_impl.inl), every device function's logic and the record encoder without a GPU; the real
parity gate is tests/test_parity_gpu.py on the B200."""
import pytest

from common import alt_contig_image, first_diff, oracle_lib, repeat_family_image, run_blob, sim_lib, simdata, synth_image
from test_oracle_kat import run_kat


def test_sim_known_answers():
    run_kat(sim_lib())


@pytest.fixture(scope="module")
def small_genome():
    return synth_image("g200k", 7, [120000, 60000, 20000], repeat_frac=0.08, tandem_frac=0.01, n_segdup=3)


CASES = [
    ("se150", dict(seed=11, n_units=400, read_len=150, paired=False), 0),
    ("pe150", dict(seed=12, n_units=300, read_len=150, paired=True), 2),
    ("pe250_noisy", dict(seed=13, n_units=100, read_len=250, paired=True, ins_mean=550, ins_std=80, err_rate=0.05,
                         mut_rate=0.01, indel_frac=0.5), 2),
    ("pe_chimeric_repeats_n", dict(seed=14, n_units=200, read_len=150, paired=True, chimeric_frac=0.1,
                                   repeat_read_frac=0.2, n_rate=0.002), 2),
    ("se_all_flag", dict(seed=15, n_units=200, read_len=150, paired=False, chimeric_frac=0.2, repeat_read_frac=0.3), 8),
]


@pytest.mark.parametrize("name,kw,flag", CASES, ids=[c[0] for c in CASES])
def test_sim_matches_oracle(small_genome, name, kw, flag):
    g, img = small_genome
    kw = dict(kw)
    blob, _ = simdata.make_reads(g, kw.pop("seed"), kw.pop("n_units"), kw.pop("read_len"), kw.pop("paired"), **kw)
    a = run_blob(oracle_lib(), img, blob, flag)
    b = run_blob(sim_lib(), img, blob, flag)
    assert a == b, first_diff(a, b, blob, bool(flag & 2))


def test_sim_rescue_fallback_path(small_genome, monkeypatch):
    """Mate rescue: every rescuing read is forced from the speculative replay onto the round-by-round kernels
    (restore of the mate's list included); repeat-heavy reads give chains of several rescuing hits."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 77, 400, 150, True, repeat_read_frac=0.5, chimeric_frac=0.05)
    a = run_blob(oracle_lib(), img, blob, 2)
    assert run_blob(sim_lib(), img, blob, 2) == a
    monkeypatch.setenv("B200_TEST_RESCUE_FALLBACK", "1")
    b = run_blob(sim_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


def test_sim_extension_wide_cells(small_genome, monkeypatch):
    """Thread-level extension with 16-bit H/E cells (what batches with scores above 255 use) on a batch that would
    run with 8-bit cells."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 79, 300, 150, True, repeat_read_frac=0.3, err_rate=0.03, indel_frac=0.4)
    a = run_blob(oracle_lib(), img, blob, 2)
    monkeypatch.setenv("B200_EXT_WIDE_CELLS", "1")
    b = run_blob(sim_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


@pytest.mark.parametrize("chunk", ["1", "7", "0"])
def test_sim_extension_rows_in_chunks(small_genome, monkeypatch, chunk):
    """Thread-level extension: a DP row is computed in visits of B200_TX_CHUNK cells (0: whole rows); the result must
    not depend on the chunk."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 78, 300, 150, True, repeat_read_frac=0.3, err_rate=0.03, indel_frac=0.4)
    a = run_blob(oracle_lib(), img, blob, 2)
    monkeypatch.setenv("B200_TX_CHUNK", chunk)
    b = run_blob(sim_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


def test_sim_repeat_family_stress():
    """Reads with hundreds of hits each (see common.repeat_family_image)."""
    g, img = repeat_family_image()
    blob, _ = simdata.make_reads(g, 91, 150, 150, True, repeat_read_frac=0.7, err_rate=0.01)
    a = run_blob(oracle_lib(), img, blob, 2)
    b = run_blob(sim_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


@pytest.mark.parametrize("env", ["B200_TEST_TINY_GA_CAP", "B200_TEST_TINY_SEED_TASKS", "B200_TEST_TINY_SEED_POOL"])
def test_sim_pool_overflows_fall_back(small_genome, monkeypatch, env):
    """Capacity fall-backs: the alignment-task pool of the two-pass finalize (items that find it full are redone by a
    warp; the task kernel must not run past the pool), the seed-task list and the seed-candidate pool (reads whose
    scans do not fit are redone by the one-machine kernel).  None of them may fail the batch."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 83, 300, 150, True, repeat_read_frac=0.3, chimeric_frac=0.1, err_rate=0.03, indel_frac=0.4)
    a = run_blob(oracle_lib(), img, blob, 2)
    monkeypatch.setenv(env, "1")
    b = run_blob(sim_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


LONG_CASES = [(800, 40, dict(err_rate=0.02)), (1500, 24, dict(err_rate=0.03, mut_rate=0.01, indel_frac=0.5)),
              (5000, 8, dict(err_rate=0.02, mut_rate=0.005, indel_frac=0.4))]


@pytest.mark.parametrize("intra", [False, True], ids=["default_options", "intra_ctg_options"])
@pytest.mark.parametrize("read_len,n,kw", LONG_CASES, ids=[f"{c[0]}bp" for c in LONG_CASES])
def test_sim_long_reads(small_genome, read_len, n, kw, intra):
    """Long queries (assembled contigs; BwaMemAligner.setIntraCtgOptions, BwaMemAligner.java:140-146): reads of >= ~730
    bases enable mem_flt_chained_seeds / mem_seed_sw (B10); reads above 1000 bases run in the long-read mode (warp-level
    kernels, DP rows and traceback in global memory)."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 12, n, read_len, False, **kw)
    setup = (lambda al: al.setIntraCtgOptions()) if intra else None
    a = run_blob(oracle_lib(), img, blob, 0, None, setup)
    b = run_blob(sim_lib(), img, blob, 0, None, setup)
    assert a == b, first_diff(a, b, blob, False)


def test_sim_long_read_pairs_and_min_chain_weight(small_genome):
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 21, 16, 1200, True, ins_mean=3000, ins_std=200, err_rate=0.02)
    assert run_blob(sim_lib(), img, blob, 2) == run_blob(oracle_lib(), img, blob, 2)
    # min_chain_weight > 0 switches the seed filter on for short reads too
    blob, _ = simdata.make_reads(g, 22, 200, 150, False, err_rate=0.03)
    setup = lambda al: al.setMinChainWeightOption(40)
    a = run_blob(oracle_lib(), img, blob, 0, None, setup)
    b = run_blob(sim_lib(), img, blob, 0, None, setup)
    assert a == b, first_diff(a, b, blob, False)


def test_sim_forced_long_mode_on_short_reads(small_genome, monkeypatch):
    """The long-read mode (every read on the warp-level kernels, rows in global memory, every finalize item by a warp)
    forced onto a paired batch of short reads: same records."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 12, 200, 150, True, repeat_read_frac=0.3, chimeric_frac=0.1)
    a = run_blob(oracle_lib(), img, blob, 2)
    monkeypatch.setenv("B200_TEST_LONG_MODE", "1")
    b = run_blob(sim_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


@pytest.mark.parametrize("paired,flag", [(False, 0), (True, 2), (False, 8), (True, 2 | 0x800)], ids=["se", "pe", "se_all", "pe_primary5"])
def test_sim_alt_contigs(paired, flag):
    """An image with ALT contigs (see common.alt_contig_image)."""
    g, img = alt_contig_image()
    blob, _ = simdata.make_reads(g, 51, 300, 150, paired, err_rate=0.01)
    a = run_blob(oracle_lib(), img, blob, flag)
    b = run_blob(sim_lib(), img, blob, flag)
    assert a == b, first_diff(a, b, blob, paired)


@pytest.mark.parametrize("k", ["2", "5", "7"])
def test_sim_prefix_tables(small_genome, monkeypatch, k):
    """Prefix tables (BWA_B200_KMER): the first extensions of every forward scan and the backward extensions that yield short
    strings are table reads (forward scans only since the sweeps measured slower with them); the intervals, hence the records, are the same."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 84, 300, 150, True, repeat_read_frac=0.3, n_rate=0.01, err_rate=0.03)
    a = run_blob(oracle_lib(), img, blob, 2)
    monkeypatch.setenv("BWA_B200_KMER", k)
    b = run_blob(sim_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


@pytest.mark.parametrize("shift", ["1", "4", "9"])
def test_sim_occurrence_super_blocks(small_genome, monkeypatch, shift):
    """The device keeps u32 occurrence counts per 32-byte sector; a text in which a base occurs 2^32 times or more makes
    them relative to u64 super-block bases (fm.cuh).  No test genome is that large, so the hook forces bases every 2^shift
    sectors: every lookup then adds a base, and the records must not change."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 85, 300, 150, True, repeat_read_frac=0.3, n_rate=0.01, err_rate=0.03)
    a = run_blob(oracle_lib(), img, blob, 2)
    monkeypatch.setenv("B200_TEST_OCC_SB", shift)
    b = run_blob(sim_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


@pytest.mark.parametrize("slices", ["2", "7"])
def test_sim_seeding_in_slices(small_genome, monkeypatch, slices):
    """A batch whose candidate pool would exceed BWA_B200_SEED_POOL_GB is seeded in slices of reads (rounds 1/2 of a
    slice complete before the next one starts, the pool holds one slice); forced here on a small batch."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 86, 301, 150, True, repeat_read_frac=0.3, n_rate=0.01, err_rate=0.03)
    a = run_blob(oracle_lib(), img, blob, 2)
    monkeypatch.setenv("B200_TEST_SEED_SLICES", slices)
    b = run_blob(sim_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


def test_sim_interval_lists_regrown(small_genome, monkeypatch):
    """A read whose SMEM list outgrows its slice (reads from families of diverged repeats: a new SMEM every base or two) has
    the list clamped, and the batch is run again with twice the room until every list fits; forced with 4 entries per read."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 87, 300, 150, True, repeat_read_frac=0.3, n_rate=0.01, err_rate=0.03)
    a = run_blob(oracle_lib(), img, blob, 2)
    monkeypatch.setenv("B200_TEST_TINY_INTV_CAP", "1")
    b = run_blob(sim_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


@pytest.mark.parametrize("extra", [None, "B200_TEST_ALL_HEAVY"])
def test_sim_warp_scratch_regrown(small_genome, monkeypatch, extra):
    """An item that outgrows a capacity of the warp-level scratch (traceback, CIGAR, MD, XA, row-maxima list, the records'
    payload pool) raises a flag; the batch is then run again with every capacity doubled instead of failing the call.
    Forced by starting at 1/64 of the capacities; with every read on the warp-level kernels as well."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 88, 300, 150, True, repeat_read_frac=0.3, n_rate=0.01, err_rate=0.03)
    a = run_blob(oracle_lib(), img, blob, 2)
    monkeypatch.setenv("B200_TEST_TINY_WARP_SCRATCH", "1")
    if extra:
        monkeypatch.setenv(extra, "1")
    b = run_blob(sim_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


def test_sim_alt_contigs_round_by_round_rescue(monkeypatch):
    """ALT contigs with every rescuing read on the round-by-round rescue kernels: those restore the mate's hit list from
    the temporary (untagged) regions and have to tag the hits on ALT contigs again (found by tests/fuzz_sim.py)."""
    g, img = alt_contig_image()
    monkeypatch.setenv("B200_TEST_RESCUE_FALLBACK", "1")
    for seed in (5005, 5011):
        blob, _ = simdata.make_reads(g, seed, 40, 50, True, err_rate=0.05, n_rate=0.1, chimeric_frac=0.2, mut_rate=0.02, ins_mean=400, ins_std=10)
        a = run_blob(oracle_lib(), img, blob, 2)
        b = run_blob(sim_lib(), img, blob, 2)
        assert a == b, first_diff(a, b, blob, True)
