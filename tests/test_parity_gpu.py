"""GPU: the CUDA engine through its C-ABI (jnibwa_createAlignments) against the CPU oracle on the same inputs,
options and batch boundaries.  Bar: the int32 record stream is byte-identical (flag, refId, pos, CIGAR, MAPQ, NM,
MD, AS, XS, XA, mate fields)."""
import pytest

from common import alt_contig_image, bwa, engine_lib, first_diff, oracle_lib, repeat_family_image, run_blob, simdata, synth_image
from test_oracle_kat import run_kat

pytestmark = pytest.mark.gpu


def test_engine_known_answers():
    run_kat(engine_lib())


@pytest.fixture(scope="module")
def small_genome():
    return synth_image("g200k", 7, [120000, 60000, 20000], repeat_frac=0.08, tandem_frac=0.01, n_segdup=3)


@pytest.fixture(scope="module")
def mid_genome():
    return synth_image("g5m", 21, [3000000, 1500000, 500000], repeat_frac=0.05, tandem_frac=0.01, n_segdup=10)


CASES = [
    ("se150", dict(seed=11, n_units=4000, read_len=150, paired=False), 0, None),
    ("pe150", dict(seed=12, n_units=3000, read_len=150, paired=True), 2, None),
    ("pe250_noisy", dict(seed=13, n_units=1000, read_len=250, paired=True, ins_mean=550, ins_std=80, err_rate=0.05,
                         mut_rate=0.01, indel_frac=0.5), 2, None),
    ("pe_chimeric_repeats_n", dict(seed=14, n_units=2000, read_len=150, paired=True, chimeric_frac=0.1,
                                   repeat_read_frac=0.2, n_rate=0.002), 2, None),
    ("se_all_flag", dict(seed=15, n_units=2000, read_len=150, paired=False, chimeric_frac=0.2, repeat_read_frac=0.3), 8, None),
    ("pe_given_stats", dict(seed=16, n_units=1000, read_len=150, paired=True), 2, (400.0, 50.0, 200, 600)),
    ("pe_no_infer", dict(seed=17, n_units=1000, read_len=150, paired=True), 2, "failed"),
    ("pe_no_rescue_primary5", dict(seed=18, n_units=1000, read_len=150, paired=True, chimeric_frac=0.2), 2 | 0x20 | 0x800, None),
    ("se_short_and_ragged", dict(seed=19, n_units=500, read_len=36, paired=False), 0, None),
]


def _stats(s):
    if s is None:
        return None
    if s == "failed":
        return bwa.BwaMemPairEndStats.DO_NOT_INFER
    return bwa.BwaMemPairEndStats.of(*s)


@pytest.mark.parametrize("name,kw,flag,stats", CASES, ids=[c[0] for c in CASES])
def test_engine_matches_oracle_small(small_genome, name, kw, flag, stats):
    g, img = small_genome
    kw = dict(kw)
    blob, _ = simdata.make_reads(g, kw.pop("seed"), kw.pop("n_units"), kw.pop("read_len"), kw.pop("paired"), **kw)
    a = run_blob(oracle_lib(), img, blob, flag, _stats(stats))
    b = run_blob(engine_lib(), img, blob, flag, _stats(stats))
    assert a == b, first_diff(a, b, blob, bool(flag & 2))


def test_edge_inputs(small_genome):
    g, img = small_genome
    reads = [b"", b"A", b"ACGT" * 5, b"N" * 150, b"ACGTNNNNACGT" * 10, b"acgtacgtnn" * 15]
    base, _ = simdata.make_reads(g, 30, 20, 150, False)
    reads += simdata.split_blob(bytes(base))
    reads.append(reads[-1][:100] + b"N" + reads[-1][101:])
    blob = bwa.pack_sequences(reads)
    for flag in (0, 2):  # odd read count in PE mode: the trailing read yields nothing (hazard H11)
        a = run_blob(oracle_lib(), img, blob, flag)
        b = run_blob(engine_lib(), img, blob, flag)
        assert a == b, first_diff(a, b, blob, bool(flag & 2))
    empty = bwa.pack_sequences([])
    assert run_blob(engine_lib(), img, empty) == run_blob(oracle_lib(), img, empty)


def test_intractg_options(small_genome):
    """hazard H1: setIntraCtgOptions changes a/b/gap-open but not the scoring matrix."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 31, 500, 250, False, err_rate=0.03)
    setup = lambda al: al.setIntraCtgOptions()
    a = run_blob(oracle_lib(), img, blob, 0, None, setup)
    b = run_blob(engine_lib(), img, blob, 0, None, setup)
    assert a == b, first_diff(a, b, blob, False)


LONG_CASES = [(1000, 400, dict(err_rate=0.02)), (5000, 120, dict(err_rate=0.03, mut_rate=0.01, indel_frac=0.5)),
              (20000, 40, dict(err_rate=0.02, mut_rate=0.005, indel_frac=0.4)), (50000, 12, dict(err_rate=0.01, mut_rate=0.005, indel_frac=0.4))]


@pytest.mark.parametrize("intra", [False, True], ids=["default_options", "intra_ctg_options"])
@pytest.mark.parametrize("read_len,n,kw", LONG_CASES, ids=[f"{c[0]}bp" for c in LONG_CASES])
def test_long_queries(mid_genome, read_len, n, kw, intra):
    """Long queries, 1-50 kbp (GATK's SV pipeline aligns assembled contigs with setIntraCtgOptions,
    BwaMemAligner.java:140-146): mem_flt_chained_seeds / mem_seed_sw on the device and the long-read mode (warp-level
    kernels with DP rows and traceback in global memory).  Hazard H1: the intra-contig preset changes a/b/gap-open
    but not the scoring matrix."""
    g, img = mid_genome
    blob, _ = simdata.make_reads(g, 12, n, read_len, False, **kw)

    def setup(al):
        al.setNThreadsOption(8)
        if intra:
            al.setIntraCtgOptions()
    a = run_blob(oracle_lib(), img, blob, 0, None, setup)
    b = run_blob(engine_lib(), img, blob, 0, None, setup)
    assert a == b, first_diff(a, b, blob, False)


def test_long_and_short_reads_in_one_batch_and_min_chain_weight(mid_genome):
    g, img = mid_genome
    short = simdata.split_blob(bytes(simdata.make_reads(g, 31, 300, 150, False)[0]))
    long_ = simdata.split_blob(bytes(simdata.make_reads(g, 32, 30, 3000, False, err_rate=0.02)[0]))
    blob = bwa.pack_sequences([r for pair in zip(short[:30], long_) for r in pair] + short[30:])
    a = run_blob(oracle_lib(), img, blob, 0)
    b = run_blob(engine_lib(), img, blob, 0)
    assert a == b, first_diff(a, b, blob, False)
    blob2, _ = simdata.make_reads(g, 33, 2000, 150, True, err_rate=0.03)
    setup = lambda al: al.setMinChainWeightOption(40)
    a = run_blob(oracle_lib(), img, blob2, 2, None, setup)
    b = run_blob(engine_lib(), img, blob2, 2, None, setup)
    assert a == b, first_diff(a, b, blob2, True)


@pytest.mark.parametrize("paired", [False, True], ids=["se", "pe"])
def test_engine_matches_oracle_mid(mid_genome, paired):
    g, img = mid_genome
    blob, _ = simdata.make_reads(g, 41 + paired, 20000, 150, paired)
    flag = 2 if paired else 0

    def setup(al):
        al.setNThreadsOption(8)
    a = run_blob(oracle_lib(), img, blob, flag, None, setup)
    b = run_blob(engine_lib(), img, blob, flag, None, setup)
    assert a == b, first_diff(a, b, blob, paired)


@pytest.mark.parametrize("env", [None, "B200_TEST_RESCUE_FALLBACK", "B200_TEST_TINY_SCRATCH", "B200_TEST_ALL_LIGHT", "B200_TEST_ALL_HEAVY", "B200_WARP_EXTEND", "B200_NO_EXT_TASKS", "B200_TEST_CHAIN_ALL_WARP", "B200_TEST_CHAIN_ALL_THREAD", "B200_FINAL_ONE_PASS", "B200_SEED_ONE_LAUNCH", "B200_SEED_FUSED_ROUNDS", "B200_TEST_TINY_SEED_POOL", "B200_TX_CHUNK=1", "B200_TX_CHUNK=0", "B200_EXT_WIDE_CELLS", "B200_SEED_NO_OVERLAP", "B200_CHAIN_NO_OVERLAP", "B200_TEST_FINAL_RING", "B200_TEST_TINY_GA_CAP", "B200_TEST_TINY_SEED_TASKS", "B200_EXT_LEGACY", "B200_EXT_NO_SORT", "B200_TEST_LONG_MODE", "B200_RESCUE_WARP_SW", "B200_EXT_TASK_MACHINE", "B200_GA_NO_SORT", "BWA_B200_KMER=6", "BWA_B200_KMER=0", "B200_TEST_OCC_SB=3", "B200_TEST_SEED_SLICES=3", "B200_TEST_TINY_INTV_CAP", "B200_TEST_TINY_WARP_SCRATCH"],
                         ids=["default", "rescue_fallback", "tiny_scratch", "extend_all_thread", "extend_all_tasks", "extend_all_warp", "extend_heavy_by_warp",
                              "chain_all_warp", "chain_all_thread", "final_one_pass", "seed_one_launch", "seed_fused_rounds", "seed_pool_overflow", "extend_rows_cell_by_cell", "extend_whole_rows", "extend_wide_cells", "seed_no_stream_overlap", "chain_no_stream_overlap", "final_passes_with_ring_smem", "alignment_task_pool_overflow", "seed_task_list_overflow", "extend_read_level_machine", "extend_tasks_unsorted", "long_read_mode_forced", "rescue_alignments_by_warps", "extend_tasks_by_lane_machine", "alignment_tasks_unsorted", "prefix_tables_of_6_mers", "no_prefix_tables", "occ_super_blocks", "seeding_in_slices", "interval_lists_regrown", "warp_scratch_regrown"])
def test_repeat_heavy_pairs_and_forced_slow_paths(small_genome, monkeypatch, env):
    """Chains of several rescuing hits per read (repeat-derived reads); the same batch with every rescuing read forced
    onto the round-by-round rescue kernels, with thread-level scratch so small that items are handed to warps, and with
    the seed extension forced onto each of its three routes (thread per read, hand-over list, warp per read)."""
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 77, 3000, 150, True, repeat_read_frac=0.5, chimeric_frac=0.05, err_rate=0.02, indel_frac=0.3)
    a = run_blob(oracle_lib(), img, blob, 2)
    if env:
        name, _, val = env.partition("=")
        monkeypatch.setenv(name, val or "1")
    b = run_blob(engine_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)


def test_repeat_family_genome(monkeypatch):
    """A genome that is 30 % repeat families: reads with hundreds of hits drive the warp-level chaining, the per-seed
    extension tasks, the warp-cooperative region de-duplication and long mate-rescue chains."""
    g, img = synth_image("g1m_rep", 9, [600000, 400000], repeat_frac=0.3, tandem_frac=0.02, n_segdup=3)
    blob, _ = simdata.make_reads(g, 81, 4000, 150, True, repeat_read_frac=0.6, chimeric_frac=0.05, err_rate=0.02)

    def setup(al):
        al.setNThreadsOption(8)
    a = run_blob(oracle_lib(), img, blob, 2, None, setup)
    b = run_blob(engine_lib(), img, blob, 2, None, setup)
    assert a == b, first_diff(a, b, blob, True)


@pytest.mark.parametrize("paired", [False, True], ids=["se", "pe"])
def test_repeat_family_stress(paired):
    """1500 copies of one element: hundreds of hits per read, rescue chains of max_matesw hits (common.repeat_family_image)."""
    g, img = repeat_family_image()
    blob, _ = simdata.make_reads(g, 92, 3000, 150, paired, repeat_read_frac=0.7, err_rate=0.01)

    def setup(al):
        al.setNThreadsOption(16)
    flag = 2 if paired else 0
    a = run_blob(oracle_lib(), img, blob, flag, None, setup)
    b = run_blob(engine_lib(), img, blob, flag, None, setup)
    assert a == b, first_diff(a, b, blob, paired)


def test_repeat_calls_are_deterministic(small_genome):
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 50, 2000, 150, True)
    idx = bwa.BwaMemIndex(img, lib=engine_lib())
    al = bwa.BwaMemAligner(idx)
    al.alignPairs()
    outs = [idx.doAlignment(bytearray(blob), al._opts_addr, None) for _ in range(3)]
    assert outs[0] == outs[1] == outs[2]
    al.close()
    idx.close()


@pytest.mark.parametrize("chunk", [None, "20000"], ids=["one_chunk", "many_chunks_with_splits"])
def test_device_index_builder_matches_host_builder(tmp_path, monkeypatch, chunk):
    """csrc/sa_build.cu (GPU suffix sorting) against the host sorter of index_build.cpp: identical image bytes, on a
    genome with repeat families, tandem repeats and segmental duplications (deep refinement rounds), also with a chunk
    capacity so small that prefix ranges have to be split."""
    import os
    g = simdata.make_genome(5, [300000, 150000, 50000], repeat_frac=0.08, tandem_frac=0.02, n_segdup=4)
    host = str(tmp_path / "host.img")
    dev = str(tmp_path / "dev.img")
    simdata.build_image(g, host)
    monkeypatch.setenv("BWA_B200_BUILD", "device")
    if chunk:
        monkeypatch.setenv("BWA_B200_BUILD_CHUNK", chunk)
    simdata.build_image(g, dev, dll=engine_lib().dll)
    assert os.path.getsize(host) == os.path.getsize(dev)
    assert open(host, "rb").read() == open(dev, "rb").read()


def test_concurrent_callers_and_two_indexes(small_genome, mid_genome):
    """The reference's usage: one shared BwaMemIndex, one BwaMemAligner per thread (BwaMemAligner.java:18); here also two
    different indexes open at once.  Callers overlap on the handle's contexts; every result must be the single-threaded one."""
    import threading
    g1, img1 = small_genome
    g2, img2 = mid_genome
    blobs1 = [simdata.make_reads(g1, 60 + i, 800, 150, True)[0] for i in range(3)]
    blobs2 = [simdata.make_reads(g2, 70 + i, 800, 150, False)[0] for i in range(3)]
    want1 = [run_blob(oracle_lib(), img1, b, 2) for b in blobs1]
    want2 = [run_blob(oracle_lib(), img2, b, 0) for b in blobs2]
    idx1 = bwa.BwaMemIndex(img1, lib=engine_lib())
    idx2 = bwa.BwaMemIndex(img2, lib=engine_lib())
    got = {}

    def worker(tag, idx, blobs, flag):
        al = bwa.BwaMemAligner(idx)
        if flag:
            al.alignPairs()
        for rep in range(2):
            for i, b in enumerate(blobs):
                got[(tag, rep, i)] = idx.doAlignment(bytearray(b), al._opts_addr, None)
        al.close()

    threads = [threading.Thread(target=worker, args=("a", idx1, blobs1, 2)), threading.Thread(target=worker, args=("b", idx1, blobs1, 2)),
               threading.Thread(target=worker, args=("c", idx2, blobs2, 0))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    idx1.close()
    idx2.close()
    for (tag, rep, i), out in got.items():
        assert out == (want2 if tag == "c" else want1)[i], (tag, rep, i)
    assert len(got) == 18


def _pool_api():
    import ctypes
    dll = engine_lib().dll
    dll.b200_contextCount.argtypes = [ctypes.c_void_p]
    dll.b200_contextDevice.argtypes = [ctypes.c_void_p, ctypes.c_int]
    return dll


def _fresh_thread(fn):
    """Run fn on a new host thread (CUDA's current device is per thread: a fresh thread starts on device 0)."""
    import threading
    box = {}

    def run():
        try:
            box["v"] = fn()
        except BaseException as e:  # noqa: BLE001
            box["e"] = e
    t = threading.Thread(target=run)
    t.start(); t.join()
    if "e" in box:
        raise box["e"]
    return box["v"]


def test_index_on_second_device_called_from_fresh_threads(small_genome, monkeypatch):
    """The handle remembers its device: an index opened with BWA_B200_DEVICE=1 is used from threads whose current CUDA
    device is 0 (every Spark task thread of the reference's deployment).  Needs two GPUs (gpurun --gpus 2)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    g, img = small_genome
    blob, _ = simdata.make_reads(g, 55, 2000, 150, True)
    want = run_blob(oracle_lib(), img, blob, 2)
    monkeypatch.setenv("BWA_B200_DEVICE", "1")
    idx = _fresh_thread(lambda: bwa.BwaMemIndex(img, lib=engine_lib()))
    assert _pool_api().b200_contextDevice(idx._addr_as_void(), 0) == 1

    def call():
        al = bwa.BwaMemAligner(idx)
        al.alignPairs()
        out = idx.doAlignment(bytearray(blob), al._opts_addr, None)
        al.close()
        return out
    for _ in range(3):
        assert _fresh_thread(call) == want
    idx.close()


def test_device_pool_single_process(small_genome, monkeypatch):
    """BWA_B200_DEVICES=all: one index replica per GPU inside one process (replicas 1.. are copied device to device),
    calls from many threads spread over the pool; every result equals the single-GPU one."""
    import threading
    import torch
    ndev = torch.cuda.device_count()
    g, img = small_genome
    blobs = [simdata.make_reads(g, 160 + i, 1500, 150, True)[0] for i in range(6)]
    want = [run_blob(oracle_lib(), img, b, 2) for b in blobs]
    monkeypatch.setenv("BWA_B200_DEVICES", "all")
    monkeypatch.setenv("BWA_B200_CONTEXTS", "2")
    idx = bwa.BwaMemIndex(img, lib=engine_lib())
    dll = _pool_api()
    assert dll.b200_contextCount(idx._addr_as_void()) == 2 * ndev
    assert sorted({dll.b200_contextDevice(idx._addr_as_void(), i) for i in range(2 * ndev)}) == list(range(ndev))
    got = {}

    def worker(t):
        al = bwa.BwaMemAligner(idx)
        al.alignPairs()
        for rep in range(2):
            for i, b in enumerate(blobs):
                got[(t, rep, i)] = idx.doAlignment(bytearray(b), al._opts_addr, None)
        al.close()
    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2 * ndev + 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    idx.close()
    assert len(got) == (2 * ndev + 1) * 12
    for (t, rep, i), out in got.items():
        assert out == want[i], (t, rep, i)


@pytest.mark.parametrize("paired,flag", [(False, 0), (True, 2), (False, 8), (True, 2 | 0x800), (True, 2 | 0x20)], ids=["se", "pe", "se_all", "pe_primary5", "pe_no_rescue"])
def test_alt_contigs(paired, flag):
    """An image with ALT contigs (is_alt from <prefix>.alt, jnibwa.c:127): ALT-aware chain filtering, the two-round primary
    marking with alt_sc, XA lists up to max_XA_hits_alt (setMaxXAHitsAltOption, BwaMemAligner.java:124-125), the extra
    supplementary record of paired output."""
    g, img = alt_contig_image()
    blob, _ = simdata.make_reads(g, 51, 4000, 150, paired, err_rate=0.01)
    a = run_blob(oracle_lib(), img, blob, flag)
    b = run_blob(engine_lib(), img, blob, flag)
    assert a == b, first_diff(a, b, blob, paired)
    n = 4000 * (2 if paired else 1)
    recs = bwa.decode_alignments(b, n)
    names = bwa.BwaMemIndex(img, lib=oracle_lib())
    alt_ids = {i for i, nm in enumerate(names.getReferenceContigNames()) if nm.startswith("alt")}
    names.close()
    assert any(al.refId in alt_ids for alns in recs for al in alns), "no record on an ALT contig: the case is not exercised"


def test_alt_contigs_round_by_round_rescue(monkeypatch):
    """ALT contigs with every rescuing read forced onto the round-by-round rescue kernels (which restore the mate's hits
    from the untagged temporary regions and tag them again)."""
    g, img = alt_contig_image()
    monkeypatch.setenv("B200_TEST_RESCUE_FALLBACK", "1")
    blob, _ = simdata.make_reads(g, 52, 3000, 100, True, err_rate=0.03, mut_rate=0.02, chimeric_frac=0.1)
    a = run_blob(oracle_lib(), img, blob, 2)
    b = run_blob(engine_lib(), img, blob, 2)
    assert a == b, first_diff(a, b, blob, True)
