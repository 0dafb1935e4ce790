"""CPU: the dynamic shared memory of the thread-level extension kernels is carved up per thread by interleaving
(csrc/kernels.cuh: thread_ext_layout).  Arrays with different element sizes interleave differently, so a region that
is private to a thread under one view can belong to other threads under another -- the kernel-debug harness (one
thread at a time, stride 1) cannot see such a clash.  Here every byte a thread may touch is enumerated for a whole
block and checked to be owned by exactly that thread, for every kernel variant (cell width x patching)."""
import ctypes

import pytest

from common import ENGINE_SO


def _layout(dll, threads, tid, rw, wm_cap, cell, patches):
    out = (ctypes.c_uint64 * 9)()
    dll.b200_extLayout(threads, tid, rw, wm_cap, cell, int(patches), out)
    return list(out)


@pytest.mark.parametrize("cell", [2, 4])
@pytest.mark.parametrize("patches", [False, True])
@pytest.mark.parametrize("read_len", [36, 100, 150, 250])
def test_no_two_threads_share_a_byte(cell, patches, read_len):
    dll = ctypes.CDLL(ENGINE_SO)
    dll.b200_extLayout.restype = None
    threads, wm_cap = 128, ((read_len + 15) // 16) * 16 + 16
    rw = read_len - 19 + 2
    if cell == 4 and rw * cell < 256:
        rw = 256 // cell  # engine_impl.inl: wide cells keep room for rings of THREAD_RING_N / 2 slots
    owner = {}
    total = None
    for tid in range(threads):
        row_off, row_st, q_off, q_st, rh, re_, r_st, ring_n, block = _layout(dll, threads, tid, rw, wm_cap, cell, patches)
        total = block
        mine = set()
        for j in range(rw):
            mine.update(range(row_off + j * row_st, row_off + j * row_st + cell))
        for j in range(wm_cap + 2):
            mine.add(q_off + j * q_st)
        if patches and cell == 4:  # (2-byte cells: the patching alignment uses rows in global scratch, no rings)
            assert ring_n >= 16 and ring_n & (ring_n - 1) == 0
            for k in range(ring_n):
                mine.update(range(rh + k * r_st, rh + k * r_st + 4))
                mine.update(range(re_ + k * r_st, re_ + k * r_st + 4))
        else:
            assert ring_n == 0
        assert max(mine) < block, "a thread reaches past the block's allocation"
        for b in mine:
            assert owner.setdefault(b, tid) == tid, f"byte {b} belongs to threads {owner[b]} and {tid}"
    assert total is not None and total <= 227 * 1024
