"""Property tests (hypothesis) of the host-side pieces of the C ABI: merge-path search and the
shard partition.  CPU only."""
import ctypes as C

import numpy as np
from hypothesis import given, settings, strategies as st

import merge_spmv_b200 as ms
from merge_spmv_b200 import sharded

row_lengths = st.lists(st.integers(min_value=0, max_value=40), min_size=1, max_size=120)


def offsets(lengths):
    ro = np.zeros(len(lengths) + 1, np.int32)
    ro[1:] = np.cumsum(lengths)
    return ro


@settings(max_examples=200, deadline=None)
@given(row_lengths, st.integers(min_value=0, max_value=6000))
def test_host_search_is_the_counting_identity(lengths, d):
    # x = #{r : row_end[r] + r < d} clamped, y = d - x; past the end -> (rows, nnz)  
    ro = offsets(lengths)
    rows, nnz = len(lengths), int(ro[-1])
    x, y = C.c_int(), C.c_int()
    ms.lib().mspmv_host_merge_path_search(ro.ctypes.data_as(C.c_void_p), rows, nnz, d, C.byref(x), C.byref(y))
    if d > rows + nnz:
        assert (x.value, y.value) == (rows, nnz)
    else:
        want = min(int(np.sum(ro[1:].astype(np.int64) + np.arange(rows) < d)), rows)
        assert (x.value, y.value) == (want, d - want)


@settings(max_examples=100, deadline=None)
@given(row_lengths, st.integers(min_value=1, max_value=9))
def test_shard_partition_invariants(lengths, p):
    ro = offsets(lengths)
    rows, nnz = len(lengths), int(ro[-1])
    coords = sharded.partition(ro, p)
    assert coords[0].tolist() == [0, 0] and coords[-1].tolist() == [rows, nnz]
    assert np.all(np.diff(coords[:, 0]) >= 0) and np.all(np.diff(coords[:, 1]) >= 0)
    share = -(-(rows + nnz) // p)
    for g in range(p):
        (x0, y0), (x1, y1) = coords[g], coords[g + 1]
        assert (x1 - x0) + (y1 - y0) <= share          # equal shares of the merge path
        lro = sharded.local_row_offsets(ro, coords, g)
        assert lro[0] == 0 and lro[-1] == y1 - y0 and np.all(np.diff(lro) >= 0)
        # complete rows of the shard keep their ends; the first local row may be a row's tail
        for i in range(x1 - x0):
            assert lro[i + 1] == ro[x0 + i + 1] - y0
