// tma_stage.cuh -- staging helpers shared by both engines: copy a slice [lo, hi) of a global array
// into shared memory with cp.async.bulk (TMA) although the slice may start and end at any element.
#pragma once

#include "merge_common.cuh"

namespace mspmv {

// ------------------------------------------------------------------------------------------------
// Producer helper: stage elements [lo, hi) of a global array into a ring.
// Ring position of element i is (i + shift) & mask where shift = element misalignment of the
// array base w.r.t. 16 bytes, so 16-byte-aligned global addresses land on 16-byte-aligned ring
// positions.  The aligned middle goes by one bulk copy (lane 0); the < 16-byte ragged ends are
// copied by lanes 1..; returns the bulk byte count (valid in lane 0).  The range never wraps the
// ring except through the ragged tail, which is masked per element.
// ------------------------------------------------------------------------------------------------
template <typename E>
__device__ __forceinline__ uint32_t stage_range(const E* __restrict__ base, int shift, int lo, int hi,
                                                E* ring, int mask, uint64_t* bar, uint64_t policy,
                                                int lane, int pos_base = 0)
{
    constexpr int GRAN = 16 / (int)sizeof(E);
    if (lo >= hi) return 0;
    int lo_al = lo + ((GRAN - ((lo + shift) & (GRAN - 1))) & (GRAN - 1));  // first aligned index >= lo
    int hi_al = hi - ((hi + shift) & (GRAN - 1));                          // last aligned index <= hi
    uint32_t bytes = 0;
    if (lo_al < hi_al) {
        bytes = (uint32_t)(hi_al - lo_al) * (uint32_t)sizeof(E);
    } else {
        lo_al = hi;  // no aligned middle: everything is ragged head
        hi_al = hi;
    }
    // ragged head [lo, lo_al) and tail [hi_al, hi): at most GRAN-1 elements each
    int nhead = lo_al - lo, ntail = hi - hi_al;
    if (lane < nhead) {
        int i = lo + lane;
        ring[((i + shift) & mask) - pos_base] = base[i];
    } else if (lane >= 8 && lane - 8 < ntail) {
        int i = hi_al + (lane - 8);
        ring[((i + shift) & mask) - pos_base] = base[i];
    }
    if (lane == 0 && bytes)
        bulk_g2s(ring + (((lo_al + shift) & mask) - pos_base), base + lo_al, bytes, bar, policy);
    return bytes;
}

// ------------------------------------------------------------------------------------------------
// Stage elements [lo, hi) of a global array of n_total elements into a LINEAR buffer with one bulk
// copy of the 16-byte-aligned superset [lo_al, hi_al) whenever that superset stays inside the
// array (always, except possibly at the very first / last elements of the array, which fall back
// to scalar copies by lanes 1..).  Element i lands at buf[(i + shift) - pos_base]; pos_base must
// be the aligned-down position of lo and the buffer needs 16 bytes of slack at its end.
// Returns the bulk byte count (valid in lane 0).
// ------------------------------------------------------------------------------------------------
template <typename E>
__device__ __forceinline__ uint32_t stage_superset(const E* __restrict__ base, int shift, int lo, int hi,
                                                   int n_total, E* buf, int pos_base, uint64_t* bar,
                                                   uint64_t policy, int lane)
{
    constexpr int GRAN = 16 / (int)sizeof(E);
    if (lo >= hi) return 0;
    int lo_al = lo - ((lo + shift) & (GRAN - 1));                          // aligned index <= lo
    int hi_al = hi + ((GRAN - ((hi + shift) & (GRAN - 1))) & (GRAN - 1));  // aligned index >= hi
    if (lo_al < 0) {        // array starts mid-granule: copy the head by hand
        lo_al += GRAN;
        const int i = lo + lane;
        if (i < min(lo_al, hi)) buf[i + shift - pos_base] = base[i];
    }
    if (hi_al > n_total) {  // array ends mid-granule: copy the tail by hand
        hi_al -= GRAN;
        const int i = max(hi_al, lo) + lane;
        if (i < hi && hi_al >= lo_al) buf[i + shift - pos_base] = base[i];
    }
    if (lo_al >= hi_al) {
        // no aligned middle at all (tiny array): everything not covered above goes by hand
        for (int i = lo + lane; i < hi; i += 32) buf[i + shift - pos_base] = base[i];
        return 0;
    }
    const uint32_t bytes = (uint32_t)(hi_al - lo_al) * (uint32_t)sizeof(E);
    if (lane == 0) bulk_g2s(buf + (lo_al + shift - pos_base), base + lo_al, bytes, bar, policy);
    return bytes;
}

// L2 prefetch of the aligned superset of elements [lo, hi) of an array (a pure hint: the tile that
// will need it starts about one block lifetime later and then finds its slice in L2, not HBM).
template <typename E>
__device__ __forceinline__ void l2_prefetch_range(const E* __restrict__ base, int shift, int lo, int hi,
                                                  int n_total)
{
    constexpr int GRAN = 16 / (int)sizeof(E);
    if (lo >= hi) return;
    int lo_al = lo - ((lo + shift) & (GRAN - 1));
    int hi_al = hi + ((GRAN - ((hi + shift) & (GRAN - 1))) & (GRAN - 1));
    if (lo_al < 0) lo_al += GRAN;
    if (hi_al > n_total) hi_al -= GRAN;
    if (lo_al >= hi_al) return;
    const uint32_t bytes = (uint32_t)(hi_al - lo_al) * (uint32_t)sizeof(E);
    bulk_prefetch_l2(base + lo_al, bytes);
}

}  // namespace mspmv
