// tma_stage.cuh -- staging helper of the producer warp: copy a slice [lo, hi) of a global array into
// shared memory with cp.async.bulk (TMA) although the slice may start and end at any element.
#pragma once

#include "merge_common.cuh"

namespace mspmv {

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

}  // namespace mspmv
