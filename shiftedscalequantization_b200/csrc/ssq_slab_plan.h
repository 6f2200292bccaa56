// Grid plan of the K1c backward (column-owner reduction): column blocks of SSQ_THREADS columns (vector kernel: float4 columns)
// x slabs of rows. Plain C++ so that the CPU test suite can compile it with g++ (tests/test_host_cpu.py).
#pragma once
#include <stdint.h>

namespace ssq {

// vec_ctas > 0 (vector kernel, that many resident CTAs per SM guaranteed by its launch bounds): ONE wave of column-block x slab
// CTAs — the slab count is rounded DOWN (rounding up gave 612 CTAs for 592 slots: a second wave of 20 CTAs that cost as much as
// the first, 0.66 instead of 0.85 of the HBM peak); slabs of at least 4 rows so the partials stay small.
// vec_ctas == 0 (scalar fallback kernel): about one wave at ctas_per_sm, rounded up.
inline void slab_plan(int64_t oc, int64_t K, int& nslab, int64_t& rows_per_slab, int vec_ctas = 0,
                      int threads = 256, int num_sms = 148, int ctas_per_sm = 8) {
    const bool vec = vec_ctas > 0;
    int64_t colblocks = ((vec ? K / 4 : K) + threads - 1) / threads;
    int64_t want = vec ? ((int64_t)num_sms * vec_ctas) / colblocks
                       : ((int64_t)num_sms * ctas_per_sm + colblocks - 1) / colblocks;
    if (vec && want > (oc + 3) / 4) want = (oc + 3) / 4;
    if (want > oc) want = oc;
    if (want < 1) want = 1;
    rows_per_slab = (oc + want - 1) / want;
    nslab = (int)((oc + rows_per_slab - 1) / rows_per_slab);
}

}  // namespace ssq
