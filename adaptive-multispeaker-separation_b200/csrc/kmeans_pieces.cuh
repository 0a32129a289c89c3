// Flat tile decomposition of the tensor-core k-means pass: the Bg * ntiles tiles of a group are cut into G equal
// contiguous ranges, one per CTA (a CTA may cross from one mixture into the next); mixture b is covered by the
// "pieces" c_first(b) .. c_first(b) + n(b) - 1, whose partial sums are reduced in that order.
#pragma once
#include <cstdint>

namespace amss {

__host__ __device__ inline int64_t kt_start(int64_t c, int64_t Tt, int G) { return Tt * c / G; }

// pieces of mixture b = the CTAs whose range [kt_start(c), kt_start(c+1)) meets [b*nt, (b+1)*nt)
__host__ __device__ inline void kt_pieces(int b, int64_t nt, int64_t Tt, int G, int& c_first, int& n) {
    const int64_t lo = (int64_t)b * nt, hi = lo + nt;
    int c = (int)(lo * G / Tt);
    while (c > 0 && kt_start(c, Tt, G) > lo) --c;
    while (c + 1 < G && kt_start(c + 1, Tt, G) <= lo) ++c;
    c_first = c;
    int d = c;
    while (d + 1 < G && kt_start(d + 1, Tt, G) < hi) ++d;
    n = d - c + 1;
}

}  // namespace amss
