// Host-side LANCZOS coefficient tables (see coeffs.cpp).
#pragma once
#include <cstdint>

namespace b200comp {
// 2*ceil(3*max(1, in/out)) + 1
int lanczos_ksize(int in_size, int out_size);
// k: out_size*ksize int32 (22-bit fixed point, zero padded), bounds: out_size*(lo, n). Returns ksize.
int build_lanczos_table(int in_size, int out_size, int32_t *k, int32_t *bounds);
// Taps of a single output sample xx: row gets ksize int32, *lo / *n its first source sample and tap count.
int build_lanczos_row(int in_size, int out_size, int xx, int32_t *row, int *lo, int *n);
}  // namespace b200comp
