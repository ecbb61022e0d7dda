// Host build of csrc/special_fold.hpp for tests/test_special_fold_cpu.py (the same source the CUDA kernels use).
#include "../avx-ecm_b200/csrc/special_fold.hpp"
using namespace ecmb200;

template <int NL>
static int run(uint32_t k, int kind, uint32_t c, const uint32_t *T, const uint32_t *n, uint32_t *r)
{
    if ((int)(k >> 5) < SpecialRange<NL>::LOW || (int)(k >> 5) >= NL) return -2;
    uint32_t t[2 * NL], out[NL];
    for (int i = 0; i < 2 * NL; i++) t[i] = T[i];
    (void)n; special_fold<NL>(out, t, k, kind, c);
    for (int i = 0; i < NL; i++) r[i] = out[i];
    return 0;
}

extern "C" int special_fold_host(int nl, uint32_t k, int kind, uint32_t c, const uint32_t *T, const uint32_t *n, uint32_t *r)
{
    switch (nl) {
    case 3: return run<3>(k, kind, c, T, n, r);
    case 6: return run<6>(k, kind, c, T, n, r);
    case 10: return run<10>(k, kind, c, T, n, r);
    case 13: return run<13>(k, kind, c, T, n, r);
    case 16: return run<16>(k, kind, c, T, n, r);
    case 20: return run<20>(k, kind, c, T, n, r);
    case 24: return run<24>(k, kind, c, T, n, r);
    case 32: return run<32>(k, kind, c, T, n, r);
    }
    return -1;
}
