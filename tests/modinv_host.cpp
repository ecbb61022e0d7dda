// modinv_host.cpp -- TEST INFRASTRUCTURE: host build of avx-ecm_b200/csrc/modinv_fast.hpp.
// stdin: "<NL> <want_inv 0|1> <y hex> <n hex>" per line; stdout: "<ok> <inv hex> <gcd hex>".
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "../avx-ecm_b200/csrc/modinv_fast.hpp"
using namespace ecmb200;
typedef std::vector<uint32_t> Big;
static Big parse(const std::string &h, int n)
{
    Big v(n, 0);
    int pos = 0;
    for (int i = (int)h.size() - 1; i >= 0; i--, pos++) {
        const char c = h[i];
        const uint32_t d = (c >= '0' && c <= '9') ? c - '0' : (c | 32) - 'a' + 10;
        if (pos / 8 < n) v[pos / 8] |= d << (4 * (pos % 8));
    }
    return v;
}
static void put(const Big &v) { for (int k = (int)v.size() - 1; k >= 0; k--) printf("%08x", v[k]); }
template <int NL> static void run(int want, const Big &y, const Big &n)
{
    Big inv(NL, 0), g(NL, 0);
    const bool ok = want ? fastinv::mod_inverse<NL, true>(inv.data(), g.data(), y.data(), n.data())
                         : fastinv::mod_inverse<NL, false>(inv.data(), g.data(), y.data(), n.data());
    printf("%d ", ok ? 1 : 0); put(inv); printf(" "); put(g); printf("\n");
}
int main()
{
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream is(line);
        int NL, want; std::string yh, nh;
        if (!(is >> NL >> want >> yh >> nh)) continue;
        const Big y = parse(yh, NL), n = parse(nh, NL);
        switch (NL) {
#define C(k) case k: run<k>(want, y, n); break;
            C(1) C(2) C(3) C(6) C(10) C(13) C(16) C(20) C(24) C(32) C(48) C(64)
#undef C
            default: printf("unsupported\n");
        }
    }
    return 0;
}
