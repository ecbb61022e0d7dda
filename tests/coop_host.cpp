// coop_host.cpp -- TEST INFRASTRUCTURE: host build of avx-ecm_b200/csrc/coop.cuh.  The L lanes of one group run as L host
// threads in lock step; shuffles and votes go through a slot array guarded by a barrier, so the routines execute exactly
// the call sequence the GPU lanes do.  stdin: "<op> <L> <M> <a> <b> <n>" (hex; op = mul | add | sub), stdout: result hex.
#include <barrier>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>
#include <iostream>
#include <sstream>
#include "../avx-ecm_b200/csrc/coop.cuh"

using namespace ecmb200;

struct Group {
    int L;
    std::vector<uint32_t> slot;
    std::barrier<> bar;
    explicit Group(int l) : L(l), slot(l), bar(l) {}
};
struct HostComm {
    uint32_t part;
    Group *g;
    int L;
    uint32_t xchg(uint32_t v, int src) const
    {
        g->slot[part] = v;
        g->bar.arrive_and_wait();
        const uint32_t r = (src >= 0 && src < L) ? g->slot[src] : 0u;
        g->bar.arrive_and_wait();
        return r;
    }
    uint32_t shfl(uint32_t v, uint32_t src) const { return xchg(v, (int)src); }
    uint32_t from_above(uint32_t v) const { return xchg(v, (int)part + 1); }
    uint32_t from_below(uint32_t v) const { return xchg(v, (int)part - 1); }
    uint32_t vote(bool p) const
    {
        g->slot[part] = p ? 1u : 0u;
        g->bar.arrive_and_wait();
        uint32_t m = 0;
        for (int i = 0; i < L; i++) m |= g->slot[i] << i;
        g->bar.arrive_and_wait();
        return m;
    }
};

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

template <int M, int L>
static Big run(const std::string &op, const Big &a, const Big &b, const Big &n)
{
    Group g(L);
    Big r(M * L, 0);
    uint32_t inv = 1;
    for (int i = 0; i < 5; i++) inv *= 2 - n[0] * inv;
    const uint32_t m0inv = 0u - inv;
    std::vector<std::thread> th;
    for (int p = 0; p < L; p++)
        th.emplace_back([&, p]() {
            HostComm cm{(uint32_t)p, &g, L};
            uint32_t A[M], B[M], N[M], R[M];
            for (int k = 0; k < M; k++) { A[k] = a[p * M + k]; B[k] = b[p * M + k]; N[k] = n[p * M + k]; }
            if (op == "mul") coop::mont_mul<M, L>(R, A, B, N, m0inv, cm);
            else if (op == "add") coop::mod_add<M, L>(R, A, B, N, cm);
            else coop::mod_sub<M, L>(R, A, B, N, cm);
            for (int k = 0; k < M; k++) r[p * M + k] = R[k];
        });
    for (auto &t : th) t.join();
    return r;
}

int main()
{
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream is(line);
        std::string op, ah, bh, nh;
        int L, M;
        if (!(is >> op >> L >> M >> ah >> bh >> nh)) continue;
        const int n = L * M;
        const Big a = parse(ah, n), b = parse(bh, n), N = parse(nh, n);
        Big r;
#define CASE(m, l) if (M == m && L == l) r = run<m, l>(op, a, b, N);
        CASE(16, 4) CASE(14, 4) CASE(12, 4) CASE(10, 4) CASE(16, 2) CASE(8, 4) CASE(2, 2) CASE(4, 8) CASE(2, 4) CASE(10, 2) CASE(16, 8)
#undef CASE
        if (r.empty()) { printf("unsupported\n"); continue; }
        for (int k = n - 1; k >= 0; k--) printf("%08x", r[k]);
        printf("\n");
        fflush(stdout);
    }
    return 0;
}
