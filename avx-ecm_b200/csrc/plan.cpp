// plan.cpp -- host planners (see plan.hpp).  Own implementation; behaviour follows the
// reference functions cited at each routine so that the device executes the same sequence of
// field operations (SURVEY facts 2, 3, 8 and Appendices A/B).
#include "plan.hpp"
#include <cstdlib>
#include <thread>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>

namespace ecmb200 {

// ---------------------------------------------------------------------------------------------
// permutations of the four physical point slots, lexicographic; byte = A | B<<2 | C<<4 | T<<6
// ---------------------------------------------------------------------------------------------
const uint8_t kPermTable[24] = {
    0xE4, 0xB4, 0xD8, 0x78, 0x9C, 0x6C, 0xE1, 0xB1, 0xC9, 0x39, 0x8D, 0x2D,
    0xD2, 0x72, 0xC6, 0x36, 0x4E, 0x1E, 0x93, 0x63, 0x87, 0x27, 0x4B, 0x1B };

int perm_index(const int s[4])
{
    const uint8_t b = (uint8_t)(s[0] | (s[1] << 2) | (s[2] << 4) | (s[3] << 6));
    for (int i = 0; i < 24; i++) if (kPermTable[i] == b) return i;
    return -1;
}

// ---------------------------------------------------------------------------------------------
// primes (replaces the YAFU sieve calls GetPRIMESRange / soe_wrapper, ecm.c:1139, main.c:568-583;
// any exact prime list is equivalent)
// ---------------------------------------------------------------------------------------------
std::vector<uint64_t> primes_in_range(uint64_t lo, uint64_t hi)
{
    std::vector<uint64_t> out;
    if (hi <= lo) return out;
    if (lo <= 2 && hi > 2) out.push_back(2);
    uint64_t root = (uint64_t)std::sqrt((double)hi) + 2;
    std::vector<uint8_t> small(root + 1, 1);
    std::vector<uint32_t> base;
    for (uint64_t p = 3; p <= root; p += 2) {
        if (!small[p]) continue;
        base.push_back((uint32_t)p);
        for (uint64_t q = p * p; q <= root; q += 2 * p) small[q] = 0;
    }
    const uint64_t SEG = 1u << 21;                  // odd numbers per segment
    uint64_t start = std::max<uint64_t>(lo | 1, 3); // first odd candidate
    std::vector<uint8_t> seg(SEG);
    for (uint64_t s0 = start; s0 < hi; s0 += 2 * SEG) {
        uint64_t s1 = std::min(hi, s0 + 2 * SEG);   // candidates s0, s0+2, ... < s1
        uint64_t cnt = (s1 - s0 + 1) / 2;
        std::fill(seg.begin(), seg.begin() + cnt, 1);
        for (uint32_t p : base) {
            uint64_t pp = (uint64_t)p * p;
            if (pp >= s1) break;
            uint64_t f = (s0 + p - 1) / p * p;
            if (f < pp) f = pp;
            if ((f & 1) == 0) f += p;
            for (uint64_t q = f; q < s1; q += 2 * (uint64_t)p) seg[(q - s0) >> 1] = 0;
        }
        for (uint64_t i = 0; i < cnt; i++) if (seg[i]) out.push_back(s0 + 2 * i);
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// PRAC chain: the Lucas chain for multiplier c started from r = round(c * v) is a sequence of
// (optional swap, rule) steps.  walk() visits it; costing and emission share the walker so that
// they cannot disagree.  Rules live: 3, 4, 5, 9 (ORIG_PRAC undefined, ecm.c:467,683-865).
// ---------------------------------------------------------------------------------------------
static const double kPracV[10] = {           // ecm.c:473-477
    0.61803398874989485, 0.72360679774997897, 0.58017872829546410, 0.63283980608870629,
    0.61242994950949500, 0.62018198080741576, 0.61721461653440386, 0.61834711965622806,
    0.61791440652881789, 0.61807966846989581 };

enum Rule { R3, R4, R5, R9 };

template <class Visitor>
static bool walk(uint64_t c, double v, Visitor &&visit)
{
    uint64_t r = (uint64_t)((double)c * v + 0.5);     // ecm.c:486 / 584
    if (r >= c) return false;
    uint64_t d = c - r, e = 2 * r - c;
    while (d != e) {
        bool swapped = false;
        if (d < e) { std::swap(d, e); swapped = true; }
        Rule rule;
        if ((d + 3) / 4 <= e) { d -= e; rule = R3; }
        else if (((d + e) & 1) == 0) { d = (d - e) / 2; rule = R4; }
        else if ((d & 1) == 0) { d /= 2; rule = R5; }
        else { e /= 2; rule = R9; }
        visit(swapped, rule);
    }
    return d == 1;
}

// cost in half-units: add = 5.5 -> 11, dup = 4.5 -> 9 (ecm.c:459-460); exact, so the arg-min and its
// first-wins tie-break equal the reference's double comparison (ecm.c:574-582).
static uint64_t chain_cost(uint64_t c, double v)
{
    uint64_t cost = 9 + 11;
    uint64_t r = (uint64_t)((double)c * v + 0.5);
    if (r >= c) return 11 * c;
    bool ok = walk(c, v, [&](bool, Rule rule) { cost += (rule == R3) ? 11 : 20; });
    return ok ? cost : 2 * 999999999ull;
}

static int best_multiplier(uint64_t c)
{
    uint64_t best = 11 * c;
    int idx = 0;
    for (int i = 0; i < 10; i++) {
        uint64_t k = chain_cost(c, kPracV[i]);
        if (k < best) { best = k; idx = i; }
    }
    return idx;
}

namespace {
struct Emitter {
    Stage1Plan &plan;
    int slot[4];          // physical point slot of logical A,B,C,T
    int pslot;            // slot holding P between primes
    void put(uint32_t type) { plan.ops.push_back((uint8_t)(type | (perm_index(slot) << 3))); }
    void make_role(int logical, int phys)
    {   // permute so that `logical` sits on `phys`
        for (int i = 0; i < 4; i++) if (slot[i] == phys) { std::swap(slot[i], slot[logical]); return; }
    }
};
enum { LA = 0, LB = 1, LC = 2, LT = 3 };
enum { M_DBL = 0, M_INIT = 1, M_C3 = 2, M_C4 = 3, M_C5 = 4, M_C9 = 5, M_FINAL = 6, M_NOP = 7 };
}  // namespace

// The reference runs stage 1 once per range of kStage1Range = 1e8 primes-by-value (vececm, ecm.c:1207-1234), and
// every call of ecm_stage1 (ecm.c:1806-1854) (a) repeats the doublings for q = 2,4,.. < B1 and (b) starts at
// PRIMES[1] of the range it was handed.  For B1 <= 1e8 that is one call that skips the prime 2 (covered by the
// doublings).  Beyond, the stream below reproduces the quirks: the doublings are emitted once per range and the
// first prime of every later range (100000007, 200000033, ...) is left out.
uint64_t stage1_prime_range()
{
    // PRIME_RANGE of the reference (main.c:585).  ECM_B200_S1_RANGE is a test hook: small ranges let the
    // tests drive the range-by-range path (checkpoints, repeated doublings, skipped primes) at small B1.
    if (const char *e = getenv("ECM_B200_S1_RANGE")) { const uint64_t v = strtoull(e, nullptr, 10); if (v >= 16) return v; }
    return 100000000ull;
}

void plan_stage1(uint64_t b1, Stage1Plan &plan)
{
    plan = Stage1Plan();
    plan.b1 = b1;
    Emitter em{plan, {1, 2, 3, 0}, 0};
    const uint64_t kStage1Range = stage1_prime_range();
    for (uint64_t lo = 0; lo < b1; lo += kStage1Range) {
        // powers of two: one doubling per q = 2,4,8,... < B1 (ecm.c:1815-1822)
        for (uint64_t q = 2; q < b1; q *= 2) {
            em.make_role(LT, em.pslot);
            em.put(M_DBL);
            plan.ptdups++;
        }
        // the primes of this range below B1 without the first one, prac(p) repeated while p^k * p < B1 (ecm.c:1824-1832)
        std::vector<uint64_t> primes = primes_in_range(lo, std::min(b1, lo + kStage1Range));
        // the multiplier search (ten trial chains per prime, ecm.c:574-584) is nine tenths of the planning time and
        // independent per prime: do it on all host threads, then emit sequentially
        std::vector<uint8_t> mult(primes.size(), 0);
        {
            const size_t np = primes.size();
            unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
            if (np < 20000) nt = 1;
            auto work = [&](size_t a, size_t b) { for (size_t i = std::max<size_t>(a, 1); i < b; i++) mult[i] = (uint8_t)best_multiplier(primes[i]); };   // primes[0] is never used
            std::vector<std::thread> th;
            for (unsigned t = 1; t < nt; t++) th.emplace_back(work, np * t / nt, np * (t + 1) / nt);
            work(0, np / nt);
            for (auto &x : th) x.join();
        }
        uint64_t last = 0;
        for (size_t i = 1; i < primes.size(); i++) {
            const uint64_t p = primes[i];
            last = p;
            uint64_t c = 1;
            do {
                const double v = kPracV[mult[i]];
                em.make_role(LB, em.pslot);           // B = P (no copy), C = copy, A = 2P
                em.put(M_INIT);
                plan.ptdups++;
                walk(p, v, [&](bool swapped, Rule rule) {
                    if (swapped) std::swap(em.slot[LA], em.slot[LB]);
                    switch (rule) {
                    case R3: {
                        em.put(M_C3);
                        int oldB = em.slot[LB];
                        em.slot[LB] = em.slot[LT]; em.slot[LT] = em.slot[LC]; em.slot[LC] = oldB;
                        plan.ptadds++;
                        break;
                    }
                    case R4: em.put(M_C4); plan.ptadds++; plan.ptdups++; break;
                    case R5: em.put(M_C5); plan.ptadds++; plan.ptdups++; break;
                    case R9: em.put(M_C9); plan.ptadds++; plan.ptdups++; break;
                    }
                });
                em.put(M_FINAL);
                plan.ptadds++;
                em.pslot = em.slot[LT];
                c *= p;
            } while (c * p < b1);
        }
        // where checkpoint.txt would be written (ecm.c:1237-1311): stream position, slot of P, last prime done;
        // padded so that the next range starts on a 16-byte boundary of the stream
        if (lo + kStage1Range < b1) while (plan.ops.size() % 16) plan.ops.push_back(M_NOP);
        plan.range_end.push_back({(uint64_t)plan.ops.size(), em.pslot, last});
    }
    plan.n_ops = plan.ops.size();
    plan.final_slot = em.pslot;
    while (plan.ops.size() % 16) plan.ops.push_back(M_NOP);
}

// ---------------------------------------------------------------------------------------------
// stage-2 geometry (thread_init, main.c:834-882; U=16 is what the reference binary ends up
// with) and the baby-step index map (ecm_work_init, ecm.c:301-329)
// ---------------------------------------------------------------------------------------------
static uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; }

Stage2Params stage2_params(uint64_t b1)
{
    static const struct { uint64_t lim; uint32_t D; } tab[] = {
        {60, 30}, {128, 60}, {256, 120}, {512, 210}, {2048, 385}, {4096, 1155} };
    Stage2Params p;
    p.D = 2310;
    for (const auto &t : tab) if (b1 <= t.lim) { p.D = t.D; break; }
    uint32_t phi = 0;
    for (uint32_t i = 0; i < 2 * p.D; i++) if (gcd_u32(i, 2 * p.D) == 1) phi++;
    p.R = phi + 3;
    p.U = 16;
    p.L = 2 * p.U;
    return p;
}

std::vector<uint32_t> stage2_map(const Stage2Params &p, uint32_t *n_stored)
{
    std::vector<uint32_t> map((size_t)p.U * (p.D + 1) + 3, 0);
    map[1] = 1; map[2] = 2;
    uint32_t next = 3;
    for (uint32_t blk = 0; blk < p.U; blk++) {
        for (uint32_t j = (blk == 0) ? 3 : 1; j < p.D; j++)
            map[blk * p.D + j] = (gcd_u32(j, p.D) == 1) ? next++ : 0;
        if (blk == 0) map[p.D] = next++;
    }
    if (n_stored) *n_stored = next;
    return map;
}

// ---------------------------------------------------------------------------------------------
// Montgomery's PAIR (pair(), ecm.c:2559-2910).  A prime s = 2*a*w + q is held in the FIFO of its
// residue class until a prime of the mirror class arrives; the two are then covered by one
// product term.  Terms are emitted as (v,u) with Pa index v - amin and baby-step index u;
// (0,0) tells the device to slide the giant-step window by 2U points.
// ---------------------------------------------------------------------------------------------
uint32_t pair_plan(uint64_t lo, uint64_t hi, const Stage2Params &p, std::vector<uint32_t> &pm_v,
                   std::vector<uint32_t> &pm_u, uint32_t *amin_final, uint32_t *npairs)
{
    const int64_t w = p.D, L = p.L, U = p.U, umax = w * U;
    std::vector<int> cls(2 * w, -1);
    std::vector<uint32_t> residue;
    for (int64_t k = 0; k < 2 * w; k++)
        if (gcd_u32((uint32_t)k, (uint32_t)(2 * w)) == 1) { cls[k] = (int)residue.size(); residue.push_back((uint32_t)k); }
    std::vector<std::deque<uint32_t>> fifo(residue.size());
    auto folded = [&](uint32_t r) { return r > (uint32_t)w ? (uint32_t)(2 * w - r) : r; };

    pm_v.clear(); pm_u.clear();
    uint32_t pairs = 0;
    uint64_t amin = (lo + w) / (2 * w);
    auto emit = [&](uint64_t v, uint64_t u) { pm_v.push_back((uint32_t)v); pm_u.push_back((uint32_t)u); };

    std::vector<uint64_t> primes = primes_in_range(lo, hi);
    for (uint64_t s : primes) {
        const uint64_t a = (s + w) / (2 * w);
        while (a >= amin + L) {                     // window must advance: flush what falls behind
            const uint64_t oldmin = amin;
            amin += L - U;
            for (size_t i = 0; i < fifo.size(); i++) {
                size_t n = fifo[i].size();
                while (n--) {
                    uint32_t ap = fifo[i].front(); fifo[i].pop_front();
                    if (ap < amin) { emit(2 * (uint64_t)ap - oldmin, folded(residue[i])); pairs++; }
                    else fifo[i].push_back(ap);
                }
            }
            emit(0, 0);
        }
        const int64_t q = (int64_t)s - 2 * (int64_t)a * w;
        const int64_t mirror = (q < 0) ? -q : 2 * w - q;
        for (;;) {
            std::deque<uint32_t> &mq = fifo[cls[mirror]];
            if (mq.empty()) {
                fifo[cls[q < 0 ? 2 * w + q : q]].push_back((uint32_t)a);
                break;
            }
            const uint64_t ap = mq.front(); mq.pop_front();
            const uint64_t u = (uint64_t)w * (a - ap) + (uint64_t)q;   // two's complement handles q < 0
            pairs++;
            if (u > (uint64_t)umax) {               // partner too far: it goes out alone
                int64_t qq = (q < 0) ? -q : q;
                if (q >= 0 && qq >= w) qq = 2 * w - qq;
                emit(2 * ap - amin, (uint64_t)qq);
                continue;
            }
            emit(a + ap - amin, u);
            break;
        }
    }
    for (size_t i = 0; i < fifo.size(); i++)
        for (uint32_t ap : fifo[i]) { emit(2 * (uint64_t)ap - amin, folded(residue[i])); pairs++; }
    if (amin_final) *amin_final = (uint32_t)amin;
    if (npairs) *npairs = pairs;
    return (uint32_t)pm_v.size();
}

}  // namespace ecmb200
