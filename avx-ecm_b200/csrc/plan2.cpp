// plan2.cpp -- stage-2 program generator (see plan2.hpp).  Own implementation; the emitted field-op
// sequence is the one the reference executes in ecm_stage2_init (ecm.c:2201-2340),
// batch_invert_pt_inplace / batch_invert_pt_to_bignum (ecm.c:1869-2136), next_pt_vec
// (ecm.c:886-976) and ecm_stage2_pair (ecm.c:2342-2540), so residues are identical.
#include "plan2.hpp"
#include <algorithm>

namespace ecmb200 {

namespace {

struct Pt { uint32_t x, z; };
const Pt PU{UX, UZ}, PV{VX, VZ}, PW{WX, WZ};

struct Asm {
    std::vector<uint64_t> &code;
    void put(uint32_t op, uint32_t d, uint32_t x, uint32_t y, uint32_t imm = 0)
    {
        code.push_back((uint64_t)(op | (d << 8) | (x << 16) | (y << 24)) | ((uint64_t)imm << 32));
    }
    void mul(uint32_t d, uint32_t x, uint32_t y) { put(V_MUL, d, x, y); }
    void sqr(uint32_t d, uint32_t x) { put(V_SQR, d, x, x); }
    void add(uint32_t d, uint32_t x, uint32_t y) { put(V_ADD, d, x, y); }
    void sub(uint32_t d, uint32_t x, uint32_t y) { put(V_SUB, d, x, y); }
    void addsub(uint32_t dsum, uint32_t ddiff, uint32_t x, uint32_t y) { put(V_ADDSUB, dsum, x, y, ddiff); }
    void copy(uint32_t d, uint32_t x) { put(V_COPY, d, x, 0); }
    void ldg(uint32_t d, uint32_t entry) { put(V_LDG, d, 0, 0, entry); }
    void stg(uint32_t x, uint32_t entry) { put(V_STG, 0, x, 0, entry); }
    void inv(uint32_t d, uint32_t x) { put(V_INV, d, x, 0); }
    void one(uint32_t d) { put(V_ONE, d, 0, 0); }
    void pair(uint32_t e_pa, uint32_t e_pb) { put(V_PAIR, 0, 0, 0, e_pa | (e_pb << 16)); }
    // d = x*y and e = u*v, independent (all four operands are read before either result is written)
    void mul2(uint32_t d, uint32_t x, uint32_t y, uint32_t e, uint32_t u, uint32_t v)
    {
        code.push_back((uint64_t)(V_MUL2 | (d << 8) | (x << 12) | (y << 16) | (e << 20) | (u << 24) | (v << 28)));
    }
    void ldpt(Pt p, uint32_t ex, uint32_t ez) { ldg(p.x, ex); ldg(p.z, ez); }
    void stpt(Pt p, uint32_t ex, uint32_t ez) { stg(p.x, ex); stg(p.z, ez); }

    // sums/differences of a point into one of the two pairs
    void sums1(Pt p) { addsub(S1_, D1_, p.x, p.z); }
    void sums2(Pt p) { addsub(S2_, D2_, p.x, p.z); }

    // vec_add (ecm.c:407-443): uses (s1,d1) and (s2,d2); clobbers s1,d1 (they hold the temporaries),
    // leaves s2,d2 intact.  out must differ from in.
    void vadd(Pt in, Pt out)
    {
        mul2(D1_, D1_, S2_, S1_, S1_, D2_);
        addsub(D1_, S1_, D1_, S1_);
        mul2(D1_, D1_, D1_, S1_, S1_, S1_);
        mul2(out.x, D1_, in.z, out.z, S1_, in.x);
    }
    // vec_duplicate (ecm.c:445-457) from sums (s,d); tmp is any dead slot; clobbers s,d
    void vdup(uint32_t s, uint32_t d, uint32_t tmp, Pt out)
    {
        mul2(d, d, d, s, s, s);
        sub(tmp, s, d);
        mul2(out.x, d, s, s, tmp, SP_);
        add(s, s, d);
        mul(out.z, s, tmp);
    }
};

// next_pt_vec (ecm.c:886-976): [c]Q by the binary ladder.  Q is read from the table, the result is
// left in point slot PU.  x1 = PU, x2 = PV, PW = Q.
void ladder(Asm &a, const Stage2Layout &L, uint64_t c, uint64_t &ptadds)
{
    a.ldpt(PU, L.qx, L.qz);
    if (c <= 1) return;          // c = 0 cannot be reached: callers reject amin = 0 (stage2_pairmap_valid, ecm_b200_stage2)
    a.sums1(PU);
    a.vdup(S1_, D1_, T1_, PV);                     // x2 = 2Q
    if (c == 2) { a.copy(UX, VX); a.copy(UZ, VZ); return; }
    a.ldpt(PW, L.qx, L.qz);
    for (int bit = 62 - __builtin_clzll(c); bit >= 0; bit--) {
        if ((c >> bit) & 1) {                      // x1 = x1 + x2 (Q) ; x2 = 2 x2
            a.sums2(PV);
            a.sums1(PU);
            a.vadd(PW, PU);
            a.vdup(S2_, D2_, T1_, PV);
        } else {                                   // x2 = x1 + x2 (Q) ; x1 = 2 x1
            a.sums1(PV);
            a.sums2(PU);
            a.vadd(PW, PV);
            a.vdup(S2_, D2_, T1_, PU);
        }
        ptadds++;
    }
}

// Montgomery's simultaneous inversion (batch_invert_pt_*, ecm.c:1869-2136) over table entries
// zs[0..n): out[i] = x[i] / z[i].  Prefix products go to the table `pref`; the running suffix
// inverse lives in slot T1.  `scratch` is one more dead slot.
// Back-substitution computes three products per entry -- 1/z[i] = B[i]*A[i-1], B[i-1] = z[i]*B[i], out[i] = x[i]/z[i] --
// and the third depends on the first.  Entry by entry that is a dual product plus a single one (half the carry chains in
// flight); taken two entries at a time the six products pair up as (1/z[i], B[i-1]), (out[i], 1/z[i-1]), (B[i-2], out[i-1]):
// the same products with the same operands, hence the same table contents, in three dual instructions.
void batch_invert(Asm &a, const std::vector<uint32_t> &xs, const std::vector<uint32_t> &zs,
                  const std::vector<uint32_t> &outs, uint32_t pref, uint32_t scratch, bool prefix_done = false)
{
    const size_t n = zs.size();
    if (!prefix_done) {                            // else the caller has A[0..n) in the table and A[n-1] in T1
        a.ldg(T1_, zs[0]);
        a.stg(T1_, pref);
        for (size_t i = 1; i < n; i++) {           // A[i] = z[i] * A[i-1]
            a.ldg(T2_, zs[i]);
            a.mul(T1_, T2_, T1_);
            a.stg(T1_, pref + (uint32_t)i);
        }
    }
    a.inv(T1_, T1_);                               // B[n-1]
    size_t i = n - 1;
    for (; i >= 2; i -= 2) {
        a.ldg(T2_, pref + (uint32_t)i - 1);
        a.ldg(D1_, zs[i]);
        a.mul2(T2_, T1_, T2_, T1_, D1_, T1_);      // 1/z[i] = B[i] * A[i-1]        |   B[i-1] = z[i] * B[i]
        a.ldg(S1_, xs[i]);
        a.ldg(scratch, pref + (uint32_t)i - 2);
        a.mul2(S1_, S1_, T2_, scratch, T1_, scratch);   // out[i] = x[i] / z[i]     |   1/z[i-1] = B[i-1] * A[i-2]
        a.stg(S1_, outs[i]);
        a.ldg(D1_, zs[i - 1]);
        a.ldg(S1_, xs[i - 1]);
        a.mul2(T1_, D1_, T1_, S1_, S1_, scratch);  // B[i-2] = z[i-1] * B[i-1]     |   out[i-1] = x[i-1] / z[i-1]
        a.stg(S1_, outs[i - 1]);
    }
    if (i == 1) {
        a.ldg(T2_, pref);
        a.ldg(D1_, zs[1]);
        a.mul2(T2_, T1_, T2_, T1_, D1_, T1_);
        a.ldg(S1_, xs[1]);
        a.mul(S1_, S1_, T2_);
        a.stg(S1_, outs[1]);
    }
    a.ldg(S1_, xs[0]);
    a.mul(S1_, S1_, T1_);
    a.stg(S1_, outs[0]);
}

}  // namespace

Stage2Layout stage2_layout(const Stage2Params &p)
{
    Stage2Layout L;
    uint32_t stored = 0;
    stage2_map(p, &stored);
    L.npb = stored;
    const uint32_t win = 2 * p.L;
    uint32_t e = 0;
    L.pbx = e; e += L.npb;
    L.pai = e; e += win;           // keep the two tables the pair loop reads below 2^16
    L.pbz = e; e += L.npb;
    L.pba = e; e += L.npb;
    L.pax = e; e += win;
    L.paz = e; e += win;
    L.paa = e; e += win;
    L.qx = e++; L.qz = e++; L.pdx = e++; L.pdz = e++;
    L.entries = e;
    return L;
}

// ecm_stage2_init (ecm.c:2201-2340)
void plan_stage2_init(uint64_t b1, Stage2Program &prog)
{
    prog.prm = stage2_params(b1);
    prog.lay = stage2_layout(prog.prm);
    prog.init.clear(); prog.ranges.clear();
    prog.ptadds = prog.numinv = prog.paired = prog.pairmap_steps = 0;
    const Stage2Params &p = prog.prm;
    const Stage2Layout &L = prog.lay;
    const std::vector<uint32_t> map = stage2_map(p, nullptr);
    Asm a{prog.init};

    // Pb[1] = Q ; Pb[2] = 2Q
    a.ldpt(PV, L.qx, L.qz);
    a.stpt(PV, L.pbx + 1, L.pbz + 1);
    a.sums1(PV);
    a.sums2(PV);                                   // (s2,d2) = sums of S1 = Q, constant during the build
    a.vdup(S1_, D1_, T1_, PU);
    a.stpt(PU, L.pbx + 2, L.pbz + 2);
    // running points: P1 = S_{j-1} (slot U), P3 = S_{j-2} (slot V), Pout (slot W); rotate names
    Pt P1 = PU, P3 = PV, Pout = PW;
    uint32_t last = 2;
    for (uint32_t j = 3; j <= p.U * p.D; j++) {
        a.sums1(P1);
        a.vadd(P3, Pout);                          // S_j = S_{j-1} + S_1, difference S_{j-2}
        prog.ptadds++;
        if (map[j] > 0) { a.stpt(Pout, L.pbx + map[j], L.pbz + map[j]); last = map[j]; }
        Pt t = P3; P3 = P1; P1 = Pout; Pout = t;
    }
    a.one(ACC);                                    // acc = one (ecm.c:2318)
    // batch_invert_pt_inplace(Pb, ..., last+1): entries 1..last
    {
        std::vector<uint32_t> xs, zs;
        for (uint32_t i = 1; i <= last; i++) { xs.push_back(L.pbx + i); zs.push_back(L.pbz + i); }
        batch_invert(a, xs, zs, xs, L.pba, WX);      // all three work points are dead here (the ladder below reloads Q)
        prog.numinv++;
    }
    // Pd = [w]Q
    ladder(a, L, p.D, prog.ptadds);
    a.stpt(PU, L.pdx, L.pdz);
}

// ecm_stage2_pair (ecm.c:2342-2540) for the primes of [lo,hi)
void plan_stage2_range(uint64_t lo, uint64_t hi, Stage2Program &prog, int index)
{
    std::vector<uint32_t> pm_v, pm_u;
    uint32_t amin_final = 0, npairs = 0;
    const uint32_t steps = pair_plan(lo, hi, prog.prm, pm_v, pm_u, &amin_final, &npairs);
    const uint32_t w = prog.prm.D;
    plan_stage2_pairmap((uint32_t)((lo + w) / (2 * (uint64_t)w)), pm_v.data(), pm_u.data(), steps, prog, index);
}

// Is (amin, pairmap) something ecm_stage2_pair could execute without leaving its tables?  (A caller-made pairmap is
// data from outside; the reference trusts it, a device program must not.)
bool stage2_pairmap_valid(const Stage2Params &p, uint32_t amin, const uint32_t *pm_v, const uint32_t *pm_u, uint32_t steps)
{
    if (amin == 0) return false;                      // A - w would be negative: B1 below D (the reference aborts there)
    const std::vector<uint32_t> map = stage2_map(p, nullptr);
    const uint32_t win = 2 * p.L;
    uint64_t a = amin;
    for (uint32_t k = 0; k < steps; k++) {
        if (pm_u[k] == 0 && pm_v[k] == 0) { a += p.U; continue; }
        if (pm_v[k] < a || pm_v[k] - a >= win) return false;
        if (pm_u[k] >= map.size() || map[pm_u[k]] == 0) return false;
    }
    return a < (1ull << 32);
}

// ecm_stage2_pair for a given starting amin and pairmap (the reference's own arguments, ecm.c:2342-2351)
void plan_stage2_pairmap(uint32_t amin, const uint32_t *pm_v, const uint32_t *pm_u, uint32_t steps, Stage2Program &prog, int index)
{
    const Stage2Params &p = prog.prm;
    const Stage2Layout &L = prog.lay;
    const std::vector<uint32_t> map = stage2_map(p, nullptr);
    const uint32_t w = p.D, U = p.U, win = 2 * p.L;
    if (index < 0) { prog.ranges.emplace_back(); index = (int)prog.ranges.size() - 1; }
    prog.ranges[index].clear();
    Asm a{prog.ranges[index]};
    prog.pairmap_steps += steps;

    const uint64_t A = (uint64_t)amin * w * 2;
    // the window is a ring of `win` table entries: logical index i lives at ring[(base+i) % win]
    uint32_t base = 0;
    auto ring = [&](uint32_t i) { return (base + i) % win; };

    // Pa[0] = [A]Q, Pad = [A-w]Q, Pa[1] = Pa[0] + Pd (Pad)
    ladder(a, L, A, prog.ptadds);
    a.stpt(PU, L.pax + ring(0), L.paz + ring(0));
    ladder(a, L, A - w, prog.ptadds);              // Pad in slot U
    a.copy(WX, UX); a.copy(WZ, UZ);                // Pad -> W
    a.ldpt(PV, L.pax + ring(0), L.paz + ring(0));  // Pa[0] -> V
    a.ldpt(PU, L.pdx, L.pdz);
    a.sums2(PU);                                   // (s2,d2) = sums of Pd, constant below
    a.sums1(PV);
    a.vadd(PW, PU);                                // Pa[1] -> U
    prog.ptadds++;
    a.stpt(PU, L.pax + ring(1), L.paz + ring(1));
    // chain: cur = Pa[i-1] (U), prev = Pa[i-2] (V), out (W)
    Pt cur = PU, prev = PV, out = PW;
    // Window points are extended and inverted in runs: the Z coordinates of the run go through Montgomery's simultaneous
    // inversion, whose prefix products A[r] = z[r] * A[r-1] are a serial chain of SINGLE products (one carry-chain pair in
    // flight per thread, and a table load in front of each: measured 11 us per product against 6.8 us per dual product in
    // the window-shift launches).  The chain only needs z[r], which sits in a work-point slot for two extensions after
    // point r was made, so the last dual product of every second extension is split and takes two prefix products along:
    // (out.x, A[r-2]) and (out.z, A[r-1]) instead of (out.x, out.z) -- same products, same operands, same table contents.
    uint32_t pf_from = 0, pf_done = 0;             // run of entries being inverted: first ring index, prefix products made
    auto extend = [&](uint32_t i) {
        a.sums1(cur);
        const uint32_t r = i - pf_from;            // entries pf_from .. i-1 exist; prev = entry i-2, cur = entry i-1
        if (r >= 2 && r - pf_done == 2) {
            a.mul2(D1_, D1_, S2_, S1_, S1_, D2_);  // vadd (ecm.c:407-443) with its last dual product split
            a.addsub(D1_, S1_, D1_, S1_);
            a.mul2(D1_, D1_, D1_, S1_, S1_, S1_);
            if (pf_done == 0) {
                a.copy(T1_, prev.z);               // A[0] = z[0]
                a.mul2(out.x, D1_, prev.z, out.z, S1_, prev.x);
            } else {
                a.mul2(out.x, D1_, prev.z, T1_, prev.z, T1_);
            }
            a.stg(T1_, L.paa + pf_done++);
            if (pf_done == 1) a.mul(T1_, cur.z, T1_);
            else a.mul2(out.z, S1_, prev.x, T1_, cur.z, T1_);
            a.stg(T1_, L.paa + pf_done++);
        } else {
            a.vadd(prev, out);
        }
        prog.ptadds++;
        a.stpt(out, L.pax + ring(i), L.paz + ring(i));
        Pt t = prev; prev = cur; cur = out; out = t;
    };
    auto invert = [&](uint32_t from, uint32_t to) {
        std::vector<uint32_t> xs, zs, outs;
        for (uint32_t i = from; i < to; i++) { xs.push_back(L.pax + ring(i)); zs.push_back(L.paz + ring(i)); outs.push_back(L.pai + ring(i)); }
        // the prefix products the extensions could not take along: entries to-2 (prev) and to-1 (cur)
        const uint32_t n = to - from;
        bool done = false;
        if (pf_from == from && n >= 2 && pf_done == n - 2) {
            if (pf_done == 0) a.copy(T1_, prev.z); else a.mul(T1_, prev.z, T1_);
            a.stg(T1_, L.paa + pf_done++);
            a.mul(T1_, cur.z, T1_);
            a.stg(T1_, L.paa + pf_done++);
            done = true;
        }
        batch_invert(a, xs, zs, outs, L.paa, out.x, done);   // `out` is the one work point the chain does not need any more
        prog.numinv++;
    };
    for (uint32_t i = 2; i < win; i++) extend(i);
    invert(0, win);
    prog.numinv++;                                 // the reference counts this one twice (ecm.c:2429)

    for (uint32_t k = 0; k < steps; k++) {
        if (pm_u[k] == 0 && pm_v[k] == 0) {        // slide the window by 2U points (ecm.c:2458-2502)
            base = (base + 2 * U) % win;
            pf_from = win - 2 * U; pf_done = 0;
            for (uint32_t i = win - 2 * U; i < win; i++) extend(i);
            amin += U;
            invert(win - 2 * U, win);
        } else {
            const uint32_t pa = pm_v[k] - amin, pb = pm_u[k];
            a.pair(L.pai + ring(pa), L.pbx + map[pb]);
            prog.paired++;
        }
    }
    prog.last_amin = amin;
}

}  // namespace ecmb200
