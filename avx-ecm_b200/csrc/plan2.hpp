// plan2.hpp -- host compiler for stage 2: turns ecm_stage2_init / ecm_stage2_pair
// (ecm.c:2201-2540) plus the PAIR output into a linear program for the stage-2 field-op machine
// (k_vm2 in kernels.cuh).  Like the PRAC plan of stage 1 it depends only on (B1, B2): every curve
// of the batch executes the same instruction stream.
#pragma once
#include "plan.hpp"
#include <cstdint>
#include <vector>

namespace ecmb200 {

// ---- instruction word: lo = op | d<<8 | x<<16 | y<<24 ; hi = imm ---------------------------------
enum V2Op : uint32_t {
    V_MUL = 0, V_SQR = 1, V_ADD = 2, V_SUB = 3,
    V_ADDSUB = 4,     // d = x+y, slot imm = x-y
    V_COPY = 5,       // d = x
    V_LDG = 6,        // d = table[imm]
    V_STG = 7,        // table[imm] = x
    V_INV = 8,        // d = 1/x (Montgomery form in and out); failure: reference lane-0 semantics
    V_ONE = 9,        // d = R mod N  (Montgomery 1)
    V_PAIR = 10,      // acc *= table[imm & 0xffff] - table[imm >> 16]      (CROSS_PRODUCT_INV, ecm.c:1857-1859)
    V_NOP = 11,
    V_MUL2 = 12,      // two independent products: lo = op | d<<8 | x<<12 | y<<16 | e<<20 | u<<24 | v<<28 (4-bit slots)
};
// slot file of the stage-2 machine
enum V2Slot : uint32_t {
    UX = 0, UZ, VX, VZ, WX, WZ,        // three work points            } kept in global memory (L2) by the
    ACC = 6,                           // stage-2 accumulator           } hybrid slot file of k_vm2
    S1_ = 7, D1_, S2_, D2_,            // sums / differences            } shared memory
    SP_ = 11,                          // (A+2)/4
    T1_ = 12, T2_ = 13,
    NSLOT_S2 = 14
};

struct Stage2Layout {                  // table entry index space, entry e of curve c: tab[(e*NL+limb)*cap+c]
    uint32_t npb;                      // stored baby-step slots (index 0 is scratch)
    uint32_t pbx, pbz, pba;            // bases of the X / Z / prefix-product tables, npb entries each
    uint32_t pax, paz, pai, paa;       // giant-step window: X, Z, X/Z, prefix products; 2L entries each
    uint32_t qx, qz, pdx, pdz;         // Q = stage-1 result, Pd = [w]Q
    uint32_t entries;
};

struct Stage2Program {
    Stage2Params prm;
    Stage2Layout lay;
    std::vector<uint64_t> init;                    // ecm_stage2_init
    std::vector<std::vector<uint64_t>> ranges;     // one per 1e8 prime range: ecm_stage2_pair
    // the reference's counters (ecm.c:1482) for cross-checking
    uint64_t ptadds = 0, numinv = 0, paired = 0, pairmap_steps = 0;
    uint32_t last_amin = 0;
};

Stage2Layout stage2_layout(const Stage2Params &p);
void plan_stage2_init(uint64_t b1, Stage2Program &prog);
// compiles the program of one prime range [lo,hi) (lo = B1 for the first; amin restarts per range) into
// prog.ranges[index] (index < 0: append).  Ranges may be compiled by a background thread while the GPU
// executes the previous one; the counters are only meaningful once all ranges are done.
void plan_stage2_range(uint64_t lo, uint64_t hi, Stage2Program &prog, int index = -1);
// the same for a caller-supplied pairmap starting at window index amin (ecm_stage2_pair's own arguments,
// ecm.c:2342-2351); validate untrusted pairmaps with stage2_pairmap_valid first
void plan_stage2_pairmap(uint32_t amin, const uint32_t *pm_v, const uint32_t *pm_u, uint32_t steps, Stage2Program &prog, int index = -1);
bool stage2_pairmap_valid(const Stage2Params &p, uint32_t amin, const uint32_t *pm_v, const uint32_t *pm_u, uint32_t steps);

}  // namespace ecmb200
