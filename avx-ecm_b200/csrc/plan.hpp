// plan.hpp -- host-side planners: the scalar control flow of avx-ecm that does not depend on
// the curve (prime generation, PRAC chain selection, Montgomery's PAIR) compiled once per
// (B1,B2) into op streams for the device.  Built with -ffp-contract=off: the PRAC multiplier
// choice is defined by plain IEEE double arithmetic (ecm.c:486,584).
#pragma once
#include <cstdint>
#include <vector>

namespace ecmb200 {

// shared with the device (vm.cuh): permutation index -> physical point slot of (A,B,C,T)
extern const uint8_t kPermTable[24];
int perm_index(const int slots[4]);

// all primes p with lo <= p < hi
std::vector<uint64_t> primes_in_range(uint64_t lo, uint64_t hi);

struct Stage1Plan {
    uint64_t b1 = 0;
    std::vector<uint8_t> ops;   // macro-op bytes (type | perm<<3), padded with M_NOP to a multiple of 16
    uint64_t n_ops = 0;         // unpadded length
    uint64_t ptadds = 0, ptdups = 0;
    int final_slot = 0;         // physical point slot that holds P after the stream
    // one entry per 1e8 prime range (ecm.c:1207-1234): ops emitted up to its end, slot of P there, last prime used
    struct RangeEnd { uint64_t ops; int slot; uint64_t last_prime; };
    std::vector<RangeEnd> range_end;
};
// width of the prime ranges stage 1 is run in (1e8 like the reference; see plan.cpp)
uint64_t stage1_prime_range();
// P starts in physical point slot 0.
void plan_stage1(uint64_t b1, Stage1Plan &plan);

struct Stage2Params { uint32_t D, U, L, R; };
Stage2Params stage2_params(uint64_t b1);
// index map of stored baby-step points (ecm.c:301-329); size U*(D+1)+3
std::vector<uint32_t> stage2_map(const Stage2Params &p, uint32_t *n_stored);
// Montgomery PAIR over the primes of [lo,hi)
uint32_t pair_plan(uint64_t lo, uint64_t hi, const Stage2Params &p, std::vector<uint32_t> &pm_v,
                   std::vector<uint32_t> &pm_u, uint32_t *amin_final, uint32_t *npairs);

}  // namespace ecmb200
