// engine_inst.cu -- instantiates every kernel for ONE limb count (compile with -DECM_NL=<n>).
#include "engine.hpp"
#include "kernels.cuh"
#include "rv.cuh"
#include "coop_s2.cuh"

#ifndef ECM_NL
#error "compile with -DECM_NL=<limbs>"
#endif

#include <type_traits>
#include <cstdlib>
namespace ecmb200 {
inline namespace ECM_VNS {

// which field policy the register-resident stage-1 machine (rv.cuh) uses for a limb count, if any: one thread per curve
// up to 16 limbs (four operand values + two accumulator pairs fit the registers of a 384-thread block), four lanes per
// curve at 48 and 64 limbs (coop.cuh).  20-32 limbs keep the slot-file machine (dedicated squaring, 0.92 of the roof).
// MAXT = block size (= lane stride) of the standard instance; BIG = a second instance for batches that fill larger blocks,
// where the registers allow it (96 / 119 registers at 10 / 13 limbs), 0 = none.
// one thread per curve: up to 16 limbs with the dual product; 20 and 24 limbs with two single products per phase (144 / 168
// registers), measured 7.16 -> 7.78 and 6.92 -> 7.43 Tprod/s against the slot-file machine with its dedicated squaring
#ifndef RV_SOLO_MAX
#define RV_SOLO_MAX (ECM_SPECIAL ? 16 : 24)
#endif
template <int NL, class Enable = void> struct RvCfg { static constexpr int MAXT = 0, BIG = 0; typedef void Field; };
template <int NL> struct RvCfg<NL, typename std::enable_if<(NL <= RV_SOLO_MAX)>::type> {
    static constexpr int MAXT = 384, BIG = (NL <= 10) ? 640 : (NL <= 13) ? 512 : 0;
    typedef SoloField<NL> Field;
};
#if !ECM_SPECIAL
template <> struct RvCfg<40> { static constexpr int MAXT = 384, BIG = 0; typedef CoopField<10, 4> Field; };
template <> struct RvCfg<48> { static constexpr int MAXT = 384, BIG = 0; typedef CoopField<12, 4> Field; };
template <> struct RvCfg<56> { static constexpr int MAXT = 384, BIG = 0; typedef CoopField<14, 4> Field; };
template <> struct RvCfg<64> { static constexpr int MAXT = 384, BIG = 0; typedef CoopField<16, 4> Field; };
#endif
// stage 2 on the cooperative layout (coop_s2.cuh) where stage 1 has it
template <int NL> struct S2CoopCfg { static constexpr int M = 0, L = 1; };
#if !ECM_SPECIAL
// (28 / 32 limbs with TWO lanes per curve in stage 2 -- a wave is limited by the tables, 32 768 curves at 1024 bits, which leaves
// one thread per curve with 7 warps per SM -- was measured and lost: pair runs 5.27 s either way, slot machine 3.29 -> 4.56 s,
// profiles/r2o_s2_1024_*.log)
template <> struct S2CoopCfg<40> { static constexpr int M = 10, L = 4; };
template <> struct S2CoopCfg<48> { static constexpr int M = 12, L = 4; };
template <> struct S2CoopCfg<56> { static constexpr int M = 14, L = 4; };
template <> struct S2CoopCfg<64> { static constexpr int M = 16, L = 4; };
#endif
template <class F> struct RvLanes { static constexpr int L = F::L, M = F::M; };
template <> struct RvLanes<void> { static constexpr int L = 1, M = 1; };

template <int NL>
struct EngineT : Engine {
    ModParams<NL> P;
    const ModParams<NL> *Pg = nullptr;
    EngineT()
    {
        nl = NL; stride_s1 = S1Cfg<NL>::STRIDE; smem_s1 = S1Cfg<NL>::smem;
        params_bytes = sizeof(ModParams<NL>);
        threads_s2 = S2Cfg<NL>::THREADS; smem_s2 = S2Cfg<NL>::smem; nslot_s2 = NSLOT_S2;
        rv_max_threads = RvCfg<NL>::BIG ? RvCfg<NL>::BIG : RvCfg<NL>::MAXT; rv_lanes = RvLanes<typename RvCfg<NL>::Field>::L;
        has_coop_s2 = S2CoopCfg<NL>::M != 0;
    }
    void set_params(const Big &n, const Big &one, const Big &r2, const Big &r3, const Big &rri, const Big &rref, uint32_t m0inv) override
    {
        for (int k = 0; k < NL; k++) { P.n[k] = n[k]; P.one[k] = one[k]; P.r2[k] = r2[k]; P.r3[k] = r3[k]; P.rrefinv[k] = rri[k]; P.rref[k] = rref[k]; }
        P.m0inv = m0inv; P.kind = 0; P.kbits = 0; P.cval = 0;
    }
    void set_special(int kind, uint32_t kbits, uint32_t cval) override { P.kind = kind; P.kbits = kbits; P.cval = cval; }
    bool serves_special(uint32_t kbits) const override
    {
        return ECM_SPECIAL && NL <= 32 && kbits >= 64 && (int)(kbits >> 5) >= SpecialRange<NL>::LOW && (int)(kbits >> 5) < NL;
    }
    int stride_for_threads(uint32_t T) const override
    {
        if (use_rv) return (RvCfg<NL>::BIG && T > (uint32_t)RvCfg<NL>::MAXT) ? RvCfg<NL>::BIG : RvCfg<NL>::MAXT;
        return T <= (uint32_t)S1Small<NL>::MAXT ? S1Small<NL>::MAXT : S1Cfg<NL>::STRIDE;
    }
    const void *params_host() const override { return &P; }
    void set_params_device(const void *d) override { Pg = static_cast<const ModParams<NL> *>(d); }
    cudaError_t prepare() override
    {
        cudaError_t e = cudaFuncSetAttribute(k_stage1<NL, S1Cfg<NL>::STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_s1);
        if (e != cudaSuccess) return e;
        if constexpr (S1Small<NL>::MAXT != S1Cfg<NL>::STRIDE) {
            e = cudaFuncSetAttribute(k_stage1<NL, S1Small<NL>::MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     S1Cfg<NL>::per_thread * S1Small<NL>::MAXT);
            if (e != cudaSuccess) return e;
        }
        if constexpr (RvCfg<NL>::MAXT != 0) {
            typedef typename RvCfg<NL>::Field F;
            e = cudaFuncSetAttribute(k_stage1_rv<F, RvCfg<NL>::MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * F::M * 4 * RvCfg<NL>::MAXT);
            if (e != cudaSuccess) return e;
            if constexpr (RvCfg<NL>::BIG != 0) {
                e = cudaFuncSetAttribute(k_stage1_rv<F, RvCfg<NL>::BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * F::M * 4 * RvCfg<NL>::BIG);
                if (e != cudaSuccess) return e;
            }
        }
        if constexpr (S2CoopCfg<NL>::M != 0) {
            if (use_coop_s2) {
                typedef CoopS2Cfg<S2CoopCfg<NL>::M, S2CoopCfg<NL>::L> C;
                threads_s2 = C::CURVES; smem_s2 = C::smem;
                threads_pair = C::PAIR_THREADS / S2CoopCfg<NL>::L;
                use_pair_kernel = true;
                e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pair_blocks_per_sm, k_pair_coop<S2CoopCfg<NL>::M, S2CoopCfg<NL>::L>, C::PAIR_THREADS, 0);
                if (e != cudaSuccess) return e;
                if (pair_blocks_per_sm < 1) pair_blocks_per_sm = 1;
                return cudaFuncSetAttribute(k_vm2_coop<S2CoopCfg<NL>::M, S2CoopCfg<NL>::L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_s2);
            }
        }
        // pair kernel: one 384-thread block per SM up to 16 limbs (measured at 415 bits, B1=1e6/B2=1e8, 65 536 curves:
        // stage 2 in 12.44 s against 13.76 s with three 128-thread blocks per SM), 128-thread blocks above
        threads_pair = (NL <= 16) ? 384 : PairCfg<NL>::THREADS;
        if (NL <= 16) if (const char *ev = getenv("ECM_B200_PAIR_THREADS")) if (atoi(ev) == 128) threads_pair = 128;
        use_pair_kernel = (NL <= 32);
        if (!use_pair_kernel) return cudaFuncSetAttribute(k_vm2<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_s2);
        if constexpr (NL <= 16) {
            if (threads_pair == 384) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pair_blocks_per_sm, k_pair<NL, 384>, 384, 0);
            else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pair_blocks_per_sm, k_pair<NL>, threads_pair, 0);
        } else if constexpr (NL <= 32) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pair_blocks_per_sm, k_pair<NL>, threads_pair, 0);
        if (e != cudaSuccess) return e;
        if (pair_blocks_per_sm < 1) pair_blocks_per_sm = 1;
        return cudaFuncSetAttribute(k_vm2<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_s2);
    }
    void vm2(cudaStream_t st, uint32_t blocks, uint32_t *state2, uint32_t cap, uint32_t *tab, const uint64_t *code,
             uint64_t ncode, uint32_t chunk_len, uint32_t groups, uint64_t item0, uint8_t *inv_fail) override
    {
        if constexpr (S2CoopCfg<NL>::M != 0) {
            if (use_coop_s2) {
                constexpr int M = S2CoopCfg<NL>::M, L = S2CoopCfg<NL>::L;
                k_vm2_coop<M, L><<<blocks, CoopS2Cfg<M, L>::THREADS, smem_s2, st>>>(P, Pg, state2, cap, tab, code, ncode, chunk_len, groups, item0, inv_fail);
                count_launch();
                return;
            }
        }
        k_vm2<NL><<<blocks, threads_s2, smem_s2, st>>>(P, Pg, state2, cap, tab, code, ncode, chunk_len, groups, item0, inv_fail);
        count_launch();
    }
    void pair_run(cudaStream_t st, uint32_t blocks, uint32_t *state2, uint32_t cap, const uint32_t *tab, const uint64_t *code,
                  uint32_t npairs, uint32_t ncurves, uint32_t chunk_len, uint32_t groups, uint64_t item0) override
    {
        if constexpr (S2CoopCfg<NL>::M != 0) {
            if (use_coop_s2) {
                constexpr int M = S2CoopCfg<NL>::M, L = S2CoopCfg<NL>::L;
                uint32_t nlanes = (ncurves * L + 31) / 32 * 32;
                if (nlanes > cap * L) nlanes = cap * L;
                k_pair_coop<M, L><<<blocks, CoopS2Cfg<M, L>::PAIR_THREADS, 0, st>>>(P, state2, cap, tab, code, npairs, nlanes, chunk_len, groups, item0);
                count_launch();
                return;
            }
        }
        if constexpr (NL <= 16) {
            if (threads_pair == 384) {
                k_pair<NL, 384><<<blocks, 384, 0, st>>>(P, state2, cap, tab, code, npairs, ncurves, chunk_len, groups, item0);
                count_launch();
                return;
            }
        }
        if constexpr (NL <= 32) {
            k_pair<NL><<<blocks, PairCfg<NL>::THREADS, 0, st>>>(P, state2, cap, tab, code, npairs, ncurves, chunk_len, groups, item0);
            count_launch();
        }
    }
    void s2_setup(cudaStream_t st, const uint32_t *state1, Geom G1, uint32_t xslot, uint32_t zslot, uint32_t spslot,
                  uint32_t first, uint32_t count, uint32_t *state2, uint32_t cap2, uint32_t *tab, uint32_t e_qx, uint32_t e_qz,
                  uint8_t *inv_fail) override
    {
        if constexpr (S2CoopCfg<NL>::M != 0) {
            if (use_coop_s2) {
                k_s2_setup_coop<S2CoopCfg<NL>::M, S2CoopCfg<NL>::L><<<(cap2 + 127) / 128, 128, 0, st>>>(state1, dg(G1), xslot, zslot, spslot, first, count, state2, cap2, tab, e_qx, e_qz, inv_fail);
                count_launch();
                return;
            }
        }
        k_s2_setup<NL><<<(cap2 + 127) / 128, 128, 0, st>>>(state1, dg(G1), xslot, zslot, spslot, first, count, state2, cap2, tab, e_qx, e_qz, inv_fail);
        count_launch();
    }
    void s2_collect(cudaStream_t st, const uint32_t *state2, uint32_t cap2, const uint8_t *inv_fail, uint32_t first, uint32_t n,
                    uint32_t count, uint32_t *acc_out, uint8_t *fail_out) override
    {
        if constexpr (S2CoopCfg<NL>::M != 0) {
            if (use_coop_s2) {
                k_s2_collect_coop<S2CoopCfg<NL>::M, S2CoopCfg<NL>::L><<<(n + 127) / 128, 128, 0, st>>>(state2, cap2, inv_fail, first, n, count, acc_out, fail_out);
                count_launch();
                return;
            }
        }
        k_s2_collect<NL><<<(n + 127) / 128, 128, 0, st>>>(state2, cap2, inv_fail, first, n, count, acc_out, fail_out);
        count_launch();
    }
    static Geom dg(const Geom &G) { return G; }
    void stage1(cudaStream_t st, uint32_t blocks, uint32_t threads, uint32_t *state, const uint8_t *ops, uint64_t nops,
                uint32_t chunk_len, uint32_t groups, uint64_t item0) override
    {
        if constexpr (RvCfg<NL>::MAXT != 0) {
            if (use_rv) {
                typedef typename RvCfg<NL>::Field F;
                if constexpr (RvCfg<NL>::BIG != 0) {
                    if (threads > (uint32_t)RvCfg<NL>::MAXT) {
                        k_stage1_rv<F, RvCfg<NL>::BIG><<<blocks, threads, 2 * F::M * 4 * RvCfg<NL>::BIG, st>>>(P, state, ops, nops, chunk_len, groups, item0);
                        count_launch();
                        return;
                    }
                }
                k_stage1_rv<F, RvCfg<NL>::MAXT><<<blocks, threads, 2 * F::M * 4 * RvCfg<NL>::MAXT, st>>>(P, state, ops, nops, chunk_len, groups, item0);
                count_launch();
                return;
            }
        }
        if (S1Small<NL>::MAXT != S1Cfg<NL>::STRIDE && threads <= (uint32_t)S1Small<NL>::MAXT)
            k_stage1<NL, S1Small<NL>::MAXT><<<blocks, threads, S1Cfg<NL>::per_thread * S1Small<NL>::MAXT, st>>>(P, state, ops, nops, chunk_len, groups, item0);
        else
            k_stage1<NL, S1Cfg<NL>::STRIDE><<<blocks, threads, smem_s1, st>>>(P, state, ops, nops, chunk_len, groups, item0);
        count_launch();
    }
    void load_curves(cudaStream_t st, uint32_t *state, Geom G, uint32_t lanes, uint32_t count, const uint32_t *x, const uint32_t *s) override
    {
        k_load_curves<NL><<<(lanes + 127) / 128, 128, 0, st>>>(Pg, state, dg(G), lanes, count, x, s);
        count_launch();
    }
    void build_curves(cudaStream_t st, uint32_t *state, Geom G, uint32_t lanes, uint32_t count, const uint32_t *uv, uint8_t *ok) override
    {
        k_build_curves<NL><<<(lanes + 63) / 64, 64, 0, st>>>(Pg, state, dg(G), lanes, count, uv, ok);
        count_launch();
    }
    void read_point(cudaStream_t st, const uint32_t *state, Geom G, uint32_t count, uint32_t xs, uint32_t zs,
                    uint32_t *x, uint32_t *z, uint8_t *flag, uint32_t *g, const uint32_t *chk) override
    {
        k_read_point<NL><<<(count + 63) / 64, 64, 0, st>>>(Pg, state, dg(G), count, xs, zs, x, z, flag, g, chk);
        count_launch();
    }
    void fieldop(cudaStream_t st, int op, uint32_t count, const uint32_t *a, const uint32_t *b, uint32_t *r, int repeat) override
    {
        k_fieldop<NL><<<(count + 127) / 128, 128, 0, st>>>(P, Pg, op, count, a, b, r, repeat);
        count_launch();
    }
};

}  // inline namespace ECM_VNS

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#if ECM_SPECIAL
Engine *CAT(make_engine_sp_, ECM_NL)() { return new EngineT<ECM_NL>(); }
#else
Engine *CAT(make_engine_, ECM_NL)() { return new EngineT<ECM_NL>(); }
#endif

}  // namespace ecmb200
