// engine.hpp -- host-side interface to the kernels of one compiled limb count.  Each limb count
// is its own translation unit (engine_inst.cu, -DECM_NL=n) so that they compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>
#include "geom.hpp"

namespace ecmb200 {

typedef std::vector<uint32_t> Big;

struct Engine {
    int nl = 0, stride_s1 = 0, smem_s1 = 0;      // stride_s1 = max threads per stage-1 block
    // Register-resident macro-op machine (rv.cuh).  rv_lanes = lanes per curve of that stage-1 kernel (> 1: the
    // warp-cooperative layout); rv_max_threads = largest block the rv kernel serves (0: not compiled for this limb count).
    int rv_lanes = 1, rv_max_threads = 0;
    bool use_rv = false;                          // set by the context before prepare()
    // stage 2 on the cooperative layout (coop_s2.cuh): compiled for this limb count / selected (before prepare())
    bool has_coop_s2 = false, use_coop_s2 = false;
    // lane stride of the stage-1 state for blocks of T threads (the stage-1 kernel instance serving T fixes it)
    virtual int stride_for_threads(uint32_t T) const = 0;
    size_t params_bytes = 0;
    virtual ~Engine() {}
    virtual void set_params(const Big &n, const Big &one, const Big &r2, const Big &r3, const Big &rri, const Big &rref, uint32_t m0inv) = 0;
    // special-form engines (make_engine_sp_*): n = 2^kbits - cval (kind > 0) or 2^kbits + 1 (kind < 0); residues are plain
    virtual void set_special(int kind, uint32_t kbits, uint32_t cval) = 0;
    virtual bool serves_special(uint32_t kbits) const = 0;   // is bit kbits inside the word range this kernel set folds at
    virtual const void *params_host() const = 0;        // ModParams<NL> image to copy to the device
    virtual void set_params_device(const void *d) = 0;  // device copy for the out-of-line kernels
    virtual cudaError_t prepare() = 0;
    virtual void stage1(cudaStream_t st, uint32_t blocks, uint32_t threads, uint32_t *state, const uint8_t *ops, uint64_t nops,
                        uint32_t chunk_len, uint32_t groups, uint64_t item0) = 0;
    virtual void load_curves(cudaStream_t st, uint32_t *state, Geom G, uint32_t lanes, uint32_t count, const uint32_t *x, const uint32_t *s) = 0;
    virtual void build_curves(cudaStream_t st, uint32_t *state, Geom G, uint32_t lanes, uint32_t count, const uint32_t *uv, uint8_t *ok) = 0;
    virtual void read_point(cudaStream_t st, const uint32_t *state, Geom G, uint32_t count, uint32_t xs, uint32_t zs,
                            uint32_t *x, uint32_t *z, uint8_t *flag, uint32_t *g, const uint32_t *chk) = 0;
    // stage 2
    int threads_s2 = 0, smem_s2 = 0, nslot_s2 = 0;       // threads_s2 = CURVES per group of the state2 layout
    virtual void vm2(cudaStream_t st, uint32_t blocks, uint32_t *state2, uint32_t cap, uint32_t *tab, const uint64_t *code,
                     uint64_t ncode, uint32_t chunk_len, uint32_t groups, uint64_t item0, uint8_t *inv_fail) = 0;
    int threads_pair = 0, pair_blocks_per_sm = 1;        // threads_pair = CURVES per block of the pair kernel
    bool use_pair_kernel = true;                  // false for wide moduli: the pair steps run inside k_vm2
    virtual void pair_run(cudaStream_t st, uint32_t blocks, uint32_t *state2, uint32_t cap, const uint32_t *tab, const uint64_t *code,
                          uint32_t npairs, uint32_t ncurves, uint32_t chunk_len, uint32_t groups, uint64_t item0) = 0;
    virtual void s2_setup(cudaStream_t st, const uint32_t *state1, Geom G1, uint32_t xslot, uint32_t zslot, uint32_t spslot,
                          uint32_t first, uint32_t count, uint32_t *state2, uint32_t cap2, uint32_t *tab, uint32_t e_qx,
                          uint32_t e_qz, uint8_t *inv_fail) = 0;
    virtual void s2_collect(cudaStream_t st, const uint32_t *state2, uint32_t cap2, const uint8_t *inv_fail, uint32_t first,
                            uint32_t n, uint32_t count, uint32_t *acc_out, uint8_t *fail_out) = 0;
    virtual void fieldop(cudaStream_t st, int op, uint32_t count, const uint32_t *a, const uint32_t *b, uint32_t *r, int repeat) = 0;
};

extern void count_launch();

}  // namespace ecmb200
