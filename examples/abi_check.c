/* abi_check.c -- plain C (not C++) client of include/ecm_b200.h.
 * Shows that the boundary is a C ABI a maintainer of the reference can link against:
 *   gcc -std=c99 -Iinclude examples/abi_check.c -Lavx-ecm_b200 -lecm_b200 -Wl,-rpath,$PWD/avx-ecm_b200
 * Without a GPU it exercises the host-side planners and checks that the engine refuses to run
 * (there is no CPU fallback); with a GPU it runs 8 curves through stage 1 and stage 2. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ecm_b200.h"

int main(void)
{
    uint64_t counts[2];
    uint64_t nops = ecm_b200_plan_stage1(1000000, NULL, 0, counts);
    printf("plan_stage1(1e6): %llu ops, %llu point-adds, %llu point-doubles\n", (unsigned long long)nops,
           (unsigned long long)counts[0], (unsigned long long)counts[1]);
    if (counts[0] != 1980817 || counts[1] != 217929) return 2;      /* ecm.c:1849 printout of the reference */

    uint32_t D, U, L, R;
    ecm_b200_stage2_params(1000000, &D, &U, &L, &R);
    printf("stage-2 geometry: D=%u U=%u L=%u R=%u\n", D, U, L, R);
    if (D != 2310 || U != 16 || L != 32 || R != 963) return 3;      /* main.c:840-882 */

    /* N = (2^89-1)*(2^107-1), 196 bits */
    uint32_t n[7] = {0};
    { /* little-endian limbs of the product, computed with 64-bit schoolbook arithmetic */
        uint32_t a[3] = {0xffffffffu, 0xffffffffu, 0x01ffffffu}, b[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0x7ffu};
        int i, j;
        for (i = 0; i < 3; i++) {
            uint64_t c = 0;
            for (j = 0; j < 4; j++) { c += (uint64_t)a[i] * b[j] + n[i + j]; n[i + j] = (uint32_t)c; c >>= 32; }
            n[i + 4] += (uint32_t)c;
        }
    }
    ecm_b200_ctx *ctx = NULL;
    int rc = ecm_b200_create(&ctx, 0, n, 7, 8);
    if (rc == ECM_B200_ENODEV) {
        printf("no GPU: %s\n", ecm_b200_last_error());
        return 0;
    }
    if (rc) { printf("create failed: %s\n", ecm_b200_last_error()); return 4; }
    {
        uint64_t sigma[8]; int i, k, Lm = ecm_b200_limbs(ctx);
        uint32_t *g = (uint32_t *)calloc((size_t)Lm * 8, 4);
        uint8_t flag[8];
        for (i = 0; i < 8; i++) sigma[i] = 1000 + i;
        if (ecm_b200_build_curves(ctx, 8, sigma) || ecm_b200_stage1(ctx, 20000) ||
            ecm_b200_stage2(ctx, 20000, 2000000) || ecm_b200_read_stage2(ctx, NULL, flag, g, NULL)) {
            printf("run failed: %s\n", ecm_b200_last_error());
            return 5;
        }
        for (i = 0; i < 8; i++) {
            printf("sigma %llu: %s", (unsigned long long)sigma[i], flag[i] ? "factor 0x" : "no factor");
            if (flag[i]) for (k = Lm - 1; k >= 0; k--) printf("%08x", g[(size_t)k * 8 + i]);
            printf("\n");
        }
        /* the same stage 2 at the reference's own granularity (ecm.c:67-72, driver loop ecm.c:1400-1476): init once,
         * then one range call per pairmap -- here the pairmap of ecm_b200_pair, where a maintainer would pass pair()'s */
        {
            uint32_t amin_final, npairs, steps, *pm_v, *pm_u, *g2 = (uint32_t *)calloc((size_t)Lm * 8, 4);
            uint8_t flag2[8];
            int found_inv = -1;
            ecm_b200_stage2_params(20000, &D, &U, &L, &R);
            steps = ecm_b200_pair(20000, 2000000, D, U, NULL, NULL, 0, &amin_final, &npairs);
            pm_v = (uint32_t *)malloc((size_t)steps * 4 + 4); pm_u = (uint32_t *)malloc((size_t)steps * 4 + 4);
            ecm_b200_pair(20000, 2000000, D, U, pm_v, pm_u, steps, &amin_final, &npairs);
            if (ecm_b200_build_curves(ctx, 8, sigma) || ecm_b200_stage1(ctx, 20000) || ecm_b200_stage2_init(ctx, 20000, &found_inv) ||
                ecm_b200_stage2_range(ctx, (20000 + D) / (2 * D), pm_v, pm_u, steps) || ecm_b200_read_stage2(ctx, NULL, flag2, g2, NULL)) {
                printf("stage2_init/range failed: %s\n", ecm_b200_last_error());
                return 6;
            }
            if (memcmp(flag, flag2, 8) || memcmp(g, g2, (size_t)Lm * 8 * 4)) { printf("stage2_init/range disagrees with stage2\n"); return 7; }
            pm_v[steps / 2] += 1000;                        /* a pairmap that leaves the window must be refused, not executed */
            if (ecm_b200_stage2_range(ctx, (20000 + D) / (2 * D), pm_v, pm_u, steps) != ECM_B200_EINVAL) { printf("bad pairmap accepted\n"); return 8; }
            printf("stage2_init + stage2_range(%u steps, %u pairs): identical factors, foundDuringInv = %d\n", steps, npairs, found_inv);
            free(pm_v); free(pm_u); free(g2);
        }
        free(g);
    }
    ecm_b200_destroy(ctx);
    return 0;
}
