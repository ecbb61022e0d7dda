// rv_prog.hpp -- phase programs of the register-resident stage-1 machine (rv.cuh): plain enums and initialisers, shared
// by the device kernel (constant memory) and the host (ecm_b200_rv_program exports them so that
// tests/test_programs_cpu.py can interpret the very table the GPU runs).
#pragma once
#include <stdint.h>

namespace ecmb200 {

// ---- phases ---------------------------------------------------------------------------------------------------------
// phase word: kind | x<<4 | y<<8 | z<<12 | flag<<16 ; x, y, z = logical points 0..3 (A, B, C, T of prac), resolved
// through the macro-op's permutation to physical point slots
enum : uint32_t { PH_A1 = 0, PH_A2, PH_A3, PH_D1L, PH_D1P, PH_D2, PH_D3, PH_COPY, PH_END = 15 };
enum : uint32_t { PA = 0, PB = 1, PC = 2, PT = 3 };
#define PHW(kind, x, y, z, flag) ((uint32_t)(kind) | ((uint32_t)(x) << 4) | ((uint32_t)(y) << 8) | ((uint32_t)(z) << 12) | ((uint32_t)(flag) << 16))
// P1 + P2 with difference Pin -> Pout (park: keep the sums of P2 for a doubling that follows)
#define RV_ADD(P1, P2, Pin, Pout, park) PHW(PH_A1, P1, P2, 0, park), PHW(PH_A2, 0, 0, 0, 0), PHW(PH_A3, Pin, Pout, 0, 0)
#define RV_DUP_PARKED(Pout) PHW(PH_D1P, 0, 0, 0, 0), PHW(PH_D2, 0, 0, 0, 0), PHW(PH_D3, Pout, 0, 0, 0)
#define RV_DUP_POINT(Psrc, Pout) PHW(PH_D1L, Psrc, 0, 0, 0), PHW(PH_D2, 0, 0, 0, 0), PHW(PH_D3, Pout, 0, 0, 0)
#define RV_MAXPROG 8
#define RV_PROGRAMS                                                                                                     \
    /* M_DBL   P (logical T) doubled in place           (ecm.c:1816-1822) */ { RV_DUP_POINT(PT, PT), PHW(PH_END, 0, 0, 0, 0) }, \
    /* M_INIT  C = B (= P) ; A = 2B                      (ecm.c:603-613)   */ { PHW(PH_COPY, PB, PC, 0, 0), RV_DUP_POINT(PB, PA), PHW(PH_END, 0, 0, 0, 0) }, \
    /* M_C3    T = B + A (C)                             (ecm.c:683-713)   */ { RV_ADD(PB, PA, PC, PT, 0), PHW(PH_END, 0, 0, 0, 0) }, \
    /* M_C4    B = B + A (C) ; A = 2A                    (ecm.c:714-726)   */ { RV_ADD(PB, PA, PC, PB, 1), RV_DUP_PARKED(PA), PHW(PH_END, 0, 0, 0, 0) }, \
    /* M_C5    C = C + A (B) ; A = 2A                    (ecm.c:728-740)   */ { RV_ADD(PC, PA, PB, PC, 1), RV_DUP_PARKED(PA), PHW(PH_END, 0, 0, 0, 0) }, \
    /* M_C9    C = C + B (A) ; B = 2B                    (ecm.c:853-865)   */ { RV_ADD(PC, PB, PA, PC, 1), RV_DUP_PARKED(PB), PHW(PH_END, 0, 0, 0, 0) }, \
    /* M_FINAL P = A + B (C), written to logical T       (ecm.c:868-873)   */ { RV_ADD(PA, PB, PC, PT, 0), PHW(PH_END, 0, 0, 0, 0) }, \
    { PHW(PH_END, 0, 0, 0, 0) },

}  // namespace ecmb200
