// emu_main.cpp — TEST INFRASTRUCTURE: builds the product's kernel sources against the SIMT emulator
// (tests/emu/simt_emu.h) into tests/emu/libconsent_emu.so.  Loaded only by `-m "not gpu"` tests.
// (built with -DCG_EMU -DCG_EMU_IMPL -include simt_emu.h)
#include "simt_emu.h"
#include "../../consent_b200/csrc/consent_b200.cu"

// Test hook: the integer forms of comparable(x, mean) / comparable(x, deciles) (cg_common.cuh) against the fp64 forms they replace, over a
// grid of the integer arguments the path can produce; returns the number of disagreements.
extern "C" unsigned long long emu_comparable_selfcheck(unsigned xmax, unsigned mmax) {
    unsigned long long bad = 0;
    for (unsigned x = 0; x <= xmax; ++x)
        for (unsigned m = 0; m <= mmax; ++m)
            bad += cg_comparable_mean_u(x, m) != cg_comparable_mean((double)x, (double)m);
    const unsigned big[] = {0u, 1u, 4u, 5u, 6u, 4999u, 5000u, 65535u, 65536u, 1u << 20, (1u << 30) - 1u, 1u << 30, (1u << 31) - 6u, (1u << 31) - 1u};
    for (unsigned x : big)
        for (unsigned m : big)
            for (int dx = -6; dx <= 6; ++dx)
                for (int dm = -6; dm <= 6; ++dm) {
                    const long long xx = (long long)x + dx, mm = (long long)m + dm;
                    if (xx < 0 || mm < 0 || xx >= (1ll << 31) || mm >= (1ll << 31)) continue;
                    bad += cg_comparable_mean_u((unsigned)xx, (unsigned)mm) != cg_comparable_mean((double)xx, (double)mm);
                    bad += cg_comparable_mean_u((unsigned)xx, (unsigned)(2 * mm < (1ll << 31) ? 2 * mm : mm)) !=
                           cg_comparable_mean((double)xx, (double)(2 * mm < (1ll << 31) ? 2 * mm : mm));
                }
    const unsigned dmax = xmax < 600u ? xmax : 600u;
    for (unsigned lo = 0; lo <= dmax; lo += (lo < 40 ? 1 : 7))
        for (unsigned hi = lo; hi <= dmax; hi += (hi < 40 ? 1 : 11))
            for (unsigned x = 0; x <= 3 * dmax; ++x)
                bad += cg_comparable_dec_u(x, lo, hi) != cg_comparable_dec((double)x, (double)lo, (double)hi);
    return bad;
}
