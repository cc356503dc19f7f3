/* Stand-in for Random123's philox.h, which the reference takes from the uncomplicate/neanderthal jar
 * (0.25.7-SNAPSHOT; not under /root/reference).  TEST INFRASTRUCTURE ONLY (oracle/): the published Philox4x32-10
 * algorithm (Salmon et al., SC'11) restated with the type and function names the reference's kernel text uses.
 * Pinned by the Random123 known-answer vectors in tests/test_oracle_golden.py through the same code path
 * (orc_philox4x32_10 in bayadera_oracle.c has the identical round function). */
#ifndef BAY_REF_SHIM_PHILOX_H
#define BAY_REF_SHIM_PHILOX_H
#include <stdint.h>

typedef struct { uint32_t v[4]; } philox4x32_ctr_t;
typedef struct { uint32_t v[2]; } philox4x32_key_t;

static __inline__ philox4x32_ctr_t philox4x32(philox4x32_ctr_t c, philox4x32_key_t k) {
    for (int round = 0; round < 10; round++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.v[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.v[2];
        const philox4x32_ctr_t n = {{(uint32_t)(p1 >> 32) ^ c.v[1] ^ k.v[0], (uint32_t)p1,
                                     (uint32_t)(p0 >> 32) ^ c.v[3] ^ k.v[1], (uint32_t)p0}};
        c = n;
        k.v[0] += 0x9E3779B9u;
        k.v[1] += 0xBB67AE85u;
    }
    return c;
}
#endif
