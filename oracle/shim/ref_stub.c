/* TEST INFRASTRUCTURE ONLY.  The reference declares but never defines
 * mpz_extrastrongbpsw_prp (eratosthenes/worker.c:361, a PRP path that the ECM
 * driver never reaches).  This stub only satisfies the linker. */
#include "gmp.h"
int mpz_extrastrongbpsw_prp(mpz_t n) { return mpz_probab_prime_p(n, 1); }
