// TEST-ONLY host build of cmc_cosmic_b200/csrc/snapshot_core.h: the rows of write_snapshot() as plain loops, compared with the compiled
// reference's rows in tests/test_oracle_vs_ref.py without a GPU.
#include <stdint.h>
#include "../../cmc_cosmic_b200/csrc/snapshot_core.h"

using namespace cmcb;

extern "C" long emul_snapshot_rows(const char *star1, const char *bin0, const double *star_r, const double *star_m, const double *star_phi, const double *sigma,
				    long n_local, long n_old_local, long first_global, long N_MAX, long N_STAR, double units_m, double units_l, double madhoc,
				    long avekernel, uint64_t nan_bits, int bh_only, uint64_t *rows)
{
	SnapArgs a{star1, bin0, star_r, star_m, star_phi, sigma, n_local, n_old_local, first_global, N_MAX, N_STAR, units_m, units_l, madhoc, avekernel, nan_bits};
	long k = 0;
	for (long i = 1; i <= n_local; i++) {
		if (bh_only && !snap_is_bh(a, i)) continue;
		snap_row(a, i, rows + k * SNAP_ROW_WORDS);
		k++;
	}
	return k;
}
