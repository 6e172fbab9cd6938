// calc_sigma_r (reference src/cmc/cmc_dynamics_helper.c:2821-2975) as ONE chain of signed operations -- the form the exact scan
// (seqscan.cuh + exactsum.h: seq_classify_signed) needs.  Not wired into a kernel yet: this header fixes the mapping and is
// checked on the host against the oracle (tests/host_emul, tests/test_host.py).
//
// The reference keeps two running sums over a window of 2p stars.  For local star si = 1..n the window is [simin(si), simin(si) + 2p - 1]
// (local indices; it stops sliding at both ends of the cluster, :2923-2931), and star si first subtracts the stars that left the window,
// then adds those that entered.  The window moves by at most one star per step, so the whole loop is the operation list
//     op 0 .. 2p-1                 + term(k),  k = simax0 + 1 + op          the first window (simax0 = 0 on the first rank, -p elsewhere, :2906-2910)
//     op 2p + 2(si-2)              - term(simin(si-1))  if the window moved, else +0
//     op 2p + 2(si-2) + 1          + term(simax(si))    if the window moved, else +0          for si = 2..n
// and the value after op 2p-1, resp. after the odd op of si, is the reference's running sum at star si.  Local indices k < 1 and k > n are the
// ghost stars of the neighbouring ranks (:2862-2900).
#pragma once
#include "fp64_exact.h"

namespace cmcb {

struct SigmaWindow {
	long p, n, first_global, N_MAX;
	int first_rank; // myid == 0

	CMCB_HD long simin(long si) const
	{
		const long g_simin = first_global + si - 1 - p;
		if (g_simin < 1) return 1;                            // :2923-2926
		if (g_simin + (2 * p - 1) > N_MAX) return n - (2 * p - 1); // :2927-2931
		return si - p;
	}
	CMCB_HD long n_ops() const { return 2 * p + 2 * (n - 1); }
	// operation j: local index *k of the star whose term is applied, *sign = +1 / -1, or 0 for "the window did not move"
	CMCB_HD void op(long j, long *k, int *sign) const
	{
		if (j < 2 * p) { *k = (first_rank ? 0 : -p) + 1 + j; *sign = +1; return; }
		const long si = 2 + (j - 2 * p) / 2;
		const long a = simin(si - 1), b = simin(si);
		if (b == a) { *k = 1; *sign = 0; return; }
		if (((j - 2 * p) & 1) == 0) { *k = a; *sign = -1; }
		else { *k = b + (2 * p - 1); *sign = +1; }
	}
	// the star whose running sums are complete after operation j (0: none)
	CMCB_HD long star_of(long j) const
	{
		if (j < 2 * p) return j == 2 * p - 1 ? 1 : 0;
		return ((j - 2 * p) & 1) ? 2 + (j - 2 * p) / 2 : 0;
	}
	// a valid configuration moves the window by at most one star per step and never reads beyond the ghost stars
	CMCB_HD bool valid() const { return p >= 1 && n >= 2 * p && (first_rank ? first_global == 1 : first_global > p); }
};

// the two terms of local star k: m_k madhoc (v_r^2 + v_t^2) and m_k madhoc, as the reference evaluates them (:2946-2947)
CMCB_HD void sigma_terms(double m, double madhoc, double vr, double vt, double *t_mv2, double *t_m)
{
	const double mm = dmul(m, madhoc);
	*t_mv2 = dmul(mm, dadd(dmul(vr, vr), dmul(vt, vt)));
	*t_m = mm;
}

} // namespace cmcb
