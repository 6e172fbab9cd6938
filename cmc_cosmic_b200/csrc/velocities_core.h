// The energy scheme's velocity correction after the potential rebuild, shared by the CUDA kernels (velocities.cu) and the host-side
// emulation used by the CPU tests.  Follows the reference's
//   energy_conservation1                  src/cmc/cmc_utils.c:2153-2174
//   set_velocities3 (energy_conservation3) src/cmc/cmc_utils.c:254-352, 360-363
// Every FP64 operation goes through fp64_exact.h in the reference's evaluation order, so the results are bit-identical.
//
// set_velocities3's first loop (:268-311) carries ONE scalar, Eexcess, serially from star to star: a star whose predicted squared speed
// vnew2 is <= 0 ("source") adds its kinetic-energy deficit; every other corrected star ("sink") gives up MIN(0.8 of its kinetic energy,
// Eexcess) while Eexcess > 0.  Eexcess is exactly 0 almost everywhere (a sink that can absorb all of it subtracts it from itself), and a
// star met with Eexcess == 0 that is not a source is left alone.  So the loop splits into
//   prepare  per star, independent: vnew2 (one potential(rOld) lookup), the rescaling to vnew2, the star's kind and its chain operand
//            (source: the deficit; sink: its v^2 after the rescaling);
//   blocks   per block of SV3_BLOCK consecutive stars, independent: the carry that leaves the block if 0 entered it (a block without a
//            source: 0, nothing to do);
//   carry    one pass over the per-block results in star order: the carry entering each block.  Where 0 enters, the block's result stands;
//            where a positive carry enters (a chain crossing a block boundary), that block is replayed from it;
//   apply    per block with a source or a positive carry entering, independent: the real walk, scaling (vr, vt) of the sinks it meets.
// Each star sees exactly the additions and subtractions of the serial loop, in its order.
#pragma once
#include "orbit_core.h"

namespace cmcb {

enum { SV3_ACTIVE = 1,   // interacted == 0: the star takes part (:273)
       SV3_SOURCE = 2,   // vnew2 <= 0 (:287)
       SV3_POSITIVE = 4  // vnew2_arr[i] = (vnew2 > 0) (:283), what the second loop tests (:329)
};
constexpr int SV3_BLOCK = 1024;

struct Sv3In { // the star_t fields the first loop reads
	long interacted;
	double rOld, vrold, vtold, Uoldrold, Uoldrnew, vr, vt;
};

// energy_conservation1's Uoldrold (:2168) / set_velocities3's Unewrnew (:276): star_phi[g] + MPI_PHI_S(star_r[g], g)
CMCB_HD double sv3_U_own(const OrbitTables &t, kint g)
{
	const RPhi own = tbl_rp(t, g);
	return dadd(own.phi, phi_self(own.r, own.r, tbl_ld(t.m, g), t.invN));
}

// Per-star part of the first loop (:270-299).  Returns the kind; *val = the chain operand.  For a sink (vr, vt) come back rescaled:
// the caller stores them to .vr/.vt and .vrnew/.vtnew.  *fail: potential()'s bracket check failed (fatal in the reference).
CMCB_HD unsigned sv3_prepare_star(const OrbitTables &t, const LadderTops &top, kint g, const Sv3In &s, double *vr, double *vt, double *val, int *fail)
{
	*val = 0.0;
	if (s.interacted != 0) return 0;
	const double m = tbl_ld(t.m, g);
	const RPhi own = tbl_rp(t, g);
	const double Unewrold = dadd(potential_at(t, top, s.rOld, fail), phi_self(s.rOld, own.r, m, t.invN));
	const double Unewrnew = dadd(own.phi, phi_self(own.r, own.r, m, t.invN));
	const double vold2 = dadd(dmul(s.vtold, s.vtold), dmul(s.vrold, s.vrold));
	// vold2 + 2.0*(1.0-q)*(Uoldrold - Uoldrnew) + 2.0*q*(Unewrold - Unewrnew) with q = 0.5: both factors are exactly 1.0 (:281-282)
	const double vnew2 = dadd(dadd(vold2, dmul(dmul(2.0, dsub(1.0, 0.5)), dsub(s.Uoldrold, s.Uoldrnew))), dmul(dmul(2.0, 0.5), dsub(Unewrold, Unewrnew)));
	const double v2cur = dadd(dmul(s.vr, s.vr), dmul(s.vt, s.vt));
	unsigned kind = SV3_ACTIVE | (vnew2 > 0.0 ? SV3_POSITIVE : 0);
	if (vnew2 <= 0.0) {
		*val = dmul(dmul(0.5, dsub(v2cur, vnew2)), m); // :288
		return kind | SV3_SOURCE;
	}
	const double alpha = dsqrt(ddiv(vnew2, v2cur)); // :291
	*vr = dmul(s.vr, alpha);
	*vt = dmul(s.vt, alpha);
	*val = dadd(dmul(*vr, *vr), dmul(*vt, *vt)); // v2, :294
	return kind;
}

// One star of the carry chain (:287-288, 300-308).  Returns true when the star's (vr, vt) are to be multiplied by *ratio.
CMCB_HD bool sv3_chain_step(unsigned kind, double val, double m, double &E, double *ratio)
{
	if (!(kind & SV3_ACTIVE)) return false;
	if (kind & SV3_SOURCE) { E = dadd(E, val); return false; }
	const double cap = dmul(dmul(dmul(0.8, 0.5), m), val); // E_dump_factor * 0.5 * m * v2
	const double dump = cap <= E ? cap : E;               // MIN, include/cmc/cmc.h:2242
	if (E > 0) {
		*ratio = dsqrt(ddiv(dsub(val, ddiv(dmul(2.0, dump), m)), val));
		E = dsub(E, dump);
		return true;
	}
	return false;
}

// get_global_idx (cmc_utils.c:2384-2407) of local star j0 (0-based)
struct Sv3Index {
	long first_global, n_old_local, N_MAX;
	CMCB_HD kint operator()(long j0) const { return (kint) (j0 + 1 > n_old_local ? 1 + N_MAX + (j0 + 1 - n_old_local) : first_global + j0); }
};

// The serial loop over stars [lo, hi) from the carry E; apply(i, ratio) is called for every star whose velocity the chain scales.
template <class Apply> CMCB_HD double sv3_walk(const unsigned char *kind, const double *val, const double *star_m, const Sv3Index &gi, long lo, long hi, double E, Apply &&apply)
{
	for (long i = lo; i < hi; i++) {
		const unsigned k = kind[i];
		if (E == 0.0 && !(k & SV3_SOURCE)) continue; // Eexcess > 0 is false and nothing is added: the reference changes nothing here
		double ratio;
		if (sv3_chain_step(k, val[i], tbl_ld(star_m, gi(i)), E, &ratio)) apply(i, ratio);
	}
	return E;
}

// One star of the second loop (:325-341): the sinks give up MIN(v2*m*0.5 - 0.8*0.5*m*v2_new, Eexcess) while Eexcess > 0.
CMCB_HD bool sv3_second_step(unsigned kind, double vr, double vt, double vrnew, double vtnew, double m, double &E, double *ratio)
{
	if ((kind & (SV3_ACTIVE | SV3_POSITIVE)) != (SV3_ACTIVE | SV3_POSITIVE)) return false;
	const double v2_new = dadd(dmul(vrnew, vrnew), dmul(vtnew, vtnew));
	const double v2 = dadd(dmul(vr, vr), dmul(vt, vt));
	const double cap = dsub(dmul(dmul(v2, m), 0.5), dmul(dmul(dmul(0.8, 0.5), m), v2_new));
	const double dump = cap <= E ? cap : E;
	if (E > 0) {
		*ratio = dsqrt(ddiv(dsub(v2, ddiv(dmul(2.0, dump), m)), v2));
		E = dsub(E, dump);
		return true;
	}
	return false;
}

} // namespace cmcb
