// One row of the reference's HDF5 snapshot table, as write_snapshot() fills it (reference src/cmc/cmc_io.c:3176-3327; the row is the
// `Snapshot` struct, include/cmc/cmc.h:136-200: 62 fields, 488 bytes).  Shared by the device kernel (snapshot.cu) and the host build the
// CPU tests compare with the compiled reference (tests/host_emul/snapshot_host.cpp): every arithmetic step is a single IEEE operation in
// the reference's order, so the rows are the reference's bit for bit.  The HDF5 calls that store the rows (H5TBmake_table /
// H5TBappend_records) stay in the reference: the library is not part of this path.
#pragma once
#include "fp64_exact.h"
#include "star_layout.h"

namespace cmcb {

constexpr int SNAP_ROW_BYTES = 488;
constexpr int SNAP_ROW_WORDS = SNAP_ROW_BYTES / 8; // 61: every field is 8 bytes except the three ints (startype + padding, bin_startype0 | bin_startype1)

// byte offsets inside the Snapshot struct (checked against the compiled reference's field_offset array in tests/test_oracle_vs_ref.py)
enum SnapOff {
	SN_ID = 0, SN_M = 8, SN_R = 16, SN_VR = 24, SN_VT = 32, SN_E = 40, SN_J = 48, SN_BINFLAG = 56, SN_M0 = 64, SN_M1 = 72, SN_ID0 = 80, SN_ID1 = 88,
	SN_A = 96, SN_ECC = 104, SN_STARTYPE = 112, SN_LUM = 120, SN_RADIUS = 128, SN_BIN_STARTYPE0 = 136, SN_BIN_STARTYPE1 = 140, SN_BIN_LUM0 = 144,
	SN_BIN_LUM1 = 152, SN_BIN_RADIUS0 = 160, SN_BIN_RADIUS1 = 168, SN_BIN_EB = 176, SN_ETA = 184, SN_PHI = 192, SN_RAD0 = 200,
	SN_OSPIN = 464, SN_B = 472, SN_FORMATION = 480
};

// star_t / binary_t fields only this file reads (include/cmc/cmc.h:409-601, :231-390; offsets checked like those of star_layout.h)
constexpr int ST_OFF_SE_OSPIN = 248, ST_OFF_SE_LUM = 304, ST_OFF_SE_SCM_B = 352, ST_OFF_SE_SCM_FORMATION = 360;
constexpr int BI_OFF_ID1 = 0, BI_OFF_ID2 = 8, BI_OFF_RAD1 = 16, BI_OFF_RAD2 = 24, BI_OFF_M1 = 32, BI_OFF_M2 = 40, BI_OFF_A = 64, BI_OFF_E = 72, BI_OFF_BSE_KW = 84,
	      BI_OFF_BSE_MASS0 = 112, BI_OFF_BSE_RADIUS = 144, BI_OFF_BSE_LUM = 160, BI_OFF_BSE_MASSC = 176, BI_OFF_BSE_RADC = 192, BI_OFF_BSE_MENV = 208,
	      BI_OFF_BSE_RENV = 224, BI_OFF_BSE_OSPIN = 240, BI_OFF_BSE_BACC = 272, BI_OFF_BSE_TACC = 288, BI_OFF_BSE_EPOCH = 304, BI_OFF_BSE_TMS = 320,
	      BI_OFF_BSE_TB = 360, BI_OFF_BSE_BCM_DMDT = 368, BI_OFF_BSE_BCM_RADROL = 384, BI_OFF_BSE_BCM_B = 400, BI_OFF_BSE_BCM_FORMATION = 416;

struct SnapArgs {
	const char *star1;   // &star[1]
	const char *bin0;    // &binary[0] (a star's binind indexes it)
	const double *star_r, *star_m, *star_phi; // the replicated tables
	const double *sigma; // sigma_array.sigma, [n_local + 1], entry 0 unused (read for binaries only; may be null without binaries)
	long n_local;        // clus.N_MAX_NEW
	long n_old_local, first_global, N_MAX; // get_global_idx (src/cmc/cmc_utils.c:2384-2407)
	long N_STAR;         // clus.N_STAR
	double units_m, units_l, madhoc; // units.m, units.l, madhoc
	long avekernel;      // AVEKERNEL (calc_average_mass_sqr)
	uint64_t nan_bits;   // bit pattern of the reference's `0.0 / 0.0` (the compiler's constant, taken from the compiled reference)
};

CMCB_HD double snap_ld(const char *rec, int off) { return *(const double *) (rec + off); }
CMCB_HD long snap_ldl(const char *rec, int off) { return *(const long *) (rec + off); }
CMCB_HD int snap_ldi(const char *rec, int off) { return *(const int *) (rec + off); }
CMCB_HD long snap_global_idx(const SnapArgs &a, long i) { return i > a.n_old_local ? 1 + a.N_MAX + (i - a.n_old_local) : a.first_global + i - 1; }

// the reference's unit constants (include/cmc/cmc.h:66-72)
constexpr double SNAP_MSUN = 1.989e+33, SNAP_RSUN = 6.9599e10, SNAP_AU = 1.496e+13;

// calc_average_mass_sqr(index, N_LIMIT) (src/cmc/cmc_dynamics_helper.c:2997-3022).  write_snapshot passes the GLOBAL clus.N_MAX as N_LIMIT for
// LOCAL indices, so on P > 1 ranks the window of a binary among the last p stars of a shard is not clamped: it runs past the shard's last
// star and get_global_idx maps the overshoot to the table entries behind N_MAX + 1 -- kept as it is (tests/ref_multirank_check.py compares
// exactly these stars on 2-4 ranks of the compiled reference)
CMCB_HD double snap_average_mass_sqr(const SnapArgs &a, long index, long N_LIMIT)
{
	const long p = a.avekernel;
	long simin = index - p, simax = simin + (2 * p - 1);
	if (simin < 1) { simin = 1; simax = simin + (2 * p - 1); }
	else if (simax > N_LIMIT) { simax = N_LIMIT; simin = simax - (2 * p - 1); }
	double M2ave = 0.0;
	for (long k = simin; k <= simax; k++) {
		const double x = dmul(a.star_m[snap_global_idx(a, k)], a.madhoc);
		M2ave = dadd(M2ave, dmul(x, x));
	}
	return ddiv(M2ave, (double) (2 * p));
}

CMCB_HD bool snap_is_bh(const SnapArgs &a, long i) // the bh_only filter (:3170-3171, :3188)
{
	const char *rec = a.star1 + (i - 1) * CMCB_SIZEOF_STAR;
	const char *bin = a.bin0 + snap_ldl(rec, CMCB_OFF_BININD) * CMCB_SIZEOF_BINARY;
	return snap_ldi(rec, CMCB_OFF_SE_K) == 14 || snap_ldi(bin, BI_OFF_BSE_KW) == 14 || snap_ldi(bin, BI_OFF_BSE_KW + 4) == 14;
}

// row of local star i (1-based) into the 61 words at `row`
CMCB_HD void snap_row(const SnapArgs &a, long i, uint64_t *row)
{
	const char *rec = a.star1 + (i - 1) * CMCB_SIZEOF_STAR;
	const long g = snap_global_idx(a, i);
	const long j = snap_ldl(rec, CMCB_OFF_BININD);
	const char *bin = a.bin0 + j * CMCB_SIZEOF_BINARY;
	const double m100 = -100.0;
	const uint64_t l100 = (uint64_t) (long) -100;
	const double mfac = ddiv(a.units_m, (double) a.N_STAR); // units.m / clus.N_STAR
	auto D = [&](int off, double v) { row[off / 8] = dbits(v); };
	auto L = [&](int off, long v) { row[off / 8] = (uint64_t) v; };
	L(SN_ID, snap_ldl(rec, CMCB_OFF_ID));
	D(SN_M, ddiv(dmul(a.star_m[g], mfac), SNAP_MSUN));             // m * (units.m / clus.N_STAR) / MSUN (:3190)
	D(SN_R, a.star_r[g]);
	D(SN_VR, snap_ld(rec, CMCB_OFF_VR)); D(SN_VT, snap_ld(rec, CMCB_OFF_VT)); D(SN_E, snap_ld(rec, CMCB_OFF_E)); D(SN_J, snap_ld(rec, CMCB_OFF_J));
	int startype, bst0, bst1;
	if (j) { // :3196-3203
		L(SN_BINFLAG, 1);
		D(SN_M0, ddiv(dmul(snap_ld(bin, BI_OFF_M1), mfac), SNAP_MSUN));
		D(SN_M1, ddiv(dmul(snap_ld(bin, BI_OFF_M2), mfac), SNAP_MSUN));
		L(SN_ID0, snap_ldl(bin, BI_OFF_ID1)); L(SN_ID1, snap_ldl(bin, BI_OFF_ID2));
		D(SN_A, ddiv(dmul(snap_ld(bin, BI_OFF_A), a.units_l), SNAP_AU));
		D(SN_ECC, snap_ld(bin, BI_OFF_E));
	} else {
		row[SN_BINFLAG / 8] = l100; D(SN_M0, m100); D(SN_M1, m100); row[SN_ID0 / 8] = l100; row[SN_ID1 / 8] = l100; D(SN_A, m100); D(SN_ECC, m100);
	}
	if (j == 0) { // :3214-3225
		startype = snap_ldi(rec, CMCB_OFF_SE_K);
		D(SN_LUM, snap_ld(rec, ST_OFF_SE_LUM));
		D(SN_RADIUS, ddiv(dmul(snap_ld(rec, CMCB_OFF_RAD), a.units_l), SNAP_RSUN));
		bst0 = bst1 = -100;
		D(SN_BIN_LUM0, m100); D(SN_BIN_LUM1, m100); D(SN_BIN_RADIUS0, m100); D(SN_BIN_RADIUS1, m100); D(SN_BIN_EB, m100); D(SN_ETA, m100);
	} else { // :3226-3240
		startype = -100;
		D(SN_LUM, m100); D(SN_RADIUS, m100);
		bst0 = snap_ldi(bin, BI_OFF_BSE_KW); bst1 = snap_ldi(bin, BI_OFF_BSE_KW + 4);
		D(SN_BIN_LUM0, snap_ld(bin, BI_OFF_BSE_LUM)); D(SN_BIN_LUM1, snap_ld(bin, BI_OFF_BSE_LUM + 8));
		D(SN_BIN_RADIUS0, ddiv(dmul(snap_ld(bin, BI_OFF_RAD1), a.units_l), SNAP_RSUN));
		D(SN_BIN_RADIUS1, ddiv(dmul(snap_ld(bin, BI_OFF_RAD2), a.units_l), SNAP_RSUN));
		const double m1 = snap_ld(bin, BI_OFF_M1), m2 = snap_ld(bin, BI_OFF_M2), sa = snap_ld(bin, BI_OFF_A);
		// -(m1 / N_STAR) * (m2 / N_STAR) / (2 * a)
		D(SN_BIN_EB, ddiv(dmul(-ddiv(m1, (double) a.N_STAR), ddiv(m2, (double) a.N_STAR)), dmul(2.0, sa)));
		// (m1 * m2 * sqr(madhoc)) / (a * sqrt(calc_average_mass_sqr(i, clus.N_MAX)) * sqr(sigma_array.sigma[i]))
		const double sg = a.sigma ? a.sigma[i] : dfrombits(a.nan_bits);
		const double num = dmul(dmul(m1, m2), dmul(a.madhoc, a.madhoc));
		const double den = dmul(dmul(sa, dsqrt(snap_average_mass_sqr(a, i, a.N_MAX))), dmul(sg, sg));
		D(SN_ETA, ddiv(num, den));
	}
	row[SN_STARTYPE / 8] = (uint64_t) (unsigned) startype;                              // int + 4 bytes of padding (zero here)
	row[SN_BIN_STARTYPE0 / 8] = (uint64_t) (unsigned) bst0 | ((uint64_t) (unsigned) bst1 << 32); // two ints
	D(SN_PHI, a.star_phi[g]);
	if (j == 0) { // :3242-3278: 33 fields of `0.0 / 0.0`, then the star's spin / field / formation
		for (int w = SN_RAD0 / 8; w < SN_OSPIN / 8; w++) row[w] = a.nan_bits;
		D(SN_OSPIN, snap_ld(rec, ST_OFF_SE_OSPIN)); D(SN_B, snap_ld(rec, ST_OFF_SE_SCM_B)); D(SN_FORMATION, snap_ld(rec, ST_OFF_SE_SCM_FORMATION));
	} else { // :3279-3315: rad, tb, lum, massc, radc, menv, renv, tms, dmdt, radrol, ospin, B, formation, bacc, tacc, mass0, epoch -- pairs [0], [1]
		int w = SN_RAD0 / 8;
		auto pair = [&](int off) { row[w++] = dbits(snap_ld(bin, off)); row[w++] = dbits(snap_ld(bin, off + 8)); };
		pair(BI_OFF_BSE_RADIUS);
		row[w++] = dbits(snap_ld(bin, BI_OFF_BSE_TB));
		pair(BI_OFF_BSE_LUM); pair(BI_OFF_BSE_MASSC); pair(BI_OFF_BSE_RADC); pair(BI_OFF_BSE_MENV); pair(BI_OFF_BSE_RENV); pair(BI_OFF_BSE_TMS);
		pair(BI_OFF_BSE_BCM_DMDT); pair(BI_OFF_BSE_BCM_RADROL); pair(BI_OFF_BSE_OSPIN); pair(BI_OFF_BSE_BCM_B); pair(BI_OFF_BSE_BCM_FORMATION);
		pair(BI_OFF_BSE_BACC); pair(BI_OFF_BSE_TACC); pair(BI_OFF_BSE_MASS0); pair(BI_OFF_BSE_EPOCH);
		D(SN_OSPIN, m100); D(SN_B, m100); D(SN_FORMATION, m100);
	}
}

} // namespace cmcb
