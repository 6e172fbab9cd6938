// Per-star new-orbit step, shared by the CUDA kernel (orbit.cu) and the host-side emulation used by the CPU tests.
//
// Follows the reference's get_positions_loop (src/cmc/cmc_evolution_thr.c:802-933) and what it calls:
//   calc_orbit_new / get_positive_Q_index / set_a_b   src/cmc/cmc_orbit.c:413-540, 30-59, 722-740
//   find_zero_Q, FindZero_r, find_root_vr             src/cmc/cmc_nr.c:316-383, 62-82, 500-603
//   function_q, potential                             src/cmc/cmc_utils.c:1654-1662, 112-165
//   MPI_PHI_S                                         include/cmc/cmc.h:2175
// Every FP64 operation is issued through fp64_exact.h in the reference's evaluation order (single rounding, no
// FMA contraction), so given the same r/m/Phi tables and the same uniform draws the outputs are bit-identical to
// the reference's.  The reference's `long double` evaluation of Q is handled by sign_Q(): FP64 plus a rigorous
// error bound on the fast path, exact x87 replay (x87emu.h) when the sign is within the bound.
#pragma once
#include "fp64_exact.h"
#include "x87emu.h"

namespace cmcb {

#define CMCB_SF_INFINITY 1.0e10 /* reference include/cmc/cmc.h:58 */
#define CMCB_ZERO 1.0e-20       /* :60 */
#define CMCB_N_TRY 50000        /* :64 */

enum OrbitStatus {
	ORBIT_OK = 0,       // new radius sampled                      cmc_evolution_thr.c:857-933
	ORBIT_MASSLESS = 1, // star_m < ZERO -> destroy_obj            :816-820
	ORBIT_UNBOUND = 2,  // E >= 0 -> remove_star                   :823-829
	ORBIT_CIRCULAR = 3, // circular_flag: keeps r, vr, vt          :835-839
	ORBIT_TIDAL = 4,    // rmax >= Rtidal -> remove_star           :848-855
	ORBIT_NTRY = 5      // N_TRY exceeded (fatal in the reference) :890-893
};

// Table indices: the library caps n_star_dim below 2^30 (cmcb200_create), so 32 bits hold every index and every
// kmin + kmax + 1 of the bisections -- half the registers of `long` in the hot loops.
typedef int kint;

struct RPhi { // one entry of the packed lookup table: every probe of the bisections needs r_k AND Phi_k -> one 16-byte load
	double r, phi;
};

// ---- ladders: coarse summaries of the sorted tables -------------------------------------------------------------
//
// Both searches of the orbit step walk the 16 N-byte (r, Phi) table with ceil(log2 N) DEPENDENT probes; beyond N ~ 4e6 the
// table does not fit in L2 and every probe of the lower levels is a DRAM sector.  The ladders answer most probes from small
// summaries instead (the coarsest level is staged in shared memory by the kernels, the next ones stay L2-resident):
//
//  QLadder, for find_zero_Q (cmc_nr.c:316-383).  Block b of a level with shift s summarises table entries
//      [b << s, min((b + 1) << s, M)] (both ends inclusive, M = N_MAX + 1) as (r_lo, r_hi, min Phi, max Phi) -- 32 bytes, one sector.
//      r is sorted, the self term phi_s is non-increasing in r, so for EVERY k of the block
//          2 (E - Phi_max - phi_s(r_lo)) - J^2 / r_lo^2  <=  Q(k)  <=  2 (E - Phi_min - phi_s(r_hi)) - J^2 / r_hi^2 .
//      When the upper bound is below (the lower bound above) zero by more than a rigorous margin, the sign the reference
//      computes at the probe index is PROVEN without touching the fine table (q_sign_on_block).  The bisection replays the
//      reference's exact probe sequence ktry = (kmin + kmax + 1) / 2, so kmin / kmax are the reference's even where Q is
//      not monotone; only the probes next to the sign change fall through to the fine table.
//      (Phi is NOT assumed monotone: rounding makes the innermost Phi_k wiggle, and Phi_0 > Phi_1 with a central mass --
//      the builder stores the true min / max.  Sortedness of r is verified by the builder; an unsorted table disables the ladders.)
//
//  RLadder, for potential()'s FindZero_r (cmc_nr.c:62-82): "largest k with star_r[k] < r" is a pure function of the sorted
//      array, so any search returns the reference's index.  Level with shift s holds star_r[b << s]; the coarsest level is
//      binary-searched (shared memory), every further level is ONE 128-byte line (16 samples), and the last step is one line of
//      the fine table (8 entries), which also holds the two entries the interpolation needs.
struct LBlock {
	double r_lo, r_hi, phi_min, phi_max;
};

constexpr int LADDER_MAX_LEVELS = 8;
#ifndef CMCB_QLAD_TOP_BLOCKS /* the CPU tests build a second copy with tiny top levels so that small tables get deep ladders */
#define CMCB_QLAD_TOP_BLOCKS 1280
#define CMCB_RLAD_TOP_ENTRIES 5120
#endif
constexpr int QLAD_MIN_SHIFT = 4, QLAD_STEP = 3, QLAD_TOP_BLOCKS = CMCB_QLAD_TOP_BLOCKS;   // coarsest level <= 40 KB
constexpr int RLAD_MIN_SHIFT = 3, RLAD_STEP = 4, RLAD_TOP_ENTRIES = CMCB_RLAD_TOP_ENTRIES; // coarsest level <= 40 KB
#ifndef CMCB_QLAD_FACTOR_LOG2
#define CMCB_QLAD_FACTOR_LOG2 2
#endif
constexpr int QLAD_FACTOR_LOG2 = CMCB_QLAD_FACTOR_LOG2; // default for QLadder::factor_log2
constexpr int LADDER_MIN_M = 64;                                          // smaller tables: plain bisection

struct QLadder {
	const LBlock *blk[LADDER_MAX_LEVELS]; // coarsest first
	int shift[LADDER_MAX_LEVELS];
	int n_levels;     // 0: no ladder
	int factor_log2;  // a level is consulted while its blocks are at most 2^-factor_log2 of the bracket (default 2)
	int max_width;    // brackets at least this wide probe the fine table directly (default: never) -- tuning knob, see orbit.cu
};

struct RLadder {
	const double *lvl[LADDER_MAX_LEVELS]; // coarsest first; every level padded with +inf to a multiple of 16 plus one spare line
	int shift[LADDER_MAX_LEVELS];
	int n_top;    // entries of the coarsest level
	int n_levels; // 0: no ladder
	int whole_lines; // != 0: read every ladder line whole (four independent sector loads) instead of sector by sector (count16_below)
};

// The SAMPLE ladder of the index bisections: level l holds the table entries (r_k, Phi_k) themselves at k = b << shift[l] -- not summaries.
// See find_zero_Q_sampled below for what it proves and why that is enough.
#ifndef CMCB_SLAD_TOP_ENTRIES
#define CMCB_SLAD_TOP_ENTRIES 2560
#endif
#ifndef CMCB_SLAD_F32_TOP
#define CMCB_SLAD_F32_TOP 0 /* measured on B200: screening the coarsest level's probes in FP32 arithmetic changes nothing (3.10 vs 3.07 ms;
                              profiles/ab_r02b_f32top.log) and its mere presence costs stage A 0.1 ms: compiled out */
#endif
#ifndef CMCB_SLAD_PAIR
#define CMCB_SLAD_PAIR 0 /* measured on B200: no faster than the two descents one after the other (2.96 vs 2.98 ms at N = 1e7; profiles/ab_r02b_pair.log) */
#endif
#ifndef CMCB_SLAD_SECANT_STEPS
#define CMCB_SLAD_SECANT_STEPS 0 /* measured on B200: 8 guided probes make stage A SLOWER (2.97 -> 3.66 ms at N = 1e7; profiles/ab_r02b_secant.log) */
#endif
constexpr int SLAD_SECANT_STEPS = CMCB_SLAD_SECANT_STEPS; // 0: plain halving on the coarsest level
constexpr int SLAD_MIN_SHIFT = 4, SLAD_STEP = 4, SLAD_TOP_ENTRIES = CMCB_SLAD_TOP_ENTRIES; // coarsest level <= 40 KB of (r, Phi) pairs
struct RPhiF;
struct SLadder {
	const RPhi *lvl[LADDER_MAX_LEVELS]; // coarsest first
	const RPhiF *topf;                  // lvl[0] rounded to FP32 (nullptr: none; set wherever n_levels is): the coarsest level's probes are screened in FP32 arithmetic
	int shift[LADDER_MAX_LEVELS];
	int n_levels; // 0: no ladder
	double margin; // relative margin of a proven sign (see side_Q_margin)
	double phi_abs_max; // max |Phi_k| over the table (= |Phi| at the innermost entries)
};
CMCB_HD int slad_plan(long M, int shift[LADDER_MAX_LEVELS], long count[LADDER_MAX_LEVELS])
{
	if (M < LADDER_MIN_M) return 0;
	int top = SLAD_MIN_SHIFT;
	while ((M >> top) + 1 > SLAD_TOP_ENTRIES) top += SLAD_STEP;
	int n = 0;
	for (int s = top; s >= SLAD_MIN_SHIFT && n < LADDER_MAX_LEVELS; s -= SLAD_STEP) { shift[n] = s; count[n] = (M >> s) + 1; n++; }
	return n;
}

// where the coarsest level of each ladder is to be read from: shared memory in the kernels (staged per CTA), the level itself on the host;
// nullptr = that ladder is not used (table too small, or the builder found it unsorted)
struct LadderTops {
	const LBlock *q;
	const double *r;
	const QLadder *qdesc; // the Q ladder's level table: a shared-memory copy in the kernels (dynamic indexing), &t.qlad on the host
	const RPhi *s;        // coarsest level of the sample ladder (nullptr: plain bisection)
	const RPhiF *sf = nullptr; // its FP32 copy (shared memory in the kernels; then s points at the global FP64 level, read only by undecided probes)
	const SLadder *sdesc;
};

// shifts (coarsest first) and block counts of the Q ladder over a table with last index M; returns the number of levels
CMCB_HD int qlad_plan(long M, int shift[LADDER_MAX_LEVELS], long count[LADDER_MAX_LEVELS])
{
	if (M < LADDER_MIN_M) return 0;
	int top = QLAD_MIN_SHIFT;
	while ((M >> top) + 1 > QLAD_TOP_BLOCKS) top += QLAD_STEP;
	int n = 0;
	for (int s = top; s >= QLAD_MIN_SHIFT && n < LADDER_MAX_LEVELS; s -= QLAD_STEP) { shift[n] = s; count[n] = (M >> s) + 1; n++; }
	return n;
}

// the same for the r ladder; count = real samples, alloc = padded length in doubles
CMCB_HD int rlad_plan(long M, int shift[LADDER_MAX_LEVELS], long count[LADDER_MAX_LEVELS], long alloc[LADDER_MAX_LEVELS])
{
	if (M < LADDER_MIN_M) return 0;
	int top = RLAD_MIN_SHIFT;
	while ((M >> top) + 1 > RLAD_TOP_ENTRIES) top += RLAD_STEP;
	int n = 0;
	for (int s = top; s >= RLAD_MIN_SHIFT && n < LADDER_MAX_LEVELS; s -= RLAD_STEP) {
		shift[n] = s; count[n] = (M >> s) + 1; alloc[n] = ((count[n] + 15) / 16) * 16 + 16; n++;
	}
	return n;
}

// what the builders store: the summary of table entries [k0, k1] (inclusive); *unsorted is set when r decreases inside the range
CMCB_HD LBlock lblock_of(const RPhi *rp, long k0, long k1, bool *unsorted)
{
	LBlock B;
	B.r_lo = rp[k0].r; B.r_hi = rp[k1].r; B.phi_min = B.phi_max = rp[k0].phi;
	for (long k = k0; k < k1; k++) {
		if (!(rp[k].r <= rp[k + 1].r)) *unsorted = true;
		const double p = rp[k + 1].phi;
		if (p < B.phi_min) B.phi_min = p;
		if (p > B.phi_max) B.phi_max = p;
		if (!(p == p)) *unsorted = true; // a NaN would escape the min / max: no ladder then
	}
	return B;
}
// the summary of a run of consecutive child blocks (a coarser level from the next finer one)
CMCB_HD LBlock lblock_merge(const LBlock *child, long n)
{
	LBlock B = child[0];
	for (long i = 1; i < n; i++) {
		B.r_hi = child[i].r_hi;
		if (child[i].phi_min < B.phi_min) B.phi_min = child[i].phi_min;
		if (child[i].phi_max > B.phi_max) B.phi_max = child[i].phi_max;
	}
	return B;
}

struct RPhiF { float r, phi; }; // the same table rounded to FP32: 8 bytes per entry, half the footprint -- the screening copy of find_zero_Q

struct OrbitTables {
	const RPhi *rp;  // (star_r[k], star_phi[k]) interleaved: [0] sentinel, [1..N_MAX], [N_MAX+1] = (SF_INFINITY, 0)
	const RPhiF *rpf; // FP32 screening copy of rp (nullptr: every probe reads rp)
	const double *m; // star_m
	QLadder qlad;    // coarse (r, Phi) summaries for find_zero_Q; n_levels == 0: every probe reads rp
	SLadder slad;    // sampled (r, Phi) entries for find_zero_Q (the default); n_levels == 0: off
	RLadder rlad;    // coarse radii for potential(); n_levels == 0: plain bisection on rp
	kint N_MAX;
	double invN;    // 1.0 / clus.N_STAR evaluated in double (the factor in MPI_PHI_S)
	double madhoc;  // 1.0 / (double) clus.N_STAR (cmc_utils.c:1183) -- same value, kept separate like the reference
	double cenma_m; // cenma.m
};

struct OrbitOut {
	double rnew, vrnew, vtnew, X, r_peri, r_apo;
	kint kmin, kmax;
	int status;
	int n_draws; // uniforms consumed
	int n_exact; // Q evaluations that needed the exact x87 replay
	int n_brent; // find_root_vr fallbacks
};

// ---- table lookups -------------------------------------------------------------------------------------------

CMCB_HD double tbl_ld(const double *p, kint k)
{
#if defined(__CUDA_ARCH__)
	return __ldg(p + k);
#else
	return p[k];
#endif
}
CMCB_HD RPhi tbl_rp(const OrbitTables &t, kint k)
{
#if defined(__CUDA_ARCH__)
	const double2 v = __ldg(reinterpret_cast<const double2 *>(t.rp + k));
	RPhi o;
	o.r = v.x;
	o.phi = v.y;
	return o;
#else
	return t.rp[k];
#endif
}
CMCB_HD double tbl_r(const OrbitTables &t, kint k) { return tbl_ld(&t.rp[k].r, 0); }
CMCB_HD double tbl_phi(const OrbitTables &t, kint k) { return tbl_ld(&t.rp[k].phi, 0); }

#if defined(__CUDA_ARCH__)
#define CMCB_COLD __device__ __noinline__
#else
#define CMCB_COLD inline
#endif

// MPI_PHI_S(rad, j) with the star's own r_j, m_j already in registers
CMCB_HD double phi_self(double rad, double r_j, double m_j, double invN)
{
	return dmul(rad >= r_j ? ddiv(m_j, rad) : ddiv(m_j, r_j), invN);
}

// FindZero_r (cmc_nr.c:62-82)
CMCB_HD kint find_zero_r(const OrbitTables &t, kint kmin, kint kmax, double r)
{
	do {
		kint ktry = (kmin + kmax + 1) / 2;
		if (tbl_r(t, ktry) < r) kmin = ktry; else kmax = ktry - 1;
	} while (kmax != kmin);
	return kmin;
}

// four consecutive doubles (one 32-byte sector, 32-byte aligned) with ONE load: sm_100's 256-bit LDG.  The rejection rounds are bound by
// the L1's sector throughput (every lane reads its own lines), so a 128-byte line costs 4 sector accesses instead of 8 with 128-bit loads.
struct Quad {
	double a, b, c, d;
};
CMCB_HD Quad ld_quad(const double *p)
{
	Quad q;
#if defined(__CUDA_ARCH__)
	asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(q.a), "=d"(q.b), "=d"(q.c), "=d"(q.d) : "l"(p));
#else
	q.a = p[0]; q.b = p[1]; q.c = p[2]; q.d = p[3];
#endif
	return q;
}

// number of entries among 16 consecutive SORTED doubles (one 128-byte line, padded with +inf) that are < r.
// The kernels that call this are bound by the L1's sector throughput (every lane reads its own line), so the line is not read whole: the
// sector holding samples 4..7 first; it settles the count alone when s4 < r <= s7, otherwise ONE more sector below, or one or two above
// (1 + 1/4 ... 3 sector reads, 2 on average, instead of 4) -- dependent loads, which a throughput-bound kernel can afford.
CMCB_HD int count16_whole(const double *line, double r)
{
	int c = 0;
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const Quad v = ld_quad(line + 4 * i);
		c += (v.a < r) + (v.b < r) + (v.c < r) + (v.d < r);
	}
	return c;
}
CMCB_HD int count16_below(const double *line, double r)
{
	const Quad q1 = ld_quad(line + 4);
	if (!(q1.d < r)) { // the answer is <= 7
		if (q1.a < r) return 5 + (q1.b < r) + (q1.c < r);      // s4 < r <= s7
		const Quad q0 = ld_quad(line);                          // r <= s4
		return (q0.a < r) + (q0.b < r) + (q0.c < r) + (q0.d < r);
	}
	const Quad q2 = ld_quad(line + 8);                              // s7 < r
	if (!(q2.d < r)) return 8 + (q2.a < r) + (q2.b < r) + (q2.c < r);
	const Quad q3 = ld_quad(line + 12);                             // s11 < r
	return 12 + (q3.a < r) + (q3.b < r) + (q3.c < r) + (q3.d < r);
}

// the coarsest ladder levels live in shared memory on the device (the kernels substitute the pointer): generic loads, no __ldg
CMCB_HD double top_ld(const double *p) { return *p; }
CMCB_HD LBlock top_ld(const LBlock *p)
{
#if defined(__CUDA_ARCH__)
	const double2 a = reinterpret_cast<const double2 *>(p)[0], b = reinterpret_cast<const double2 *>(p)[1];
	LBlock o;
	o.r_lo = a.x; o.r_hi = a.y; o.phi_min = b.x; o.phi_max = b.y;
	return o;
#else
	return *p;
#endif
}
CMCB_HD LBlock blk_ld(const LBlock *p)
{
#if defined(__CUDA_ARCH__)
	LBlock o; // one 256-bit load (sm_100): the block is exactly one 32-byte sector
	asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o.r_lo), "=d"(o.r_hi), "=d"(o.phi_min), "=d"(o.phi_max) : "l"(p));
	return o;
#else
	return *p;
#endif
}

// FindZero_r(1, N_MAX + 1, r) through the r ladder: largest k in [1, N_MAX+1] with star_r[k] < r, 1 if there is none.
// The last step reads the fine-table line that holds the answer (entries base .. base + 7, four 256-bit loads) and only counts; the
// caller then re-reads the two entries it interpolates between -- L1 hits, and 30 registers fewer than carrying the line along.
CMCB_HD kint find_zero_r_ladder(const OrbitTables &t, const double *top, double r)
{
	const RLadder &L = t.rlad;
	int lo = 0, hi = L.n_top; // samples [0, lo) are < r, samples [hi, n_top) are >= r
	while (lo < hi) {
		const int mid = (lo + hi) >> 1;
		if (top_ld(top + mid) < r) lo = mid + 1; else hi = mid;
	}
	long b = lo > 0 ? lo - 1 : 0; // last sample below r
	for (int l = 1; l < L.n_levels; l++) {
		const int c = L.whole_lines ? count16_whole(L.lvl[l] + 16 * b, r) : count16_below(L.lvl[l] + 16 * b, r);
		b = 16 * b + (c > 0 ? c - 1 : 0);
	}
	const kint base = (kint) (b << RLAD_MIN_SHIFT), M = t.N_MAX + 1;
	// the fine-table line that holds the answer: entries base .. base + 7 (base is a multiple of 8: 128-byte aligned, two (r, Phi) entries per
	// 256-bit load), sorted; entries beyond M do not count.  Same sector-sparing order as count16_below: the sector with entries 2, 3 first.
	int c;
	if (L.whole_lines) {
		c = 0;
#pragma unroll
		for (int i = 0; i < 4; i++) {
			const Quad v = ld_quad(&t.rp[base + 2 * i].r);
			c += (base + 2 * i <= M && v.a < r) + (base + 2 * i + 1 <= M && v.c < r);
		}
	} else {
		auto below = [&](kint k, double rk) { return (int) (k <= M && rk < r); };
		const Quad v1 = ld_quad(&t.rp[base + 2].r);
		if (!below(base + 3, v1.c)) { // the answer is <= 3
			if (below(base + 2, v1.a)) c = 3;
			else { const Quad v0 = ld_quad(&t.rp[base].r); c = below(base, v0.a) + below(base + 1, v0.c); }
		} else {
			const Quad v2 = ld_quad(&t.rp[base + 4].r);
			if (!below(base + 5, v2.c)) c = 4 + below(base + 4, v2.a);
			else { const Quad v3 = ld_quad(&t.rp[base + 6].r); c = 6 + below(base + 6, v3.a) + below(base + 7, v3.c); }
		}
	}
	const kint k = base + c - 1; // largest k in [0, M] with star_r[k] < r (-1: none)
	return k >= 1 ? k : 1;
}

// potential(r) (cmc_utils.c:112-165, SEARCH_GRID = 0).  *fail is set when the bracket check (:144-152) would abort.
CMCB_HD double potential_at(const OrbitTables &t, const LadderTops &top, double r, int *fail)
{
	double phi0 = tbl_phi(t, 0);
	if (r == 0.0) return phi0;
	const RPhi p1 = tbl_rp(t, 1);
	if (r < p1.r) return dsub(phi0, ddiv(dmul(dsub(phi0, p1.phi), p1.r), r));
	kint kmax = t.N_MAX + 1, kmin = 1;
	const kint i = (kmin == kmax - 1) ? kmin : (top.r ? find_zero_r_ladder(t, top.r, r) : find_zero_r(t, kmin, kmax, r));
	const RPhi a = tbl_rp(t, i), b = tbl_rp(t, i + 1);
	if (a.r > r || b.r < r) { *fail = 1; return 0.0; }
	double inv_i = ddiv(1.0, a.r);
	return dadd(a.phi, ddiv(dmul(dsub(b.phi, a.phi), dsub(inv_i, ddiv(1.0, r))), dsub(inv_i, ddiv(1.0, b.r))));
}

// ---- Q(r): sign with the reference's long-double semantics ---------------------------------------------------

struct StarConst {
	double E, J;     // energy including the self term (cmc_evolution_thr.c:806), angular momentum
	double r_j, m_j; // star_r[g_j], star_m[g_j]
	kint g_j;
	// derived once per star for the division-free sign test
	double JJ;       // fl(J*J)
	double mi;       // fl(m_j * invN):            phis * r^2 = mi * r        for r >= r_j
	double c_in;     // fl(fl(m_j / r_j) * invN):  phis       = c_in          for r <  r_j
	double absE;
};

CMCB_HD void star_const_derive(const OrbitTables &t, StarConst &s)
{
	s.JJ = dmul(s.J, s.J);
	s.mi = dmul(s.m_j, t.invN);
	s.c_in = dmul(ddiv(s.m_j, s.r_j), t.invN);
	s.absE = dabs(s.E);
}

// FP64 evaluation of function_q and a bound B with |Qd - (x87 result before its final rounding)| <= B.
// Roundings: ratio, phis, pot+phis, E-(..), J/r (one ulp more for the x87 double rounding), its square, the final
// subtraction -- each <= 2^-52 relative to a term no larger than |E|+|pot|+|phis|+Jr; 2^-48 covers them with slack.
// Used where the VALUE matters (the qmin < qmax direction test of find_zero_Q); the bisection probes use sign_Q.
CMCB_HD double q_fp64(const OrbitTables &t, const StarConst &s, double r, double pot, double *B)
{
	double phis = phi_self(r, s.r_j, s.m_j, t.invN);
	double jr = ddiv(s.J, r);
	double Jr = dmul(jr, jr);
	double Q = dsub(dmul(2.0, dsub(s.E, dadd(pot, phis))), Jr);
	*B = dmul(3.552713678800501e-15 /* 2^-48 */, dadd(dadd(dabs(s.E), dabs(pot)), dadd(dabs(phis), Jr)));
	return Q;
}

// value the reference's function_q returns (a double), exact.  Cold: reached when a sign is within the error bound.
// (scalars by value: a by-reference struct argument of a non-inlined function would pin the caller's copy in local memory)
CMCB_COLD double q_exact_cold(double r, double pot, double E, double J, double m_j, double r_j, double invN)
{
	return x87_function_q(r, pot, E, J, m_j, r_j, invN);
}
CMCB_HD double q_exact(const OrbitTables &t, const StarConst &s, double r, double pot)
{
	return q_exact_cold(r, pot, s.E, s.J, s.m_j, s.r_j, t.invN);
}

// Sign of function_q(j, r, pot, E, J) -- what every bisection probe and the linear scan consume -- WITHOUT a division:
//      Q r^2 = 2 (E - pot) r^2 - 2 phis r^2 - J^2,     phis r^2 = (m_j invN) r  if r >= r_j,  (m_j / r_j invN) r^2  otherwise
// evaluated in FP64 as S with a rigorous bound.  Error budget in units of 2^-53 (first order):
//   a = E - pot: 1 (of |a|);  r2: 1;  t1 = a r2: 1  -> 3 |t1|;   t2: <= 4 |t2|;   JJ: 1;   d = t1 - t2: 1 (|t1| + |t2|);
//   S = 2 d - JJ: 1 (2|d| + JJ)   =>  |S - S*| <= (10 |t1| + 12 |t2| + 2 JJ) 2^-53
// and the reference's own deviation from the real-number Q, scaled by r^2: SQR(J/r) is a double square of a double-rounded
// quotient, <= 3.1 * 2^-53 JJ; its 80-bit operations <= 2^-62 (|E| + |pot| + |phis|) r^2.
// B = 2^-48 (|t1| + |t2| + JJ) + 2^-58 (|E| + |pot|) r^2 covers all of it with a factor >= 2 to spare (32 > 12, 10, 5.1).
// Inside the bound the x87 arithmetic is replayed exactly (q_exact), so the returned sign is ALWAYS the reference's.
CMCB_HD int sign_Q_at(const OrbitTables &t, const StarConst &s, const RPhi p, int *n_exact)
{
	const double a = dsub(s.E, p.phi);
	const double r2 = dmul(p.r, p.r);
	const double t1 = dmul(a, r2);
	const double t2 = (p.r >= s.r_j) ? dmul(s.mi, p.r) : dmul(s.c_in, r2);
	const double S = dsub(dmul(2.0, dsub(t1, t2)), s.JJ);
	const double B = dadd(dmul(3.552713678800501e-15 /* 2^-48 */, dadd(dadd(dabs(t1), dabs(t2)), s.JJ)),
			      dmul(dmul(3.469446951953614e-18 /* 2^-58 */, dadd(s.absE, dabs(p.phi))), r2));
	if (S > B) return 1;
	if (S < -B) return -1;
	(*n_exact)++;
	const double q = q_exact(t, s, p.r, p.phi);
	return q > 0.0 ? 1 : (q < 0.0 ? -1 : 0);
}
CMCB_HD int sign_Q(const OrbitTables &t, const StarConst &s, kint k, int *n_exact) { return sign_Q_at(t, s, tbl_rp(t, k), n_exact); }

// The same sign decided, where possible, from the FP32 copy of the table entry (8 bytes instead of 16: the copy of a 1e7-star table is 80 MB
// and largely stays in L2, the FP64 table does not).  With r^ = fl32(r), phi^ = fl32(phi): r = r^ (1 + dr), phi = phi^ (1 + dp), |dr|, |dp| <= 2^-24
// (round to nearest, both inside FP32's normal range -- checked), so for the real-number value S*(r, phi) = 2 ((E - phi) r^2 - t2(r)) - J^2
//   |S*(r, phi) - S*(r^, phi^)| <= 2 |phi - phi^| r^2 + 2 |E - phi^| |r^2 - r^^2| + 2 |t2(r) - t2(r^)|
//                               <= 2^-22 (|phi^| r^^2 + |t1^| + |t2^|)          (2^-23 (1 + 2^-23) < 2^-22; t2 is continuous across its branch at r_j,
//                                                                                so a branch taken on the wrong side of r_j costs <= 2^-23 |t2|)
// S^ is evaluated in FP64 like sign_Q_at, whose bound B (evaluation error + the reference's own noise) is taken at the rounded entry and doubled.
// |S^| > 2^-21 (|phi^| r^^2 + |t1^| + |t2^|) + 2 B therefore PROVES the sign the reference computes at the FP64 entry; otherwise the probe is
// repeated on the FP64 table (sign_Q: the last levels of a bisection, next to the sign change).  Same probe sequence, same signs, same result.
#ifdef CMCB_SCREEN_STATS
static long g_screen_stats[2]; // host-only instrumentation: [0] probes decided on the FP32 copy, [1] repeated on the FP64 table
#endif
CMCB_HD int sign_Q_screened(const OrbitTables &t, const StarConst &s, kint k, int *n_exact)
{
	if (t.rpf) {
#if defined(__CUDA_ARCH__)
		const float2 f = __ldg(reinterpret_cast<const float2 *>(t.rpf + k));
		const double r = (double) f.x, phi = (double) f.y;
#else
		const double r = (double) t.rpf[k].r, phi = (double) t.rpf[k].phi;
#endif
		const double aphi = dabs(phi);
		if (r > 1.0e-30 && r < 1.0e30 && aphi > 1.0e-30 && aphi < 1.0e30) { // FP32 normal range with room: the relative rounding bound holds
			const double a = dsub(s.E, phi);
			const double r2 = dmul(r, r);
			const double t1 = dmul(a, r2);
			const double t2 = (r >= s.r_j) ? dmul(s.mi, r) : dmul(s.c_in, r2);
			const double S = dsub(dmul(2.0, dsub(t1, t2)), s.JJ);
			const double B = dadd(dmul(3.552713678800501e-15 /* 2^-48 */, dadd(dadd(dabs(t1), dabs(t2)), s.JJ)),
					      dmul(dmul(3.469446951953614e-18 /* 2^-58 */, dadd(s.absE, aphi)), r2));
			const double Bf = dadd(dmul(4.76837158203125e-07 /* 2^-21 */, dadd(dadd(dmul(aphi, r2), dabs(t1)), dabs(t2))), dmul(2.0, B));
#ifdef CMCB_SCREEN_STATS
			g_screen_stats[(S > Bf || S < -Bf) ? 0 : 1]++;
#endif
			if (S > Bf) return 1;
			if (S < -Bf) return -1;
		}
	}
	return sign_Q(t, s, k, n_exact);
}

// the direction test of find_zero_Q (cmc_nr.c:342): function_q(kmin) < function_q(kmax) on the VALUES
CMCB_HD bool q_increasing(const OrbitTables &t, const StarConst &s, kint kmin, kint kmax, int *n_exact)
{
	double Bmin, Bmax;
	const RPhi lo = tbl_rp(t, kmin), hi = tbl_rp(t, kmax);
	double qmin = q_fp64(t, s, lo.r, lo.phi, &Bmin), qmax = q_fp64(t, s, hi.r, hi.phi, &Bmax);
	// 2^-52 |q| covers the final rounding to double of each value
	double slack = dadd(dadd(Bmin, Bmax), dmul(2.220446049250313e-16, dadd(dabs(qmin), dabs(qmax))));
	if (dabs(dsub(qmin, qmax)) > slack) return qmin < qmax;
	(*n_exact) += 2;
	return q_exact(t, s, lo.r, lo.phi) < q_exact(t, s, hi.r, hi.phi);
}

// Can function_q be > 0 (or undecided) anywhere on table entries [k_lo, k_hi]?  false = PROVEN <= 0 on the whole range:
//   Phi is non-decreasing in k (each phi_k = fl(phi_{k+1} - (a non-negative term)), cmc_utils.c:719), r is sorted, and the self
//   term is non-increasing in r, so for every k in the range  Q(k) <= U := 2 (E - Phi_lo - phis(r_hi)) - J^2 / r_hi^2,
//   and |Q(k)| >= |U| + (J^2/r_k^2 - J^2/r_hi^2) keeps the reference's computed sign on the side of the real-number sign
//   (its noise is a 2^-51 multiple of |E| + |Phi_k| + phis + J^2/r_k^2).  U is evaluated in the division-free form of sign_Q_at
//   with a 64x wider margin (2^-44): the test only has to dismiss ranges that are nowhere near a sign change.
CMCB_HD bool q_maybe_positive_on(const StarConst &s, const RPhi lo, const RPhi hi)
{
	const double a = dsub(s.E, lo.phi);
	const double r2 = dmul(hi.r, hi.r);
	const double t1 = dmul(a, r2);
	const double t2 = (hi.r >= s.r_j) ? dmul(s.mi, hi.r) : dmul(s.c_in, r2);
	const double S = dsub(dmul(2.0, dsub(t1, t2)), s.JJ);
	const double mag = dadd(dadd(dmul(dadd(dadd(s.absE, dabs(lo.phi)), s.c_in), r2), s.JJ), dadd(dabs(t1), dabs(t2)));
	return !(S < -dmul(5.684341886080802e-14 /* 2^-44 */, mag));
}

// The two-sided form on a ladder block (see "ladders" above): -1 = the reference's function_q is PROVEN < 0 at every table entry
// of the block, +1 = proven > 0, 0 = undecided (the caller descends to a finer level).
//   Upper side (the argument of q_maybe_positive_on with Phi_lo -> min Phi, |Phi| -> max |Phi| of the block):
//     U r_hi^2 := 2 ((E - Phi_min) r_hi^2 - phis(r_hi) r_hi^2) - J^2 >= Q*(k) r_hi^2 + [(J^2/r_k^2 - J^2/r_hi^2) + 2 (phis(r_k) - phis(r_hi))] r_hi^2;
//     the reference's noise at k is <= 2^-51 (J^2/r_k^2 + |E| + |Phi_k| + phis(r_k)), whose excess over the same expression at r_hi
//     is 2^-51 times the bracket, so computed Q(k) < 0 whenever -U r_hi^2 > 2^-51 (J^2 + (|E| + max|Phi| + c_in) r_hi^2).
//   Lower side:  L r_lo^2 := 2 ((E - Phi_max) r_lo^2 - phis(r_lo) r_lo^2) - J^2 <= Q*(k) r_lo^2 (Phi_k <= Phi_max, phis non-increasing,
//     r_k >= r_lo), and the noise at k times r_lo^2 is <= 2^-51 (J^2 (r_lo/r_k)^2 + (|E| + |Phi_k| + phis(r_k)) r_lo^2)
//     <= 2^-51 (J^2 + (|E| + max|Phi| + c_in) r_lo^2).
//   Each side is evaluated in the division-free FP64 form of sign_Q_at, S = 2 (t1 - t2) - JJ with |S - exact| <= 2^-50 (|t1| + |t2| + JJ),
//   against the margin 2^-44 mag, mag := (|E| + max|Phi| + c_in) r^2 + JJ  (>= (|t1| + |t2| + JJ) (1 - 2^-50) since |t1| <= (|E| + |Phi|) r^2
//   and t2 <= c_in r^2): 40x the evaluation error and the reference's noise (2^-50 + 2^-51) mag together.
CMCB_HD int q_sign_on_block(const StarConst &s, const LBlock B)
{
	const double pabs = dmax(dabs(B.phi_min), dabs(B.phi_max));
	const double base = dadd(dadd(s.absE, pabs), s.c_in);
	{
		const double a = dsub(s.E, B.phi_min);
		const double r2 = dmul(B.r_hi, B.r_hi);
		const double t1 = dmul(a, r2);
		const double t2 = (B.r_hi >= s.r_j) ? dmul(s.mi, B.r_hi) : dmul(s.c_in, r2);
		const double S = dsub(dmul(2.0, dsub(t1, t2)), s.JJ);
		if (S < -dmul(5.684341886080802e-14 /* 2^-44 */, dadd(dmul(base, r2), s.JJ))) return -1;
	}
	{
		const double a = dsub(s.E, B.phi_max);
		const double r2 = dmul(B.r_lo, B.r_lo);
		const double t1 = dmul(a, r2);
		const double t2 = (B.r_lo >= s.r_j) ? dmul(s.mi, B.r_lo) : dmul(s.c_in, r2);
		const double S = dsub(dmul(2.0, dsub(t1, t2)), s.JJ);
		if (S > dmul(5.684341886080802e-14 /* 2^-44 */, dadd(dmul(base, r2), s.JJ))) return 1;
	}
	return 0;
}

#ifdef CMCB_LADDER_STATS
static long g_lad_stats[17]; // host-only instrumentation: [2l] proven at level l, [2l+1] undecided at level l, [16] fine probes
#endif
// sign of function_q at table index k during a bisection whose bracket is `width` = kmax - kmin entries wide: answered from the
// coarsest ladder level whose blocks are at most an eighth of the bracket (the sign change sits somewhere in the bracket, the probe in
// its middle: a block that small is rarely the one that straddles it), descending on "undecided", the fine table last.
CMCB_HD int floor_log2(unsigned v) // v >= 1
{
#if defined(__CUDA_ARCH__)
	return 31 - __clz((int) v);
#else
	return 31 - __builtin_clz(v);
#endif
}

CMCB_HD int sign_Q_laddered(const OrbitTables &t, const QLadder &L, const LBlock *top, const StarConst &s, kint k, kint width, int *n_exact)
{
	if (width < L.max_width) {
		// first level with (1 << (shift + factor_log2)) <= width: shifts go top, top - STEP, ...
		const int want = floor_log2((unsigned) width) - L.factor_log2; // largest usable shift
		int l = want >= L.shift[0] ? 0 : (L.shift[0] - want + QLAD_STEP - 1) / QLAD_STEP;
		for (; l < L.n_levels; l++) {
			const int sh = L.shift[l];
			const LBlock B = (l == 0) ? top_ld(top + (k >> sh)) : blk_ld(L.blk[l] + (k >> sh));
			const int sg = q_sign_on_block(s, B);
#ifdef CMCB_LADDER_STATS
			g_lad_stats[sg ? 2 * l : 2 * l + 1]++;
#endif
			if (sg) return sg;
		}
	}
#ifdef CMCB_LADDER_STATS
	g_lad_stats[16]++;
#endif
	return sign_Q(t, s, k, n_exact);
}

// ---- find_zero_Q through the sample ladder ---------------------------------------------------------------------------------------------
//
// The reference's bisection (cmc_nr.c:316-383) probes ktry = (kmin + kmax + 1) / 2 about log2(N) times, every probe a dependent load from a
// table that does not fit the cache.  Almost all of those probes are far from the sign change, and their signs follow from a handful of
// entries:
//
//   U_k := Phi_k + phis(r_k) (the potential of all OTHER stars at r_k) is, as a function of x = 1 / r, convex over the entries 1 .. N_MAX + 1:
//   potential_calculate builds Phi_k = Phi_{k+1} - M_k (x_k - x_{k+1}) with the enclosed mass M_k non-decreasing in k (cmc_utils.c:716-720), and
//   the self term removes the star's own share of M_k.  Hence Q*(k) = 2 (E - U_k) - J^2 x_k^2 is CONCAVE over the entries: for k1 < k2 < k3,
//   Q*(k2) >= min(Q*(k1), Q*(k3)), and left of an entry with larger Q* it is non-decreasing, right of one non-increasing.
//
// So with an entry p where Q > 0 (the bracket end that get_positive_Q_index found) and an entry a on the other side of the sign change with
// Q(a) < -margin, EVERY entry beyond a (away from p) has Q < 0; and with an entry b between the sign change and p with Q(b) > +margin and
// Q(p) > +margin, every entry between b and p has Q > 0.  The margin absorbs what separates the stored table and the reference's evaluation
// from that idealisation:
//   - the stored Phi_k are sequentially rounded: over a span of n entries they leave an exactly convex sequence by at most n 2^-49 max|Phi|
//     (five roundings per step of the recurrence, each <= 2^-52 (|Phi| + M/r));
//   - the reference's own evaluation noise at an entry k, 2^-51 (|E| + |U_k| + J^2 / r_k^2): by the argument in the note below it never
//     exceeds the margin at a plus the gap Q*(a) - Q*(k);
//   - the FP64 evaluation of the margin test itself (sign_Q_at's bound B).
//   margin = (N_MAX 2^-47 + 2^-40) (|E| + max|Phi| + J^2 / r^2), times r^2 in the division-free form: 7e-8 relative at N = 1e7 -- the band of
//   entries it leaves undecided is a fraction of an index wide.
//   (Note: for k beyond a, -Q*(k) = -Q*(a) + G with G = [J^2 (x_k^2 - x_a^2) - 2 (U_a - U_k)] >= 0, and the reference's noise at k is
//    2^-51 (|E| + |U_k| + J^2 x_a^2 + J^2 (x_k^2 - x_a^2)) <= 2^-51 (|E| + 3 max|Phi| + J^2 x_a^2) + 2^-51 G < margin(a) + G.)
//
// The search: from the coarsest level (in shared memory) down, a binary search over the level's entries inside the current window moves the
// window's ends onto entries whose sign is proven with that margin; an entry inside the margin ends the descent.  Then the reference's
// bisection is replayed: a probe outside the window takes the proven sign WITHOUT touching memory, a probe inside is evaluated exactly like
// before (sign_Q: FP64 with a rigorous bound, the x87 arithmetic replayed inside it).  Same probe sequence, same signs, same k_min / k_max --
// with about 3 dependent line fetches instead of 24.  Whatever the proof needs and cannot be established (a bracket end inside the margin:
// a star at its turning point; an unsorted table) falls back to the plain bisection.
CMCB_HD int side_Q_margin(const StarConst &s, const RPhi p, double margin, double phi_abs_max, float *q_approx = nullptr)
{
	const double a = dsub(s.E, p.phi);
	const double r2 = dmul(p.r, p.r);
	const double t1 = dmul(a, r2);
	const double t2 = (p.r >= s.r_j) ? dmul(s.mi, p.r) : dmul(s.c_in, r2);
	const double S = dsub(dmul(2.0, dsub(t1, t2)), s.JJ);
	const double Bm = dmul(margin, dadd(dadd(dmul(dadd(s.absE, phi_abs_max), r2), s.JJ), dadd(dabs(t1), dabs(t2))));
	if (q_approx) *q_approx = (float) S / (float) r2; // Q itself, roughly: only steers the search below, never decides a sign
	if (S > Bm) return 1;
	if (S < -Bm) return -1;
	return 0;
}

// The same test in FP32 ARITHMETIC on an FP32 copy of the sample, for the coarsest level: its ~12 probes per bisection are far from the sign
// change (|Q| is a sizeable fraction of its terms), stage A is bound by the FP64 pipe and by the shared-memory wavefronts of 16-byte entries
// at unrelated addresses, and an FP32 probe costs a third of either.  Every input is rounded to FP32 once (2^-24 relative), the six
// operations add as much again, the r >= r_j branch may be taken the other way within 2^-23 of r_j (where both forms of the self term
// agree to that much): |S32 - S| <= 2^-20 (the sum of the terms' magnitudes), so with the margin raised by 2^-19 a decided FP32 probe
// implies the decided FP64 probe of the same sign.  0 = undecided here: the caller evaluates side_Q_margin on the FP64 sample.
struct StarConstF {
	float E, JJ, mi, c_in, r_j, scaleA, margin;
};
CMCB_HD StarConstF star_const_f32(const StarConst &s, double margin, double phi_abs_max)
{
	StarConstF f;
	f.E = (float) s.E; f.JJ = (float) s.JJ; f.mi = (float) s.mi; f.c_in = (float) s.c_in; f.r_j = (float) s.r_j;
	f.scaleA = (float) dadd(s.absE, phi_abs_max) * 1.000001f;
	f.margin = (float) margin * 1.000001f + 1.9073486e-6f /* 2^-19 */;
	return f;
}
CMCB_HD int side_Q_margin_f32(const StarConstF &f, float r, float phi)
{
	const float a = f.E - phi;
	const float r2 = r * r;
	const float t1 = a * r2;
	const float t2 = (r >= f.r_j) ? f.mi * r : f.c_in * r2;
	const float S = 2.0f * (t1 - t2) - f.JJ;
	const float at1 = t1 < 0.0f ? -t1 : t1, at2 = t2 < 0.0f ? -t2 : t2;
	const float scale = f.scaleA * r2 + f.JJ + at1 + at2;
	if (!(scale > 1.0e-30f && scale < 1.0e30f && r2 > 1.0e-30f)) return 0; // outside FP32's comfortable range (or NaN): FP64 decides
	const float B = f.margin * scale;
	if (S > B) return 1;
	if (S < -B) return -1;
	return 0;
}

CMCB_HD RPhi slad_ld(const RPhi *p)
{
#if defined(__CUDA_ARCH__)
	const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
	RPhi o; o.r = v.x; o.phi = v.y;
	return o;
#else
	return *p;
#endif
}

CMCB_HD RPhi slad_top_ld(const RPhi *p)
{
#if defined(__CUDA_ARCH__)
	const double2 v = *reinterpret_cast<const double2 *>(p);
	RPhi o; o.r = v.x; o.phi = v.y;
	return o;
#else
	return *p;
#endif
}

CMCB_HD kint find_zero_Q_plain(const OrbitTables &t, const StarConst &s, kint kmin, kint kmax, bool increasing, int *n_exact)
{
	do {
		kint ktry = (kmin + kmax + 1) / 2;
		int sg = sign_Q_screened(t, s, ktry, n_exact);
		bool up = increasing ? (sg < 0) : (sg > 0); // :349 / :362
		if (up) kmin = ktry; else kmax = ktry - 1;
	} while (kmax != kmin);
	return kmin;
}

CMCB_HD kint find_zero_Q_sampled(const OrbitTables &t, const LadderTops &top, const StarConst &s, kint kmin, kint kmax, bool increasing, int *n_exact)
{
	const SLadder &L = *top.sdesc;
	// the kmin side of the sign change has sign `low` (the probes there answer "up"), the kmax side -low
	const int low = increasing ? -1 : 1;
	// both bracket ends must carry their sign with the margin: the end where Q > 0 anchors the concavity argument.  Entry 0 (r = ZERO, the
	// central-mass slot) is never probed by the bisection and lies outside the convex sequence: Q there is -J^2 / ZERO^2.
	if (kmax - kmin < 2 * (1 << SLAD_MIN_SHIFT)) return find_zero_Q_plain(t, s, kmin, kmax, increasing, n_exact);
	float q_lo = 0.0f, q_hi = 0.0f; // approximate Q at the bracket ends (0: unknown)
	if (kmin > 0 && side_Q_margin(s, tbl_rp(t, kmin), L.margin, L.phi_abs_max, &q_lo) != low) return find_zero_Q_plain(t, s, kmin, kmax, increasing, n_exact);
	if (kmin == 0 && !(increasing && s.JJ > 0.0)) return find_zero_Q_plain(t, s, kmin, kmax, increasing, n_exact);
	if (side_Q_margin(s, tbl_rp(t, kmax), L.margin, L.phi_abs_max, &q_hi) != -low) return find_zero_Q_plain(t, s, kmin, kmax, increasing, n_exact);
	kint lo = kmin, hi = kmax; // every entry in [kmin, lo] has the sign `low`, every entry in [hi, kmax] the sign -low
#if CMCB_SLAD_F32_TOP
	StarConstF sf32 = {};
	if (top.sf) sf32 = star_const_f32(s, L.margin, L.phi_abs_max);
#endif
	for (int l = 0; l < L.n_levels; l++) {
		const int sh = L.shift[l];
		const RPhi *lvl = (l == 0) ? top.s : L.lvl[l];
		kint a = lo >> sh, b = ((hi - 1) >> sh) + 1; // level entries a+1 .. b-1 lie strictly inside (lo, hi)
		const kint a0 = a, b0 = b;
		bool undecided = false;
		// Any entry strictly between a and b may be probed next: the invariant (entries <= a carry `low`, entries >= b carry -low) does not care.
		// On the coarsest level (thousands of entries, ~12 halvings) the next probe is therefore taken where a secant through the approximate
		// Q values at the window's ends crosses zero (Illinois variant: an end that stays put has its value halved), falling back to halving
		// after SLAD_SECANT_STEPS probes or whenever the values do not bracket zero.  Q over the entries is concave (see above): 4-6 probes.
		float xa = (float) lo / (float) (1 << sh), xb = (float) hi / (float) (1 << sh), qa = q_lo, qb = q_hi;
		int steps = 0, last = 0;
		while (b - a > 1) {
			kint mid = (a + b) >> 1;
			if (l == 0 && steps < SLAD_SECANT_STEPS && b - a > 4 && qa * qb < 0.0f) {
				const float xm = xa + (xb - xa) * (qa / (qa - qb));
				const kint g = (kint) xm;
				if (g > a && g < b) mid = g;
				else if (xm <= (float) a) mid = a + 1;
				else if (xm >= (float) b) mid = b - 1;
			}
			steps++;
			// the coarsest level sits in shared memory in the kernels (a generic load), the others are read through the read-only path
			float qm = 0.0f;
			int sd = 0;
#if CMCB_SLAD_F32_TOP
			if (l == 0 && top.sf) { // FP32 screening of the coarsest level
#if defined(__CUDA_ARCH__)
				const float2 v = *reinterpret_cast<const float2 *>(top.sf + mid);
				sd = side_Q_margin_f32(sf32, v.x, v.y);
#else
				sd = side_Q_margin_f32(sf32, top.sf[mid].r, top.sf[mid].phi);
#endif
			}
#endif
			if (sd == 0) {
				const RPhi smp = (l == 0) ? slad_top_ld(lvl + mid) : slad_ld(lvl + mid);
				sd = side_Q_margin(s, smp, L.margin, L.phi_abs_max, SLAD_SECANT_STEPS ? &qm : nullptr);
			}
			if (sd == low) {
				a = mid; xa = (float) mid; qa = qm;
				if (last == 1) qb *= 0.5f;
				last = 1;
			} else if (sd == -low) {
				b = mid; xb = (float) mid; qb = qm;
				if (last == 2) qa *= 0.5f;
				last = 2;
			} else { undecided = true; break; }
		}
		if (a > a0) lo = a << sh;
		if (b < b0) hi = b << sh;
		if (undecided) break;
	}
	// the reference's probe sequence; only the probes strictly inside (lo, hi) read the table
	do {
		kint ktry = (kmin + kmax + 1) / 2;
		bool up;
		if (ktry <= lo) up = true;
		else if (ktry >= hi) up = false;
		else {
			const int sg = sign_Q(t, s, ktry, n_exact);
			up = increasing ? (sg < 0) : (sg > 0);
		}
		if (up) kmin = ktry; else kmax = ktry - 1;
	} while (kmax != kmin);
	return kmin;
}

// Both index bisections of calc_orbit_new (cmc_orbit.c:441, :444) through the sample ladder with their descents run in LOCKSTEP: the two
// searches are independent (k_min left of the star's own entry, k_max right of it), so their probes on every level are issued together --
// two independent chains of sample load -> margin test per thread instead of one after the other.  Each search is exactly
// find_zero_Q_sampled: same proven-sign windows, same replay of the reference's probe sequence, same fall-backs.
struct SladOne {
	kint kmin, kmax, lo, hi, a, b, a0, b0;
	int low;
	bool increasing, sampled, descending, undecided;
};
CMCB_HD void slad_one_init(const OrbitTables &t, const SLadder &L, const StarConst &s, SladOne &q, kint kmin, kint kmax, int *n_exact)
{
	q.kmin = kmin; q.kmax = kmax; q.lo = kmin; q.hi = kmax;
	q.increasing = q_increasing(t, s, kmin, kmax, n_exact);
	q.low = q.increasing ? -1 : 1;
	q.sampled = kmax - kmin >= 2 * (1 << SLAD_MIN_SHIFT);
	if (q.sampled && kmin > 0 && side_Q_margin(s, tbl_rp(t, kmin), L.margin, L.phi_abs_max) != q.low) q.sampled = false;
	if (q.sampled && kmin == 0 && !(q.increasing && s.JJ > 0.0)) q.sampled = false;
	if (q.sampled && side_Q_margin(s, tbl_rp(t, kmax), L.margin, L.phi_abs_max) != -q.low) q.sampled = false;
	q.descending = q.sampled;
	q.undecided = false;
	q.a = q.b = q.a0 = q.b0 = 0;
}
CMCB_HD kint slad_one_finish(const OrbitTables &t, const StarConst &s, SladOne &q, int *n_exact)
{
	if (!q.sampled) return find_zero_Q_plain(t, s, q.kmin, q.kmax, q.increasing, n_exact);
	kint kmin = q.kmin, kmax = q.kmax;
	do {
		kint ktry = (kmin + kmax + 1) / 2;
		bool up;
		if (ktry <= q.lo) up = true;
		else if (ktry >= q.hi) up = false;
		else {
			const int sg = sign_Q(t, s, ktry, n_exact);
			up = q.increasing ? (sg < 0) : (sg > 0);
		}
		if (up) kmin = ktry; else kmax = ktry - 1;
	} while (kmax != kmin);
	return kmin;
}
CMCB_HD void find_zero_Q_sampled_pair(const OrbitTables &t, const LadderTops &top, const StarConst &s, kint ktemp, kint *kmin_out, kint *kmax_out, int *n_exact)
{
	const SLadder &L = *top.sdesc;
	SladOne q1, q2;
	slad_one_init(t, L, s, q1, 0, ktemp, n_exact);
	slad_one_init(t, L, s, q2, ktemp, t.N_MAX + 1, n_exact);
	for (int l = 0; l < L.n_levels && (q1.descending || q2.descending); l++) {
		const int sh = L.shift[l];
		const RPhi *lvl = (l == 0) ? top.s : L.lvl[l];
		q1.a = q1.a0 = q1.lo >> sh; q1.b = q1.b0 = ((q1.hi - 1) >> sh) + 1;
		q2.a = q2.a0 = q2.lo >> sh; q2.b = q2.b0 = ((q2.hi - 1) >> sh) + 1;
		for (;;) {
			const bool act1 = q1.descending && !q1.undecided && q1.b - q1.a > 1, act2 = q2.descending && !q2.undecided && q2.b - q2.a > 1;
			if (!act1 && !act2) break;
			const kint m1 = act1 ? (q1.a + q1.b) >> 1 : q1.a0, m2 = act2 ? (q2.a + q2.b) >> 1 : q2.a0; // an idle search re-reads an entry it owns: no branch around the loads
			const RPhi s1 = (l == 0) ? slad_top_ld(lvl + m1) : slad_ld(lvl + m1);
			const RPhi s2 = (l == 0) ? slad_top_ld(lvl + m2) : slad_ld(lvl + m2);
			const int d1 = side_Q_margin(s, s1, L.margin, L.phi_abs_max), d2 = side_Q_margin(s, s2, L.margin, L.phi_abs_max);
			if (act1) { if (d1 == q1.low) q1.a = m1; else if (d1 == -q1.low) q1.b = m1; else q1.undecided = true; }
			if (act2) { if (d2 == q2.low) q2.a = m2; else if (d2 == -q2.low) q2.b = m2; else q2.undecided = true; }
		}
		if (q1.descending) { if (q1.a > q1.a0) q1.lo = q1.a << sh; if (q1.b < q1.b0) q1.hi = q1.b << sh; if (q1.undecided) q1.descending = false; }
		if (q2.descending) { if (q2.a > q2.a0) q2.lo = q2.a << sh; if (q2.b < q2.b0) q2.hi = q2.b << sh; if (q2.undecided) q2.descending = false; }
	}
	*kmin_out = slad_one_finish(t, s, q1, n_exact);
	*kmax_out = slad_one_finish(t, s, q2, n_exact);
}

// find_zero_Q (cmc_nr.c:316-383): the same probe sequence ktry = (kmin+kmax+1)/2
CMCB_HD kint find_zero_Q(const OrbitTables &t, const LadderTops &top, const StarConst &s, kint kmin, kint kmax, int *n_exact)
{
	const bool increasing = q_increasing(t, s, kmin, kmax, n_exact);
	if (top.s != nullptr) return find_zero_Q_sampled(t, top, s, kmin, kmax, increasing, n_exact);
	const bool laddered = top.q != nullptr;
	do {
		kint ktry = (kmin + kmax + 1) / 2;
		int sg = laddered ? sign_Q_laddered(t, *top.qdesc, top.q, s, ktry, kmax - kmin, n_exact) : sign_Q_screened(t, s, ktry, n_exact);
		bool up = increasing ? (sg < 0) : (sg > 0); // :349 / :362
		if (up) kmin = ktry; else kmax = ktry - 1;
	} while (kmax != kmin);
	return kmin;
}

// the linear scan of get_positive_Q_index (cmc_orbit.c:46-57): first k from 0 with Q > 0; -1 if there is none below N_MAX.
// Sequential form (host emulation, and the device's last resort); the CUDA path runs it as a grid-wide find-first.
CMCB_HD kint scan_positive_Q(const OrbitTables &t, const StarConst &s, int *n_exact)
{
	kint ktemp = -1;
	int sg;
	do {
		ktemp++;
		sg = sign_Q(t, s, ktemp, n_exact);
	} while (sg <= 0 && ktemp <= t.N_MAX);
	return ktemp >= t.N_MAX ? -1 : ktemp;
}

// set_a_b (cmc_orbit.c:722-740)
CMCB_HD void set_a_b(const OrbitTables &t, const StarConst &s, kint k, double *a, double *b)
{
	if (k == 0) {
		*a = dmul(-t.cenma_m, t.madhoc);
		*b = dadd(dadd(tbl_phi(t, 1), ddiv(dmul(t.cenma_m, t.madhoc), tbl_r(t, 1))), ddiv(dmul(s.m_j, t.madhoc), s.r_j));
	} else {
		const RPhi pk = tbl_rp(t, k), pk1 = tbl_rp(t, k + 1);
		double rk = pk.r, rk1 = pk1.r;
		double Uk = dadd(pk.phi, phi_self(rk, s.r_j, s.m_j, t.invN));
		double Uk1 = dadd(pk1.phi, phi_self(rk1, s.r_j, s.m_j, t.invN));
		double den = dsub(ddiv(1., rk1), ddiv(1., rk));
		*a = ddiv(dsub(Uk1, Uk), den);
		*b = ddiv(dsub(ddiv(Uk, rk1), ddiv(Uk1, rk)), den);
	}
}

// calc_vr_within_interval (cmc_nr.c:393-405 + 279-302): vr^2 inside [r_k, r_k+1]; Brent needs the VALUE, so this is
// always the exact x87 replay (rare path)
CMCB_HD double vr2_within_interval(const OrbitTables &t, const StarConst &s, kint k, double r)
{
	const double EPS = 2.220446049250313e-16; // DBL_EPSILON
	double rk = tbl_r(t, k), rk1 = tbl_r(t, k + 1), pot;
	if (dabs(dsub(r, rk)) < EPS) pot = tbl_phi(t, k);
	else if (dabs(dsub(r, rk1)) < EPS) pot = tbl_phi(t, k + 1);
	else if (r < tbl_r(t, 1)) pot = dsub(tbl_phi(t, 0), ddiv(dmul(t.cenma_m, t.madhoc), r));
	else {
		double pk = tbl_phi(t, k), inv_k = ddiv(1.0, rk);
		pot = dadd(pk, ddiv(dmul(dsub(tbl_phi(t, k + 1), pk), dsub(inv_k, ddiv(1.0, r))), dsub(inv_k, ddiv(1.0, rk1))));
	}
	return q_exact(t, s, r, pot);
}

// find_root_vr (cmc_nr.c:500-603): Brent-Dekker as GSL's gsl_root_fsolver_brent iterates it (un-vendored
// dependency of the reference, version unpinned -> this branch is "parity unpinned"), APSIDES_PRECISION = 1e-11,
// APSIDES_MAX_ITER = 100, APSIDES_CONVERGENCE = 5e-13 (cmc_io.c:1752-1754).
CMCB_COLD double find_root_vr(const RPhi *rp, double invN, double madhoc, double cenma_m, const StarConst s, kint k) // by value: see q_exact_cold
{
	OrbitTables t; // only what vr2_within_interval reads
	t.rp = rp; t.m = nullptr; t.N_MAX = 0; t.invN = invN; t.madhoc = madhoc; t.cenma_m = cenma_m;
	t.qlad.n_levels = 0; t.rlad.n_levels = 0;
	const double PREC = 1.0e-11, CONV = 5.e-13, EPS = 2.220446049250313e-16;
	int iter = 100;
	double a = tbl_r(t, k), b = tbl_r(t, k + 1), c = b;
	double fa = vr2_within_interval(t, s, k, a), fb = vr2_within_interval(t, s, k, b), fc = fb;
	double d = dsub(b, a), e = d, root = dmul(0.5, dadd(a, b)), lo = a, hi = b, prev_apsis = -1.0;
	bool not_converged = true;
	while (not_converged && iter) {
		bool ac_equal = false;
		if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) { ac_equal = true; c = a; fc = fa; d = dsub(b, a); e = d; }
		if (dabs(fc) < dabs(fb)) { ac_equal = true; a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
		double tol = dmul(dmul(0.5, EPS), dabs(b));
		double m = dmul(0.5, dsub(c, b));
		if (fb == 0) { root = lo = hi = b; }
		else if (dabs(m) <= tol) { root = b; if (b < c) { lo = b; hi = c; } else { lo = c; hi = b; } }
		else {
			if (dabs(e) < tol || dabs(fa) <= dabs(fb)) { d = m; e = m; }
			else {
				double p, q, r, sf = ddiv(fb, fa);
				if (ac_equal) { p = dmul(dmul(2, m), sf); q = dsub(1, sf); }
				else {
					q = ddiv(fa, fc); r = ddiv(fb, fc);
					p = dmul(sf, dsub(dmul(dmul(dmul(2, m), q), dsub(q, r)), dmul(dsub(b, a), dsub(r, 1))));
					q = dmul(dmul(dsub(q, 1), dsub(r, 1)), dsub(sf, 1));
				}
				if (p > 0) q = -q; else p = -p;
				double lim1 = dsub(dmul(dmul(3, m), q), dabs(dmul(tol, q))), lim2 = dabs(dmul(e, q));
				if (dmul(2, p) < (lim1 < lim2 ? lim1 : lim2)) { e = d; d = ddiv(p, q); } else { d = m; e = m; }
			}
			a = b; fa = fb;
			if (dabs(d) > tol) b = dadd(b, d); else b = dadd(b, (m > 0 ? tol : -tol));
			fb = vr2_within_interval(t, s, k, b);
			root = b;
			if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) c = a;
			if (b < c) { lo = b; hi = c; } else { lo = c; hi = b; }
		}
		{
			double al = dabs(lo), ah = dabs(hi);
			double min_abs = ((lo > 0.0 && hi > 0.0) || (lo < 0.0 && hi < 0.0)) ? (al < ah ? al : ah) : 0;
			not_converged = !(dabs(dsub(hi, lo)) < dadd(PREC, dmul(PREC, min_abs)));
		}
		iter--;
		if (iter < 10) {
			if (prev_apsis < 0.) prev_apsis = root;
			else {
				bool delta_continue = !(dabs(dsub(root, prev_apsis)) < dadd(CONV, dmul(CONV, dabs(root))) || root == prev_apsis);
				not_converged = not_converged && delta_continue;
				prev_apsis = root;
			}
		}
	}
	double apsis = root;
	if (vr2_within_interval(t, s, k, apsis) < 0.) {
		if (vr2_within_interval(t, s, k, lo) < 0.) apsis = hi;
		if (vr2_within_interval(t, s, k, hi) < 0.) apsis = lo;
	}
	return apsis;
}

struct OrbitRS { // orbit_rs_t, reference include/cmc/cmc.h:1621-1650
	double rp, ra, dQdrp, dQdra;
	kint kmin, kmax;
	int circular;
};

// calc_orbit_new (cmc_orbit.c:413-540) incl. get_positive_Q_index (:30-59)
// ktemp_in: KTEMP_COMPUTE, or the result of the linear scan found elsewhere (>= -1).  need_scan != nullptr: instead of running the
// O(N) scan in this thread, set *need_scan and return (the caller queues the star for the cooperative scan).
constexpr kint KTEMP_COMPUTE = -2;
CMCB_HD OrbitRS calc_orbit_new(const OrbitTables &t, const LadderTops &top, const StarConst &s, bool is_new, double rnew_of_new, int *n_exact, int *n_brent,
			       kint ktemp_in = KTEMP_COMPUTE, bool *need_scan = nullptr)
{
	OrbitRS o;
	kint ktemp = ktemp_in;
	if (ktemp_in == KTEMP_COMPUTE) {
		ktemp = s.g_j;
		if (is_new) ktemp = find_zero_r(t, 0, t.N_MAX, rnew_of_new) + 1; // cmc_orbit.c:39-41
		if (sign_Q(t, s, ktemp, n_exact) <= 0) {                          // :46-57 linear scan from 0
			if (need_scan) {
				*need_scan = true;
				o.rp = o.ra = o.dQdrp = o.dQdra = 0.0; o.kmin = o.kmax = -1; o.circular = 0;
				return o;
			}
			ktemp = scan_positive_Q(t, s, n_exact);
		}
	}
	bool circular = ktemp <= -1;
	double rmin = 0, rmax = 0, dQdr_min = 0, dQdr_max = 0;
	kint kmin = 0, kmax = 0;
	if (!circular) {
#if CMCB_SLAD_PAIR
		if (top.s != nullptr) find_zero_Q_sampled_pair(t, top, s, ktemp, &kmin, &kmax, n_exact); // :441, :444 with their descents in lockstep
		else
#endif
		{
			kmin = find_zero_Q(t, top, s, 0, ktemp, n_exact);               // :441
			kmax = find_zero_Q(t, top, s, ktemp, t.N_MAX + 1, n_exact);     // :444
		}
		double a, b;
		const double J = s.J, E = s.E;
		double JJ2 = dmul(dmul(2.0, J), J);
		set_a_b(t, s, kmin, &a, &b);                               // :456
		double inside = dsub(dmul(a, a), dmul(JJ2, dsub(b, E)));
		if (inside < 0.0) {                                        // :459-474
			rmin = ddiv(dmul(dmul(-1.0, J), J), a);
			if (rmin < tbl_r(t, kmin) || rmin > tbl_r(t, kmin + 1)) { rmin = find_root_vr(t.rp, t.invN, t.madhoc, t.cenma_m, s, kmin); (*n_brent)++; }
		} else {
			rmin = ddiv(dmul(J, J), dadd(-a, dsqrt(inside)));      // :476
		}
		dQdr_min = dadd(ddiv(JJ2, dmul(dmul(rmin, rmin), rmin)), ddiv(dmul(2.0, a), dmul(rmin, rmin))); // :479
		set_a_b(t, s, kmax, &a, &b);                               // :482
		inside = dsub(dmul(a, a), dmul(JJ2, dsub(b, E)));
		if (inside < 0.0) {                                        // :485-500
			rmax = ddiv(-a, dmul(2.0, dsub(b, E)));
			if (rmax < tbl_r(t, kmax) || rmax > tbl_r(t, kmax + 1)) { rmax = find_root_vr(t.rp, t.invN, t.madhoc, t.cenma_m, s, kmax); (*n_brent)++; }
		} else {
			rmax = ddiv(dadd(-a, dsqrt(inside)), dmul(2.0, dsub(b, E))); // :502
		}
		dQdr_max = dadd(ddiv(JJ2, dmul(dmul(rmax, rmax), rmax)), ddiv(dmul(2.0, a), dmul(rmax, rmax))); // :505
		if (rmin > rmax || dQdr_min <= 0. || dQdr_max >= 0.) circular = true; // :510
	}
	if (circular) { // :520-528
		o.rp = o.ra = s.r_j; o.kmin = o.kmax = s.g_j; o.dQdrp = o.dQdra = 0.0; o.circular = 1;
	} else {
		o.rp = rmin; o.ra = rmax; o.kmin = kmin; o.kmax = kmax; o.dQdrp = dQdr_min; o.dQdra = dQdr_max; o.circular = 0;
	}
	return o;
}

// One iteration of the loop body of get_positions_loop (cmc_evolution_thr.c:802-933).
// E_star is star[j].E (without the self term), Draws provides `double next()` and `bool has(int n)`.
// the star's constants as the loop body sets them up (:803-807)
CMCB_HD StarConst star_const(const OrbitTables &t, kint g_j, double E_star, double J)
{
	StarConst s;
	s.g_j = g_j;
	s.r_j = tbl_r(t, g_j);
	s.m_j = tbl_ld(t.m, g_j);
	s.J = J;
	s.E = dadd(E_star, phi_self(s.r_j, s.r_j, s.m_j, t.invN)); // :806
	star_const_derive(t, s);
	return s;
}

constexpr int ORBIT_DEFERRED = 6; // internal to the CUDA path: the star waits for the cooperative linear scan; never leaves the library
constexpr int ORBIT_PENDING = 7;  // internal: the orbit is known (job), the rejection sampling of the new radius is still to do

// What the rejection loop (:862-888) needs besides the star's constants: the orbit's turning points and the envelope F.
struct OrbitJob {
	double rmin, rmax, F;
};

// Everything of one loop iteration of get_positions_loop BEFORE the rejection sampling (:803-860): the early exits (massless,
// unbound, circular, tidally stripped) are final and fill `out`; otherwise the status is ORBIT_PENDING and `job` is set.
// (for ORBIT_CIRCULAR the caller fills vrnew / vtnew with the star's current vr / vt, :837-838)
CMCB_HD void orbit_prepare(const OrbitTables &t, const LadderTops &top, const StarConst &s, bool is_new, double rnew_of_new, double Rtidal,
			   kint ktemp_in, bool allow_defer, OrbitOut &out, OrbitJob &job)
{
	out.n_draws = 0; out.n_exact = 0; out.n_brent = 0;
	out.kmin = out.kmax = -1;
	out.X = out.r_peri = out.r_apo = 0.0;
	if (s.m_j < CMCB_ZERO) {               // :816-820 (destroy_obj: rnew = SF_INFINITY, vrnew = vtnew = 0, :760-765)
		out.status = ORBIT_MASSLESS; out.rnew = CMCB_SF_INFINITY; out.vrnew = 0.0; out.vtnew = 0.0;
		return;
	}
	if (s.E >= 0.0) {                      // :823-829 (remove_star :662-664)
		out.status = ORBIT_UNBOUND; out.rnew = CMCB_SF_INFINITY; out.vrnew = 0.0; out.vtnew = 0.0;
		return;
	}
	bool need_scan = false;
	OrbitRS o = calc_orbit_new(t, top, s, is_new, rnew_of_new, &out.n_exact, &out.n_brent, ktemp_in, allow_defer ? &need_scan : nullptr); // :832
	if (need_scan) { out.status = ORBIT_DEFERRED; out.rnew = out.vrnew = out.vtnew = 0.0; return; }
	out.kmin = o.kmin; out.kmax = o.kmax;
	if (o.circular) {                      // :835-839
		out.status = ORBIT_CIRCULAR; out.rnew = s.r_j; out.vrnew = 0.0; out.vtnew = 0.0;
		return;
	}
	const double rmin = o.rp, rmax = o.ra;
	if (rmax >= Rtidal) {                  // :848-855
		out.status = ORBIT_TIDAL; out.r_apo = rmax; out.r_peri = rmin;
		out.rnew = CMCB_SF_INFINITY; out.vrnew = 0.0; out.vtnew = 0.0;
		return;
	}
	const double span = dsub(rmax, rmin);
	const double g1 = dsqrt(ddiv(dmul(3.0, span), o.dQdrp));   // :857
	const double g2 = dsqrt(ddiv(dmul(-3.0, span), o.dQdra));  // :858
	job.F = dmul(1.2, (g1 >= g2 ? g1 : g2));                   // :860
	job.rmin = rmin; job.rmax = rmax;
	out.r_peri = rmin; out.r_apo = rmax;                       // what :919-920 will store once a radius is accepted
	out.status = ORBIT_PENDING;
}

// One pass of the rejection loop body (:863-887) with the two uniforms X, u of that try.  1: accepted (r, vr set), 0: rejected,
// -1: the bracket check inside potential() failed (fatal in the reference, cmc_utils.c:144-152).
// the radius a try proposes (:864-867): a function of the orbit's turning points and the uniform X alone
CMCB_HD double orbit_try_radius(double rmin, double rmax, double X)
{
	const double span = dsub(rmax, rmin);
	const double mid = dmul(0.5, dadd(rmin, rmax)), quarter = dmul(0.25, span);
	const double s0 = dsub(dmul(2.0, X), 1.0);
	return dadd(mid, dmul(quarter, dsub(dmul(3.0, s0), dmul(dmul(s0, s0), s0))));
}

// the rest of the try once potential(r) is known (pot_raw; fail: its bracket check failed)
CMCB_HD int orbit_try_finish(const OrbitTables &t, const StarConst &s, const OrbitJob &job, double X, double u, double r, double pot_raw, int fail, double *vr_out)
{
	const double span = dsub(job.rmax, job.rmin);
	const double quarter = dmul(0.25, span), twoE = dmul(2.0, s.E);
	const double s0 = dsub(dmul(2.0, X), 1.0);
	const double g0 = dmul(job.F, u);
	const double pot = dadd(pot_raw, phi_self(r, s.r_j, s.m_j, t.invN));
	const double drds = dmul(quarter, dsub(3.0, dmul(dmul(3.0, s0), s0)));
	const double Q = dsub(dsub(twoE, dmul(2.0, pot)), ddiv(ddiv(s.JJ, r), r));
	const double vr = (Q >= 0.0) ? dsqrt(Q) : 0.0;
	*vr_out = vr;
	if (fail) return -1;
	return g0 < dmul(ddiv(1.0, vr), drds) ? 1 : 0;
}

CMCB_HD int orbit_try(const OrbitTables &t, const LadderTops &top, const StarConst &s, const OrbitJob &job, double X, double u, double *r_out, double *vr_out)
{
	const double r = orbit_try_radius(job.rmin, job.rmax, X);
	int fail = 0;
	const double pot_raw = potential_at(t, top, r, &fail);
	*r_out = r;
	return orbit_try_finish(t, s, job, X, u, r, pot_raw, fail, vr_out);
}

// :918-931 once try number `tries` (1-based) was accepted with (X, r, vr); usign = the next uniform of the stream
CMCB_HD void orbit_finish(const StarConst &s, const OrbitJob &job, double X, double r, double vr, double usign, int tries, OrbitOut &out)
{
	out.X = X; out.r_peri = job.rmin; out.r_apo = job.rmax;  // :918-920
	if (usign < 0.5) vr = -vr;                               // :923-924
	out.n_draws = 2 * tries + 1;
	out.status = ORBIT_OK;
	out.rnew = r; out.vrnew = vr; out.vtnew = ddiv(s.J, r);  // :926-931
}

// One iteration of the loop body of get_positions_loop (cmc_evolution_thr.c:802-933), start to end, on a sequential stream.
// E_star is star[j].E (without the self term), Draws provides `double next()` and `bool has(int n)`.
template <class Draws>
CMCB_HD void orbit_step_star(const OrbitTables &t, kint g_j, bool is_new, double rnew_of_new, double E_star, double J,
			     double vr_in, double vt_in, double Rtidal, Draws &draws, OrbitOut &out,
			     kint ktemp_in = KTEMP_COMPUTE, bool allow_defer = false)
{
	const StarConst s = star_const(t, g_j, E_star, J);
	LadderTops top;
	top.q = t.qlad.n_levels ? t.qlad.blk[0] : nullptr; top.r = t.rlad.n_levels ? t.rlad.lvl[0] : nullptr; top.qdesc = &t.qlad;
	top.s = t.slad.n_levels ? t.slad.lvl[0] : nullptr; top.sdesc = &t.slad;
	top.sf = t.slad.n_levels ? t.slad.topf : nullptr;
	OrbitJob job;
	orbit_prepare(t, top, s, is_new, rnew_of_new, Rtidal, ktemp_in, allow_defer, out, job);
	if (out.status == ORBIT_CIRCULAR) { out.vrnew = vr_in; out.vtnew = vt_in; }
	if (out.status != ORBIT_PENDING) return;
	double r = 0.0, vr = 0.0, X = 0.0;
	int k, rc = 0;
	for (k = 1; k <= CMCB_N_TRY; k++) {    // :862-888
		if (!draws.has(2)) { rc = -1; break; }
		X = draws.next();
		const double u = draws.next();
		rc = orbit_try(t, top, s, job, X, u, &r, &vr);
		if (rc != 0) break;
	}
	if (k == CMCB_N_TRY + 1 || rc < 0 || !draws.has(1)) { // :890-893
		out.status = ORBIT_NTRY; out.rnew = s.r_j; out.vrnew = vr_in; out.vtnew = vt_in;
		out.X = out.r_peri = out.r_apo = 0.0;
		out.n_draws = 2 * (k > CMCB_N_TRY ? CMCB_N_TRY : k);
		return;
	}
	orbit_finish(s, job, X, r, vr, draws.next(), k, out);
}

} // namespace cmcb
