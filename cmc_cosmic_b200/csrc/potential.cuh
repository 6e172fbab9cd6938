#pragma once
#include <cuda_runtime.h>
#include "seqscan.cuh"

namespace cmcb {

struct PotScalars { // device-resident scalars of one potential rebuild
	long n_max;        // new clus.N_MAX
	long first_inf;    // first k with r >= SF_INFINITY
	double cenma_m, cenma_m_new;
	double mass_sum;   // sequential sum of star_m[1..N_MAX]
	double Mtotal;
	double mprev_end;  // mprev after the last subtraction (unused by the reference; kept for tests)
	double phi_sum;    // -phi_1
};

struct PotArgs {
	double *r;         // star_r (device), [N_MAX+1] is written
	const double *m;   // star_m
	double *phi;       // star_phi (out)
	double *mprev;     // scratch, n_dim doubles
	PotScalars *sc;
	SeqWork work;
	long n_max_in;     // clus.N_MAX on entry
	double n_star;     // (double) clus.N_STAR
	double madhoc;
	double cenma_m, cenma_m_new;
};

size_t pot_work_bytes(long n_dim);
void pot_work_bind(SeqWork &w, void *base, long n_dim);
int potential_calculate_device(const PotArgs &a, cudaStream_t st);

} // namespace cmcb
