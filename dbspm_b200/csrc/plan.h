// plan.h -- host-side FFT plan: radix schedule, twiddle table, frequency->position table.
#pragma once
#include <vector>
#include "fft_core.h"

struct HostPlan {
    FftPlanDev dev;
    std::vector<cplx> tw;       // tw[e] = exp(-2 pi i e / n)
    std::vector<int> pos_of;    // pos_of[K] = in-place position holding frequency K after the DIF stages
    std::vector<int> freq_of;   // inverse table: freq_of[pos_of[K]] = K
};

// returns false if n has a prime factor > DBSPM_MAX_RADIX or needs too many stages.  The schedule has the
// fewest stages possible with radices {2..10,12,14,15,16} (and generic primes 11..31), ties broken by
// butterfly cost; max_pow2_radix caps the fast radices (smaller radices need fewer registers)
bool dbspm_make_plan(int n, HostPlan &out, int max_pow2_radix = 16);
