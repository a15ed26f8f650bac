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

// returns false if n has a prime factor > DBSPM_MAX_RADIX or needs too many stages
// max_pow2_radix (16, 8, 4 or 2) caps the power-of-two radices: smaller radices need fewer registers
bool dbspm_make_plan(int n, HostPlan &out, int max_pow2_radix = 16);
