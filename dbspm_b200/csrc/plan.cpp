// plan.cpp -- see plan.h.  Pure host C++ (compiled by nvcc into the library and by g++
// into the CPU emulation used by the tests).
#include "plan.h"
#include <math.h>

bool dbspm_make_plan(int n, HostPlan &out, int max_pow2_radix)
{
    if (n < 1) return false;
    FftPlanDev &d = out.dev;
    d.n = n;
    d.nstages = 0;
    int rem = n;
    auto push = [&](int r) -> bool {
        if (d.nstages >= DBSPM_MAX_STAGES) return false;
        d.radix[d.nstages++] = r;
        return true;
    };
    // odd prime factors first (large strides, twiddle-heavy stages early), then the 2-power
    for (int p = 3; p <= DBSPM_MAX_RADIX; p += 2)
        while (rem % p == 0) {
            if (!push(p)) return false;
            rem /= p;
        }
    int twos = 0;
    while (rem % 2 == 0) { rem /= 2; ++twos; }
    if (rem != 1) return false;                    // prime factor > DBSPM_MAX_RADIX
    // 2^twos over ceil(twos/maxbits) stages of radix <= max_pow2_radix, bits spread evenly, larger radices first
    if (twos > 0) {
        const int maxbits = max_pow2_radix >= 16 ? 4 : (max_pow2_radix >= 8 ? 3 : (max_pow2_radix >= 4 ? 2 : 1));
        const int ns = (twos + maxbits - 1) / maxbits;
        for (int i = 0; i < ns; ++i) {
            const int bits = twos / ns + (i < twos % ns ? 1 : 0);
            if (!push(1 << bits)) return false;
        }
    }
    if (n == 1) { d.nstages = 0; }

    out.tw.resize(n);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int e = 0; e < n; ++e) {
        long double a = two_pi * (long double)e / (long double)n;
        out.tw[e].x = (double)cosl(a);
        out.tw[e].y = (double)(-sinl(a));
    }
    out.pos_of.assign(n, 0);
    out.freq_of.assign(n, 0);
    for (int p = 0; p < n; ++p) {
        int r0 = p, sub = n, K = 0, mult = 1;
        for (int st = 0; st < d.nstages; ++st) {
            sub /= d.radix[st];
            int k = r0 / sub;
            r0 -= k * sub;
            K += k * mult;
            mult *= d.radix[st];
        }
        out.pos_of[K] = p;
        out.freq_of[p] = K;
    }
    return true;
}
