// plan.cpp -- see plan.h.  Pure host C++ (compiled by nvcc into the library and by g++
// into the CPU emulation used by the tests).
#include "plan.h"
#include <math.h>
#include <vector>

// estimated real additions + multiplications per point of each register butterfly (fft_core.h); only
// used to choose between factorisations with the same number of stages
static double radix_cost(int r)
{
    switch (r) {
    case 2: return 2.0;   case 3: return 5.3;   case 4: return 4.0;   case 5: return 9.6;
    case 6: return 7.3;   case 7: return 13.7;  case 8: return 6.5;   case 9: return 13.3;
    case 10: return 11.6; case 12: return 9.3;  case 14: return 19.0; case 15: return 14.9;   // 14: measured (n = 112: [7,16] beats [14,8])
    case 16: return 11.4;
    default: return 2.0 * r;                       // generic O(r^2) stage
    }
}

// Fewest stages first (every stage is a shared-memory round trip and a block barrier), then the
// cheapest butterflies.  Radices: the fast set of fft_core.h capped at max_radix, plus the primes
// 11..31 (generic stage) when n needs them.  Non-increasing sequences only; the order is fixed below.
static void search_split(int rem, int max_r, int cap, std::vector<int> &cur, double cost,
                         std::vector<int> &best, double &best_cost)
{
    if (rem == 1) {
        if (best.empty() || cur.size() < best.size() || (cur.size() == best.size() && cost < best_cost - 1e-9)) {
            best = cur; best_cost = cost;
        }
        return;
    }
    if (!best.empty() && cur.size() + 1 > best.size()) return;
    for (int r = max_r; r >= 2; --r) {
        if (rem % r) continue;
        const bool prime_big = r > 16 || r == 11 || r == 13;
        if (!prime_big && (!fast_radix(r) || r > cap)) continue;
        if (prime_big) { bool pr = true; for (int d = 2; d * d <= r; ++d) if (r % d == 0) pr = false; if (!pr) continue; }
        cur.push_back(r);
        search_split(rem / r, r, cap, cur, cost + radix_cost(r), best, best_cost);
        cur.pop_back();
    }
}

bool dbspm_make_plan(int n, HostPlan &out, int max_pow2_radix)
{
    if (n < 1) return false;
    FftPlanDev &d = out.dev;
    d.n = n;
    d.nstages = 0;
    if (n > 1) {
        std::vector<int> cur, best;
        double best_cost = 0.0;
        search_split(n, DBSPM_MAX_RADIX - 1, max_pow2_radix < 2 ? 2 : max_pow2_radix, cur, 0.0, best, best_cost);
        if (best.empty() || (int)best.size() > DBSPM_MAX_STAGES) return false;   // prime factor > 31 or too many stages
        // stage order: radices that are not powers of two first (large strides, twiddle-heavy stages early),
        // then the powers of two, each group in descending order
        auto pow2 = [](int r) { return (r & (r - 1)) == 0; };
        for (int pass = 0; pass < 2; ++pass)
            for (int r : best)
                if ((pass == 1) == pow2(r)) d.radix[d.nstages++] = r;
    }

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
