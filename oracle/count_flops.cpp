// count_flops.cpp -- dynamic operation census of the reference surfaces:
// compiles the oracle's PES sources with `real` = counting type and evaluates them at
// TS-neighbourhood geometries.  Output: one JSON object on stdout.  TEST INFRASTRUCTURE.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "count_real.hpp"
cnt_counters g_cnt = {0, 0, 0, 0, 0, 0};
#include "pes_h3.c"
#include "pes_oh3.c"
#include "pes_ch4h.c"
#include "pes_brh2.c"
#include "pes_o3.c"
// the 7-atom members of the CBE family are the CH4 + H source compiled again (pes_ch4oh.c, pes_geh4oh.c)
// (ipow takes the counting type, which lives in the global namespace: renamed so that argument-dependent lookup
// does not see the copy of the first inclusion)
namespace cbe_ch4oh {
#define ipow ipow_ch4oh
#include "pes_ch4oh.c"
#undef ipow
}
namespace cbe_geh4oh {
#define ipow ipow_geh4oh
#include "pes_geh4oh.c"
#undef ipow
}

namespace cbe_ch4cn {
#define ipow ipow_ch4cn
#include "pes_ch4cn.c"
#undef ipow
}

// the three-hydrogen members (pes_clnh3.c, and pes_nh3oh.c = the same source with CBE3_NH3OH; for nh3oh the census
// counts the 19 energy evaluations of the difference quotient -- the analytic derivative code the reference still
// runs 19 times and discards is not restated, so its operations are not in the figure)
namespace cbe_clnh3 {
#include "pes_clnh3.c"
}
namespace cbe_nh3oh {
#include "pes_nh3oh.c"
}

#include "pes_h2co.c"

typedef void (*egrad_fn)(const real*, int, int, real*, real*, int*);

static void census(const char* name, egrad_fn fn, int nat, const double* ts, int n, bool last)
{
    std::mt19937_64 gen(20261017);
    std::normal_distribution<double> nd(0.0, 0.15);
    cnt_counters z = {0, 0, 0, 0, 0, 0};
    g_cnt = z;
    for (int s = 0; s < n; s++) {
        real q[21], V[1], g[21];
        int info;
        for (int i = 0; i < 3 * nat; i++) q[i] = cnt_real(ts[i] + nd(gen));
        fn(q, nat, 1, V, g, &info);
    }
    printf("  \"%s\": {\"samples\": %d, \"add\": %.1f, \"mul\": %.1f, \"div\": %.1f, \"sqrt\": %.1f, \"libm\": %.1f, "
           "\"arith\": %.1f, \"flops\": %.1f}%s\n",
           name, n, (double)g_cnt.add / n, (double)g_cnt.mul / n, (double)g_cnt.div / n, (double)g_cnt.sqrt_ / n,
           (double)g_cnt.libm / n, (double)g_cnt.arith() / n, (double)g_cnt.total() / n, last ? "" : ",");
}

int main()
{
    const double b = 0.52917721092;
    const double h3[9] = {0, 0, -0.929764359586 / b, 0, 0, 0, 0, 0, 0.929764359586 / b};
    const double oh3[12] = {0, 0, 0, 0.97 * -0.2272020946930871 / b, 0.97 * 0.9738476308781951 / b, 0,
                            1.35 / b, 0, 0, 2.11 / b, 0, 0};
    const double s3 = 0.5773502691896258;
    const double ch5[18] = {1.39 * s3 / b, 1.39 * s3 / b, 1.39 * s3 / b, 0, 0, 0,
                            1.09 * s3 / b, -1.09 * s3 / b, -1.09 * s3 / b, -1.09 * s3 / b, 1.09 * s3 / b, -1.09 * s3 / b,
                            -1.09 * s3 / b, -1.09 * s3 / b, 1.09 * s3 / b,
                            2.263 * s3 / b, 2.263 * s3 / b, 2.263 * s3 / b};
    printf("{\n");
    census("h3", oracle_egrad_h3_real, 3, h3, 2000, false);
    census("oh3", oracle_egrad_oh3_real, 4, oh3, 2000, false);
    const double brh2[9] = {0, 0, 0, 0, 0, -2.72158888, 0, 0, 2.64056088};
    census("ch4h", oracle_egrad_ch4h_real, 6, ch5, 2000, false);
    const double o3[9] = {0, 0, 0, 1.60 / b, 0, 0, -0.4017 / b, 1.2364 / b, 0};
    census("brh2", oracle_egrad_brh2_real, 3, brh2, 2000, false);
    census("o3", oracle_egrad_o3_real, 3, o3, 2000, false);
    // examples/explore/ts_irc_ch4oh/ts_start.xyz; tests/common.py geh4oh_ts
    const double ch4oh[21] = {-4.62878267 / b, 1.25606861 / b, 0.95459788 / b, -4.85261637 / b, 2.15380812 / b, 0.37457524 / b,
                              -4.27740626 / b, 2.99438311 / b, 0.76831501 / b, -4.53003946 / b, 1.95346377 / b, -0.88649386 / b,
                              -5.91912714 / b, 2.37708958 / b, 0.44643735 / b, -4.21407574 / b, 1.75722671 / b, -2.12170961 / b,
                              -3.93964920 / b, 2.62885961 / b, -2.44704966 / b};
    census("ch4oh", cbe_ch4oh::oracle_egrad_ch4oh_real, 7, ch4oh, 2000, false);
    const double t0 = 1.62 * s3 / b, t1 = 1.525 * s3 / b, to = 2.97 * s3 / b;
    const double geh4oh[21] = {t0, t0, t0, 0, 0, 0, t1, -t1, -t1, -t1, t1, -t1, -t1, -t1, t1, to, to, to,
                               to + 0.97 * 0.6650 / b, to - 0.97 * 0.6820 / b, to - 0.97 * 0.3040 / b};
    census("geh4oh", cbe_geh4oh::oracle_egrad_geh4oh_real, 7, geh4oh, 2000, false);
    // caracal_b200/systems.py ch4cn_ts
    const double u0 = 1.20 * s3 / b, u1 = 1.094 * s3 / b, uc = 2.75 * s3 / b;
    const double ch4cn[21] = {u0, u0, u0, 0, 0, 0, u1, -u1, -u1, -u1, u1, -u1, -u1, -u1, u1, uc, uc, uc,
                              uc + 1.172 * 0.6385 / b, uc + 1.172 * 0.5384 / b, uc + 1.172 * 0.5384 / b};
    census("ch4cn", cbe_ch4cn::oracle_egrad_ch4cn_real, 7, ch4cn, 2000, false);
    // caracal_b200/systems.py clnh3_ts, nh3oh_ts
    const double clnh3[15] = {2.3329132995, 0.0, 0.9350786638, 0.0, 0.0, 0.0, -0.8893135660, 1.5403362803, 0.7129095978,
                              -0.8893135660, -1.5403362803, 0.7129095978, 4.7886115095, 0.0, 1.9193719941};
    census("clnh3", cbe_clnh3::oracle_egrad_clnh3_real, 5, clnh3, 2000, false);
    const double nh3oh[18] = {2.0171806725, 0.0, 0.8085266642, 0.0, 0.0, 0.0, -0.8893135660, 1.5403362803, 0.7129095978,
                              -0.8893135660, -1.5403362803, 0.7129095978, 4.2974718675, 0.0, 1.7225133281,
                              3.9213112830, 0.0, 3.5165362123};
    census("nh3oh", cbe_nh3oh::oracle_egrad_nh3oh_real, 6, nh3oh, 2000, false);
    // caracal_b200/systems.py h2co_ts
    const double h2co[12] = {0.0, 0.0, 0.0, 0.0, 0.0, 1.17 / b, 1.05 / b, 0.0, -0.33 / b,
                             (1.05 + 1.25 * 0.384615) / b, 0.0, (-0.33 - 1.25 * 0.923077) / b};
    census("h2co", oracle_egrad_h2co_real, 4, h2co, 50, true);
    printf("}\n");
    return 0;
}
