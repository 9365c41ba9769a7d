// count_qmdff.cpp -- operation-counting build of the QMDFF / DG-EVB / water oracle: the same
// sources as liboracle.so with `double` replaced by the counting scalar of count_real.hpp (one 8-byte member, so the
// C API keeps its layout and oracle.py can drive it with the tables it already builds).  TEST INFRASTRUCTURE.
//   g++ -O1 -shared -fPIC -o liboracle_count.so count_qmdff.cpp
#include <cstdlib>
#include <cstring>
#include <cmath>
#include "count_real.hpp"
cnt_counters g_cnt = {0, 0, 0, 0, 0, 0};
#define CNT_LIBM1X(f) inline cnt_real f(const cnt_real& a) { g_cnt.libm++; return cnt_real(std::f(a.v)); }
CNT_LIBM1X(erf)
CNT_LIBM1X(erfc)
CNT_LIBM1X(asin)
CNT_LIBM1X(atan)
inline cnt_real floor(const cnt_real& a) { return cnt_real(std::floor(a.v)); }
inline cnt_real atan2(const cnt_real& a, const cnt_real& b) { g_cnt.libm++; return cnt_real(std::atan2(a.v, b.v)); }
#define CNT_ASSIGN(op, field)                                                                         \
    inline cnt_real& operator op##=(cnt_real& a, const cnt_real& b) { g_cnt.field++; a.v op##= b.v; return a; } \
    inline cnt_real& operator op##=(cnt_real& a, double b) { g_cnt.field++; a.v op##= b; return a; }
CNT_ASSIGN(+, add)
CNT_ASSIGN(-, add)
CNT_ASSIGN(*, mul)
CNT_ASSIGN(/, div)
static cnt_counters g_mark = {0, 0, 0, 0, 0, 0};
#define ORC_MARK() (g_mark = g_cnt)
#define ORC_REJECT() (g_cnt = g_mark)
#define double cnt_real
#include "qmdff.c"
#include "dgevb.c"
#undef double
extern "C" void oracle_count_reset(void) { g_cnt = cnt_counters{0, 0, 0, 0, 0, 0}; }
extern "C" void oracle_count_get(unsigned long long out[6])
{
    out[0] = g_cnt.add; out[1] = g_cnt.mul; out[2] = g_cnt.div; out[3] = g_cnt.sqrt_; out[4] = g_cnt.libm; out[5] = g_cnt.cmp;
}
