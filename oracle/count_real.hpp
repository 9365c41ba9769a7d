// count_real.hpp -- operation-counting scalar for the oracle's flop census (TEST INFRASTRUCTURE).
// Counting convention of BASELINE.md section 4: add/sub/mul/div/sqrt = 1, every libm call = 1
// (FMA never appears: the oracle is built without contraction, so a*b+c counts 2).
#pragma once
#include <cmath>

struct cnt_counters {
    unsigned long long add, mul, div, sqrt_, libm, cmp;
    unsigned long long arith() const { return add + mul + div + sqrt_; }
    unsigned long long total() const { return add + mul + div + sqrt_ + libm; }
};
extern cnt_counters g_cnt;

struct cnt_real {
    double v;
    cnt_real() : v(0.0) {}
    cnt_real(double x) : v(x) {}
    cnt_real(float x) : v(x) {}
    cnt_real(int x) : v(x) {}
    cnt_real& operator=(double x) { v = x; return *this; }
};
#define CNT_BIN(op, field)                                                                        \
    inline cnt_real operator op(const cnt_real& a, const cnt_real& b) { g_cnt.field++; return cnt_real(a.v op b.v); } \
    inline cnt_real operator op(const cnt_real& a, double b) { g_cnt.field++; return cnt_real(a.v op b); }           \
    inline cnt_real operator op(double a, const cnt_real& b) { g_cnt.field++; return cnt_real(a op b.v); }
CNT_BIN(+, add)
CNT_BIN(-, add)
CNT_BIN(*, mul)
CNT_BIN(/, div)
inline cnt_real operator-(const cnt_real& a) { return cnt_real(-a.v); }
#define CNT_CMP(op)                                                                        \
    inline bool operator op(const cnt_real& a, const cnt_real& b) { g_cnt.cmp++; return a.v op b.v; } \
    inline bool operator op(const cnt_real& a, double b) { g_cnt.cmp++; return a.v op b; }           \
    inline bool operator op(double a, const cnt_real& b) { g_cnt.cmp++; return a op b.v; }
CNT_CMP(<)
CNT_CMP(>)
CNT_CMP(<=)
CNT_CMP(>=)
CNT_CMP(==)
CNT_CMP(!=)
inline cnt_real sqrt(const cnt_real& a) { g_cnt.sqrt_++; return cnt_real(std::sqrt(a.v)); }
#define CNT_LIBM1(f) inline cnt_real f(const cnt_real& a) { g_cnt.libm++; return cnt_real(std::f(a.v)); }
CNT_LIBM1(exp)
CNT_LIBM1(log)
CNT_LIBM1(tanh)
CNT_LIBM1(cosh)
CNT_LIBM1(acos)
CNT_LIBM1(sin)
CNT_LIBM1(cos)
inline cnt_real fabs(const cnt_real& a) { return cnt_real(std::fabs(a.v)); }
inline cnt_real pow(const cnt_real& a, const cnt_real& b) { g_cnt.libm++; return cnt_real(std::pow(a.v, b.v)); }
inline cnt_real pow(const cnt_real& a, double b) { g_cnt.libm++; return cnt_real(std::pow(a.v, b)); }
