/*
 * oracle_real.h -- scalar type used by the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be linked, imported or
 * executed by the product path (caracal_b200/); only tests/, smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it, as the checker.
 *
 * The oracle sources are written in the common subset of C and C++ so that the
 * same files can be compiled
 *   (a) with gcc as plain C            -> liboracle.so (the parity checker), and
 *   (b) with g++ -DORACLE_COUNTING     -> flop-counting build (oracle/count_flops.cpp)
 * which replaces `real` by a class that counts every arithmetic operation and
 * libm call (BASELINE.md section 4 counting convention).
 *
 * LITERALS: gfortran reads a real literal without a D exponent as REAL*4 and
 * rounds it to single precision before widening.  F(x) reproduces
 * that: default "gfortran" mode pastes an `f` suffix, -DORACLE_LITERALS_EXACT
 * keeps the decimal value in double precision.
 */
#ifndef ORACLE_REAL_H
#define ORACLE_REAL_H

#include <math.h>

#ifdef ORACLE_LITERALS_EXACT
#define F(x) ((double)(x))
#else
#define F(x) ((double)(x##f))
#endif

#if defined(ORACLE_COUNTING) && defined(__cplusplus)
#include "count_real.hpp"
typedef cnt_real real;
#else
typedef double real;
#endif

#ifdef __cplusplus
#define ORACLE_API extern "C"
#else
#define ORACLE_API
#endif

#endif
