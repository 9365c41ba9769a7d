/*
 * pes_ch4cn.c -- CPU oracle: CH4 + CN -> CH3 + HCN surface of Espinosa-Garcia, Rangel and Suleimanov
 * (Phys. Chem. Chem. Phys. 19, 19341 (2017); POTLIB form), /root/reference/src/egrad_ch4cn.f.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity UNPINNED by the reference (no golden vectors, cannot be
 * compiled here); pinned by finite differences and the properties in tests/test_oracle_ch4oh.py.
 *
 * egrad_ch4cn.f is egrad_ch4oh.f line for line (the same coorden / refangles / stretch / opbend / ipbend / calcdelta /
 * opforce / ipforce / switchf / PREPOT routines; whitespace and comments aside the two files differ in 96 lines) with the
 * abstracting atom "b" the carbon of the CN radical and the seventh atom its nitrogen: its own BLOCK DATA (:2074-2114:
 * d3ch, a1ch, b1ch, c1ch, r0hh, d1hh, d3hh, ahh, r0cb, d3cbi, acb, aa1, fkh2oeq, alph2o, anh2oeq differ), the H-C-N
 * "bend" reference angle 180 degrees, and the C-N Morse bond on literal constants r0 = 1.172 A, a = 0.80 A^-1,
 * D = 80.0 (:625-627, :659-660; in the routine's 1e5 J/mol, i.e. not passed through PREPOT's kcal/mol scaling).
 *   egrad_ch4cn :73-130, POT_ch4cn :161-293, stretch :524-787, ipbend :945-1182, BLOCK DATA :2014-2117.
 * Atom order H, C, H, H, H, C(N), N.
 */
#define CBE_CH4OH 1
#define CBE_CH4CN 1
#include "pes_ch4h.c"
