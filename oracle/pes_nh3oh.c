/*
 * pes_nh3oh.c -- CPU oracle: NH3 + OH -> NH2 + H2O surface (Monge-Palacios, Rangel, Espinosa-Garcia, J. Chem. Phys. 138,
 * 084305 (2013); POTLIB form), /root/reference/src/egrad_nh3oh.f.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity UNPINNED by the reference (no golden vectors, cannot be
 * compiled here); pinned by the properties and known answers of tests/test_oracle_clnh3.py.
 *
 * egrad_nh3oh.f is egrad_clnh3.f with the abstracting atom "b" the oxygen of OH and a sixth atom its hydrogen; the
 * additions and the numeric gradient (POT_nh3oh :283-296) are listed in the header of pes_clnh3.c and marked CBE3_NH3OH
 * there.  Atom order H, N, H, H, O, H(O).
 */
#define CBE3_NH3OH 1
#include "pes_clnh3.c"
