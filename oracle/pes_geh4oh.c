/*
 * pes_geh4oh.c -- CPU oracle: GeH4 + OH -> GeH3 + H2O surface (Espinosa-Garcia, Corchado et al.; POTLIB form),
 * /root/reference/src/egrad_geh4oh.f.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity UNPINNED by the reference (no golden vectors, cannot be
 * compiled here); pinned by finite differences and the properties in tests/test_oracle_ch4oh.py.
 *
 * egrad_geh4oh.f is egrad_ch4oh.f with the "C" atom a germanium: its own BLOCK DATA (:2002-2042), the in-plane
 * reference angle on taugeh = 0.678 pi (:368), sphi without the 3.8 A cut (:1793-1804).  Its driver (:36-67) copies
 * bead 1 only; every call site passes one bead (gradient.f90:193), and the oracle loops over the images it is given.
 *   egrad_geh4oh :36-67, pot_geh4oh :84-217, coorden :219-303, refangles :305-448, stretch :450-713,
 *   opbend :715-869, ipbend :871-1108, calcdelta :1110-1366, opforce :1368-1480, ipforce :1482-1679,
 *   switchf :1681-1821, initialize (PREPOT) :1823-1940, BLOCK DATA :1942-2043.
 * Atom order H, Ge, H, H, H, O, H(O).
 */
#define CBE_CH4OH 1
#define CBE_GEH4OH 1
#include "pes_ch4h.c"
