/*
 * pes_ch4oh.c -- CPU oracle: CH4 + OH -> CH3 + H2O surface of Espinosa-Garcia and Corchado
 * (J. Chem. Phys. 112, 5731 (2000); POTLIB form), /root/reference/src/egrad_ch4oh.f.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity UNPINNED by the reference (no golden
 * vectors, cannot be compiled here); pinned by finite differences and the properties in
 * tests/test_oracle_ch4oh.py.
 *
 * egrad_ch4oh.f is the CH4 + H template of egrad_ch4h.f (the same coorden / refangles / stretch /
 * opbend / ipbend / calcdelta / opforce / ipforce / switchf routines, the abstracting atom an oxygen)
 * with its own constants and three added terms; the literal restatement of both lives in
 * pes_ch4h.c, whose CBE_CH4OH blocks cite the egrad_ch4oh.f lines they follow.
 *   egrad_ch4oh :69-124, POT_ch4oh :157-286, coorden :288-372, refangles :374-515, stretch :517-780,
 *   opbend :782-936, ipbend :938-1175, calcdelta :1177-1433, opforce :1435-1547, ipforce :1549-1746,
 *   switchf :1748-1886 (constants :1808-1811), PREPOT :1888-2005, BLOCK DATA :2007-2109.
 * Atom order H, C, H, H, H, O, H(O) (nnc=2, nnb=6, nnh=3,4,5,1, nno=7; :2066-2069).
 */
#define CBE_CH4OH 1
#include "pes_ch4h.c"
