// pes_ch4oh.cuh -- CH4 + OH -> CH3 + H2O and GeH4 + OH -> GeH3 + H2O; first: the CH4 + OH surface of Espinosa-Garcia and Corchado (J. Chem. Phys. 112, 5731
// (2000); POTLIB form), one thread per image, FP64.  SURVEY.md 8(f) row N4.
//
// Replaces /root/reference/src/egrad_ch4oh.f: egrad_ch4oh :69-124, POT_ch4oh :157-286 and the routines below it.
// That file is the CH4 + H template of egrad_ch4h.f with the abstracting atom an oxygen, its own BLOCK DATA
// (:2066-2106, scaled once as PREPOT_ch4oh :1989-2002 does), its own switching constants (:1808-1811) and three
// added terms; the evaluation is PesCBE1 (one thread per image, crcl_egrad) and PesCBE4 (four lanes per
// bead, trajectory kernels) of pes_ch4h.cuh with the constants below (K::HAS_OH selects the added terms).  Atom order H, C, H, H, H, O, H(O) (nnc=2, nnb=6, nnh=3,4,5,1, nno=7); the four methane hydrogens are
// equivalent, any of them can be the one abstracted (the shipped examples/explore/ts_irc_ch4oh/ts_start.xyz
// transfers atom 4).
#pragma once
#include "pes_ch4h.cuh"

namespace crcl {
namespace ch4oh {

struct K7 {
    static constexpr int NATOMS = 7, ID = CRCL_PES_CH4OH;
    static constexpr bool HAS_OH = true;
    // fact1 = 0.041840 (kcal/mol -> 1e5 J/mol), fact2 = 6.022045 (mdyn A -> 1e5 J/mol), fact3 = 2 pi / 360
    static constexpr double R0CH = 1.09397, A1CH = 1.78000, B1CH = 0.15000, C1CH = 15.00000;
    static constexpr double R0HH = 0.97060, AHH = 2.15000, R0CB = 1.49092, ACB = 2.98688;
    static constexpr double D1CH = 112.17000 * 0.041840, D3CH = 32.65328 * 0.041840;
    static constexpr double D1HH = 125.44000 * 0.041840, D3HH = 20.41017 * 0.041840;
    static constexpr double D1CB = 91.47526 * 0.041840;
    // d3cb = (d3cbi - a3cb) + a3cb exp(-(4 (rav - rcbsp) / b3cb)^4) (:590-592) with a3cb = 0 (:2104): a constant,
    // and the "derivative of D3cb" terms (:681-699) are products with dd3cb = 0
    static constexpr double A3CB = 0.000000 * 0.041840;
    static constexpr double D3CB = (112.69509 * 0.041840 - A3CB) + A3CB;
    static constexpr double A3S = 0.1419100, B3S = -0.3068400;
    static constexpr double APHI = 0.5287900, BPHI = 0.4006600, CPHI = 1.9209900;
    static constexpr double ATHETA = 0.9078700, BTHETA = 0.3548900, CTHETA = 1.8915500;
    static constexpr double FCH3 = 0.0740000 * 6.022045, HCH3 = 0.1915000 * 6.022045;
    static constexpr double FKINF = 0.4400000 * 6.022045, AK = 0.1260000 * 6.022045;
    static constexpr double AA1 = 0.303746, AA2 = 1.599960, AA3 = 3.216595, AA4 = 11.569980;
    static constexpr double A1S = 1.5313681e-7, B1S = -4.6696246, A2S = 1.0147402e-7, B2S = -12.363798;
    // H-O-H bends: fkh2oeq scaled by fact2, anh2oeq by fact3 (:2001-2002)
    static constexpr double FKH2OEQ = 0.7300000 * 6.022045, ALPH2O = 1.1080000;
    static constexpr double ANH2OEQ = 104.7132000 * (2.0 * 3.141592653589793 / 360.0);
    // Morse bond of the seventh atom: O-H(O) on the H-H... constants r0hh, ahh, d1hh (stretch_ch4oh :616-621)
    static constexpr double MORSE_R0 = R0HH, MORSE_A = AHH, MORSE_D = D1HH;
    static constexpr bool SPHI_ANY_R = false;
    static constexpr double TAU_PLANAR = 0.5;
};
static_assert(K7::A3CB == 0.0, "a switched C-O triplet depth needs the dd3cb terms of stretch_ch4oh :681-699");

}  // namespace ch4oh

using PesCH4OH = PesCBE1<ch4oh::K7>;

// GeH4 + OH -> GeH3 + H2O, /root/reference/src/egrad_geh4oh.f: the CH4 + OH file with the central atom a germanium --
// BLOCK DATA :2002-2042, in-plane reference angle on taugeh = 0.678 pi (:368, :382-440: pyramidal GeH3), sphi
// evaluated at every distance (:1793-1804).  Atom order H, Ge, H, H, H, O, H(O).
namespace geh4oh {
struct K7 : ch4oh::K7 {
    static constexpr int ID = CRCL_PES_GEH4OH;
    static constexpr double R0CH = 1.52500, A1CH = 1.43925, B1CH = 0.12330, C1CH = 2.00400;
    static constexpr double AHH = 2.18200, R0CB = 1.90035, ACB = 0.67621;
    static constexpr double D1CH = 86.50000 * 0.041840, D3CH = 41.50000 * 0.041840;
    static constexpr double D1HH = 120.94800 * 0.041840, D3HH = 31.86417 * 0.041840;
    static constexpr double MORSE_R0 = R0HH, MORSE_A = AHH, MORSE_D = D1HH;
    static constexpr double D1CB = 41.50283 * 0.041840;
    static constexpr double D3CB = (10.50589 * 0.041840 - A3CB) + A3CB;
    static constexpr double A3S = 0.2019100, B3S = -0.6068400;
    static constexpr double CPHI = 11.8809900;
    static constexpr double FCH3 = 0.0150000 * 6.022045;
    static constexpr double FKINF = 0.3060000 * 6.022045;
    static constexpr double AA1 = 0.173746, AA3 = 2.166595;
    static constexpr bool SPHI_ANY_R = true;
    static constexpr double TAU_PLANAR = 0.678;
};
}  // namespace geh4oh

using PesGeH4OH = PesCBE1<geh4oh::K7>;

// CH4 + CN -> CH3 + HCN (Espinosa-Garcia, Rangel, Suleimanov, PCCP 19, 19341 (2017)), /root/reference/src/egrad_ch4cn.f:
// egrad_ch4oh.f line for line with the abstracting atom the carbon of CN and the seventh atom its nitrogen -- its own
// BLOCK DATA (:2074-2114), the H-C-N "bend" on 180 degrees, and the C-N Morse bond on literal constants (:625-627,
// :659-660: r0 = 1.172 A, a = 0.80 / A, D = 80.0 in the routine's 1e5 J/mol, NOT scaled by PREPOT's fact1).
// Atom order H, C, H, H, H, C(N), N.
namespace ch4cn {
struct K7 : ch4oh::K7 {
    static constexpr int ID = CRCL_PES_CH4CN;
    static constexpr double A1CH = 1.75000, B1CH = 0.12000, C1CH = 5.00000;
    static constexpr double R0HH = 1.06497, AHH = 1.70000, R0CB = 1.72592, ACB = 3.08688;
    static constexpr double D3CH = 14.65328 * 0.041840;
    static constexpr double D1HH = 132.17000 * 0.041840, D3HH = 44.63017 * 0.041840;
    static constexpr double D3CB = (88.69509 * 0.041840 - A3CB) + A3CB;
    static constexpr double AA1 = 0.273746;
    static constexpr double FKH2OEQ = 0.2600000 * 6.022045, ALPH2O = 3.1080000;
    static constexpr double ANH2OEQ = 180.0000000 * (2.0 * 3.141592653589793 / 360.0);
    static constexpr double MORSE_R0 = 1.172, MORSE_A = 0.80, MORSE_D = 80.0;
};
}  // namespace ch4cn
using PesCH4CN = PesCBE1<ch4cn::K7>;
#ifdef __CUDACC__
using PesCH4CN4 = PesCBE4<ch4cn::K7>;
#endif
#ifdef __CUDACC__
// four lanes per bead in the trajectory kernels (pes_ch4h.cuh, PesCBE4): lane x owns methane / germane hydrogen x
using PesCH4OH4 = PesCBE4<ch4oh::K7>;
using PesGeH4OH4 = PesCBE4<geh4oh::K7>;
#endif

}  // namespace crcl
