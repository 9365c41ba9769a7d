// pes_host.cu -- TEST INFRASTRUCTURE: compiles the product's __host__ __device__ PES
// functors for the CPU so that the derivations can be checked against the oracle in the
// GPU-less container (tests -m "not gpu").  Never loaded by caracal_b200 itself.
#include "../../caracal_b200/csrc/pes_h3.cuh"
#include "../../caracal_b200/csrc/pes_oh3.cuh"
#include "../../caracal_b200/csrc/pes_ch4h.cuh"
#include "../../caracal_b200/csrc/pes_brh2.cuh"
#include "../../caracal_b200/csrc/pes_o3.cuh"
#include "../../caracal_b200/csrc/pes_ch4oh.cuh"
#include "../../caracal_b200/csrc/pes_nh3x.cuh"
#include "../../caracal_b200/csrc/pes_h2co.cuh"

template <class PES>
static int run(const double* q, int nimg, double* V, double* g)
{
    int info = 0;
    for (int i = 0; i < nimg; i++)
        info |= PES::eval(q + (size_t)i * 3 * PES::NATOMS, V[i], g + (size_t)i * 3 * PES::NATOMS);
    return info;
}

extern "C" int hh_egrad(int pes, const double* q, int nimg, double* V, double* g)
{
    switch (pes) {
    case CRCL_PES_H3: return run<crcl::PesH3>(q, nimg, V, g);
    case CRCL_PES_OH3: return run<crcl::PesOH3>(q, nimg, V, g);
    case CRCL_PES_CH4H: return run<crcl::PesCH4H>(q, nimg, V, g);
    case CRCL_PES_BRH2: return run<crcl::PesBrH2>(q, nimg, V, g);
    case CRCL_PES_O3: return run<crcl::PesO3>(q, nimg, V, g);
    case CRCL_PES_CH4OH: return run<crcl::PesCH4OH>(q, nimg, V, g);
    case CRCL_PES_GEH4OH: return run<crcl::PesGeH4OH>(q, nimg, V, g);
    case CRCL_PES_CH4CN: return run<crcl::PesCH4CN>(q, nimg, V, g);
    case CRCL_PES_CLNH3: return run<crcl::PesClNH3>(q, nimg, V, g);
    case CRCL_PES_NH3OH: return run<crcl::PesNH3OH>(q, nimg, V, g);
    case CRCL_PES_H2CO: return run<crcl::PesH2CO>(q, nimg, V, g);
    }
    return -1;
}

#include "../../caracal_b200/csrc/xi.cuh"
#include "../../caracal_b200/csrc/rng.cuh"

template <int NAT>
static void xi_run(const crcl::Mech& M, const double* mass, const double* x, double xi_ideal, int mode,
                   double* xi, double* dxi, double* hams, double beta)
{
    crcl::calc_xi<NAT>(M, mass, x, xi_ideal, mode, *xi, dxi, hams, beta);
}

extern "C" int hh_calc_xi(int natoms, const double* mass, int form_num, const int* bond_form, int break_num,
                          const int* bond_break, const double* form_ref, const double* break_ref, int sum_reacs,
                          const int* n_reac, const int* at_reac, double R_inf, const double* x, double xi_ideal,
                          int mode, double beta, double* xi, double* dxi, double* hams)
{
    crcl::Mech M;
    int rc = crcl::build_mech(M, natoms, mass, form_num, bond_form, break_num, bond_break, form_ref, break_ref,
                              sum_reacs, n_reac, at_reac, R_inf);
    if (rc) return rc;
    switch (natoms) {
    case 3: xi_run<3>(M, mass, x, xi_ideal, mode, xi, dxi, hams, beta); break;
    case 4: xi_run<4>(M, mass, x, xi_ideal, mode, xi, dxi, hams, beta); break;
    case 6: xi_run<6>(M, mass, x, xi_ideal, mode, xi, dxi, hams, beta); break;
    default: return -3;
    }
    return 0;
}

extern "C" void hh_normal_pair(unsigned long long seed, unsigned traj, unsigned event, unsigned bead,
                               unsigned pair, double* z)
{
    crcl::normal_pair(seed, traj, event, bead, pair, z[0], z[1]);
}

#ifdef CRCL_FM_ACTIVE
// the branch-free elementary functions of crcl_common.cuh (namespace fm) with the host model of the MUFU seeds
extern "C" void hh_fm(int func, const double* x, const double* y, int n, double* out)
{
    for (int i = 0; i < n; i++) {
        switch (func) {
        case 0: out[i] = crcl::fm::rcp(x[i]); break;
        case 1: out[i] = crcl::fm::div(x[i], y[i]); break;
        case 2: out[i] = crcl::fm::rsqrt(x[i]); break;
        case 3: out[i] = crcl::fm::sqrt<>(x[i]); break;
        case 4: out[i] = crcl::fm::exp(x[i]); break;
        case 5: out[i] = crcl::fm::acos(x[i]); break;
        case 6: { double s, is; crcl::sqrt_rsqrt(x[i], s, is); out[i] = s; } break;
        case 7: out[i] = crcl::fm::log(x[i]); break;
        case 8: out[i] = crcl::fm::pow(x[i], y[i]); break;
        case 9: out[i] = crcl::fm::sqrt<true>(x[i]); break;
        }
    }
}
#endif
