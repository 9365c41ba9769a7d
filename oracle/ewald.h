/* ewald.h -- smooth particle-mesh Ewald reciprocal sum as the reference holds it in module pbc_mod
 * (set_periodic.f90:114-231) and evaluates it in ewald_recip.f90.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_EWALD_H
#define ORACLE_EWALD_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_ewald {
    double box[3], volbox, a_ewald;
    int nfft, bsorder;
    double *bsmod1, *bsmod2, *bsmod3; /* [nfft] */
} orc_ewald;
orc_ewald *orc_ewald_setup(const double box[3]);
void orc_ewald_free(orc_ewald *E);
int orc_ewald_nfft(const orc_ewald *E);
double orc_ewald_alpha(const orc_ewald *E);
void orc_ewald_bsmod(const orc_ewald *E, double *out3n);
void orc_ewald_recip(const orc_ewald *E, int n, const double *xyz, const double *q, double *energy, double *grad);
/* reference sum: reciprocal-space part of the plain Ewald sum, |m_d| <= mmax (validation only) */
void orc_ewald_direct_recip(const double box[3], double alpha, int mmax, int n, const double *xyz, const double *q,
                            double *energy, double *grad);
#ifdef __cplusplus
}
#endif
#endif
