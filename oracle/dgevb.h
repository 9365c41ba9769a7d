/* dgevb.h -- DG-EVB parameters as module evb_mod holds them (read_pes.f90:2046-2225).
 * TEST INFRASTRUCTURE ONLY (see oracle.h). */
#ifndef ORACLE_DGEVB_H
#define ORACLE_DGEVB_H
#include "qmdff.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_dgevb {
    int mode, npoints, nat6, natoms;
    const int *coord_def;    /* [nat6][5]: type (1 dist, 2 angle, 3 dihedral, 4 oop), atoms 1-based */
    const double *point_int; /* point_int(nat6, npoints) Fortran == C [point][coordinate] */
    const double *alph;      /* alph_opt(npoints) */
    const double *b_vec;     /* b_vec(mat_size) */
    double g_thres;
} orc_dgevb;
void orc_qmdff_two_one(const orc_qmdff *f2, const double *xyz, double *e, double *g);
void orc_xyz_2int(const orc_dgevb *d, const double *xyz, double *internal);
void orc_dgevb_egrad(const orc_qmdff *f1, const orc_qmdff *f2, const orc_dgevb *d, const double *xyz, int nimg,
                     double *V, double *g);
#ifdef __cplusplus
}
#endif
#endif
