/* qmdff.h -- tables of one QMDFF as the reference holds them in module qmdff (qmdff.f90:49-110) and
 * pbc_mod; TEST INFRASTRUCTURE ONLY (see oracle.h). */
#ifndef ORACLE_QMDFF_H
#define ORACLE_QMDFF_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_qmdff {
    int n;
    const int *at;        /* atomic numbers */
    const double *q;      /* charges */
    int nbond, nangl, ntors, nnci, ldvt, nmols;
    const int *bond;      /* (2,nbond) 1-based */
    const double *vbond;  /* (3,nbond) r0, k, a */
    const int *angl;      /* (3,nangl): centre first */
    const double *vangl;  /* (2,nangl) theta0, k */
    const int *tors;      /* (6,ntors): i,j,k,l,nt,type */
    const double *vtors;  /* (ldvt,ntors): phi0, k, nt x (n, phase, V) */
    const int *nci;       /* (3,nnci): i,j,class */
    const int *molnum;    /* (n) */
    const double *c6xy;   /* (n,n) Fortran order */
    const double *r0ab, *zab, *r094, *sr42; /* (94,94) Fortran order */
    const double *rad;    /* (94) */
    double eps1[6], eps2[6];
    int periodic, zahn;
    double box[3], coul_cut, vdw_cut, cut_low, zahn_a, zahn_par;
    double e_zero;
    /* H/X-bond terms (ff_hb.f90) */
    int nhb;
    const int *hb;            /* (3,nhb): A, B, H */
    const double *vhb;        /* (2,nhb) */
    const double *scalehb, *scalexb; /* scalehb_glob(94), scalexb_glob(94) */
    const double *q_glob;     /* (n) */
} orc_qmdff;
void orc_ff_eg(const orc_qmdff *f, const double *xyz, double *e, double *g);
void orc_ff_nonb(const orc_qmdff *f, const double *xyz, double *e_io, double *g);
void orc_ff_hb(const orc_qmdff *f, double *xyz, double *e_io, double *g);
void orc_qmdff_egrad(const orc_qmdff *f, const double *xyz, int nimg, double *V, double *g);
#ifdef __cplusplus
}
#endif
#endif
