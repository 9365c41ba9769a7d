/*
 * water.h -- CPU oracle of the flexible SPC water model (pes WATER_SPC).  TEST INFRASTRUCTURE ONLY
 * (see oracle_real.h): only tests/, smoke() and bench.py's CPU legs may load it.
 */
#ifndef ORACLE_WATER_H
#define ORACLE_WATER_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int natoms;          /* 3 * nwater, atoms ordered O,H,H per molecule (water_init.f90:58-69) */
    int periodic, zahn;  /* pbc_mod switches (set_periodic.f90:66-68) */
    double box[3];       /* boxlen_x/y/z in bohr */
    double coul_cut, zahn_a, zahn_par;
    double pars[11];     /* water_pars(1:11) in atomic units (water_init.f90:75-101) */
    const double *q;     /* [natoms] charges (water_init.f90:103-107) */
    const int *is_O;     /* [natoms] name(i) == "O" */
} orc_water;

/* water_init.f90:75-101: the parameter set in atomic units */
void orc_water_default_pars(double pars[11]);
/* egrad_water.f90:36-333 for nimg structures xyz[nimg][natoms][3]; V[nimg], g like xyz */
void orc_water_egrad(const orc_water *w, const double *xyz, int nimg, double *V, double *g);

#ifdef __cplusplus
}
#endif
#endif
