/*
 * oracle_api.c -- double-typed entry points of the CPU oracle for ctypes (tests, smoke,
 * bench cpu_baseline) plus a pthread driver that reproduces the reference's MPI
 * master/worker decomposition of recrossing child pairs (recross.f90:334-417,512-628:
 * whole +/- pairs per worker).  TEST INFRASTRUCTURE ONLY.
 */
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "oracle.h"
#include "rpmd.h"

/* PES seam, same signature as the reference's egrad_<pes>(q,Natoms,Nbeads,V,dVdq,info) */
void oracle_egrad_h3(const double *q, int natoms, int nbeads, double *V, double *dVdq, int *info)
{
    oracle_egrad_h3_real(q, natoms, nbeads, V, dVdq, info);
}
void oracle_egrad_oh3(const double *q, int natoms, int nbeads, double *V, double *dVdq, int *info)
{
    oracle_egrad_oh3_real(q, natoms, nbeads, V, dVdq, info);
}
void oracle_egrad_ch4h(const double *q, int natoms, int nbeads, double *V, double *dVdq, int *info)
{
    oracle_egrad_ch4h_real(q, natoms, nbeads, V, dVdq, info);
}
void oracle_egrad_brh2(const double *q, int natoms, int nbeads, double *V, double *dVdq, int *info)
{
    oracle_egrad_brh2_real(q, natoms, nbeads, V, dVdq, info);
}
void oracle_brh2_pot(const double R[3], double *V, double dVdR[3], int *ierr) { oracle_brh2_pot_real(R, V, dVdR, ierr); }
void oracle_h3_pote(const double R[3], double *pe, double dpe[3]) { oracle_h3_pote_real(R, pe, dpe); }
void oracle_oh3_pot(const double R[6], double *V, double dVdR[6]) { oracle_oh3_pot_real(R, V, dVdR); }
void oracle_ch4h_parts(const double *q18, double parts[3], double *V)
{
    oracle_ch4h_parts_real(q18, parts, V);
}
void oracle_ch4h_parts_grad(const double *q18, double parts[3], double *gparts) { oracle_ch4h_parts_grad_real(q18, parts, gparts); }
void oracle_ch4oh_parts_grad(const double *q21, double parts[3], double *gparts) { oracle_ch4oh_parts_grad_real(q21, parts, gparts); }
void oracle_geh4oh_parts_grad(const double *q21, double parts[3], double *gparts) { oracle_geh4oh_parts_grad_real(q21, parts, gparts); }
void oracle_clnh3_parts(const double *q15, double parts[3], double *V) { oracle_clnh3_parts_real(q15, parts, V); }
void oracle_clnh3_parts_grad(const double *q15, double parts[3], double *gparts) { oracle_clnh3_parts_grad_real(q15, parts, gparts); }
void oracle_clnh3_parts_frozen(const double *q15, double r0, double parts[3], double *V) { clnh3_parts_frozen_real(q15, r0, parts, V); }
void oracle_h2co_energy(const double *q12, double *V, int *far) { oracle_h2co_energy_real(q12, V, far); }
void oracle_nh3oh_parts(const double *q18, double parts[3], double *V) { oracle_nh3oh_parts_real(q18, parts, V); }
void oracle_ch4cn_parts_grad(const double *q21, double parts[3], double *gparts) { oracle_ch4cn_parts_grad_real(q21, parts, gparts); }
void oracle_ch4oh_parts(const double *q21, double parts[3], double *V)
{
    oracle_ch4oh_parts_real(q21, parts, V);
}

int oracle_egrad(int pes, const double *q, int natoms, int nimg, double *V, double *dVdq)
{
    int info = 0;
    switch (pes) {
    case ORC_PES_H3: oracle_egrad_h3_real(q, natoms, nimg, V, dVdq, &info); break;
    case ORC_PES_OH3: oracle_egrad_oh3_real(q, natoms, nimg, V, dVdq, &info); break;
    case ORC_PES_CH4H: oracle_egrad_ch4h_real(q, natoms, nimg, V, dVdq, &info); break;
    case ORC_PES_BRH2: oracle_egrad_brh2_real(q, natoms, nimg, V, dVdq, &info); break;
    case ORC_PES_O3: oracle_egrad_o3_real(q, natoms, nimg, V, dVdq, &info); break;
    case ORC_PES_CH4OH: oracle_egrad_ch4oh_real(q, natoms, nimg, V, dVdq, &info); break;
    case ORC_PES_GEH4OH: oracle_egrad_geh4oh_real(q, natoms, nimg, V, dVdq, &info); break;
    case ORC_PES_CH4CN: oracle_egrad_ch4cn_real(q, natoms, nimg, V, dVdq, &info); break;
    case ORC_PES_CLNH3: oracle_egrad_clnh3_real(q, natoms, nimg, V, dVdq, &info); break;
    case ORC_PES_NH3OH: oracle_egrad_nh3oh_real(q, natoms, nimg, V, dVdq, &info); break;
    case ORC_PES_H2CO: oracle_egrad_h2co_real(q, natoms, nimg, V, dVdq, &info); break;
    default: return -1;
    }
    return info;
}

/* ---- threaded recrossing driver: pairs [pair0, pair0+npairs) distributed over nthreads ---- */
typedef struct {
    const orc_sys *proto;
    const double *q_parents; /* [nparent][nbeads][natoms][3] */
    int nparent, pair0, npairs, child_evol, tid, nthreads;
    double xi_ideal;
    uint64_t seed;
    double *num, denom;
    int status;
} rc_job;

static orc_sys *clone_sys(const orc_sys *p)
{
    orc_sys *s = oracle_sys_create(p->natoms, p->nbeads, p->mass, p->at_move, p->beta, p->dt, p->pes);
    int keepn = s->natoms, keepb = s->nbeads;
    double *m = s->mass, *q = s->q, *pp = s->p;
    int *am = s->at_move;
    memcpy(s, p, sizeof(orc_sys));
    s->mass = m; s->q = q; s->p = pp; s->at_move = am; s->natoms = keepn; s->nbeads = keepb;
    return s;
}

static void *rc_worker(void *arg)
{
    rc_job *jb = (rc_job *)arg;
    orc_sys *s = clone_sys(jb->proto);
    size_t tot = (size_t)3 * s->natoms * s->nbeads;
    int g;
    jb->denom = 0.0;
    jb->status = 0;
    for (g = jb->tid; g < jb->npairs; g += jb->nthreads) {
        int pair = jb->pair0 + g, st;
        memcpy(s->q, jb->q_parents + (size_t)(pair % jb->nparent) * tot, sizeof(double) * tot);
        s->thermostat = 0; s->andersen_step = 0; s->k_force = 0.0; /* recross_serial.f90:157-160 */
        oracle_sys_set_rng(s, jb->seed, (uint32_t)pair, 0);
        st = orc_recross_pair(s, jb->xi_ideal, jb->child_evol, jb->num, &jb->denom);
        jb->status |= st;
    }
    oracle_sys_free(s);
    return 0;
}

/* kappa numerators/denominator for child pairs pair0..pair0+npairs-1; pair g starts from
 * parent snapshot g % nparent with RNG stream (seed, traj=g, event 0).  Returns status. */
int oracle_recross_children(orc_sys *proto, const double *q_parents, int nparent, int pair0,
                            int npairs, int child_evol, double xi_ideal, uint64_t seed,
                            int nthreads, double *kappa_num, double *kappa_denom)
{
    pthread_t *th;
    rc_job *jobs;
    int t, l, status = 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > npairs) nthreads = npairs > 0 ? npairs : 1;
    th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    jobs = (rc_job *)calloc(nthreads, sizeof(rc_job));
    for (t = 0; t < nthreads; t++) {
        jobs[t].proto = proto; jobs[t].q_parents = q_parents; jobs[t].nparent = nparent;
        jobs[t].pair0 = pair0; jobs[t].npairs = npairs; jobs[t].child_evol = child_evol;
        jobs[t].tid = t; jobs[t].nthreads = nthreads; jobs[t].xi_ideal = xi_ideal;
        jobs[t].seed = seed; jobs[t].num = (double *)calloc(child_evol, sizeof(double));
        pthread_create(&th[t], 0, rc_worker, &jobs[t]);
    }
    for (l = 0; l < child_evol; l++) kappa_num[l] = 0.0;
    *kappa_denom = 0.0;
    for (t = 0; t < nthreads; t++) {
        pthread_join(th[t], 0);
        for (l = 0; l < child_evol; l++) kappa_num[l] += jobs[t].num[l];
        *kappa_denom += jobs[t].denom;
        status |= jobs[t].status;
        free(jobs[t].num);
    }
    free(th);
    free(jobs);
    return status;
}
