/*
 * oracle.h -- CPU oracle for the Caracal RPMD hot path (TEST INFRASTRUCTURE ONLY).
 *
 * A literal C restatement of the reference's Fortran routines, used exclusively as
 * the parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product (caracal_b200/) never links or loads it.
 * Parity status: UNPINNED by the reference (no Fortran toolchain here,
 * reference ships no tests or golden outputs) -- see DESIGN.md "Oracle".
 *
 * Layout everywhere: Fortran X(3,natoms,nbeads) column-major == C [bead][atom][xyz].
 */
#ifndef ORACLE_H
#define ORACLE_H
#include "oracle_real.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- `real`-typed internals (double in the checker build, counting type otherwise) ---- */
void oracle_egrad_h3_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info);
void oracle_h3_pote_real(const real R[3], real *pe, real dpe[3]);
void oracle_egrad_oh3_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info);
void oracle_oh3_pot_real(const real R[6], real *V, real dVdR[6]);
void oracle_egrad_ch4h_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info);
void oracle_ch4h_parts_real(const real *q18, real parts[3], real *V);
void oracle_ch4h_parts_grad_real(const real *q18, real parts[3], real *gparts);
void oracle_ch4oh_parts_grad_real(const real *q21, real parts[3], real *gparts);
void oracle_geh4oh_parts_grad_real(const real *q21, real parts[3], real *gparts);
void oracle_ch4cn_parts_grad_real(const real *q21, real parts[3], real *gparts);
void oracle_egrad_brh2_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info);
void oracle_egrad_o3_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info);
void oracle_egrad_ch4oh_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info);
void oracle_ch4oh_parts_real(const real *q21, real parts[3], real *V);
void oracle_egrad_geh4oh_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info);
void oracle_geh4oh_parts_real(const real *q21, real parts[3], real *V);
void oracle_egrad_ch4cn_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info);
void oracle_ch4cn_parts_real(const real *q21, real parts[3], real *V);
void oracle_egrad_h2co_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info);
void oracle_h2co_energy_real(const real *q12, real *V, int *far);
void oracle_egrad_clnh3_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info);
void oracle_clnh3_parts_real(const real *q15, real parts[3], real *V);
void oracle_clnh3_parts_grad_real(const real *q15, real parts[3], real *gparts);
void clnh3_parts_frozen_real(const real *q15, double r0ch_frozen, real parts[3], real *V);
void oracle_egrad_nh3oh_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info);
void oracle_nh3oh_parts_real(const real *q18, real parts[3], real *V);
void oracle_nh3oh_parts_grad_real(const real *q18, real parts[3], real *gparts);
void nh3oh_parts_frozen_real(const real *q18, double r0ch_frozen, real parts[3], real *V);
void oracle_brh2_pot_real(const real R[3], real *V, real dVdR[3], int *ierr);

#ifdef __cplusplus
}
#endif
#endif
