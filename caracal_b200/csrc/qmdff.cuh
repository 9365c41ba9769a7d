// qmdff.cuh -- device tables and launch interface of the QMDFF force-field kernels
// (qmdff_kernels.cu).  Replaces ff_eg.f90, ff_nonb.f90:33-512 and the QMDFF1 branch of
// gradient.f90:341-362 for one QMDFF (nqmdff = 1).
#pragma once
#include <cuda_runtime.h>
#include "../../include/caracal_gpu.h"

namespace crcl {

constexpr int QM_MAXTYPE = 12;  // distinct elements in one force field
constexpr int QM_MAXCLS = 256;  // distinct rows of the c6 table that are still kept as classes

struct QmdffDev {
    int n, nbond, nangl, ntors, nnci, ldvt, nmols, ntype;
    // per atom
    int* type;      // element type index 0..ntype-1
    int* molnum;
    double* q;
    double* radsum2;  // unused placeholder (abdamp uses rad per type)
    // lists, 0-based atom indices
    int* bond;
    double* vbond;
    int* angl;
    double* vangl;
    int* tors;
    double* vtors;
    int* nci;
    double* c6;  // [n][n] symmetric: c6xy(max,min) of the reference
    int ncls;    // > 0: c6(i,j) == c6c[cls[i]][cls[j]] for every pair (verified at upload)
    int* cls;    // [n]
    double* c6c; // [ncls][ncls]
    // H/X-bond terms (ff_hb.f90)
    int nhb, ndonor, use_hb;
    int is_two;       // evaluate with the *_two semantics (second diabatic state)
    int* hb;          // [nhb][3] A,B,H 0-based
    double* vhb;      // [nhb][2]
    int* isH;         // [nhb]: third atom is hydrogen -> eabhag, else eabxag
    int* donor;       // [ndonor][3]: H, A, kind (1 halogen / 2 hydrogen), from the bond list
    double* dthr;     // [ndonor]: dum1 of ff_hb.f90:157,239
    double* dcoef;    // [ndonor]: halogen: scalexb*hbpara(-6.5,1,q_H); hydrogen: hbpara(10,5,q_A)*scalehb(A)
    double* dscal;    // [ndonor]: hydrogen: scalehb(at(A))
    double* acc_c1;   // [n]: hbpara(10,5,q_j)*scalehb(at(j))
    double* acc_s;    // [n]: scalehb(at(j))
    int* acc_no;      // [n]: at(j) is N or O (halogen-bond acceptor)
    double radH;      // rad(1)
    // per element-type tables
    double rad[QM_MAXTYPE];
    double r0ab[QM_MAXTYPE][QM_MAXTYPE], zab[QM_MAXTYPE][QM_MAXTYPE], r094[QM_MAXTYPE][QM_MAXTYPE],
        sr42[QM_MAXTYPE][QM_MAXTYPE];
    double eps1[6], eps2[6];
    int periodic, zahn;
    double box[3], coul_cut, vdw_cut, cut_low, zahn_a, zahn_par, e_zero;
    // grow-only scratch of qmdff_egrad: SoA copies of the positions, xs[img][c][n] (FP64), xf (FP32)
    double* xs;
    float4* xf;
    size_t xs_cap;
    // cell sweep of the inter-molecular part (qm_cellsort_kernel / qm_inter_cell_kernel): atoms sorted by cell
    float4* sf;       // [img][pos] {wrapped x, y, z, atom number}
    int* smol;        // [img][pos] molnum
    int* cstart;      // [img][ncell+1]
    size_t cell_cap_a, cell_cap_c;
    int cells_enabled;
};

// host: build / free the device copy; evaluate nimg images (AoS [img][atom][xyz])
int qmdff_upload(const crcl_qmdff_tables* T, QmdffDev** out, const char** err, bool is_two = false);
void qmdff_free(QmdffDev* D);
cudaError_t qmdff_egrad(QmdffDev* D, const double* d_xyz, int nimg, double* d_V, double* d_g,
                        cudaStream_t s, long long* launches);

}  // namespace crcl
