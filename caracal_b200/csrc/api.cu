// api.cu -- C-ABI of libcaracal_gpu.so (include/caracal_gpu.h): handle management, host-side
// set-up that the Fortran drivers do once (free ring-polymer kernels, mechanism tables), and
// kernel dispatch.  No CPU compute path exists behind these entry points.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "crcl_common.cuh"
#include "pes_h3.cuh"
#include "pes_oh3.cuh"
#include "pes_ch4h.cuh"
#include "pes_brh2.cuh"
#include "pes_o3.cuh"
#include "pes_nh3x.cuh"
#include "pes_h2co.cuh"
#include "transform_bench.cuh"
#include "pes_ch4oh.cuh"
#include "traj_inst.cuh"
#include "split_kernels.cuh"
#include "split_xi.cuh"
#include "qmdff.cuh"
#include "dgevb.cuh"
#include "ewald.cuh"
#include "water.cuh"
#include "comm.cuh"

using namespace crcl;

struct crcl_handle_s {
    int device = 0, natoms = 0, nbeads = 0, pes = 0;
    double beta = 0, dt = 0, kelvin = 0, nose_q = 0;
    int thermostat = 0, andersen_step = 0, transform = CRCL_TRANSFORM_REFERENCE;
    int spread_max = CRCL_SPREAD_MAX_BEADS;   // crcl_set_spread_max_beads
    uint64_t seed = 0;
    std::vector<double> mass;
    std::vector<int> at_move;
    Mech mech{};
    cudaStream_t stream = nullptr, own_stream = nullptr;
    // ring of CUDA-event pairs bracketing every trajectory/egrad kernel launch on h->stream
    static constexpr int NEV = 256;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // the pair of the most recent launch
    cudaEvent_t evs[2 * NEV] = {nullptr};
    int ev_head = 0, ev_count = 0;
    bool timed = false;
    bool capturing = false;   // inside a stream capture: no event records (they would become graph nodes)
    double* d_fker = nullptr;
    bool fker_dirty = true;
    crcl_host_grad_fn cb = nullptr;
    void* cb_user = nullptr;
    int path = CRCL_PATH_AUTO;
    // pbc_mod (wrap of verlet.f90:591-641) and the rpmd_check settings
    int periodic = 0;
    double box[3] = {0.0, 0.0, 0.0};
    int chk_on = 0;
    double chk_energy_ts = 0.0, chk_energy_tol = 0.0, chk_xi_tol = 0.0;
    bool use_graph = true;   // split path: replay steps from a CUDA graph (crcl_set_graph)
    QmdffDev* qmdff = nullptr;
    QmdffDev* qmdff2 = nullptr;
    DgevbDev* dgevb = nullptr;
    EwaldDev* ewald = nullptr;
    WaterDev* water = nullptr;
    // split path: per-atom tables on the device, generic-size mechanism
    double *d_mass = nullptr, *d_wfrag = nullptr;
    int *d_atmove = nullptr, *d_frag = nullptr;
    MechDev mechd{};
    bool mechd_valid = false;
    std::vector<int> frag_h;
    std::vector<double> wfrag_h;
    std::vector<double> hostbuf_q, hostbuf_g, hostbuf_v;
    long long launches = 0;
    Comm* comm = nullptr;   // NCCL communicator of the job (crcl_comm_init); null = single process
    std::string err;
    // grow-only device scratch
    void* scratch[32] = {nullptr};
    size_t scratch_sz[32] = {0};
};

#define CK(call)                                                                     \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);             \
            return CRCL_ECUDA;                                                       \
        }                                                                            \
    } while (0)

static void next_event_pair(crcl_handle h)
{
    const int i = h->ev_head;
    h->ev0 = h->evs[2 * i];
    h->ev1 = h->evs[2 * i + 1];
    h->ev_head = (i + 1) % crcl_handle_s::NEV;
    if (h->ev_count < crcl_handle_s::NEV) h->ev_count++;
}

static int fail(crcl_handle h, int code, const char* msg)
{
    if (h) h->err = msg;
    return code;
}

template <class T>
static int scratch(crcl_handle h, int slot, size_t n, T** out)
{
    const size_t bytes = n * sizeof(T);
    if (h->scratch_sz[slot] < bytes) {
        if (h->scratch[slot]) cudaFree(h->scratch[slot]);
        h->scratch[slot] = nullptr;
        h->scratch_sz[slot] = 0;
        if (cudaMalloc(&h->scratch[slot], bytes ? bytes : 8) != cudaSuccess) {
            h->err = "cudaMalloc failed";
            return CRCL_ENOMEM;
        }
        h->scratch_sz[slot] = bytes;
    }
    *out = static_cast<T*>(h->scratch[slot]);
    return CRCL_OK;
}

// ---- free ring-polymer kernels f_c, f_a, f_b (see traj_kernel.cuh header) ------------------
// d_k as verlet.f90:405-433: w_k = (2 nbeads/beta) sin(k pi/nbeads), mirrored k <-> N-k.
static void build_fker(int N, double beta, double dt, std::vector<double>& f)
{
    f.assign((size_t)3 * N, 0.0);
    std::vector<double> dc(N), da(N), db(N);
    dc[0] = 1.0;
    da[0] = 0.0;
    db[0] = dt;
    const double beta_n = beta / N, twown = 2.0 / beta_n, pi_n = PI_QMDFF / N;
    for (int k = 1; k <= N / 2; k++) {
        const double wk = twown * std::sin(k * pi_n), wt = wk * dt;
        dc[k] = std::cos(wt);
        da[k] = -wk * std::sin(wt);
        db[k] = std::sin(wt) / wk;
    }
    for (int k = 1; k <= (N - 1) / 2; k++) {
        dc[N - k] = dc[k];
        da[N - k] = da[k];
        db[N - k] = db[k];
    }
    for (int j = 0; j < N; j++) {
        long double sc = 0, sa = 0, sb = 0;
        for (int k = 0; k < N; k++) {
            const long double c = cosl(2.0L * 3.14159265358979323846264338327950288L * ((k * j) % N) / N);
            sc += dc[k] * c;
            sa += da[k] * c;
            sb += db[k] * c;
        }
        f[j] = (double)(sc / N);
        f[N + j] = (double)(sa / N);
        f[2 * N + j] = (double)(sb / N);
    }
}

static int ensure_fker(crcl_handle h)
{
    if (!h->fker_dirty) return CRCL_OK;
    std::vector<double> f;
    build_fker(h->nbeads, h->beta, h->dt, f);
    if (h->d_fker) cudaFree(h->d_fker);
    CK(cudaMalloc(&h->d_fker, f.size() * sizeof(double)));
    CK(cudaMemcpyAsync(h->d_fker, f.data(), f.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->fker_dirty = false;
    return CRCL_OK;
}

// ---- small kernels ---------------------------------------------------------------------------
// One thread per image, each walking its own 3 NATOMS doubles of q and g ([image][atom][xyz]).
// Tried and rejected on measurement (profiles/r2x_bench_egrad.json against r2t_bench_egrad.json): the CTA's tile moved
// with coalesced 16-byte accesses through shared memory -- every surface but one got slower (H3 0.157 -> 0.193 ms per 2^20
// images, OH3 0.081 -> 0.101, CH4 + OH 0.58 -> 0.88): the per-thread accesses already hit full sectors through L1, and the
// two barriers plus the tile's registers cost more than they save.  What that build did show: Br + H2 gains from three
// resident CTAs (168 registers instead of 188: 0.418 -> 0.347 ms), hence the launch bound below for that surface.
template <class PES>
__device__ __forceinline__ void egrad_body(const double* __restrict__ q, int nimg, double* __restrict__ V,
                                           double* __restrict__ g, int* __restrict__ info)
{
    constexpr int NC = 3 * PES::NATOMS;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nimg) return;
    double x[NC], gr[NC], e;
#pragma unroll
    for (int c = 0; c < NC; c++) x[c] = q[(size_t)i * NC + c];
    const int w = PES::eval(x, e, gr);
    V[i] = e;
#pragma unroll
    for (int c = 0; c < NC; c++) g[(size_t)i * NC + c] = gr[c];
    if (w && info) atomicOr(info, w);
}
template <class PES>
__global__ void egrad_kernel(const double* __restrict__ q, int nimg, double* __restrict__ V,
                             double* __restrict__ g, int* __restrict__ info)
{
    egrad_body<PES>(q, nimg, V, g, info);
}
constexpr int EGRAD_TPB = 128;
template <class PES>
__global__ void __launch_bounds__(EGRAD_TPB, 3) egrad_kernel_mb3(const double* __restrict__ q, int nimg, double* __restrict__ V,
                                                                 double* __restrict__ g, int* __restrict__ info)
{
    egrad_body<PES>(q, nimg, V, g, info);
}

template <int NAT>
__global__ void calc_xi_kernel(const __grid_constant__ TrajArgs A, int n, const double* coords,
                               const double* xi_ideal, int mode, double* xi, double* dxi, double* hams)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x[3 * NAT], d[3 * NAT], hh[3 * NAT], v;
#pragma unroll
    for (int c = 0; c < 3 * NAT; c++) x[c] = coords[(size_t)i * 3 * NAT + c];
    calc_xi<NAT>(A.mech, A.mass, x, xi_ideal[i], mode, v, d, (hams && mode == 1) ? hh : nullptr, A.beta);
    xi[i] = v;
#pragma unroll
    for (int c = 0; c < 3 * NAT; c++) {
        dxi[(size_t)i * 3 * NAT + c] = d[c];
        if (hams && mode == 1) hams[(size_t)i * 3 * NAT + c] = hh[c];
    }
}

__global__ void rng_kernel(uint64_t seed, uint32_t traj, uint32_t event, uint32_t bead, int n, double* out)
{
    const int pr = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * pr >= n) return;
    double z0, z1;
    normal_pair(seed, traj, event, bead, (uint32_t)pr, z0, z1);
    out[2 * pr] = z0;
    if (2 * pr + 1 < n) out[2 * pr + 1] = z1;
}

// DFMA throughput probe: 8 independent chains per thread
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
           x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b);
        x1 = fma(x1, a, b);
        x2 = fma(x2, a, b);
        x3 = fma(x3, a, b);
        x4 = fma(x4, a, b);
        x5 = fma(x5, a, b);
        x6 = fma(x6, a, b);
        x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// DMMA throughput probe: mma.sync.m8n8k4.f64 (the only FP64 tensor-core shape; tcgen05 has no FP64 kind),
// 4 independent accumulator pairs per warp.  2*8*8*4 = 512 flops per instruction per warp.
__global__ void dmma_peak_kernel(double* out, int iters, double a, double b)
{
    double c0[2] = {0.0, 0.0}, c1[2] = {1.0, 0.0}, c2[2] = {0.0, 1.0}, c3[2] = {1.0, 1.0};
    const double av = a + threadIdx.x * 1e-9, bv = b;
    for (int i = 0; i < iters; i++) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[0]), "+d"(c0[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c1[0]), "+d"(c1[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c2[0]), "+d"(c2[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c3[0]), "+d"(c3[1]) : "d"(av), "d"(bv));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

namespace crcl {
// kappa_num[l] = sum_t weight[t]*theta[l][t], kappa_denom = sum_t denom_part[t]; fixed
// summation order -> bit-reproducible for a given (pair0, npairs) regardless of launch shape.
__global__ void reduce_kappa_kernel(const unsigned char* theta, const double* weight,
                                    const double* denom_part, int ntraj, int nsteps,
                                    double* kappa_num, double* kappa_denom)
{
    __shared__ double sh[256];
    const int l = blockIdx.x;  // l == nsteps -> denominator
    double s = 0.0;
    if (l < nsteps) {
        for (int t = threadIdx.x; t < ntraj; t += 256)
            if (theta[(size_t)l * ntraj + t]) s += weight[t];
    } else {
        for (int t = threadIdx.x; t < ntraj; t += 256) s += denom_part[t];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (l < nsteps)
            kappa_num[l] = sh[0];
        else
            *kappa_denom = sh[0];
    }
}

}  // namespace crcl

// ---- dispatch --------------------------------------------------------------------------------
static int launch_traj(crcl_handle h, int kind, const TrajArgs& A, int bias_mode = 0)
{
    static const traj_launch_fn table[11][3] = {
        {launch_h3_verlet, launch_h3_mdinit, launch_h3_recross},
        {launch_oh3_verlet, launch_oh3_mdinit, launch_oh3_recross},
        {launch_ch4h_verlet, launch_ch4h_mdinit, launch_ch4h_recross},
        {launch_brh2_verlet, launch_brh2_mdinit, launch_brh2_recross},
        {launch_o3_verlet, launch_o3_mdinit, launch_o3_recross},
        {launch_ch4oh_verlet, launch_ch4oh_mdinit, launch_ch4oh_recross},
        {launch_geh4oh_verlet, launch_geh4oh_mdinit, launch_geh4oh_recross},
        {launch_ch4cn_verlet, launch_ch4cn_mdinit, launch_ch4cn_recross},
        {launch_clnh3_verlet, launch_clnh3_mdinit, launch_clnh3_recross},
        {launch_nh3oh_verlet, launch_nh3oh_mdinit, launch_nh3oh_recross},
        {launch_h2co_verlet, launch_h2co_mdinit, launch_h2co_recross}};
    int row;
    switch (h->pes) {
    case CRCL_PES_H3: row = 0; break;
    case CRCL_PES_OH3: row = 1; break;
    case CRCL_PES_CH4H: row = 2; break;
    case CRCL_PES_BRH2: row = 3; break;
    case CRCL_PES_O3: row = 4; break;
    case CRCL_PES_CH4OH: row = 5; break;
    case CRCL_PES_GEH4OH: row = 6; break;
    case CRCL_PES_CH4CN: row = 7; break;
    case CRCL_PES_CLNH3: row = 8; break;
    case CRCL_PES_NH3OH: row = 9; break;
    case CRCL_PES_H2CO: row = 10; break;
    default: return fail(h, CRCL_ENOSUP, "no device trajectory kernel for this PES id");
    }
    if (A.ntraj <= 0) return CRCL_OK;
    int nosup = 0;
    if (h->timed && !h->capturing) {
        next_event_pair(h);
        cudaEventRecord(h->ev0, h->stream);
    }
    cudaError_t e = table[row][kind](h->nbeads, A, bias_mode, h->nose_q, h->stream, &nosup);
    if (nosup) return fail(h, CRCL_ENOSUP, "fused trajectory kernel: nbeads must be a power of two <= 128");
    if (h->timed && !h->capturing) cudaEventRecord(h->ev1, h->stream);
    h->launches++;
    if (e != cudaSuccess) {
        h->err = std::string("trajectory kernel launch: ") + cudaGetErrorString(e);
        return CRCL_ECUDA;
    }
    return CRCL_OK;
}

static void fill_args(crcl_handle h, TrajArgs& A)
{
    memset(&A, 0, sizeof(A));
    A.nbeads = h->nbeads;
    A.beta = h->beta;
    A.dt = h->dt;
    A.kelvin = h->kelvin;
    A.thermostat = h->thermostat;
    A.andersen_step = h->andersen_step;
    A.symmetrize = (h->transform == CRCL_TRANSFORM_REFERENCE) ? 1 : 0;
    A.spread_max = h->spread_max;
    for (int i = 0; i < h->natoms && i < TRAJ_MAXNAT; i++) {
        A.mass[i] = h->mass[i];
        A.at_move[i] = h->at_move[i];
    }
    A.mech = h->mech;
    A.seed = h->seed;
    A.fker = h->d_fker;
    A.chk_on = h->chk_on;
    A.chk_emax = (h->chk_energy_ts + h->chk_energy_tol) * h->nbeads;   // rpmd_check.f90:100
    A.chk_xi_tol = h->chk_xi_tol;
}

static int pes_natoms(int pes)
{
    switch (pes) {
    case CRCL_PES_H3: return 3;
    case CRCL_PES_OH3: return 4;
    case CRCL_PES_CH4H: return 6;
    case CRCL_PES_BRH2: return 3;
    case CRCL_PES_O3: return 3;
    case CRCL_PES_CH4OH: return 7;
    case CRCL_PES_GEH4OH: return 7;
    case CRCL_PES_CH4CN: return 7;
    case CRCL_PES_CLNH3: return 5;
    case CRCL_PES_NH3OH: return 6;
    case CRCL_PES_H2CO: return 4;
    }
    return -1;
}

// ---- split (HBM-resident) path ----------------------------------------------------------------
namespace crcl {
__global__ void sp_transrot_apply(const SplitArgs A, double mt, const double* sums)
{
    const int na = A.natoms, nb = A.nbeads, rb = gridDim.x / A.ntraj, t = blockIdx.x / rb, bx = blockIdx.x - t * rb;
    const size_t nab = (size_t)na * nb;
    const double* s = sums + (size_t)t * 16;
    double totmass, vtot[3], ctr[3];
    sp_centre(A, s, mt, totmass, vtot, ctr);
    double mang[3];
    mang[0] = s[6] - (ctr[1] * vtot[2] - ctr[2] * vtot[1]) * totmass;
    mang[1] = s[7] - (ctr[2] * vtot[0] - ctr[0] * vtot[2]) * totmass;
    mang[2] = s[8] - (ctr[0] * vtot[1] - ctr[1] * vtot[0]) * totmass;
    const double xx = s[9], xy = s[10], xz = s[11], yy = s[12], yz = s[13], zz = s[14];
    double ten[3][3] = {{yy + zz, -xy, -xz}, {-xy, xx + zz, -yz}, {-xz, -yz, xx + yy}};
    if (na <= 2) {
        ten[0][0] += 0.000001;
        ten[1][1] += 0.000001;
        ten[2][2] += 0.000001;
    }
    if (invert3(ten)) {
        if (bx == 0 && threadIdx.x == 0) atomicOr(&A.status[t], CRCL_TRAJ_SINGULAR);
        return;
    }
    double vang[3];
    for (int i = 0; i < 3; i++) vang[i] = ten[i][0] * mang[0] + ten[i][1] * mang[1] + ten[i][2] * mang[2];
    for (size_t e = (size_t)bx * blockDim.x + threadIdx.x; e < nab; e += (size_t)rb * blockDim.x) {
        const int atom = (int)(e % na);
        const size_t k = ((size_t)t * nab + e) * 3;
        const double w = A.mass[atom];
        const bool mv = A.at_move[atom] != 0;
        const double xd = A.q[k] - ctr[0], yd = A.q[k + 1] - ctr[1], zd = A.q[k + 2] - ctr[2];
        const double v0 = (A.p[k] / w - vtot[0]) - vang[1] * zd + vang[2] * yd;
        const double v1 = (A.p[k + 1] / w - vtot[1]) - vang[2] * xd + vang[0] * zd;
        const double v2 = (A.p[k + 2] / w - vtot[2]) - vang[0] * yd + vang[1] * xd;
        A.p[k] = mv ? v0 * w : 0.0;
        A.p[k + 1] = mv ? v1 * w : 0.0;
        A.p[k + 2] = mv ? v2 * w : 0.0;
    }
}
}  // namespace crcl

static bool fused_ok(crcl_handle h)
{
    const int nb = h->nbeads;
    const bool pow2 = (nb & (nb - 1)) == 0 && nb <= 128;
    return h->natoms <= TRAJ_MAXNAT && pow2 && pes_natoms(h->pes) == h->natoms;
}
static bool use_split(crcl_handle h)
{
    if (h->path == CRCL_PATH_SPLIT || h->periodic) return true;
    if (h->path == CRCL_PATH_FUSED) return false;
    return !fused_ok(h);
}

static int ensure_split_tables(crcl_handle h)
{
    if (h->d_mass) return CRCL_OK;
    CK(cudaMalloc(&h->d_mass, h->natoms * sizeof(double)));
    CK(cudaMalloc(&h->d_atmove, h->natoms * sizeof(int)));
    CK(cudaMemcpy(h->d_mass, h->mass.data(), h->natoms * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_atmove, h->at_move.data(), h->natoms * sizeof(int), cudaMemcpyHostToDevice));
    return CRCL_OK;
}

static int upload_mechd(crcl_handle h)
{
    if (!h->mechd_valid) return CRCL_OK;
    if (h->d_frag) cudaFree(h->d_frag);
    if (h->d_wfrag) cudaFree(h->d_wfrag);
    CK(cudaMalloc(&h->d_frag, h->natoms * sizeof(int)));
    CK(cudaMalloc(&h->d_wfrag, h->natoms * sizeof(double)));
    CK(cudaMemcpy(h->d_frag, h->frag_h.data(), h->natoms * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_wfrag, h->wfrag_h.data(), h->natoms * sizeof(double), cudaMemcpyHostToDevice));
    h->mechd.frag = h->d_frag;
    h->mechd.wfrag = h->d_wfrag;
    return CRCL_OK;
}

// PES for all images of the split path: device kernel for the analytic surfaces, otherwise the host
// callback (custom_grad.f90:35 / external_grad.f90:33 seam): D2H positions, one call per image as
// gradient.f90 does per bead (verlet.f90:772-777), H2D gradients.
static int split_forces(crcl_handle h, int nimg, const double* dq, double* dg, double* dV)
{
    if (h->pes != CRCL_PES_HOSTCB && h->pes != CRCL_PES_NONE)
        return crcl_egrad_dev(h, h->pes, dq, h->natoms, nimg, dV, dg, nullptr);
    if (!h->cb) return fail(h, CRCL_ESTATE, "PES is CRCL_PES_HOSTCB but crcl_set_host_gradient_cb was not called");
    const size_t n = (size_t)nimg * 3 * h->natoms;
    h->hostbuf_q.resize(n);
    h->hostbuf_g.resize(n);
    h->hostbuf_v.resize(nimg);
    CK(cudaMemcpyAsync(h->hostbuf_q.data(), dq, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < nimg; i++)
        h->cb(h->hostbuf_q.data() + (size_t)i * 3 * h->natoms, &h->hostbuf_v[i],
              h->hostbuf_g.data() + (size_t)i * 3 * h->natoms, h->natoms, h->cb_user);
    CK(cudaMemcpyAsync(dg, h->hostbuf_g.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dV, h->hostbuf_v.data(), nimg * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return CRCL_OK;
}

template <int NB>
static void launch_kfr_reg(const SplitArgs& A, dim3 grid, cudaStream_t s)
{
    sp_kick_freerp_reg<NB><<<grid, 128, 0, s>>>(A);
}

static int launch_kick_freerp(crcl_handle h, const SplitArgs& A)
{
    const int nc = 3 * A.natoms;
    const size_t ncomp = (size_t)nc * A.ntraj;
    dim3 grid((unsigned)((ncomp + 127) / 128));
    if (h->timed && !h->capturing) {
        next_event_pair(h);
        cudaEventRecord(h->ev0, h->stream);
    }
    switch (A.nbeads) {
    case 1: launch_kfr_reg<1>(A, grid, h->stream); break;
    case 2: launch_kfr_reg<2>(A, grid, h->stream); break;
    case 3: launch_kfr_reg<3>(A, grid, h->stream); break;
    case 4: launch_kfr_reg<4>(A, grid, h->stream); break;
    case 6: launch_kfr_reg<6>(A, grid, h->stream); break;
    case 8: launch_kfr_reg<8>(A, grid, h->stream); break;
    case 12: launch_kfr_reg<12>(A, grid, h->stream); break;
    case 16: launch_kfr_reg<16>(A, grid, h->stream); break;
    default: {
        int bd = 128;
        while (bd > 32 && (size_t)(3 * A.nbeads + 2 * (size_t)A.nbeads * bd) * sizeof(double) > 200 * 1024) bd >>= 1;
        const size_t smem = (size_t)(3 * A.nbeads + 2 * (size_t)A.nbeads * bd) * sizeof(double);
        if (smem > 200 * 1024) return fail(h, CRCL_ENOSUP, "split path: nbeads too large for the shared-memory transform");
        CK(cudaFuncSetAttribute(sp_kick_freerp_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 g2((unsigned)((ncomp + bd - 1) / bd));
        sp_kick_freerp_smem<<<g2, bd, smem, h->stream>>>(A);
    }
    }
    if (h->timed && !h->capturing) cudaEventRecord(h->ev1, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    return CRCL_OK;
}

// Device-resident state of a batch on the split path (all pointers device memory).
struct SplitCall {
    int ntraj = 0;
    double *q = nullptr, *p = nullptr, *g = nullptr, *dxi = nullptr, *epot = nullptr, *xi_real = nullptr, *nhc = nullptr;
    int* status = nullptr;
    const uint32_t* tid = nullptr;
    uint32_t* event = nullptr;
    const double *xi_ideal = nullptr, *k_force = nullptr;   // per trajectory, or null -> scalars
    double xi_ideal_s = 0.0, k_force_s = 0.0;
    double *xi_sum = nullptr, *xi_sum2 = nullptr;           // umbrella sampling accumulators (may be null)
    unsigned char* theta = nullptr;                         // [nsteps][ntraj] (recrossing; may be null)
};

static int nfree_of(crcl_handle h)
{
    int nf = 0;
    for (int i = 0; i < h->natoms; i++)
        if (h->at_move[i]) nf += 3;   // mdinit.f90:131-137: not multiplied by nbeads
    return nf;
}

// scratch of the biased / constrained modes; fills S
static int split_traj_scratch(crcl_handle h, const SplitCall& C, SpTraj& S)
{
    const int na = h->natoms, nc = 3 * na, nt = C.ntraj;
    double *dh, *dw, *dco;
    int* dbad;
    int rc;
    if ((rc = scratch(h, 16, (size_t)nt * nc, &dh)) || (rc = scratch(h, 17, (size_t)nt * SPX_NWORK * nc, &dw)) ||
        (rc = scratch(h, 18, (size_t)nt * 2, &dco)) || (rc = scratch(h, 19, (size_t)nt, &dbad)))
        return rc;
    S.ntraj = nt;
    S.natoms = na;
    S.nbeads = h->nbeads;
    S.dt = h->dt;
    S.beta = h->beta;
    S.mass = h->d_mass;
    S.at_move = h->d_atmove;
    S.xi_ideal = C.xi_ideal;
    S.k_force = C.k_force;
    S.xi_ideal_s = C.xi_ideal_s;
    S.k_force_s = C.k_force_s;
    S.xi_real = C.xi_real;
    S.dxi = C.dxi;
    S.hams = dh;
    S.work = dw;
    S.status = C.status;
    S.bad = dbad;
    S.coeff = dco;
    return CRCL_OK;
}

// nsteps verlet steps on device-resident state (split path): every constrain mode of verlet.f90
// (-1 plain MD, 0 / 3 umbrella, 1 SHAKE / RATTLE, 2 child trajectory), thermostats 0 / 1 / 2.
static int verlet_split(crcl_handle h, const SplitCall& C, int nsteps, int istep0, int constrain)
{
    if (constrain >= 0 && !h->mechd_valid) return fail(h, CRCL_ESTATE, "crcl_set_mechanism has not been called");
    int rc;
    if ((rc = ensure_split_tables(h)) || (rc = ensure_fker(h))) return rc;
    const int na = h->natoms, nb = h->nbeads, nc = 3 * na, ntraj = C.ntraj;
    const size_t per = (size_t)nb * nc;
    double *dcen, *dV, *dsums;
    if ((rc = scratch(h, 8, (size_t)ntraj * nc, &dcen)) || (rc = scratch(h, 9, (size_t)ntraj * nb, &dV)) ||
        (rc = scratch(h, 10, (size_t)ntraj * 16, &dsums)))
        return rc;
    SpTraj S{};
    if (constrain >= 0 || h->thermostat == 2)
        if ((rc = split_traj_scratch(h, C, S))) return rc;
    SplitArgs A;
    A.ntraj = ntraj;
    A.natoms = na;
    A.nbeads = nb;
    A.symmetrize = (h->transform == CRCL_TRANSFORM_REFERENCE) ? 1 : 0;
    A.dt = h->dt;
    A.beta = h->beta;
    A.mass = h->d_mass;
    A.at_move = h->d_atmove;
    A.fker = h->d_fker;
    A.q = C.q;
    A.p = C.p;
    A.g = C.g;
    A.cen = dcen;
    A.status = C.status;
    A.seed = h->seed;
    A.traj_id = C.tid;
    A.traj_id0 = 0;
    A.event = C.event;
    A.periodic = h->periodic;
    for (int d = 0; d < 3; d++) A.box[d] = h->box[d];
    double mt = 0.0;
    for (int i = 0; i < na; i++) mt += h->mass[i];
    cudaStream_t s = h->stream;
    const dim3 gel((unsigned)(((per + 255) / 256) * ntraj));
    const int rblocks = (int)std::min<size_t>(64, ((size_t)na * nb + 255) / 256);
    const int tb = (ntraj + 63) / 64, nfree = nfree_of(h);
    const bool nhc_on = (constrain != 2 && h->thermostat == 2);
    // One step as a sequence of 6-14 small launches.  Small systems are launch-bound, so steps 2..nsteps are
    // replayed from a CUDA graph captured once per call (two variants: with / without the Andersen draw);
    // step 1 runs eagerly so that every grow-only scratch buffer has its final size before the capture.
    // (a capture cannot start on the legacy default stream: a handle that was given stream 0 launches eagerly)
    bool can_graph = h->use_graph && nsteps >= 4 && h->pes != CRCL_PES_HOSTCB && h->pes != CRCL_PES_NONE && s != nullptr &&
                     s != cudaStreamLegacy;
    uint32_t* dctr = nullptr;   // device step counter: where sp_theta writes when the step is a graph replay
    if (can_graph && C.theta) {
        if ((rc = scratch(h, 27, (size_t)2, &dctr))) return rc;
        CK(cudaMemsetAsync(dctr, 0, 2 * sizeof(uint32_t), s));
    }
    auto one_step = [&](int st, bool andersen_now) -> int {
        int rc;
        if (nhc_on) {                                                        // 1
            sp_nhc<<<ntraj, 256, 0, s>>>(S, C.p, C.nhc, h->kelvin, nfree);
            h->launches++;
        }
        if ((rc = launch_kick_freerp(h, A))) return rc;                       // 2,3,4,6,7
        if (constrain == 1) {                                                 // 9
            sp_shake_solve<<<tb, 64, 0, s>>>(h->mechd, S, dcen);
            sp_shake_apply<<<gel, 256, 0, s>>>(S, C.q, C.p);
            h->launches += 2;
        }
        if ((rc = split_forces(h, ntraj * nb, C.q, C.g, dV))) return rc;      // 10
        sp_epot<<<(ntraj + 127) / 128, 128, 0, s>>>(dV, nb, ntraj, C.epot);
        h->launches++;
        if (constrain == 1) {
            sp_shake_penalty<<<(ntraj + 127) / 128, 128, 0, s>>>(ntraj, S.bad, C.epot);
            h->launches++;
        }
        if (constrain == 0 || constrain == 3) {                               // 12
            sp_calc_xi_kernel<<<tb, 64, 0, s>>>(h->mechd, S, dcen, 1, 1);
            sp_add_bias<<<gel, 256, 0, s>>>(S, C.g);
            h->launches += 2;
        } else if (constrain == 1) {
            sp_calc_xi_kernel<<<tb, 64, 0, s>>>(h->mechd, S, dcen, 2, 0);
            h->launches++;
        } else if (constrain == 2) {
            sp_xi_value<<<tb, 64, 0, s>>>(h->mechd, na, ntraj, dcen, C.xi_ideal, C.xi_ideal_s, 2, C.xi_real);
            h->launches++;
        }
        if (h->chk_on && constrain >= 0 && constrain != 2) {                  // rpmd_check.f90:88-116
            sp_rpmd_check<<<(ntraj + 127) / 128, 128, 0, s>>>(ntraj, C.epot, C.xi_real, C.xi_ideal, C.xi_ideal_s,
                                                              (h->chk_energy_ts + h->chk_energy_tol) * nb,
                                                              h->chk_xi_tol, constrain != 1, C.status);
            h->launches++;
        }
        sp_kick<<<gel, 256, 0, s>>>(A);                                       // 13, 18
        h->launches++;
        if (constrain == 1) {                                                 // 14
            sp_rattle<<<ntraj, 256, 0, s>>>(S, C.p);
            h->launches++;
        }
        if (nhc_on) {                                                         // 15
            sp_nhc<<<ntraj, 256, 0, s>>>(S, C.p, C.nhc, h->kelvin, nfree);
            h->launches++;
        }
        if (andersen_now) {
            sp_andersen<<<gel, 256, 0, s>>>(A);                               // 16
            sp_bump_event<<<(ntraj + 127) / 128, 128, 0, s>>>(C.event, ntraj);
            h->launches += 2;
        }
        if (constrain <= 0) {                                                 // 19
            CK(cudaMemsetAsync(dsums, 0, (size_t)ntraj * 16 * sizeof(double), s));
            sp_transrot_sums1<<<rblocks * ntraj, 128, 0, s>>>(A, dsums);
            sp_transrot_sums2<<<rblocks * ntraj, 128, 0, s>>>(A, mt, dsums);
            sp_transrot_apply<<<rblocks * ntraj, 128, 0, s>>>(A, mt, dsums);
            h->launches += 3;
        }
        if (C.xi_sum && constrain >= 0) {
            sp_accum_xi<<<(ntraj + 127) / 128, 128, 0, s>>>(ntraj, C.xi_real, C.xi_sum, C.xi_sum2);
            h->launches++;
        }
        if (C.theta) {
            sp_theta<<<(ntraj + 127) / 128, 128, 0, s>>>(ntraj, C.xi_real, C.theta + (dctr ? 0 : (size_t)(st - 1) * ntraj), dctr);
            h->launches++;
            if (dctr) {
                sp_bump_event<<<1, 32, 0, s>>>(dctr, 1);
                h->launches++;
            }
        }
        CK(cudaGetLastError());
        return CRCL_OK;
    };
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    long long glaunches[2] = {0, 0};
    auto drop_graphs = [&]() {
        for (auto& g : gexec)
            if (g) {
                cudaGraphExecDestroy(g);
                g = nullptr;
            }
    };
    for (int st = 1; st <= nsteps; st++) {
        const int istep = istep0 + st;
        const bool an = (constrain != 2 && h->thermostat == 1 && h->andersen_step > 0 && (istep % h->andersen_step) == 0);
        if (!can_graph || st == 1) {
            if ((rc = one_step(st, an))) {
                drop_graphs();
                return rc;
            }
            continue;
        }
        const int v = an ? 1 : 0;
        if (!gexec[v]) {
            const long long l0 = h->launches;
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
            if (e != cudaSuccess) {
                // this stream cannot be captured (e.g. it is itself part of somebody else's capture): eager launches
                cudaGetLastError();
                can_graph = false;
                if ((rc = one_step(st, an))) {
                    drop_graphs();
                    return rc;
                }
                continue;
            }
            h->capturing = true;
            rc = one_step(st, an);
            h->capturing = false;
            e = cudaStreamEndCapture(s, &graph);
            glaunches[v] = h->launches - l0;
            h->launches = l0;
            if (rc == CRCL_OK && e == cudaSuccess) e = cudaGraphInstantiate(&gexec[v], graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (rc || e != cudaSuccess) {
                drop_graphs();
                if (!rc) h->err = std::string("CUDA graph capture of a split-path step: ") + cudaGetErrorString(e);
                return rc ? rc : CRCL_ECUDA;
            }
        }
        cudaError_t e = cudaGraphLaunch(gexec[v], s);
        if (e != cudaSuccess) {
            drop_graphs();
            h->err = std::string("cudaGraphLaunch: ") + cudaGetErrorString(e);
            return CRCL_ECUDA;
        }
        h->launches += glaunches[v];
    }
    if (gexec[0] || gexec[1]) {
        // the executable graphs must outlive their last replay
        CK(cudaStreamSynchronize(s));
        drop_graphs();
    }
    // child steps only evaluate the value of xi; leave dxi as verlet.f90:1049-1050 would
    if (constrain == 2 && nsteps > 0 && C.dxi) {
        sp_calc_xi_kernel<<<tb, 64, 0, s>>>(h->mechd, S, dcen, 2, 0);
        h->launches++;
        CK(cudaGetLastError());
    }
    return CRCL_OK;
}

// mdinit on the split path (mdinit.f90:40-172): forces of all beads, centroid, umbrella (bias_mode 1:
// xi and dxi in the recrossing form; 2: bias applied), Andersen draw, NHC reset
static int mdinit_split(crcl_handle h, const SplitCall& C, int bias_mode)
{
    if (bias_mode && !h->mechd_valid) return fail(h, CRCL_ESTATE, "crcl_set_mechanism has not been called");
    int rc;
    if ((rc = ensure_split_tables(h))) return rc;
    const int na = h->natoms, nb = h->nbeads, nc = 3 * na, ntraj = C.ntraj;
    const size_t per = (size_t)nb * nc;
    double *dV, *dcen;
    if ((rc = scratch(h, 8, (size_t)ntraj * nc, &dcen)) || (rc = scratch(h, 9, (size_t)ntraj * nb, &dV))) return rc;
    if ((rc = split_forces(h, ntraj * nb, C.q, C.g, dV))) return rc;
    cudaStream_t s = h->stream;
    const dim3 gel((unsigned)(((per + 255) / 256) * ntraj));
    if (bias_mode) {
        SpTraj S{};
        if ((rc = split_traj_scratch(h, C, S))) return rc;
        sp_centroid<<<(unsigned)(((size_t)ntraj * nc + 255) / 256), 256, 0, s>>>(ntraj, na, nb, C.q, dcen);
        if (bias_mode == 1) {
            sp_calc_xi_kernel<<<(ntraj + 63) / 64, 64, 0, s>>>(h->mechd, S, dcen, 2, 0);
            h->launches += 2;
        } else {
            sp_calc_xi_kernel<<<(ntraj + 63) / 64, 64, 0, s>>>(h->mechd, S, dcen, 1, 1);
            sp_add_bias<<<gel, 256, 0, s>>>(S, C.g);
            h->launches += 3;
        }
    }
    SplitArgs A{};
    A.ntraj = ntraj;
    A.natoms = na;
    A.nbeads = nb;
    A.beta = h->beta;
    A.mass = h->d_mass;
    A.at_move = h->d_atmove;
    A.p = C.p;
    A.seed = h->seed;
    A.traj_id = C.tid;
    A.event = C.event;
    sp_andersen<<<gel, 256, 0, s>>>(A);
    sp_bump_event<<<(ntraj + 127) / 128, 128, 0, s>>>(C.event, ntraj);
    h->launches += 2;
    if (h->thermostat == 2 && C.nhc) {
        sp_nhc_init<<<(ntraj + 127) / 128, 128, 0, s>>>(ntraj, C.nhc, h->kelvin, h->nose_q, nfree_of(h));
        h->launches++;
    }
    CK(cudaGetLastError());
    return CRCL_OK;
}

// recrossing children on the split path (recross.f90:515-628 / recross_serial.f90:172-229): the same
// sequence as recross_kernel, composed from the split kernels
static int recross_split(crcl_handle h, const double* d_q_parents, int nparent, int pair0, int npairs, int child_evol,
                         double xi_ideal, double* d_kappa_num, double* d_kappa_denom, int* d_status)
{
    if (!h->mechd_valid) return fail(h, CRCL_ESTATE, "crcl_set_mechanism has not been called");
    int rc;
    if ((rc = ensure_split_tables(h)) || (rc = ensure_fker(h))) return rc;
    const int ntraj = 2 * npairs, na = h->natoms, nb = h->nbeads, nc = 3 * na;
    if (ntraj == 0) {
        CK(cudaMemsetAsync(d_kappa_num, 0, (size_t)child_evol * sizeof(double), h->stream));
        CK(cudaMemsetAsync(d_kappa_denom, 0, sizeof(double), h->stream));
        return CRCL_OK;
    }
    const size_t per = (size_t)nb * nc, n = per * ntraj;
    double *dq, *dp, *dg, *ddxi, *dep, *dw, *dcen, *dV;
    unsigned char* dth;
    int* dst;
    uint32_t* dtid;
    if ((rc = scratch(h, 22, n, &dq)) || (rc = scratch(h, 1, n, &dg)) || (rc = scratch(h, 2, n, &dp)) ||
        (rc = scratch(h, 3, (size_t)ntraj * nc, &ddxi)) || (rc = scratch(h, 23, (size_t)ntraj * 2, &dep)) ||
        (rc = scratch(h, 6, (size_t)ntraj * 2, &dtid)) ||
        (rc = scratch(h, 24, (size_t)ntraj * (child_evol > 0 ? child_evol : 1), &dth)) ||
        (rc = scratch(h, 25, (size_t)ntraj * 2 + 2, &dw)) || (rc = scratch(h, 26, (size_t)ntraj + 1, &dst)) ||
        (rc = scratch(h, 8, (size_t)ntraj * nc, &dcen)) || (rc = scratch(h, 9, (size_t)ntraj * nb, &dV)))
        return rc;
    cudaStream_t s = h->stream;
    int* status = d_status ? d_status : dst;
    CK(cudaMemsetAsync(status, 0, ntraj * sizeof(int), s));
    SplitCall C;
    C.ntraj = ntraj;
    C.q = dq;
    C.p = dp;
    C.g = dg;
    C.dxi = ddxi;
    C.epot = dep;
    C.xi_real = dep + ntraj;
    C.status = status;
    C.tid = dtid;
    C.event = dtid + ntraj;
    C.xi_ideal_s = xi_ideal;
    C.theta = dth;
    SpTraj S{};
    if ((rc = split_traj_scratch(h, C, S))) return rc;
    const dim3 gel((unsigned)(((per + 255) / 256) * ntraj));
    sp_recross_init<<<gel, 256, 0, s>>>(ntraj, na, nb, nparent, pair0, d_q_parents, dq, dtid, dtid + ntraj);
    SplitArgs A{};
    A.ntraj = ntraj;
    A.natoms = na;
    A.nbeads = nb;
    A.beta = h->beta;
    A.mass = h->d_mass;
    A.at_move = h->d_atmove;
    A.p = dp;
    A.seed = h->seed;
    A.traj_id = dtid;
    A.event = dtid + ntraj;
    sp_andersen<<<gel, 256, 0, s>>>(A);                       // the draw of recross_serial.f90:151
    sp_bump_event<<<(ntraj + 127) / 128, 128, 0, s>>>(dtid + ntraj, ntraj);
    sp_flip_odd<<<gel, 256, 0, s>>>(ntraj, per, dp);           // child 2: p = -p_save (:160)
    sp_centroid<<<(unsigned)(((size_t)ntraj * nc + 255) / 256), 256, 0, s>>>(ntraj, na, nb, dq, dcen);
    sp_calc_xi_kernel<<<(ntraj + 63) / 64, 64, 0, s>>>(h->mechd, S, dcen, 2, 0);   // :163-164
    h->launches += 6;
    if ((rc = split_forces(h, ntraj * nb, dq, dg, dV))) return rc;                  // :165-169
    sp_recross_weights<<<ntraj, 256, 0, s>>>(S, dp, dw, dw + ntraj);                // :170-190
    h->launches++;
    CK(cudaGetLastError());
    const int th_save = h->thermostat, as_save = h->andersen_step;
    h->thermostat = 0;                                          // no thermostat for the children (:131-135)
    h->andersen_step = 0;
    rc = verlet_split(h, C, child_evol, 0, 2);
    h->thermostat = th_save;
    h->andersen_step = as_save;
    if (rc) return rc;
    reduce_kappa_kernel<<<child_evol + 1, 256, 0, s>>>(dth, dw, dw + ntraj, ntraj, child_evol, d_kappa_num,
                                                       d_kappa_denom);
    h->launches++;
    CK(cudaGetLastError());
    return CRCL_OK;
}

// ---- C-ABI -----------------------------------------------------------------------------------
extern "C" {

int crcl_create(crcl_handle* out, int device, int natoms, int nbeads, const double* mass,
                const int* at_move, double beta, double dt, int pes_id)
{
    if (!out || !mass || natoms <= 0 || nbeads <= 0) return CRCL_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return CRCL_ENODEV;
    if (device < 0 || device >= ndev) return CRCL_EINVAL;
    if (pes_id != CRCL_PES_HOSTCB && pes_id != CRCL_PES_NONE && pes_id != CRCL_PES_QMDFF &&
        pes_id != CRCL_PES_DGEVB && pes_id != CRCL_PES_WATER) {
        const int n = pes_natoms(pes_id);
        if (n < 0 || n != natoms) return CRCL_EINVAL;
    }
    crcl_handle h = new crcl_handle_s();
    h->device = device;
    h->natoms = natoms;
    h->nbeads = nbeads;
    h->pes = pes_id;
    h->beta = beta;
    h->dt = dt;
    h->mass.assign(mass, mass + natoms);
    h->at_move.assign(natoms, 1);
    if (at_move)
        for (int i = 0; i < natoms; i++) h->at_move[i] = at_move[i] ? 1 : 0;
    h->mech.valid = 0;
    bool ok = cudaSetDevice(device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 2 * crcl_handle_s::NEV; i++) ok = cudaEventCreate(&h->evs[i]) == cudaSuccess;
    if (!ok) {
        delete h;
        return CRCL_ECUDA;
    }
    h->ev0 = h->evs[0];
    h->ev1 = h->evs[1];
    h->stream = h->own_stream;
    h->timed = true;
    *out = h;
    return CRCL_OK;
}

int crcl_destroy(crcl_handle h)
{
    if (!h) return CRCL_EINVAL;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    crcl_comm_destroy(h);
    for (auto& s : h->scratch)
        if (s) cudaFree(s);
    if (h->d_fker) cudaFree(h->d_fker);
    qmdff_free(h->qmdff);
    qmdff_free(h->qmdff2);
    dgevb_free(h->dgevb);
    ewald_free(h->ewald);
    water_free(h->water);
    if (h->d_mass) cudaFree(h->d_mass);
    if (h->d_atmove) cudaFree(h->d_atmove);
    if (h->d_frag) cudaFree(h->d_frag);
    if (h->d_wfrag) cudaFree(h->d_wfrag);
    for (auto& e : h->evs)
        if (e) cudaEventDestroy(e);
    cudaStreamDestroy(h->own_stream);
    delete h;
    return CRCL_OK;
}

const char* crcl_last_error(crcl_handle h) { return h ? h->err.c_str() : "null handle"; }

int crcl_set_stream(crcl_handle h, void* s)
{
    if (!h) return CRCL_EINVAL;
    h->stream = (cudaStream_t)s;   // used as given; NULL is CUDA's (legacy) default stream
    return CRCL_OK;
}

int crcl_synchronize(crcl_handle h)
{
    if (!h) return CRCL_EINVAL;
    CK(cudaStreamSynchronize(h->stream));
    return CRCL_OK;
}

int crcl_set_beta_dt(crcl_handle h, double beta, double dt)
{
    if (!h) return CRCL_EINVAL;
    if (beta != h->beta || dt != h->dt) h->fker_dirty = true;
    h->beta = beta;
    h->dt = dt;
    return CRCL_OK;
}

int crcl_set_transform(crcl_handle h, int mode)
{
    if (!h || (mode != CRCL_TRANSFORM_REFERENCE && mode != CRCL_TRANSFORM_EXACT)) return CRCL_EINVAL;
    h->transform = mode;
    return CRCL_OK;
}

int crcl_set_spread_max_beads(crcl_handle h, int max_beads)
{
    if (!h || max_beads < 0) return CRCL_EINVAL;
    h->spread_max = max_beads;
    return CRCL_OK;
}

int crcl_set_host_gradient_cb(crcl_handle h, crcl_host_grad_fn fn, void* user)
{
    if (!h) return CRCL_EINVAL;
    h->cb = fn;
    h->cb_user = user;
    return CRCL_OK;
}

int crcl_set_mechanism(crcl_handle h, int form_num, const int* bond_form, int break_num,
                       const int* bond_break, const double* form_ref, const double* break_ref,
                       int sum_reacs, const int* n_reac, const int* at_reac, double R_inf)
{
    if (!h) return CRCL_EINVAL;
    if (!bond_form && form_num > 0) return CRCL_EINVAL;
    // generic-size tables for the split path (any natoms; <= 8 bonds of each kind, <= 4 fragments)
    if (form_num < 0 || break_num < 0 || form_num > 8 || break_num > 8 || sum_reacs < 2 || sum_reacs > 4)
        return fail(h, CRCL_ENOSUP, "mechanism limits: <= 8 forming / breaking bonds, 2..4 reactant fragments");
    MechDev& D = h->mechd;
    D = MechDev();
    D.form_num = form_num;
    D.break_num = break_num;
    D.sum_reacs = sum_reacs;
    for (int i = 0; i < form_num; i++) {
        D.bf[i][0] = bond_form[2 * i] - 1;
        D.bf[i][1] = bond_form[2 * i + 1] - 1;
        D.fref[i] = form_ref[i];
        if (D.bf[i][0] < 0 || D.bf[i][0] >= h->natoms || D.bf[i][1] < 0 || D.bf[i][1] >= h->natoms)
            return fail(h, CRCL_EINVAL, "mechanism atom index out of range");
    }
    for (int i = 0; i < break_num; i++) {
        D.bb[i][0] = bond_break[2 * i] - 1;
        D.bb[i][1] = bond_break[2 * i + 1] - 1;
        D.bref[i] = break_ref[i];
        if (D.bb[i][0] < 0 || D.bb[i][0] >= h->natoms || D.bb[i][1] < 0 || D.bb[i][1] >= h->natoms)
            return fail(h, CRCL_EINVAL, "mechanism atom index out of range");
    }
    h->frag_h.assign(h->natoms, -1);
    h->wfrag_h.assign(h->natoms, 0.0);
    {
        int off = 0;
        double mr[4] = {0, 0, 0, 0};
        for (int k = 0; k < sum_reacs; k++) {
            for (int i = 0; i < n_reac[k]; i++) {
                const int a = at_reac[off + i] - 1;
                if (a < 0 || a >= h->natoms) return fail(h, CRCL_EINVAL, "mechanism atom index out of range");
                h->frag_h[a] = k;
                mr[k] += h->mass[a];
            }
            off += n_reac[k];
        }
        for (int a = 0; a < h->natoms; a++)
            if (h->frag_h[a] >= 0) h->wfrag_h[a] = h->mass[a] / mr[h->frag_h[a]];
    }
    D.R_inf = R_inf;
    h->mechd_valid = true;
    CK(cudaSetDevice(h->device));
    int rc2 = upload_mechd(h);
    if (rc2) return rc2;
    // in-register tables of the fused kernels (small systems only)
    h->mech.valid = 0;
    if (h->natoms <= XI_MAXAT && form_num <= XI_MAXBOND && break_num <= XI_MAXBOND) {
        const int rc = build_mech(h->mech, h->natoms, h->mass.data(), form_num, bond_form, break_num, bond_break,
                                  form_ref, break_ref, sum_reacs, n_reac, at_reac, R_inf);
        if (rc == -2) return fail(h, CRCL_EINVAL, "mechanism atom index out of range");
    }
    return CRCL_OK;
}

// shared by the two entry points below: bond lists of the unimolecular mechanisms, no fragments
static int set_bond_lists(crcl_handle h, int form_num, const int* bond_form, int break_num, const int* bond_break,
                          const double* form_ref, const double* break_ref)
{
    if (form_num < 0 || break_num < 0 || form_num > 8 || break_num > 8)
        return fail(h, CRCL_ENOSUP, "mechanism limits: <= 8 forming / breaking bonds");
    if ((form_num > 0 && (!bond_form || !form_ref)) || (break_num > 0 && (!bond_break || !break_ref))) return CRCL_EINVAL;
    MechDev& D = h->mechd;
    D = MechDev();
    D.form_num = form_num;
    D.break_num = break_num;
    for (int i = 0; i < form_num; i++) {
        D.bf[i][0] = bond_form[2 * i] - 1;
        D.bf[i][1] = bond_form[2 * i + 1] - 1;
        D.fref[i] = form_ref[i];
        if (D.bf[i][0] < 0 || D.bf[i][0] >= h->natoms || D.bf[i][1] < 0 || D.bf[i][1] >= h->natoms)
            return fail(h, CRCL_EINVAL, "mechanism atom index out of range");
    }
    for (int i = 0; i < break_num; i++) {
        D.bb[i][0] = bond_break[2 * i] - 1;
        D.bb[i][1] = bond_break[2 * i + 1] - 1;
        D.bref[i] = break_ref[i];
        if (D.bb[i][0] < 0 || D.bb[i][0] >= h->natoms || D.bb[i][1] < 0 || D.bb[i][1] >= h->natoms)
            return fail(h, CRCL_EINVAL, "mechanism atom index out of range");
    }
    h->frag_h.assign(h->natoms, -1);
    h->wfrag_h.assign(h->natoms, 0.0);
    h->mech = Mech();
    h->mech.form_num = form_num;
    h->mech.break_num = break_num;
    if (h->natoms <= XI_MAXAT && form_num <= XI_MAXBOND && break_num <= XI_MAXBOND) {
        for (int i = 0; i < form_num; i++) {
            h->mech.bf[i][0] = D.bf[i][0];
            h->mech.bf[i][1] = D.bf[i][1];
            h->mech.fref[i] = D.fref[i];
        }
        for (int i = 0; i < break_num; i++) {
            h->mech.bb[i][0] = D.bb[i][0];
            h->mech.bb[i][1] = D.bb[i][1];
            h->mech.bref[i] = D.bref[i];
        }
        for (int a = 0; a < XI_MAXAT; a++) h->mech.frag[a] = -1;
        h->mech.inv_form = form_num ? 1.0 / form_num : 0.0;
        h->mech.inv_break = break_num ? 1.0 / break_num : 0.0;
    }
    return CRCL_OK;
}

int crcl_set_mechanism_unimol(crcl_handle h, int form_num, const int* bond_form, int break_num, const int* bond_break,
                              const double* form_ref, const double* break_ref, const double* form_reac,
                              const double* break_reac)
{
    if (!h) return CRCL_EINVAL;
    if ((form_num > 0 && !form_reac) || (break_num > 0 && !break_reac) || form_num + break_num <= 0) return CRCL_EINVAL;
    int rc = set_bond_lists(h, form_num, bond_form, break_num, bond_break, form_ref, break_ref);
    if (rc) return rc;
    MechDev& D = h->mechd;
    D.type = 1;
    for (int i = 0; i < form_num; i++) D.freac[i] = form_reac[i];
    for (int i = 0; i < break_num; i++) D.breac[i] = break_reac[i];
    h->mechd_valid = true;
    CK(cudaSetDevice(h->device));
    if ((rc = upload_mechd(h))) return rc;
    if (h->natoms <= XI_MAXAT && form_num <= XI_MAXBOND && break_num <= XI_MAXBOND) {
        mech_set_unimol(h->mech, form_reac, break_reac);
        h->mech.valid = 1;
    }
    return CRCL_OK;
}

int crcl_set_mechanism_atom_shift(crcl_handle h, int shift_atom, int shift_coord, double shift_lo, double shift_hi,
                                  double shift2_lo, double shift2_hi)
{
    if (!h) return CRCL_EINVAL;
    int rc = set_bond_lists(h, 0, nullptr, 0, nullptr, nullptr, nullptr);
    if (rc) return rc;
    Mech tmp = Mech();
    if (mech_set_atom_shift(tmp, h->natoms, shift_atom, shift_coord, shift_lo, shift_hi, shift2_lo, shift2_hi))
        return fail(h, CRCL_EINVAL, "ATOM_SHIFT: atom must be 1..natoms and the coordinate code 1..6");
    MechDev& D = h->mechd;
    D.type = 2;
    D.shift_atom = tmp.shift_atom;
    D.shift_c1 = tmp.shift_c1;
    D.shift_c2 = tmp.shift_c2;
    D.shift_lo = shift_lo;
    D.shift_hi = shift_hi;
    D.shift2_lo = shift2_lo;
    D.shift2_hi = shift2_hi;
    h->mechd_valid = true;
    CK(cudaSetDevice(h->device));
    if ((rc = upload_mechd(h))) return rc;
    if (h->natoms <= XI_MAXAT)
        mech_set_atom_shift(h->mech, h->natoms, shift_atom, shift_coord, shift_lo, shift_hi, shift2_lo, shift2_hi);
    return CRCL_OK;
}

int crcl_set_qmdff(crcl_handle h, const crcl_qmdff_tables* T)
{
    if (!h || !T) return CRCL_EINVAL;
    if (T->n != h->natoms) return fail(h, CRCL_EINVAL, "crcl_set_qmdff: T->n differs from the handle's natoms");
    CK(cudaSetDevice(h->device));
    qmdff_free(h->qmdff);
    h->qmdff = nullptr;
    const char* msg = "";
    const int rc = qmdff_upload(T, &h->qmdff, &msg);
    if (rc) return fail(h, rc, msg);
    // `periodic` and boxlen_* are one global state in the reference (pbc_mod): the integrator wraps when the PES is periodic
    return crcl_set_box(h, T->periodic, T->box);
}

int crcl_set_qmdff2(crcl_handle h, const crcl_qmdff_tables* T)
{
    if (!h || !T) return CRCL_EINVAL;
    if (T->n != h->natoms) return fail(h, CRCL_EINVAL, "crcl_set_qmdff2: T->n differs from the handle's natoms");
    CK(cudaSetDevice(h->device));
    qmdff_free(h->qmdff2);
    h->qmdff2 = nullptr;
    const char* msg = "";
    const int rc = qmdff_upload(T, &h->qmdff2, &msg, true);
    if (rc) return fail(h, rc, msg);
    return CRCL_OK;
}

int crcl_set_water(crcl_handle h, const crcl_water_params* P)
{
    if (!h || !P) return CRCL_EINVAL;
    if (P->n != h->natoms) return fail(h, CRCL_EINVAL, "crcl_set_water: P->n differs from the handle's natoms");
    CK(cudaSetDevice(h->device));
    water_free(h->water);
    h->water = nullptr;
    const char* msg = "";
    const int rc = water_upload(P, &h->water, &msg);
    if (rc) return fail(h, rc, msg);
    return crcl_set_box(h, P->periodic, P->box);
}

int crcl_set_dgevb(crcl_handle h, const crcl_dgevb_params* P)
{
    if (!h || !P || !P->coord_def || !P->point_int || !P->alph || !P->b_vec) return CRCL_EINVAL;
    CK(cudaSetDevice(h->device));
    dgevb_free(h->dgevb);
    h->dgevb = nullptr;
    const char* msg = "";
    const int rc = dgevb_upload(P, h->natoms, &h->dgevb, &msg);
    if (rc) return fail(h, rc, msg);
    return CRCL_OK;
}

int crcl_set_ewald(crcl_handle h, const crcl_ewald_params* P)
{
    if (!h || !P) return CRCL_EINVAL;
    CK(cudaSetDevice(h->device));
    ewald_free(h->ewald);
    h->ewald = nullptr;
    const char* msg = "";
    const int rc = ewald_upload(P, &h->ewald, &msg);
    if (rc) return fail(h, rc, msg);
    return CRCL_OK;
}

int crcl_ewald_recip(crcl_handle h, int n, int nimg, const double* xyz, const double* q, double* energy, double* grad)
{
    if (!h || !xyz || !q || !energy || !grad || n < 0 || nimg < 0) return CRCL_EINVAL;
    if (!h->ewald) return fail(h, CRCL_ESTATE, "crcl_set_ewald has not been called");
    if (n == 0 || nimg == 0) return CRCL_OK;
    CK(cudaSetDevice(h->device));
    const size_t nx = (size_t)nimg * 3 * n;
    double *dx, *dg, *de, *dq;
    int rc;
    if ((rc = scratch(h, 0, nx, &dx)) || (rc = scratch(h, 1, nx, &dg)) || (rc = scratch(h, 2, (size_t)nimg, &de)) ||
        (rc = scratch(h, 3, (size_t)n, &dq)))
        return rc;
    CK(cudaMemcpyAsync(dx, xyz, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dq, q, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (h->timed && !h->capturing) {
        next_event_pair(h);
        cudaEventRecord(h->ev0, h->stream);
    }
    const char* msg = "";
    rc = ewald_recip(h->ewald, n, nimg, dx, dq, de, dg, h->stream, &h->launches, &msg);
    if (h->timed && !h->capturing) cudaEventRecord(h->ev1, h->stream);
    if (rc) return fail(h, rc, msg);
    CK(cudaMemcpyAsync(energy, de, (size_t)nimg * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(grad, dg, nx * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return CRCL_OK;
}

int crcl_comm_unique_id(void* id_out)
{
    if (!id_out) return CRCL_EINVAL;
    static_assert(sizeof(ncclUniqueId) == CRCL_UNIQUE_ID_BYTES, "CRCL_UNIQUE_ID_BYTES must be sizeof(ncclUniqueId)");
    NcclApi* N = nccl_api(nullptr);
    if (!N) return CRCL_ESTATE;
    ncclUniqueId id;
    if (N->GetUniqueId(&id) != ncclSuccess) return CRCL_ECUDA;
    memcpy(id_out, &id, sizeof(id));
    return CRCL_OK;
}

int crcl_comm_init(crcl_handle h, int nranks, int rank, const void* unique_id)
{
    if (!h || !unique_id || nranks < 1 || rank < 0 || rank >= nranks) return CRCL_EINVAL;
    if (h->comm) return fail(h, CRCL_ESTATE, "crcl_comm_init: the handle already has a communicator");
    std::string err;
    NcclApi* N = nccl_api(&err);
    if (!N) return fail(h, CRCL_ESTATE, err.c_str());
    CK(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    Comm* c = new Comm();
    c->nranks = nranks;
    c->rank = rank;
    const ncclResult_t r = N->CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        delete c;
        h->err = std::string("ncclCommInitRank: ") + N->GetErrorString(r);
        return CRCL_ECUDA;
    }
    h->comm = c;
    return CRCL_OK;
}

int crcl_comm_destroy(crcl_handle h)
{
    if (!h) return CRCL_EINVAL;
    if (h->comm) {
        NcclApi* N = nccl_api(nullptr);
        cudaStreamSynchronize(h->stream);
        if (N && h->comm->comm) N->CommDestroy(h->comm->comm);
        delete h->comm;
        h->comm = nullptr;
    }
    return CRCL_OK;
}

int crcl_comm_info(crcl_handle h, int* nranks, int* rank, int* nccl_version)
{
    if (!h) return CRCL_EINVAL;
    if (nranks) *nranks = h->comm ? h->comm->nranks : 1;
    if (rank) *rank = h->comm ? h->comm->rank : 0;
    if (nccl_version) {
        *nccl_version = 0;
        NcclApi* N = nccl_api(nullptr);
        if (N) N->GetVersion(nccl_version);
    }
    return CRCL_OK;
}

int crcl_set_box(crcl_handle h, int periodic, const double* boxlen)
{
    if (!h || (periodic && !boxlen)) return CRCL_EINVAL;
    if (periodic && !(boxlen[0] > 0 && boxlen[1] > 0 && boxlen[2] > 0)) return fail(h, CRCL_EINVAL, "box lengths must be positive");
    h->periodic = periodic ? 1 : 0;
    for (int d = 0; d < 3; d++) h->box[d] = periodic ? boxlen[d] : 0.0;
    return CRCL_OK;
}

int crcl_set_rpmd_check(crcl_handle h, int on, double energy_ts, double energy_tol, double xi_tol)
{
    if (!h) return CRCL_EINVAL;
    h->chk_on = on ? 1 : 0;
    h->chk_energy_ts = energy_ts;
    h->chk_energy_tol = energy_tol;
    h->chk_xi_tol = xi_tol;
    return CRCL_OK;
}

int crcl_set_path(crcl_handle h, int path)
{
    if (!h || path < CRCL_PATH_AUTO || path > CRCL_PATH_SPLIT) return CRCL_EINVAL;
    h->path = path;
    return CRCL_OK;
}

int crcl_set_graph(crcl_handle h, int on)
{
    if (!h) return CRCL_EINVAL;
    h->use_graph = on != 0;
    return CRCL_OK;
}

int crcl_set_thermostat(crcl_handle h, int thermostat, int andersen_step, double kelvin, double nose_q)
{
    if (!h || thermostat < 0 || thermostat > 2) return CRCL_EINVAL;
    h->thermostat = thermostat;
    h->andersen_step = andersen_step;
    h->kelvin = kelvin;
    h->nose_q = nose_q;
    return CRCL_OK;
}

int crcl_set_seed(crcl_handle h, uint64_t seed)
{
    if (!h) return CRCL_EINVAL;
    h->seed = seed;
    return CRCL_OK;
}

// ---- PES seam ---------------------------------------------------------------------------------
int crcl_egrad_dev(crcl_handle h, int pes_id, const double* d_q, int natoms, int nimg, double* d_V,
                   double* d_dVdq, int* d_info)
{
    if (!h || !d_q || !d_V || !d_dVdq || nimg < 0) return CRCL_EINVAL;
    if (pes_id == CRCL_PES_DGEVB) {
        if (!h->qmdff || !h->qmdff2 || !h->dgevb)
            return fail(h, CRCL_ESTATE, "DG-EVB needs crcl_set_qmdff, crcl_set_qmdff2 and crcl_set_dgevb");
        if (h->qmdff->n != natoms) return fail(h, CRCL_EINVAL, "natoms does not match the QMDFF tables");
        if (nimg == 0) return CRCL_OK;
        CK(cudaSetDevice(h->device));
        if (d_info) CK(cudaMemsetAsync(d_info, 0, sizeof(int), h->stream));
        const size_t n = (size_t)nimg * 3 * natoms;
        double *g1, *g2, *v12;
        int rc;
        if ((rc = scratch(h, 12, n, &g1)) || (rc = scratch(h, 13, n, &g2)) || (rc = scratch(h, 14, (size_t)2 * nimg, &v12)))
            return rc;
        if (h->timed && !h->capturing) {
            next_event_pair(h);
            cudaEventRecord(h->ev0, h->stream);
        }
        cudaError_t e = qmdff_egrad(h->qmdff, d_q, nimg, v12, g1, h->stream, &h->launches);
        if (e == cudaSuccess) e = qmdff_egrad(h->qmdff2, d_q, nimg, v12 + nimg, g2, h->stream, &h->launches);
        if (e == cudaSuccess)
            e = dgevb_mix(h->dgevb, natoms, d_q, nimg, v12, g1, v12 + nimg, g2, d_V, d_dVdq, h->stream);
        h->launches++;
        if (h->timed && !h->capturing) cudaEventRecord(h->ev1, h->stream);
        if (e != cudaSuccess) {
            h->err = std::string("dg-evb kernels: ") + cudaGetErrorString(e);
            return CRCL_ECUDA;
        }
        return CRCL_OK;
    }
    if (pes_id == CRCL_PES_WATER) {
        if (!h->water) return fail(h, CRCL_ESTATE, "crcl_set_water has not been called");
        if (h->water->n != natoms) return fail(h, CRCL_EINVAL, "natoms does not match the water box");
        if (nimg == 0) return CRCL_OK;
        CK(cudaSetDevice(h->device));
        if (d_info) CK(cudaMemsetAsync(d_info, 0, sizeof(int), h->stream));
        if (h->timed && !h->capturing) {
            next_event_pair(h);
            cudaEventRecord(h->ev0, h->stream);
        }
        cudaError_t e = water_egrad(h->water, d_q, nimg, d_V, d_dVdq, h->stream, &h->launches);
        if (h->timed && !h->capturing) cudaEventRecord(h->ev1, h->stream);
        if (e != cudaSuccess) {
            h->err = std::string("water kernels: ") + cudaGetErrorString(e);
            return CRCL_ECUDA;
        }
        return CRCL_OK;
    }
    if (pes_id == CRCL_PES_QMDFF) {
        if (!h->qmdff) return fail(h, CRCL_ESTATE, "crcl_set_qmdff has not been called");
        if (h->qmdff->n != natoms) return fail(h, CRCL_EINVAL, "natoms does not match the QMDFF tables");
        if (nimg == 0) return CRCL_OK;
        CK(cudaSetDevice(h->device));
        if (d_info) CK(cudaMemsetAsync(d_info, 0, sizeof(int), h->stream));
        if (h->timed && !h->capturing) {
            next_event_pair(h);
            cudaEventRecord(h->ev0, h->stream);
        }
        cudaError_t e = qmdff_egrad(h->qmdff, d_q, nimg, d_V, d_dVdq, h->stream, &h->launches);
        if (h->timed && !h->capturing) cudaEventRecord(h->ev1, h->stream);
        if (e != cudaSuccess) {
            h->err = std::string("qmdff kernels: ") + cudaGetErrorString(e);
            return CRCL_ECUDA;
        }
        return CRCL_OK;
    }
    if (pes_natoms(pes_id) != natoms) return fail(h, CRCL_EINVAL, "natoms does not match the PES");
    if (nimg == 0) return CRCL_OK;
    CK(cudaSetDevice(h->device));
    const int tpb = EGRAD_TPB, grid = (nimg + tpb - 1) / tpb;
    if (d_info) CK(cudaMemsetAsync(d_info, 0, sizeof(int), h->stream));
    if (h->timed && !h->capturing) {
        next_event_pair(h);
        cudaEventRecord(h->ev0, h->stream);
    }
    switch (pes_id) {
    case CRCL_PES_H3: egrad_kernel<PesH3><<<grid, tpb, 0, h->stream>>>(d_q, nimg, d_V, d_dVdq, d_info); break;
    case CRCL_PES_OH3: egrad_kernel<PesOH3><<<grid, tpb, 0, h->stream>>>(d_q, nimg, d_V, d_dVdq, d_info); break;
    case CRCL_PES_CH4H: egrad_kernel<PesCH4H><<<grid, tpb, 0, h->stream>>>(d_q, nimg, d_V, d_dVdq, d_info); break;
    case CRCL_PES_BRH2: egrad_kernel_mb3<PesBrH2><<<grid, tpb, 0, h->stream>>>(d_q, nimg, d_V, d_dVdq, d_info); break;
    case CRCL_PES_O3: egrad_kernel<PesO3><<<grid, tpb, 0, h->stream>>>(d_q, nimg, d_V, d_dVdq, d_info); break;
    case CRCL_PES_CH4OH: egrad_kernel<PesCH4OH><<<grid, tpb, 0, h->stream>>>(d_q, nimg, d_V, d_dVdq, d_info); break;
    case CRCL_PES_GEH4OH: egrad_kernel<PesGeH4OH><<<grid, tpb, 0, h->stream>>>(d_q, nimg, d_V, d_dVdq, d_info); break;
    case CRCL_PES_CH4CN: egrad_kernel<PesCH4CN><<<grid, tpb, 0, h->stream>>>(d_q, nimg, d_V, d_dVdq, d_info); break;
    case CRCL_PES_CLNH3: egrad_kernel<PesClNH3><<<grid, tpb, 0, h->stream>>>(d_q, nimg, d_V, d_dVdq, d_info); break;
    case CRCL_PES_NH3OH: egrad_kernel<PesNH3OH><<<grid, tpb, 0, h->stream>>>(d_q, nimg, d_V, d_dVdq, d_info); break;
    case CRCL_PES_H2CO: egrad_kernel<PesH2CO><<<grid, tpb, 0, h->stream>>>(d_q, nimg, d_V, d_dVdq, d_info); break;
    default: return fail(h, CRCL_ENOSUP, "unknown PES id");
    }
    if (h->timed && !h->capturing) cudaEventRecord(h->ev1, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    return CRCL_OK;
}

int crcl_egrad(crcl_handle h, int pes_id, const double* q, int natoms, int nimg, double* V,
               double* dVdq, int* info)
{
    if (!h || !q || !V || !dVdq || nimg < 0) return CRCL_EINVAL;
    if (info) *info = 0;
    if (nimg == 0) return CRCL_OK;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)nimg * 3 * natoms;
    double *dq, *dg, *dV;
    int* di;
    int rc;
    if ((rc = scratch(h, 0, n, &dq)) || (rc = scratch(h, 1, n, &dg)) || (rc = scratch(h, 2, (size_t)nimg, &dV)) ||
        (rc = scratch(h, 3, (size_t)1, &di)))
        return rc;
    CK(cudaMemcpyAsync(dq, q, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if ((rc = crcl_egrad_dev(h, pes_id, dq, natoms, nimg, dV, dg, di))) return rc;
    CK(cudaMemcpyAsync(V, dV, (size_t)nimg * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(dVdq, dg, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    int hi = 0;
    CK(cudaMemcpyAsync(&hi, di, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (info) *info = hi;
    return CRCL_OK;
}

// ---- integrator seam --------------------------------------------------------------------------
static int check_traj_call(crcl_handle h, int constrain)
{
    if (!h) return CRCL_EINVAL;
    if (constrain < -1 || constrain > 3) return fail(h, CRCL_EINVAL, "constrain must be -1..3");
    if (h->path == CRCL_PATH_FUSED && !fused_ok(h))
        return fail(h, CRCL_ENOSUP, "fused trajectory kernels need <= 8 atoms, a device PES and power-of-two nbeads <= 128");
    if (constrain >= 0 && !use_split(h) && !h->mech.valid)
        return fail(h, CRCL_ESTATE, "crcl_set_mechanism has not been called");
    return CRCL_OK;
}

int crcl_verlet(crcl_handle h, int ntraj, int nsteps, int istep0, int constrain, const double* xi_ideal,
                const double* k_force, double* q, double* p, double* derivs, double* epot,
                double* xi_real, double* dxi, int* status, const uint32_t* traj_id, uint32_t* event0)
{
    int rc = check_traj_call(h, constrain);
    if (rc) return rc;
    if (!q || !p || !derivs || ntraj < 0 || nsteps < 0) return CRCL_EINVAL;
    if (ntraj == 0) return CRCL_OK;
    CK(cudaSetDevice(h->device));
    if ((rc = ensure_fker(h))) return rc;
    const size_t n = (size_t)ntraj * h->nbeads * h->natoms * 3, nd = (size_t)ntraj * h->natoms * 3;
    double *dq, *dp, *dg, *ddxi, *dep, *dxr, *dxid, *dkf, *dnhc;
    int* dst;
    uint32_t *dtid, *dev;
    if ((rc = scratch(h, 0, n, &dq)) || (rc = scratch(h, 1, n, &dg)) || (rc = scratch(h, 2, n, &dp)) ||
        (rc = scratch(h, 3, nd, &ddxi)) || (rc = scratch(h, 4, (size_t)ntraj * 4, &dep)) ||
        (rc = scratch(h, 5, (size_t)ntraj, &dst)) || (rc = scratch(h, 6, (size_t)ntraj * 2, &dtid)) ||
        (rc = scratch(h, 7, (size_t)ntraj * 8, &dnhc)))
        return rc;
    dxr = dep + ntraj;
    dxid = dep + 2 * ntraj;
    dkf = dep + 3 * ntraj;
    dev = dtid + ntraj;
    cudaStream_t s = h->stream;
    CK(cudaMemsetAsync(dep, 0, (size_t)ntraj * 2 * sizeof(double), s));   // epot, xi_real: 0 unless the mode sets them
    CK(cudaMemcpyAsync(dq, q, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dp, p, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dg, derivs, n * sizeof(double), cudaMemcpyHostToDevice, s));
    if (dxi) CK(cudaMemcpyAsync(ddxi, dxi, nd * sizeof(double), cudaMemcpyHostToDevice, s));
    else CK(cudaMemsetAsync(ddxi, 0, nd * sizeof(double), s));
    if (xi_ideal) CK(cudaMemcpyAsync(dxid, xi_ideal, ntraj * sizeof(double), cudaMemcpyHostToDevice, s));
    if (k_force) CK(cudaMemcpyAsync(dkf, k_force, ntraj * sizeof(double), cudaMemcpyHostToDevice, s));
    if (traj_id) CK(cudaMemcpyAsync(dtid, traj_id, ntraj * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    if (event0) CK(cudaMemcpyAsync(dev, event0, ntraj * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    else CK(cudaMemsetAsync(dev, 0, ntraj * sizeof(uint32_t), s));
    if (status) CK(cudaMemcpyAsync(dst, status, ntraj * sizeof(int), cudaMemcpyHostToDevice, s));
    else CK(cudaMemsetAsync(dst, 0, ntraj * sizeof(int), s));
    TrajArgs A;
    fill_args(h, A);
    A.ntraj = ntraj;
    A.nsteps = nsteps;
    A.istep0 = istep0;
    A.constrain = constrain;
    A.xi_ideal = xi_ideal ? dxid : nullptr;
    A.k_force = k_force ? dkf : nullptr;
    A.q = dq;
    A.p = dp;
    A.g = dg;
    A.dxi = ddxi;
    A.epot = dep;
    A.xi_real = dxr;
    A.status = dst;
    A.nhc = dnhc;  // NHC chain state persists on the device between calls of one handle
    A.traj_id = traj_id ? dtid : nullptr;
    A.event = dev;
    if (use_split(h)) {
        if (!traj_id) {
            std::vector<uint32_t> ids(ntraj);
            for (int t = 0; t < ntraj; t++) ids[t] = (uint32_t)t;
            CK(cudaMemcpyAsync(dtid, ids.data(), ntraj * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));
        }
        SplitCall C;
        C.ntraj = ntraj;
        C.q = dq;
        C.p = dp;
        C.g = dg;
        C.dxi = ddxi;
        C.epot = dep;
        C.xi_real = dxr;
        C.nhc = dnhc;
        C.status = dst;
        C.tid = dtid;
        C.event = dev;
        C.xi_ideal = xi_ideal ? dxid : nullptr;
        C.k_force = k_force ? dkf : nullptr;
        if ((rc = verlet_split(h, C, nsteps, istep0, constrain))) return rc;
    } else if ((rc = launch_traj(h, K_VERLET, A)))
        return rc;
    CK(cudaMemcpyAsync(q, dq, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(p, dp, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(derivs, dg, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (dxi) CK(cudaMemcpyAsync(dxi, ddxi, nd * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (epot) CK(cudaMemcpyAsync(epot, dep, ntraj * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (xi_real) CK(cudaMemcpyAsync(xi_real, dxr, ntraj * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (status) CK(cudaMemcpyAsync(status, dst, ntraj * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (event0) CK(cudaMemcpyAsync(event0, dev, ntraj * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return CRCL_OK;
}

// crcl_verlet on state resident in device memory (asynchronous on the handle's stream, nothing crosses PCIe)
int crcl_verlet_dev(crcl_handle h, int ntraj, int nsteps, int istep0, int constrain, const double* d_xi_ideal,
                    const double* d_k_force, double* d_q, double* d_p, double* d_derivs, double* d_epot,
                    double* d_xi_real, double* d_dxi, int* d_status, const uint32_t* d_traj_id, uint32_t* d_event)
{
    int rc = check_traj_call(h, constrain);
    if (rc) return rc;
    if (!d_q || !d_p || !d_derivs || !d_epot || !d_xi_real || !d_dxi || !d_status || !d_event || ntraj < 0 || nsteps < 0)
        return CRCL_EINVAL;
    if (ntraj == 0) return CRCL_OK;
    CK(cudaSetDevice(h->device));
    if ((rc = ensure_fker(h))) return rc;
    double* dnhc;
    if ((rc = scratch(h, 7, (size_t)ntraj * 8, &dnhc))) return rc;
    if (use_split(h)) {
        uint32_t* dtid = nullptr;
        if (!d_traj_id) {
            if ((rc = scratch(h, 29, (size_t)ntraj, &dtid))) return rc;
            std::vector<uint32_t> ids(ntraj);
            for (int t = 0; t < ntraj; t++) ids[t] = (uint32_t)t;
            CK(cudaMemcpyAsync(dtid, ids.data(), ntraj * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
        SplitCall C;
        C.ntraj = ntraj;
        C.q = d_q;
        C.p = d_p;
        C.g = d_derivs;
        C.dxi = d_dxi;
        C.epot = d_epot;
        C.xi_real = d_xi_real;
        C.nhc = dnhc;
        C.status = d_status;
        C.tid = d_traj_id ? d_traj_id : dtid;
        C.event = d_event;
        C.xi_ideal = d_xi_ideal;
        C.k_force = d_k_force;
        return verlet_split(h, C, nsteps, istep0, constrain);
    }
    TrajArgs A;
    fill_args(h, A);
    A.ntraj = ntraj;
    A.nsteps = nsteps;
    A.istep0 = istep0;
    A.constrain = constrain;
    A.xi_ideal = d_xi_ideal;
    A.k_force = d_k_force;
    A.q = d_q;
    A.p = d_p;
    A.g = d_derivs;
    A.dxi = d_dxi;
    A.epot = d_epot;
    A.xi_real = d_xi_real;
    A.status = d_status;
    A.nhc = dnhc;
    A.traj_id = d_traj_id;
    A.event = d_event;
    return launch_traj(h, K_VERLET, A);
}

int crcl_mdinit(crcl_handle h, int ntraj, int bias_mode, const double* xi_ideal, const double* k_force,
                const double* q, double* p, double* derivs, double* dxi, const uint32_t* traj_id,
                uint32_t* event0)
{
    int rc = check_traj_call(h, bias_mode ? 0 : -1);
    if (rc) return rc;
    if (!q || !p || !derivs || ntraj < 0 || bias_mode < 0 || bias_mode > 2) return CRCL_EINVAL;
    if (ntraj == 0) return CRCL_OK;
    CK(cudaSetDevice(h->device));
    if ((rc = ensure_fker(h))) return rc;
    const size_t n = (size_t)ntraj * h->nbeads * h->natoms * 3, nd = (size_t)ntraj * h->natoms * 3;
    double *dq, *dp, *dg, *ddxi, *dep, *dxid, *dkf, *dnhc;
    uint32_t *dtid, *dev;
    if ((rc = scratch(h, 0, n, &dq)) || (rc = scratch(h, 1, n, &dg)) || (rc = scratch(h, 2, n, &dp)) ||
        (rc = scratch(h, 3, nd, &ddxi)) || (rc = scratch(h, 4, (size_t)ntraj * 4, &dep)) ||
        (rc = scratch(h, 6, (size_t)ntraj * 2, &dtid)) || (rc = scratch(h, 7, (size_t)ntraj * 8, &dnhc)))
        return rc;
    dxid = dep + 2 * ntraj;
    dkf = dep + 3 * ntraj;
    dev = dtid + ntraj;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dq, q, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dp, p, n * sizeof(double), cudaMemcpyHostToDevice, s));
    if (xi_ideal) CK(cudaMemcpyAsync(dxid, xi_ideal, ntraj * sizeof(double), cudaMemcpyHostToDevice, s));
    if (k_force) CK(cudaMemcpyAsync(dkf, k_force, ntraj * sizeof(double), cudaMemcpyHostToDevice, s));
    if (traj_id) CK(cudaMemcpyAsync(dtid, traj_id, ntraj * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    if (event0) CK(cudaMemcpyAsync(dev, event0, ntraj * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    else CK(cudaMemsetAsync(dev, 0, ntraj * sizeof(uint32_t), s));
    TrajArgs A;
    fill_args(h, A);
    A.ntraj = ntraj;
    A.xi_ideal = xi_ideal ? dxid : nullptr;
    A.k_force = k_force ? dkf : nullptr;
    A.q = dq;
    A.p = dp;
    A.g = dg;
    A.dxi = ddxi;
    A.nhc = dnhc;
    A.traj_id = traj_id ? dtid : nullptr;
    A.event = dev;
    if (use_split(h)) {
        if (!traj_id) {
            std::vector<uint32_t> ids(ntraj);
            for (int t = 0; t < ntraj; t++) ids[t] = (uint32_t)t;
            CK(cudaMemcpyAsync(dtid, ids.data(), ntraj * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));
        }
        SplitCall C;
        C.ntraj = ntraj;
        C.q = dq;
        C.p = dp;
        C.g = dg;
        C.dxi = ddxi;
        C.epot = dep;
        C.xi_real = dep + ntraj;
        C.nhc = dnhc;
        C.tid = dtid;
        C.event = dev;
        C.xi_ideal = xi_ideal ? dxid : nullptr;
        C.k_force = k_force ? dkf : nullptr;
        int* dst0;
        if ((rc = scratch(h, 5, (size_t)ntraj, &dst0))) return rc;
        CK(cudaMemsetAsync(dst0, 0, ntraj * sizeof(int), s));
        C.status = dst0;
        if ((rc = mdinit_split(h, C, bias_mode))) return rc;
    } else if ((rc = launch_traj(h, K_MDINIT, A, bias_mode)))
        return rc;
    CK(cudaMemcpyAsync(p, dp, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(derivs, dg, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (dxi && bias_mode) CK(cudaMemcpyAsync(dxi, ddxi, nd * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (event0) CK(cudaMemcpyAsync(event0, dev, ntraj * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return CRCL_OK;
}

int crcl_calc_xi(crcl_handle h, int ncoord, const double* coords, const double* xi_ideal, int mode,
                 double* xi, double* dxi, double* hams)
{
    if (!h || !coords || !xi || !dxi || ncoord < 0 || (mode != 1 && mode != 2)) return CRCL_EINVAL;
    if (!h->mech.valid && !h->mechd_valid) return fail(h, CRCL_ESTATE, "crcl_set_mechanism has not been called");
    if (ncoord == 0) return CRCL_OK;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)ncoord * h->natoms * 3;
    double *dc, *dd, *dh, *dx;
    int rc;
    if ((rc = scratch(h, 0, n, &dc)) || (rc = scratch(h, 1, n, &dd)) || (rc = scratch(h, 2, n, &dh)) ||
        (rc = scratch(h, 4, (size_t)ncoord * 2, &dx)))
        return rc;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dc, coords, n * sizeof(double), cudaMemcpyHostToDevice, s));
    std::vector<double> xid(ncoord, 0.0);
    if (xi_ideal) xid.assign(xi_ideal, xi_ideal + ncoord);
    CK(cudaMemcpyAsync(dx + ncoord, xid.data(), ncoord * sizeof(double), cudaMemcpyHostToDevice, s));
    TrajArgs A;
    fill_args(h, A);
    const int tpb = 64, grid = (ncoord + tpb - 1) / tpb;
    double* hp = hams ? dh : nullptr;
    switch (h->natoms) {
    case 3: calc_xi_kernel<3><<<grid, tpb, 0, s>>>(A, ncoord, dc, dx + ncoord, mode, dx, dd, hp); break;
    case 4: calc_xi_kernel<4><<<grid, tpb, 0, s>>>(A, ncoord, dc, dx + ncoord, mode, dx, dd, hp); break;
    case 5: calc_xi_kernel<5><<<grid, tpb, 0, s>>>(A, ncoord, dc, dx + ncoord, mode, dx, dd, hp); break;
    case 6: calc_xi_kernel<6><<<grid, tpb, 0, s>>>(A, ncoord, dc, dx + ncoord, mode, dx, dd, hp); break;
    case 7: calc_xi_kernel<7><<<grid, tpb, 0, s>>>(A, ncoord, dc, dx + ncoord, mode, dx, dd, hp); break;
    case 8: calc_xi_kernel<8><<<grid, tpb, 0, s>>>(A, ncoord, dc, dx + ncoord, mode, dx, dd, hp); break;
    default: {
        // any number of atoms: the split path's Hessian-free evaluation (split_xi.cuh)
        if (!h->mechd_valid) return fail(h, CRCL_ESTATE, "crcl_set_mechanism has not been called");
        if ((rc = ensure_split_tables(h))) return rc;
        SpTraj S{};
        double* dw;
        if ((rc = scratch(h, 17, (size_t)ncoord * SPX_NWORK * 3 * h->natoms, &dw))) return rc;
        S.ntraj = ncoord;
        S.natoms = h->natoms;
        S.nbeads = 1;
        S.beta = h->beta;
        S.mass = h->d_mass;
        S.xi_ideal = dx + ncoord;
        S.xi_real = dx;
        S.dxi = dd;
        S.hams = dh;
        S.work = dw;
        sp_calc_xi_kernel<<<grid, tpb, 0, s>>>(h->mechd, S, dc, mode, (hams && mode == 1) ? 1 : 0);
    }
    }
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(xi, dx, ncoord * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(dxi, dd, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (hams) CK(cudaMemcpyAsync(hams, dh, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return CRCL_OK;
}

// ---- work-unit seam ---------------------------------------------------------------------------
// all-reduce (sum) over the ranks of the handle's communicator, in place, on the handle's stream
static int comm_allreduce(crcl_handle h, void* a, size_t na, void* b, size_t nb, bool b_is_int)
{
    std::string err;
    NcclApi* N = nccl_api(&err);
    if (!N) return fail(h, CRCL_ESTATE, err.c_str());
    ncclResult_t r = N->GroupStart();
    if (r == ncclSuccess && na) r = N->AllReduce(a, a, na, ncclDouble, ncclSum, h->comm->comm, h->stream);
    if (r == ncclSuccess && nb) r = N->AllReduce(b, b, nb, b_is_int ? ncclInt : ncclDouble, ncclSum, h->comm->comm, h->stream);
    const ncclResult_t r2 = N->GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) {
        h->err = std::string("ncclAllReduce: ") + N->GetErrorString(r);
        return CRCL_ECUDA;
    }
    return CRCL_OK;
}

// the rank-local part of the recrossing work unit: pairs [pair0, pair0+npairs)
static int recross_local(crcl_handle h, const double* d_q_parents, int nparent, int pair0, int npairs,
                         int child_evol, double xi_ideal, double* d_kappa_num, double* d_kappa_denom,
                         int* d_status)
{
    int rc;
    const int ntraj = 2 * npairs;
    unsigned char* dth;
    double* dw;
    int* dst;
    if (use_split(h)) return recross_split(h, d_q_parents, nparent, pair0, npairs, child_evol, xi_ideal, d_kappa_num,
                                           d_kappa_denom, d_status);
    if ((rc = scratch(h, 8, (size_t)ntraj * (child_evol > 0 ? child_evol : 1), &dth)) ||
        (rc = scratch(h, 9, (size_t)ntraj * 2 + 2, &dw)) || (rc = scratch(h, 10, (size_t)ntraj + 1, &dst)))
        return rc;
    TrajArgs A;
    fill_args(h, A);
    A.ntraj = ntraj;
    A.nsteps = child_evol;
    A.constrain = 2;
    A.thermostat = 0;
    A.andersen_step = 0;
    A.xi_ideal_s = xi_ideal;
    A.k_force_s = 0.0;
    A.q_parents = d_q_parents;
    A.nparent = nparent;
    A.pair0 = pair0;
    A.theta = dth;
    A.weight = dw;
    A.denom_part = dw + ntraj;
    A.status = d_status ? d_status : dst;
    if (ntraj > 0 && (rc = launch_traj(h, K_RECROSS, A))) return rc;
    reduce_kappa_kernel<<<child_evol + 1, 256, 0, h->stream>>>(dth, dw, dw + ntraj, ntraj, child_evol,
                                                            d_kappa_num, d_kappa_denom);
    h->launches++;
    CK(cudaGetLastError());
    return CRCL_OK;
}

int crcl_recross_children_dev(crcl_handle h, const double* d_q_parents, int nparent, int pair0, int npairs,
                              int child_evol, double xi_ideal, double* d_kappa_num, double* d_kappa_denom,
                              int* d_status)
{
    int rc = check_traj_call(h, 2);
    if (rc) return rc;
    if (!d_q_parents || !d_kappa_num || !d_kappa_denom || nparent <= 0 || npairs < 0 || child_evol < 0)
        return CRCL_EINVAL;
    CK(cudaSetDevice(h->device));
    if ((rc = ensure_fker(h))) return rc;
    if (!h->comm) return recross_local(h, d_q_parents, nparent, pair0, npairs, child_evol, xi_ideal, d_kappa_num,
                                       d_kappa_denom, d_status);
    // collective form (recross.f90:334-417 master/worker + :390,411 result messages): every rank passes the GLOBAL
    // pair range, runs its contiguous block and receives the sums of the whole job
    long long lo, cnt;
    shard_range(npairs, h->comm->rank, h->comm->nranks, &lo, &cnt);
    if (d_status) CK(cudaMemsetAsync(d_status, 0, (size_t)2 * npairs * sizeof(int), h->stream));
    if ((rc = recross_local(h, d_q_parents, nparent, pair0 + (int)lo, (int)cnt, child_evol, xi_ideal, d_kappa_num,
                            d_kappa_denom, d_status ? d_status + 2 * lo : nullptr)))
        return rc;
    if ((rc = comm_allreduce(h, d_kappa_num, (size_t)child_evol, d_kappa_denom, 1, false))) return rc;
    if (d_status && npairs && (rc = comm_allreduce(h, nullptr, 0, d_status, (size_t)2 * npairs, true))) return rc;
    return CRCL_OK;
}

int crcl_recross_children(crcl_handle h, const double* q_parents, int nparent, int pair0, int npairs,
                          int child_evol, double xi_ideal, double* kappa_num, double* kappa_denom, int* status)
{
    if (!h || !q_parents || !kappa_num || !kappa_denom || nparent <= 0 || npairs < 0 || child_evol < 0)
        return CRCL_EINVAL;
    CK(cudaSetDevice(h->device));
    const size_t np = (size_t)nparent * h->nbeads * h->natoms * 3;
    double *dqp, *dk;
    int* dst;
    int rc;
    if ((rc = scratch(h, 0, np, &dqp)) || (rc = scratch(h, 4, (size_t)child_evol + 2, &dk)) ||
        (rc = scratch(h, 5, (size_t)2 * npairs + 1, &dst)))
        return rc;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(dqp, q_parents, np * sizeof(double), cudaMemcpyHostToDevice, s));
    if ((rc = crcl_recross_children_dev(h, dqp, nparent, pair0, npairs, child_evol, xi_ideal, dk,
                                        dk + child_evol, dst)))
        return rc;
    if (child_evol) CK(cudaMemcpyAsync(kappa_num, dk, child_evol * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(kappa_denom, dk + child_evol, sizeof(double), cudaMemcpyDeviceToHost, s));
    if (status && npairs) {
        std::vector<int> st(2 * (size_t)npairs);
        CK(cudaMemcpyAsync(st.data(), dst, st.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (int g = 0; g < npairs; g++) status[g] = st[2 * g] | st[2 * g + 1];
    }
    CK(cudaStreamSynchronize(s));
    return CRCL_OK;
}

int crcl_umbrella_windows(crcl_handle h, int nwin, const double* q0, const double* xi0, const double* k_force,
                          int ntraj, int equi_steps, int sample_steps, int constrain, uint32_t traj_id0, double* avg,
                          double* var, int* status)
{
    int rc = check_traj_call(h, 0);
    if (rc) return rc;
    if (!q0 || !xi0 || !k_force || !avg || !var || nwin < 0 || ntraj < 0 || equi_steps < 0 || sample_steps <= 0 ||
        (constrain != 0 && constrain != 3))
        return CRCL_EINVAL;
    const int nglob = nwin * ntraj;
    if (nglob == 0) return CRCL_OK;
    CK(cudaSetDevice(h->device));
    if ((rc = ensure_fker(h))) return rc;
    // With a communicator the (window, trajectory) units of the whole job are partitioned over the ranks
    // (calc_rate.f90:1351-1376 hands whole windows to MPI workers); this rank runs units [lo, lo+ntot)
    long long lo = 0, cnt = nglob;
    if (h->comm) shard_range(nglob, h->comm->rank, h->comm->nranks, &lo, &cnt);
    const int ntot = (int)cnt;
    const size_t per = (size_t)h->nbeads * h->natoms * 3, n = per * ntot, nd = (size_t)ntot * h->natoms * 3;
    double *dq, *dp, *dg, *ddxi, *dep, *dnhc, *dwin, *dv;
    int* dst;
    uint32_t* dev;
    if ((rc = scratch(h, 0, n, &dq)) || (rc = scratch(h, 1, n, &dg)) || (rc = scratch(h, 2, n, &dp)) ||
        (rc = scratch(h, 3, nd, &ddxi)) || (rc = scratch(h, 4, (size_t)ntot * 4, &dep)) ||
        (rc = scratch(h, 5, (size_t)ntot, &dst)) || (rc = scratch(h, 6, (size_t)ntot * 2, &dev)) ||
        (rc = scratch(h, 7, (size_t)ntot * 8, &dnhc)) || (rc = scratch(h, 20, (size_t)ntot * 2, &dwin)) ||
        (rc = scratch(h, 21, (size_t)ntot * h->nbeads, &dv)))
        return rc;
    cudaStream_t s = h->stream;
    std::vector<double> sums(2 * (size_t)ntot);
    std::vector<int> st(ntot);
    if (ntot > 0) {
    // every trajectory starts from its window's equilibrated structure (calc_rate.f90:1383-1387)
    std::vector<double> rep(n), win(2 * (size_t)ntot);
    for (int i = 0; i < ntot; i++) {
        const int w = (int)((lo + i) / ntraj);
        memcpy(rep.data() + (size_t)i * per, q0 + (size_t)w * per, per * sizeof(double));
        win[i] = xi0[w];
        win[ntot + i] = k_force[w];
    }
    CK(cudaMemcpyAsync(dq, rep.data(), n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dwin, win.data(), win.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(dp, 0, n * sizeof(double), s));
    CK(cudaMemsetAsync(dev, 0, ntot * 2 * sizeof(uint32_t), s));
    CK(cudaMemsetAsync(dst, 0, ntot * sizeof(int), s));
    CK(cudaMemsetAsync(dep, 0, ntot * 4 * sizeof(double), s));
    const uint32_t tid0 = traj_id0 + (uint32_t)lo;   // RNG streams are keyed by the GLOBAL unit index
    if (use_split(h)) {
        std::vector<uint32_t> ids(ntot);
        for (int t = 0; t < ntot; t++) ids[t] = tid0 + (uint32_t)t;
        uint32_t* dtid;
        if ((rc = scratch(h, 27, (size_t)ntot, &dtid))) return rc;
        CK(cudaMemcpyAsync(dtid, ids.data(), ntot * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        SplitCall C;
        C.ntraj = ntot;
        C.q = dq;
        C.p = dp;
        C.g = dg;
        C.dxi = ddxi;
        C.epot = dep;
        C.xi_real = dep + ntot;
        C.nhc = dnhc;
        C.status = dst;
        C.tid = dtid;
        C.event = dev;
        C.xi_ideal = dwin;
        C.k_force = dwin + ntot;
        if ((rc = mdinit_split(h, C, 2))) return rc;
        if (equi_steps > 0 && (rc = verlet_split(h, C, equi_steps, 0, constrain))) return rc;
        // plain forces before sampling (calc_rate.f90:1619-1623); split_forces also serves the host-callback PES
        if ((rc = split_forces(h, ntot * h->nbeads, dq, dg, dv))) return rc;
        C.xi_sum = dep + 2 * ntot;
        C.xi_sum2 = dep + 3 * ntot;
        if ((rc = verlet_split(h, C, sample_steps, 0, constrain))) return rc;
    } else {
    TrajArgs A;
    fill_args(h, A);
    A.ntraj = ntot;
    A.constrain = constrain;
    A.xi_ideal = dwin;
    A.k_force = dwin + ntot;
    A.q = dq;
    A.p = dp;
    A.g = dg;
    A.dxi = ddxi;
    A.epot = dep;
    A.xi_real = dep + ntot;
    A.status = dst;
    A.nhc = dnhc;
    A.traj_id0 = tid0;
    A.event = dev;
    if ((rc = launch_traj(h, K_MDINIT, A, 2))) return rc;
    A.nsteps = equi_steps;
    A.istep0 = 0;
    if (equi_steps > 0 && (rc = launch_traj(h, K_VERLET, A))) return rc;
    // calc_rate.f90:1619-1623 recomputes derivs with plain `gradient` calls before the sampling
    // loop: the first half kick of the sampling phase sees the forces WITHOUT the umbrella bias
    if ((rc = crcl_egrad_dev(h, h->pes, dq, h->natoms, ntot * h->nbeads, dv, dg, nullptr))) return rc;
    A.nsteps = sample_steps;
    A.istep0 = 0;
    A.xi_sum = dep + 2 * ntot;
    A.xi_sum2 = dep + 3 * ntot;
    if ((rc = launch_traj(h, K_VERLET, A))) return rc;
    }
    CK(cudaMemcpyAsync(sums.data(), dep + 2 * ntot, sums.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(st.data(), dst, ntot * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    }
    if (h->comm) {
        for (int t = 0; t < nglob; t++) {
            avg[t] = var[t] = 0.0;
            if (status) status[t] = 0;
        }
    }
    for (int t = 0; t < ntot; t++) {
        // calc_rate.f90:1660-1664: av = sum/n, var = sum2/n - av^2
        const double a = sums[t] / sample_steps;
        avg[lo + t] = a;
        var[lo + t] = sums[ntot + t] / sample_steps - a * a;
        if (status) status[lo + t] = st[t];
    }
    if (h->comm) {
        // the statistics files of calc_rate.f90:1690-1734 become one all-reduce: every rank fills its own slice of
        // zero vectors (average, variance, status as doubles) and the vectors are summed
        double* dred;
        if ((rc = scratch(h, 28, (size_t)3 * nglob, &dred))) return rc;
        std::vector<double> pack(3 * (size_t)nglob);
        for (int t = 0; t < nglob; t++) {
            pack[t] = avg[t];
            pack[nglob + t] = var[t];
            pack[2 * (size_t)nglob + t] = status ? (double)status[t] : 0.0;
        }
        CK(cudaMemcpyAsync(dred, pack.data(), pack.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        if ((rc = comm_allreduce(h, dred, pack.size(), nullptr, 0, false))) return rc;
        CK(cudaMemcpyAsync(pack.data(), dred, pack.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (int t = 0; t < nglob; t++) {
            avg[t] = pack[t];
            var[t] = pack[nglob + t];
            if (status) status[t] = (int)pack[2 * (size_t)nglob + t];
        }
    }
    return CRCL_OK;
}

int crcl_umbrella_window(crcl_handle h, const double* q0, double xi0, double k_force, int ntraj,
                         int equi_steps, int sample_steps, uint32_t traj_id0, double* avg, double* var,
                         int* status)
{
    return crcl_umbrella_windows(h, 1, q0, &xi0, &k_force, ntraj, equi_steps, sample_steps, 0, traj_id0, avg, var,
                                 status);
}

// ---- hooks ------------------------------------------------------------------------------------
int crcl_rng_normals(crcl_handle h, uint64_t seed, uint32_t traj, uint32_t event, uint32_t bead, int n,
                     double* out)
{
    if (!h || !out || n < 0) return CRCL_EINVAL;
    if (n == 0) return CRCL_OK;
    CK(cudaSetDevice(h->device));
    double* d;
    int rc;
    if ((rc = scratch(h, 0, (size_t)n + 1, &d))) return rc;
    const int pairs = (n + 1) / 2;
    rng_kernel<<<(pairs + 127) / 128, 128, 0, h->stream>>>(seed, traj, event, bead, n, d);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return CRCL_OK;
}

int crcl_bench_propagate(crcl_handle h, int ntraj, int reps, double* ms_out)
{
    if (!h || ntraj <= 0 || reps <= 0 || !ms_out) return CRCL_EINVAL;
    CK(cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure_split_tables(h)) || (rc = ensure_fker(h))) return rc;
    const int na = h->natoms, nb = h->nbeads, nc = 3 * na;
    const size_t n = (size_t)ntraj * nb * nc;
    double *dq, *dp, *dg, *dcen;
    int* dst;
    if ((rc = scratch(h, 0, n, &dq)) || (rc = scratch(h, 1, n, &dg)) || (rc = scratch(h, 2, n, &dp)) ||
        (rc = scratch(h, 8, (size_t)ntraj * nc, &dcen)) || (rc = scratch(h, 5, (size_t)ntraj, &dst)))
        return rc;
    // finite, non-trivial contents: zero momenta/forces, positions = byte pattern 0x3f (1.2e-4 ...)
    CK(cudaMemsetAsync(dq, 0x3f, n * sizeof(double), h->stream));
    CK(cudaMemsetAsync(dp, 0, n * sizeof(double), h->stream));
    CK(cudaMemsetAsync(dg, 0, n * sizeof(double), h->stream));
    SplitArgs A{};
    A.ntraj = ntraj;
    A.natoms = na;
    A.nbeads = nb;
    A.symmetrize = (h->transform == CRCL_TRANSFORM_REFERENCE) ? 1 : 0;
    A.dt = h->dt;
    A.beta = h->beta;
    A.mass = h->d_mass;
    A.at_move = h->d_atmove;
    A.fker = h->d_fker;
    A.q = dq;
    A.p = dp;
    A.g = dg;
    A.cen = dcen;
    A.status = dst;
    double best = 1e30, sum = 0.0;
    for (int r = 0; r < reps + 2; r++) {
        if ((rc = launch_kick_freerp(h, A))) return rc;
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (r >= 2) {
            sum += ms;
            if (ms < best) best = ms;
        }
    }
    ms_out[0] = sum / reps;
    ms_out[1] = best;
    return CRCL_OK;
}

long long crcl_launch_count(crcl_handle h) { return h ? h->launches : -1; }

double crcl_last_kernel_ms(crcl_handle h)
{
    if (!h) return -1.0;
    float ms = -1.0f;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0;
    return (double)ms;
}

int crcl_kernel_timings(crcl_handle h, double* ms_out, int max_n)
{
    if (!h || (!ms_out && max_n > 0)) return CRCL_EINVAL;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return CRCL_ECUDA;
    const int n = h->ev_count < max_n ? h->ev_count : max_n;
    // oldest first among the last n launches
    for (int k = 0; k < n; k++) {
        const int i = ((h->ev_head - n + k) % crcl_handle_s::NEV + crcl_handle_s::NEV) % crcl_handle_s::NEV;
        float ms = -1.0f;
        if (cudaEventElapsedTime(&ms, h->evs[2 * i], h->evs[2 * i + 1]) != cudaSuccess) ms = -1.0f;
        ms_out[k] = (double)ms;
    }
    h->ev_count = 0;
    return n;
}

double crcl_measure_fp64_tflops(crcl_handle h, int iters)
{
    if (!h) return -1.0;
    if (iters <= 0) iters = 4096;
    cudaSetDevice(h->device);
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, h->device) != cudaSuccess) return -1.0;
    const int tpb = 256, grid = pr.multiProcessorCount * 8;
    double* d;
    if (scratch(h, 11, (size_t)grid * tpb, &d)) return -1.0;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(h->ev0, h->stream);
        fp64_peak_kernel<<<grid, tpb, 0, h->stream>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(h->ev1, h->stream);
        h->launches++;
        if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0;
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        const double flops = 2.0 * 8.0 * (double)iters * grid * tpb;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    return best;
}

}  // extern "C" (templates below)

template <int NB, int NC, int TPC, int MODE>
static int run_transform_bench(crcl_handle h, const double* d_f, const double* d_m, double2* d_pq, int ntraj, int reps, float* ms)
{
    using B = TransformBench<NB, NC, TPC, MODE>;
    auto kern = transform_bench_kernel<NB, NC, TPC, MODE>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B::smem_bytes()) != cudaSuccess) return CRCL_ECUDA;
    cudaEventRecord(h->ev0, h->stream);
    kern<<<ntraj / TPC, B::THREADS, B::smem_bytes(), h->stream>>>(d_f, d_m, d_pq, reps);
    cudaEventRecord(h->ev1, h->stream);
    h->launches++;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess || cudaGetLastError() != cudaSuccess) return CRCL_ECUDA;
    cudaEventElapsedTime(ms, h->ev0, h->ev1);
    return CRCL_OK;
}

template <int NB, int NC, int TPC>
static int transform_bench(crcl_handle h, int ntraj, int reps, double* out)
{
    ntraj = (ntraj / TPC) * TPC;
    if (ntraj <= 0 || reps <= 0) return fail(h, CRCL_EINVAL, "crcl_bench_transform: ntraj >= 4 and reps >= 1");
    std::vector<double> f, m(NC);
    build_fker(NB, h->beta, h->dt, f);
    for (int c = 0; c < NC; c++) m[c] = ((c / 3) == 1 ? 12.0 : 1.00782503207) * 1822.888486;   // one heavy atom among hydrogens
    const size_t n = (size_t)ntraj * NC * NB;
    std::vector<double2> x(n);
    uint64_t sd = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < n; i++) {
        sd = sd * 6364136223846793005ull + 1442695040888963407ull;
        const double u = (double)(sd >> 11) / 9007199254740992.0 - 0.5;
        sd = sd * 6364136223846793005ull + 1442695040888963407ull;
        const double v = (double)(sd >> 11) / 9007199254740992.0 - 0.5;
        x[i] = make_double2(10.0 * u, 2.0 * v);      // momenta ~ sqrt(m kT), positions ~ bohr
    }
    double *d_f, *d_m;
    double2 *d_a, *d_b;
    int rc;
    if ((rc = scratch(h, 11, f.size(), &d_f)) || (rc = scratch(h, 12, m.size(), &d_m)) || (rc = scratch(h, 13, n, &d_a)) ||
        (rc = scratch(h, 14, n, &d_b)))
        return rc;
    CK(cudaMemcpy(d_f, f.data(), f.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_m, m.data(), m.size() * sizeof(double), cudaMemcpyHostToDevice));
    // the two forms on the same input, three steps: they must agree to rounding
    float ms;
    CK(cudaMemcpy(d_a, x.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, x.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
    if ((rc = run_transform_bench<NB, NC, TPC, 0>(h, d_f, d_m, d_a, ntraj, 3, &ms))) return fail(h, rc, "transform bench (DFMA) failed");
    if ((rc = run_transform_bench<NB, NC, TPC, 1>(h, d_f, d_m, d_b, ntraj, 3, &ms))) return fail(h, rc, "transform bench (DMMA) failed");
    std::vector<double2> ya(n), yb(n);
    CK(cudaMemcpy(ya.data(), d_a, n * sizeof(double2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(yb.data(), d_b, n * sizeof(double2), cudaMemcpyDeviceToHost));
    double dmax = 0.0, pmax = 0.0, qmax = 0.0, moved = 0.0;
    for (size_t i = 0; i < n; i++) {
        pmax = fmax(pmax, fabs(ya[i].x));
        qmax = fmax(qmax, fabs(ya[i].y));
        moved = fmax(moved, fabs(ya[i].y - x[i].y));
    }
    for (size_t i = 0; i < n; i++) dmax = fmax(dmax, fmax(fabs(ya[i].x - yb[i].x) / pmax, fabs(ya[i].y - yb[i].y) / qmax));
    out[2] = dmax;
    out[4] = moved;       // the transform did something (max |q' - q|)
    for (int mode = 0; mode < 2; mode++) {
        double best = 1e30;
        for (int rep = 0; rep < 4; rep++) {
            rc = mode ? run_transform_bench<NB, NC, TPC, 1>(h, d_f, d_m, d_b, ntraj, reps, &ms)
                      : run_transform_bench<NB, NC, TPC, 0>(h, d_f, d_m, d_a, ntraj, reps, &ms);
            if (rc) return fail(h, rc, "transform bench failed");
            if (rep > 0 && ms < best) best = ms;
        }
        out[mode] = best;
    }
    out[3] = (double)reps * ntraj * NC * 4.0 * 2.0 * NB * NB;
    return CRCL_OK;
}

extern "C" {

int crcl_bench_transform(crcl_handle h, int nbeads, int ntraj, int reps, double* out)
{
    if (!h || !out) return CRCL_EINVAL;
    cudaSetDevice(h->device);
    if (nbeads == 16) return transform_bench<16, 18, 4>(h, ntraj, reps, out);
    if (nbeads == 64) return transform_bench<64, 12, 4>(h, ntraj, reps, out);
    return fail(h, CRCL_ENOSUP, "crcl_bench_transform: 16 beads (6 atoms) or 64 beads (4 atoms)");
}

double crcl_measure_dmma_tflops(crcl_handle h, int iters)
{
    if (!h) return -1.0;
    if (iters <= 0) iters = 4096;
    cudaSetDevice(h->device);
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, h->device) != cudaSuccess) return -1.0;
    const int tpb = 256, grid = pr.multiProcessorCount * 8;
    double* d;
    if (scratch(h, 11, (size_t)grid * tpb, &d)) return -1.0;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(h->ev0, h->stream);
        dmma_peak_kernel<<<grid, tpb, 0, h->stream>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(h->ev1, h->stream);
        h->launches++;
        if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0;
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        const double flops = 4.0 * 512.0 * (double)iters * grid * (tpb / 32);
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    return best;
}

}  // extern "C"
