// Bundle adjustment of libcubemap_b200.so: Optimizer::LocalBundleAdjustment and Optimizer::PoseOptimization on sm_100a.
//
// Reference (CPU, g2o): src/Optimizer.cpp:48-451; edges src/g2o_cubemap_vertices_edges.cpp:61-233; LM loop
// ThirdParty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189; Schur solve core/block_solver.hpp:354-486.
// The hypergraph is replaced by flat fp64 arrays; all arithmetic stays fp64 (the gate is 1e-5 relative on poses/points),
// including the reference's float32 round trip inside the projection (ba_math.cuh).
//
// Per LM iteration:  k_ba_errors -> k_ba_linearize -> [per trial: k_ba_dinv, k_ba_schur (landmark-sharded across ranks,
// ncclAllReduce of [S | g | scalars]), k_ba_solve (blocked LDL^T of the <=6P x 6P reduced camera system, one CTA),
// k_ba_backsub, k_ba_update, k_ba_errors, k_ba_scale].  Control flow (accept/reject, lambda schedule, the ORB-SLAM2
// stop rule, pbStopFlag) stays on the host like in the reference; one small D2H per trial.
#include <dlfcn.h>
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <vector>
#include "ba_math.cuh"
#include "common.cuh"

namespace cslam {

struct BADev {
    int nKF, nMP, nE, nP, n;        // nP free active poses, n = 6 nP
    double f;                        // fx=fy=cx=cy
    Pose* pose; Pose* poseBak;
    double* X; double* Xbak;
    const int* eMP; const int* eKF;  // edges sorted by landmark
    const double* obs;               // 3 per edge: mx, my, w
    const int8_t* face;
    double* err; uint8_t* level;
    const int* poseIdx; const uint8_t* ptAct;   // per KF: compact index or -1; per MP: 1 if it has an active edge
    const int* lmStart;              // nMP+1, CSR over sorted edges
    const int* peStart; const int* peList;      // CSR of edge ids by KF
    double *Hpp, *bp, *Hll, *bl, *Hpl, *Dinv, *db, *xp, *xl;
    double *S, *g, *bpr;             // one contiguous all-reduce buffer: [S n*n | g n | bpr n] (bpr = full pose gradient for computeScale)
    double* scal;                    // [0] chi2, [1] scale, [2] max diag (as bits), [3] solve flag
    int robust; double delta, dsqr;
    int rank, nranks;                // landmark l is owned by rank l % nranks
};

__device__ __forceinline__ bool owned(const BADev& D, int l) { return (l % D.nranks) == D.rank; }

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0;
    if (w == 0) {
        r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;   // valid in thread 0
}

// computeActiveErrors + activeRobustChi2 (partial sum over the landmarks this rank owns)
__global__ void __launch_bounds__(256) k_ba_errors(BADev D) {
    __shared__ double sh[8];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    double c = 0;
    if (e < D.nE && D.level[e] == 0 && owned(D, D.eMP[e])) {
        double Xc[3], er[2];
        pose_map(D.pose[D.eKF[e]], D.X + 3 * D.eMP[e], Xc);
        edge_error(D.face[e], D.f, D.obs[3 * e], D.obs[3 * e + 1], Xc, er);
        D.err[2 * e] = er[0]; D.err[2 * e + 1] = er[1];
        const double chi = D.obs[3 * e + 2] * (er[0] * er[0] + er[1] * er[1]);
        if (D.robust) { double r0, r1; huber(D.delta, D.dsqr, chi, r0, r1); c = r0; } else c = chi;
    }
    const double s = block_sum(c, sh);
    if (threadIdx.x == 0 && s != 0.0) atomicAdd(&D.scal[0], s);
}

// buildSystem: linearizeOplus + constructQuadraticForm of every active edge (base_binary_edge.hpp:55-120)
__global__ void __launch_bounds__(256) k_ba_linearize(BADev D, int useSmem) {
    extern __shared__ double shH[];   // nP * 42 doubles when useSmem
    if (useSmem) { for (int i = threadIdx.x; i < D.nP * 42; i += blockDim.x) shH[i] = 0.0; __syncthreads(); }
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < D.nE && D.level[e] == 0 && owned(D, D.eMP[e])) {
        const int k = D.eKF[e], l = D.eMP[e], pi = D.poseIdx[k];
        const Pose T = D.pose[k];
        double Xc[3], G[2][3], Jp[2][6], Jx[2][3], R[3][3];
        pose_map(T, D.X + 3 * l, Xc);
        edge_G(D.face[e], D.f, Xc, G);
        quat_to_matrix(T.q, R);
        for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) Jx[i][j] = G[i][0] * R[0][j] + G[i][1] * R[1][j] + G[i][2] * R[2][j];
        const double e0 = D.err[2 * e], e1 = D.err[2 * e + 1], w0 = D.obs[3 * e + 2];
        double rho1 = 1.0;
        if (D.robust) { double r0; huber(D.delta, D.dsqr, w0 * (e0 * e0 + e1 * e1), r0, rho1); }
        const double w = rho1 * w0, r0 = -w * e0, r1 = -w * e1;
        for (int i = 0; i < 3; i++) {
            atomicAdd(&D.bl[3 * l + i], Jx[0][i] * r0 + Jx[1][i] * r1);
            for (int j = 0; j < 3; j++) atomicAdd(&D.Hll[9 * l + 3 * i + j], w * (Jx[0][i] * Jx[0][j] + Jx[1][i] * Jx[1][j]));
        }
        if (pi >= 0) {
            edge_Jpose(G, Xc, Jp);
            double* Hp = useSmem ? shH + 42 * pi : D.Hpp + 36 * pi;
            double* bpp = useSmem ? shH + 42 * pi + 36 : D.bp + 6 * pi;
            for (int i = 0; i < 6; i++) {
                atomicAdd(&bpp[i], Jp[0][i] * r0 + Jp[1][i] * r1);
                for (int j = 0; j < 6; j++) atomicAdd(&Hp[6 * i + j], w * (Jp[0][i] * Jp[0][j] + Jp[1][i] * Jp[1][j]));
                for (int j = 0; j < 3; j++) D.Hpl[18 * (size_t)e + 3 * i + j] = w * (Jp[0][i] * Jx[0][j] + Jp[1][i] * Jx[1][j]);
            }
        }
    }
    if (useSmem) {
        __syncthreads();
        for (int i = threadIdx.x; i < D.nP * 42; i += blockDim.x) {
            const double v = shH[i];
            if (v != 0.0) { const int p = i / 42, r = i - 42 * p; atomicAdd(r < 36 ? &D.Hpp[36 * p + r] : &D.bp[6 * p + r - 36], v); }
        }
    }
}

// computeLambdaInit: max |H_jj| over all active free vertices
__global__ void __launch_bounds__(256) k_ba_maxdiag(BADev D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0;
    if (i < D.nP * 6) m = fabs(D.Hpp[36 * (i / 6) + 7 * (i % 6)]);
    const int j = i - D.nP * 6;
    if (j >= 0 && j < D.nMP * 3) { const int l = j / 3; if (D.ptAct[l] && owned(D, l)) m = fabs(D.Hll[9 * l + 4 * (j % 3)]); }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(reinterpret_cast<unsigned long long*>(&D.scal[2]), (unsigned long long)__double_as_longlong(m));
}

// Dinv = (Hll + lambda I)^-1 (3x3 cofactors, like Eigen), db = Dinv * bl
__global__ void __launch_bounds__(256) k_ba_dinv(BADev D, double lambda) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= D.nMP || !D.ptAct[l] || !owned(D, l)) return;
    double M[9];
    for (int i = 0; i < 9; i++) M[i] = D.Hll[9 * l + i];
    M[0] += lambda; M[4] += lambda; M[8] += lambda;
    const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    const double id = 1.0 / (M[0] * c00 + M[1] * c01 + M[2] * c02);
    double Di[9];
    Di[0] = c00 * id; Di[1] = (M[2] * M[7] - M[1] * M[8]) * id; Di[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    Di[3] = c01 * id; Di[4] = (M[0] * M[8] - M[2] * M[6]) * id; Di[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    Di[6] = c02 * id; Di[7] = (M[1] * M[6] - M[0] * M[7]) * id; Di[8] = (M[0] * M[4] - M[1] * M[3]) * id;
    for (int i = 0; i < 9; i++) D.Dinv[9 * l + i] = Di[i];
    for (int i = 0; i < 3; i++) D.db[3 * l + i] = Di[3 * i] * D.bl[3 * l] + Di[3 * i + 1] * D.bl[3 * l + 1] + Di[3 * i + 2] * D.bl[3 * l + 2];
}

// S = Hpp(partial) ; g = bp(partial)   (lambda is added to the diagonal after the all-reduce)
__global__ void __launch_bounds__(256) k_ba_s_init(BADev D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D.n * D.n) {
        const int r = i / D.n, c = i - r * D.n;
        D.S[i] = (r / 6 == c / 6) ? D.Hpp[36 * (r / 6) + 6 * (r % 6) + (c % 6)] : 0.0;
    }
    if (i < D.n) { D.g[i] = D.bp[i]; D.bpr[i] = D.bp[i]; }
}
__global__ void __launch_bounds__(256) k_ba_add_lambda(BADev D, double lambda) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D.n) D.S[(size_t)i * D.n + i] += lambda;
}

// Schur complement, reference block_solver.hpp:381-439: for every landmark and every ordered pair of its observations
// (a1,a2): S[p1,p2] -= (B1 Dinv) B2^T, g[p1] -= B1 (Dinv bl). One CTA = a chunk of the edges of one pose p1 (row block of S
// accumulated in shared memory), one warp = one edge a1 at a time.
__global__ void __launch_bounds__(256) k_ba_schur(BADev D, int chunks) {
    extern __shared__ double shS[];   // 6 x n row block + 6 (g)
    const int k1 = blockIdx.y, p1 = D.poseIdx[k1];
    if (p1 < 0) return;
    const int n = D.n;
    for (int i = threadIdx.x; i < 6 * n + 6; i += blockDim.x) shS[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int beg = D.peStart[k1], end = D.peStart[k1 + 1];
    for (int q = beg + blockIdx.x * nw + warp; q < end; q += chunks * nw) {
        const int a1 = D.peList[q];
        if (D.level[a1] != 0) continue;
        const int l = D.eMP[a1];
        if (!owned(D, l)) continue;
        const double* B1 = D.Hpl + 18 * (size_t)a1;
        const double* Di = D.Dinv + 9 * l;
        // BD (6x3): lanes 0..17
        double bd = 0;
        if (lane < 18) { const int i = lane / 3, j = lane % 3; bd = B1[3 * i] * Di[j] + B1[3 * i + 1] * Di[3 + j] + B1[3 * i + 2] * Di[6 + j]; }
        if (lane < 6) atomicAdd(&shS[6 * n + lane], -(B1[3 * lane] * D.db[3 * l] + B1[3 * lane + 1] * D.db[3 * l + 1] + B1[3 * lane + 2] * D.db[3 * l + 2]));
        const int s = D.lmStart[l], t = D.lmStart[l + 1];
        // 36 outputs per partner: lane handles elements lane and lane+32 (<36)
        for (int a2 = s; a2 < t; a2++) {
            if (D.level[a2] != 0) continue;
            const int p2 = D.poseIdx[D.eKF[a2]];
            if (p2 < 0) continue;
            const double* B2 = D.Hpl + 18 * (size_t)a2;
#pragma unroll
            for (int rep = 0; rep < 2; rep++) {
                const int el = lane + 32 * rep;
                const int i = el / 6, j = el - 6 * i;   // rows of B1 x rows of B2
                double v = 0;
#pragma unroll
                for (int c = 0; c < 3; c++) v += __shfl_sync(0xffffffffu, bd, (el < 36 ? i : 0) * 3 + c) * (el < 36 ? B2[3 * j + c] : 0.0);
                if (el < 36) atomicAdd(&shS[i * n + 6 * p2 + j], -v);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 6 * n; i += blockDim.x) {
        const double v = shS[i];
        if (v != 0.0) atomicAdd(&D.S[(size_t)(6 * p1 + i / n) * n + (i % n)], v);
    }
    if (threadIdx.x < 6) { const double v = shS[6 * n + threadIdx.x]; if (v != 0.0) atomicAdd(&D.g[6 * p1 + threadIdx.x], v); }
}

// Dense LDL^T (no pivoting) + solve of the n x n reduced camera system in one CTA, blocked (panel NB) so that the O(n^3)
// work reads the panel from shared memory. scal[3] = 1 on a zero / non-finite pivot (LinearSolverEigen's failure).
static const int LD_NB = 32;
// Blocked right-looking LDL^T in one CTA. The right-hand side g is stored right behind S, i.e. it is row n of an (n+1) x n
// matrix: carrying it through the factorisation as one more row performs the forward substitution for free (row n of L =
// D^-1 L^-1 g). Per panel of NB columns: (1) warp 0 factors the NB x NB diagonal block with warp-level sync only,
// (2) every row below is a NB-step forward substitution against it, one thread per row, no block syncs,
// (3) trailing update with 4x2 register tiles from the shared-memory panel. Back substitution is blocked the same way.
__global__ void __launch_bounds__(1024) k_ba_solve(BADev D) {
    extern __shared__ double sm[];   // panel L: (n+1) x (NB+1) ; panel L*d: same ; d: n ; x: n
    const int n = D.n, T = blockDim.x, tid = threadIdx.x, PS = LD_NB + 1, lane = tid & 31, warp = tid >> 5;
    double* A = D.S;                 // (n+1) x n, lower triangle + row n used; overwritten by L (unit diagonal implied)
    double* P = sm; double* PD = P + (size_t)(n + 1) * PS; double* dvec = PD + (size_t)(n + 1) * PS; double* y = dvec + n;
    __shared__ int fail;
    if (tid == 0) fail = 0;
    __syncthreads();
    for (int jb = 0; jb < n; jb += LD_NB) {
        const int nb = min(LD_NB, n - jb), rows = n + 1 - jb;
        for (int i = tid; i < rows * nb; i += T) { const int r = i / nb, c = i - r * nb; P[r * PS + c] = A[(size_t)(jb + r) * n + jb + c]; }
        __syncthreads();
        // (1) diagonal block: lane = row
        if (warp == 0) {
            for (int c = 0; c < nb; c++) {
                const double dc = P[c * PS + c];
                if (lane == 0) { if (dc == 0.0 || !isfinite(dc)) fail = 1; dvec[jb + c] = dc; }
                if (lane > c && lane < nb) {
                    const double v = P[lane * PS + c];
                    PD[lane * PS + c] = v; P[lane * PS + c] = v / dc;
                }
                __syncwarp();
                if (lane > c && lane < nb)
                    for (int c2 = c + 1; c2 <= lane; c2++) P[lane * PS + c2] -= PD[lane * PS + c] * P[c2 * PS + c];
                __syncwarp();
            }
        }
        __syncthreads();
        if (fail) break;
        // (2) rows below the diagonal block (including the rhs row): v_c = a_c - sum_{c'<c} (L d)[r][c'] L[c][c'] ; L[r][c] = v_c / d_c
        for (int r = nb + tid; r < rows; r += T) {
            double* Pr = P + (size_t)r * PS; double* PDr = PD + (size_t)r * PS;
            for (int c = 0; c < nb; c++) {
                double v = Pr[c];
                const double* Lc = P + (size_t)c * PS;
                for (int c2 = 0; c2 < c; c2++) v -= PDr[c2] * Lc[c2];
                PDr[c] = v; Pr[c] = v / dvec[jb + c];
            }
        }
        __syncthreads();
        // write L panel back, then trailing update A[i][k] -= sum_c (L[i][c] d_c) L[k][c] for rows i >= cols k >= jb+nb, 4x2 tiles
        for (int i = tid; i < rows * nb; i += T) { const int r = i / nb, c = i - r * nb; if (r > c) A[(size_t)(jb + r) * n + jb + c] = P[r * PS + c]; }
        const int trR = rows - nb, trC = n - jb - nb;      // rows include the rhs row, columns do not
        const int tR = (trR + 3) >> 2, tC = (trC + 1) >> 1;
        for (int i = tid; i < tR * tC; i += T) {
            const int br = i / tC, bc = i - br * tC;
            const int r0 = 4 * br, k0 = 2 * bc;
            if (k0 > r0 + 3) continue;                      // tile entirely above the diagonal
            const double* Lr[4]; const double* Lk[2];
#pragma unroll
            for (int a = 0; a < 4; a++) Lr[a] = PD + (size_t)(nb + min(r0 + a, trR - 1)) * PS;
#pragma unroll
            for (int bq = 0; bq < 2; bq++) Lk[bq] = P + (size_t)(nb + min(k0 + bq, trC - 1)) * PS;
            double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
            for (int c = 0; c < nb; c++) {
                const double b0 = Lk[0][c], b1 = Lk[1][c];
#pragma unroll
                for (int a = 0; a < 4; a++) { const double av = Lr[a][c]; acc[a][0] += av * b0; acc[a][1] += av * b1; }
            }
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const int r = r0 + a;
                if (r >= trR) break;
                double* dst = A + (size_t)(jb + nb + r) * n + jb + nb;
                if (k0 <= r && k0 < trC) dst[k0] -= acc[a][0];
                if (k0 + 1 <= r && k0 + 1 < trC) dst[k0 + 1] -= acc[a][1];
            }
        }
        __syncthreads();
    }
    if (fail) { if (tid == 0) D.scal[3] = 1.0; return; }
    // backward: L^T x = z, z = row n of L (= D^-1 L^-1 g); processed in blocks of NB columns from the bottom
    for (int i = tid; i < n; i += T) y[i] = A[(size_t)n * n + i];
    __syncthreads();
    for (int je = n; je > 0; je -= LD_NB) {
        const int j0 = max(je - LD_NB, 0), nb = je - j0;
        // load the nb x nb diagonal block of L into P (row-major), solve it with warp 0: x_j final for j in [j0, je)
        for (int i = tid; i < nb * nb; i += T) { const int r = i / nb, c = i - r * nb; P[r * PS + c] = (r > c) ? A[(size_t)(j0 + r) * n + j0 + c] : 0.0; }
        __syncthreads();
        if (warp == 0) {
            for (int j = nb - 1; j >= 0; j--) {
                const double xj = y[j0 + j];
                if (lane < j) y[j0 + lane] -= P[j * PS + lane] * xj;
                __syncwarp();
            }
        }
        __syncthreads();
        // rows above: y[i] -= sum_{j in block} L[j][i] x_j   (rows of L are contiguous: coalesced over i)
        for (int i = tid; i < j0; i += T) {
            double s = 0;
            for (int j = 0; j < nb; j++) s += A[(size_t)(j0 + j) * n + i] * y[j0 + j];
            y[i] -= s;
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += T) D.xp[i] = y[i];
}

// xl = Dinv (bl - Hpl^T xp)   (block_solver.hpp:461-481)
__global__ void __launch_bounds__(256) k_ba_backsub(BADev D) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= D.nMP) return;
    if (!D.ptAct[l] || !owned(D, l)) { D.xl[3 * l] = 0; D.xl[3 * l + 1] = 0; D.xl[3 * l + 2] = 0; return; }
    double cl[3] = {D.bl[3 * l], D.bl[3 * l + 1], D.bl[3 * l + 2]};
    for (int a = D.lmStart[l]; a < D.lmStart[l + 1]; a++) {
        if (D.level[a] != 0) continue;
        const int p = D.poseIdx[D.eKF[a]];
        if (p < 0) continue;
        const double* B = D.Hpl + 18 * (size_t)a;
        for (int j = 0; j < 3; j++) { double s = 0; for (int i = 0; i < 6; i++) s += B[3 * i + j] * D.xp[6 * p + i]; cl[j] -= s; }
    }
    const double* Di = D.Dinv + 9 * l;
    for (int i = 0; i < 3; i++) D.xl[3 * l + i] = Di[3 * i] * cl[0] + Di[3 * i + 1] * cl[1] + Di[3 * i + 2] * cl[2];
}

__global__ void __launch_bounds__(256) k_ba_update(BADev D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D.nKF) { const int p = D.poseIdx[i]; if (p >= 0) D.pose[i] = pose_oplus(D.pose[i], D.xp + 6 * p); }
    if (i < D.nMP && D.ptAct[i] && owned(D, i)) { D.X[3 * i] += D.xl[3 * i]; D.X[3 * i + 1] += D.xl[3 * i + 1]; D.X[3 * i + 2] += D.xl[3 * i + 2]; }
}

// computeScale: sum_j x_j (lambda x_j + b_j); the pose part is added by rank 0 only (replicated), landmarks by their owner
__global__ void __launch_bounds__(256) k_ba_scale(BADev D, double lambda) {
    __shared__ double sh[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0;
    if (i < D.n && D.rank == 0) v = D.xp[i] * (lambda * D.xp[i] + D.bpr[i]);
    const int j = i - D.n;
    if (j >= 0 && j < 3 * D.nMP) { const int l = j / 3; if (D.ptAct[l] && owned(D, l)) v = D.xl[j] * (lambda * D.xl[j] + D.bl[j]); }
    const double s = block_sum(v, sh);
    if (threadIdx.x == 0 && s != 0.0) atomicAdd(&D.scal[1], s);
}


// outlier test of src/Optimizer.cpp:384 / :411: chi2 of the LAST COMPUTED error (g2o keeps e->_error from the last
// computeActiveErrors, which may belong to a rejected trial) or non-positive depth in the rig frame
__global__ void __launch_bounds__(256) k_ba_classify(BADev D, uint8_t* flag) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D.nE) return;
    double Xc[3];
    pose_map(D.pose[D.eKF[e]], D.X + 3 * D.eMP[e], Xc);
    const double chi = D.obs[3 * e + 2] * (D.err[2 * e] * D.err[2 * e] + D.err[2 * e + 1] * D.err[2 * e + 1]);
    flag[e] = (chi > 5.991 || !(Xc[2] > 0.0)) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------- PoseOptimization
// One CTA per frame; the whole schedule of src/Optimizer.cpp:138-181 (4 rounds x optimize(10), dense 6x6 LM) runs inside
// the kernel. Block reductions of chi2 and of the 27 unique entries of (H,b); thread 0 does the 6x6 LDL^T and the LM logic.
struct PoseOptArgs {
    const int* offset; float* Tcw; const float* Xw; const float* kpxy; const float* invSigma2;
    int faceW, faceH; uint8_t* outlier; int32_t* inliers; double* pose64; double* err; uint8_t* level;
};

__device__ double po_block_sum(double v, double* sh) {
    const double s = block_sum(v, sh);
    __shared__ double bc;
    if (threadIdx.x == 0) bc = s;
    __syncthreads();
    const double r = bc;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) k_pose_opt(PoseOptArgs A) {
    __shared__ double sh[8];
    __shared__ double Hs[27];
    __shared__ Pose pose, pose0, backup;
    __shared__ double x[6], lambda, ni;
    __shared__ int ctrl;   // LM control word broadcast by thread 0
    const int f = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int beg = A.offset[f], n = A.offset[f + 1] - beg;
    const double fc = A.faceW / 2.0;
    const float dlt = (float)sqrt(5.991); const double delta = (double)dlt, dsqr = delta * delta;
    if (tid == 0) { pose0 = pose_from_Tcw32(A.Tcw + 16 * f); pose = pose0; }
    for (int i = tid; i < n; i += T) { A.outlier[beg + i] = 0; A.level[beg + i] = 0; A.err[2 * (beg + i)] = 0; A.err[2 * (beg + i) + 1] = 0; }
    __syncthreads();
    if (n < 3) { if (tid == 0) { A.inliers[f] = 0; if (A.pose64) { for (int i = 0; i < 3; i++) A.pose64[7 * f + i] = pose0.t[i]; for (int i = 0; i < 4; i++) A.pose64[7 * f + 3 + i] = pose0.q[i]; } } return; }
    auto obs_face = [&](int e, double& mx, double& my, double& w) -> int {
        const float kx = A.kpxy[2 * e], ky = A.kpxy[2 * e + 1];
        const float fi = kx / (float)A.faceW, fj = ky / (float)A.faceH;
        int face = -1;
        if (fi >= 0 && fi < 1 && fj >= 1 && fj < 2) face = 1;
        else if (fi >= 1 && fi < 2 && fj >= 0 && fj < 1) face = 3;
        else if (fi >= 1 && fi < 2 && fj >= 1 && fj < 2) face = 0;
        else if (fi >= 1 && fi < 2 && fj >= 2 && fj < 3) face = 4;
        else if (fi >= 2 && fi < 3 && fj >= 1 && fj < 2) face = 2;
        mx = (double)kx - floor((double)kx / A.faceW) * A.faceW; my = (double)ky - floor((double)ky / A.faceH) * A.faceH;
        w = (double)A.invSigma2[e];
        return face;
    };
    auto compute_errors = [&](bool robust, bool onlyActive) -> double {   // returns activeRobustChi2
        double c = 0;
        for (int i = tid; i < n; i += T) {
            const int e = beg + i;
            if (onlyActive && A.level[e] != 0) continue;
            double mx, my, w; const int face = obs_face(e, mx, my, w);
            const double Xw[3] = {(double)A.Xw[3 * e], (double)A.Xw[3 * e + 1], (double)A.Xw[3 * e + 2]};
            double Xc[3], er[2];
            pose_map(pose, Xw, Xc); edge_error(face, fc, mx, my, Xc, er);
            A.err[2 * e] = er[0]; A.err[2 * e + 1] = er[1];
            const double chi = w * (er[0] * er[0] + er[1] * er[1]);
            if (robust) { double r0, r1; huber(delta, dsqr, chi, r0, r1); c += r0; } else c += chi;
        }
        return po_block_sum(c, sh);
    };
    int nBadEdges = 0;
    bool robust = true;
    for (int it = 0; it < 4; it++) {
        if (tid == 0) pose = pose0;
        __syncthreads();
        int nAct = 0;
        for (int i = tid; i < n; i += T) nAct += A.level[beg + i] == 0;
        nAct = (int)po_block_sum((double)nAct, sh);
        if (nAct > 0) {
            int nBad = 0;
            for (int iter = 0; iter < 10; iter++) {
                double currentChi = compute_errors(robust, true);
                const double iniChi = currentChi;
                // buildSystem
                double acc[27];
                for (int k = 0; k < 27; k++) acc[k] = 0;
                for (int i = tid; i < n; i += T) {
                    const int e = beg + i;
                    if (A.level[e] != 0) continue;
                    double mx, my, w0; const int face = obs_face(e, mx, my, w0);
                    const double Xw[3] = {(double)A.Xw[3 * e], (double)A.Xw[3 * e + 1], (double)A.Xw[3 * e + 2]};
                    double Xc[3], G[2][3], Jp[2][6];
                    pose_map(pose, Xw, Xc); edge_G(face, fc, Xc, G); edge_Jpose(G, Xc, Jp);
                    const double e0 = A.err[2 * e], e1 = A.err[2 * e + 1];
                    double rho1 = 1.0;
                    if (robust) { double r0; huber(delta, dsqr, w0 * (e0 * e0 + e1 * e1), r0, rho1); }
                    int k = 0;
                    for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++) acc[k++] += (rho1 * w0) * (Jp[0][a] * Jp[0][b] + Jp[1][a] * Jp[1][b]);
                    for (int a = 0; a < 6; a++) acc[21 + a] -= rho1 * (Jp[0][a] * w0 * e0 + Jp[1][a] * w0 * e1);
                }
                for (int k = 0; k < 27; k++) { const double s = block_sum(acc[k], sh); if (tid == 0) Hs[k] = s; }
                __syncthreads();
                if (tid == 0 && iter == 0) {
                    double m = 0; int k = 0;
                    for (int a = 0; a < 6; a++) { m = fmax(m, fabs(Hs[k])); k += 6 - a; }
                    lambda = 1e-5 * m; ni = 2; 
                }
                if (iter == 0) nBad = 0;
                __syncthreads();
                double rho = 0; int qmax = 0;
                do {
                    if (tid == 0) {
                        backup = pose;
                        double M[6][6], d[6], b[6]; int k = 0; bool ok = true;
                        for (int a = 0; a < 6; a++) for (int c = a; c < 6; c++) { M[a][c] = Hs[k]; M[c][a] = Hs[k]; k++; }
                        for (int a = 0; a < 6; a++) { M[a][a] += lambda; b[a] = Hs[21 + a]; }
                        for (int j = 0; j < 6 && ok; j++) {   // LDL^T, LinearSolverDense: fails unless positive (LDLT::isPositive)
                            double dj = M[j][j];
                            for (int c = 0; c < j; c++) dj -= M[j][c] * M[j][c] * d[c];
                            if (!(dj > 0.0) || !isfinite(dj)) { ok = false; break; }
                            d[j] = dj;
                            for (int i2 = j + 1; i2 < 6; i2++) { double s = M[i2][j]; for (int c = 0; c < j; c++) s -= M[i2][c] * M[j][c] * d[c]; M[i2][j] = s / dj; }
                        }
                        if (ok) {
                            for (int i2 = 0; i2 < 6; i2++) { double s = b[i2]; for (int c = 0; c < i2; c++) s -= M[i2][c] * x[c]; x[i2] = s; }
                            for (int i2 = 0; i2 < 6; i2++) x[i2] /= d[i2];
                            for (int i2 = 5; i2 >= 0; i2--) { double s = x[i2]; for (int c = i2 + 1; c < 6; c++) s -= M[c][i2] * x[c]; x[i2] = s; }
                            pose = pose_oplus(pose, x);
                        } else { for (int i2 = 0; i2 < 6; i2++) x[i2] = 0; }
                        ctrl = ok ? 1 : 0;
                    }
                    __syncthreads();
                    const bool ok2 = ctrl != 0;
                    double tempChi = compute_errors(robust, true);
                    if (!ok2) tempChi = DBL_MAX;
                    double scale = 0;
                    for (int a = 0; a < 6; a++) scale += x[a] * (lambda * x[a] + Hs[21 + a]);
                    scale += 1e-3;
                    rho = (currentChi - tempChi) / scale;
                    const bool good = rho > 0 && isfinite(tempChi);
                    __syncthreads();
                    if (tid == 0) {
                        if (good) { double alpha = 1. - pow(2 * rho - 1, 3); alpha = fmin(alpha, 2. / 3.); lambda *= fmax(1. / 3., alpha); ni = 2; }
                        else { lambda *= ni; ni *= 2; pose = backup; }
                    }
                    if (good) currentChi = tempChi;
                    __syncthreads();
                    qmax++;
                } while (rho < 0 && qmax < 10);
                if (qmax == 10 || rho == 0) break;
                if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
                if (nBad >= 3) break;
            }
        }
        // classification (src/Optimizer.cpp:149-177): outliers get a fresh error, inliers keep the last computed one
        int bad = 0;
        for (int i = tid; i < n; i += T) {
            const int e = beg + i;
            double mx, my, w; const int face = obs_face(e, mx, my, w);
            if (A.outlier[e]) {
                const double Xw[3] = {(double)A.Xw[3 * e], (double)A.Xw[3 * e + 1], (double)A.Xw[3 * e + 2]};
                double Xc[3], er[2];
                pose_map(pose, Xw, Xc); edge_error(face, fc, mx, my, Xc, er);
                A.err[2 * e] = er[0]; A.err[2 * e + 1] = er[1];
            }
            const float chi = (float)(w * (A.err[2 * e] * A.err[2 * e] + A.err[2 * e + 1] * A.err[2 * e + 1]));
            if (chi > 5.991f) { A.outlier[e] = 1; A.level[e] = 1; bad++; } else { A.outlier[e] = 0; A.level[e] = 0; }
        }
        nBadEdges = (int)po_block_sum((double)bad, sh);
        if (it == 2) robust = false;
        if (n < 10) break;
    }
    if (tid == 0) {
        A.inliers[f] = n - nBadEdges;
        pose_to_Tcw32(pose, A.Tcw + 16 * f);
        if (A.pose64) { for (int i = 0; i < 3; i++) A.pose64[7 * f + i] = pose.t[i]; for (int i = 0; i < 4; i++) A.pose64[7 * f + 3 + i] = pose.q[i]; }
    }
}

}  // namespace cslam

// =================================================================================================== host side
using namespace cslam;

// ---- NCCL through dlopen (libnccl.so.2 is already mapped when torch.distributed is in use); no link-time dependency
typedef struct { char internal[128]; } nccl_uid_t;
typedef void* nccl_comm_t;
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(nccl_uid_t*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi* nccl_api() {
    static NcclApi api; static bool tried = false;
    if (tried) return api.h ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.h) break; }
    if (!api.h) return nullptr;
    api.GetUniqueId = (int (*)(nccl_uid_t*))dlsym(api.h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_uid_t, int))dlsym(api.h, "ncclCommInitRank");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(api.h, "ncclAllReduce");
    api.AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t))dlsym(api.h, "ncclAllGather");
    api.CommDestroy = (int (*)(nccl_comm_t))dlsym(api.h, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(api.h, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) { dlclose(api.h); api.h = nullptr; return nullptr; }
    return &api;
}
enum { NCCL_F64 = 8, NCCL_SUM = 0, NCCL_MAX = 2 };

struct cslam_optimizer {
    int device = 0;
    cudaStream_t stream = nullptr;
    nccl_comm_t comm = nullptr; int rank = 0, nranks = 1;
    int64_t launches = 0;
    // device arena: chunks are kept across calls (cudaMalloc/cudaFree per BA call cost more than the solve itself)
    struct Chunk { char* p; size_t size, used; };
    std::vector<Chunk> chunks;
    double* h_scal = nullptr;  // pinned
};

extern "C" int cslam_optimizer_create(cslam_optimizer** out, int device) {
    if (!out) return CSLAM_E_BADARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device (this library has no CPU fallback)"); return CSLAM_E_NODEVICE; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return CSLAM_E_BADARG; }
    CSLAM_CUDA(cudaSetDevice(device));
    cslam_optimizer* o = new cslam_optimizer; o->device = device;
    if (cudaStreamCreateWithFlags(&o->stream, cudaStreamNonBlocking) != cudaSuccess || cudaMallocHost(&o->h_scal, 8 * sizeof(double)) != cudaSuccess) {
        set_error("optimizer: stream / pinned allocation failed"); delete o; return CSLAM_E_CUDA;
    }
    cudaFuncSetAttribute(k_ba_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_ba_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_ba_linearize, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    *out = o;
    return CSLAM_OK;
}
static void free_pool(cslam_optimizer* o) { for (auto& c : o->chunks) c.used = 0; }
static void release_pool(cslam_optimizer* o) { for (auto& c : o->chunks) cudaFree(c.p); o->chunks.clear(); }
extern "C" void cslam_optimizer_destroy(cslam_optimizer* o) {
    if (!o) return;
    cudaSetDevice(o->device);
    if (o->stream) cudaStreamSynchronize(o->stream);
    release_pool(o);
    if (o->comm && nccl_api()) nccl_api()->CommDestroy(o->comm);
    if (o->stream) cudaStreamDestroy(o->stream);
    if (o->h_scal) cudaFreeHost(o->h_scal);
    delete o;
}
extern "C" int64_t cslam_optimizer_launches(const cslam_optimizer* o) { return o ? o->launches : 0; }

extern "C" int cslam_nccl_unique_id(uint8_t id128[128]) {
    NcclApi* a = nccl_api();
    if (!a) { set_error("libnccl.so.2 not found"); return CSLAM_E_NCCL; }
    nccl_uid_t id; int rc = a->GetUniqueId(&id);
    if (rc) { set_error("ncclGetUniqueId failed (%d)", rc); return CSLAM_E_NCCL; }
    std::memcpy(id128, id.internal, 128);
    return CSLAM_OK;
}
extern "C" int cslam_optimizer_init_nccl(cslam_optimizer* o, const uint8_t id128[128], int rank, int nranks) {
    if (!o || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return CSLAM_E_BADARG;
    NcclApi* a = nccl_api();
    if (!a) { set_error("libnccl.so.2 not found"); return CSLAM_E_NCCL; }
    CSLAM_CUDA(cudaSetDevice(o->device));
    nccl_uid_t id; std::memcpy(id.internal, id128, 128);
    int rc = a->CommInitRank(&o->comm, nranks, id, rank);
    if (rc) { set_error("ncclCommInitRank failed: %s", a->GetErrorString ? a->GetErrorString(rc) : "?"); return CSLAM_E_NCCL; }
    o->rank = rank; o->nranks = nranks;
    return CSLAM_OK;
}

template <class T>
static int dalloc(cslam_optimizer* o, T** p, size_t count, bool zero = false) {
    const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255;
    char* q = nullptr;
    for (auto& c : o->chunks) if (c.size - c.used >= bytes) { q = c.p + c.used; c.used += bytes; break; }
    if (!q) {
        cslam_optimizer::Chunk c; c.size = std::max<size_t>(bytes, (size_t)64 << 20); c.used = bytes;
        CSLAM_CUDA(cudaMalloc((void**)&c.p, c.size));
        o->chunks.push_back(c); q = c.p;
    }
    if (zero) CSLAM_CUDA(cudaMemsetAsync(q, 0, bytes, o->stream));
    *p = (T*)q;
    return 0;
}
template <class T>
static int dupload(cslam_optimizer* o, const T** p, const std::vector<T>& v) {
    T* q = nullptr; int rc = dalloc(o, &q, v.size());
    if (rc) return rc;
    if (!v.empty()) CSLAM_CUDA(cudaMemcpyAsync(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, o->stream));
    *p = q;
    return 0;
}

namespace {
struct BAHost {
    cslam_optimizer* o; BADev D; cslam_ba_result* res;
    std::vector<int> perm;            // sorted edge position -> caller's edge index
    std::vector<uint8_t> level; std::vector<int> poseIdx; std::vector<uint8_t> ptAct, fixed;
    int* d_poseIdx = nullptr; uint8_t* d_ptAct = nullptr; uint8_t* d_flag = nullptr;
    double lambda = -1, ni = 2; int nBad = 0; int iterations = 0, trials = 0;
    const volatile uint8_t* stop = nullptr;
    bool terminate() const { return stop ? (*stop != 0) : false; }
    int grid(int n) const { return std::max(1, cdiv(n, 256)); }

    int allreduce(double* buf, size_t count, int op) {
        if (o->nranks <= 1) return 0;
        int rc = nccl_api()->AllReduce(buf, buf, count, NCCL_F64, op, o->comm, o->stream);
        if (rc) { set_error("ncclAllReduce failed (%d)", rc); return CSLAM_E_NCCL; }
        return 0;
    }
    int fetch_scal() {   // device scal[0..3] -> host (with the cross-rank reductions the quantities need)
        CSLAM_CUDA(cudaMemcpyAsync(o->h_scal, D.scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, o->stream));
        CSLAM_CUDA(cudaStreamSynchronize(o->stream));
        return 0;
    }
    int zero_scal(int i, int n) { CSLAM_CUDA(cudaMemsetAsync(D.scal + i, 0, n * sizeof(double), o->stream)); return 0; }

    // initializeOptimization(level 0): active edges, active vertices, compact pose indices (sparse_optimizer.cpp:206-267,166-190)
    int initialize() {
        std::vector<uint8_t> pAct(D.nKF, 0);
        std::fill(ptAct.begin(), ptAct.end(), 0);
        const int* eMP = h_eMP.data(); const int* eKF = h_eKF.data();
        for (int e = 0; e < D.nE; e++) if (level[e] == 0) { pAct[eKF[e]] = 1; ptAct[eMP[e]] = 1; }
        int nP = 0;
        for (int k = 0; k < D.nKF; k++) poseIdx[k] = (pAct[k] && !fixed[k]) ? nP++ : -1;
        D.nP = nP; D.n = 6 * nP;
        CSLAM_CUDA(cudaMemcpyAsync(d_poseIdx, poseIdx.data(), D.nKF * sizeof(int), cudaMemcpyHostToDevice, o->stream));
        CSLAM_CUDA(cudaMemcpyAsync(d_ptAct, ptAct.data(), D.nMP, cudaMemcpyHostToDevice, o->stream));
        CSLAM_CUDA(cudaMemcpyAsync(D.level, level.data(), D.nE, cudaMemcpyHostToDevice, o->stream));
        nActive = 0; for (int e = 0; e < D.nE; e++) nActive += level[e] == 0;
        return 0;
    }
    std::vector<int> h_eMP, h_eKF; int nActive = 0;

    int errors_chi2(double* chi) {   // computeActiveErrors + activeRobustChi2
        int rc = zero_scal(0, 1); if (rc) return rc;
        k_ba_errors<<<grid(D.nE), 256, 0, o->stream>>>(D); o->launches++;
        if ((rc = allreduce(D.scal, 1, NCCL_SUM))) return rc;
        if ((rc = fetch_scal())) return rc;
        *chi = o->h_scal[0];
        return 0;
    }
    int build_system() {
        CSLAM_CUDA(cudaMemsetAsync(D.Hpp, 0, (size_t)std::max(D.nP, 1) * 36 * 8, o->stream));
        CSLAM_CUDA(cudaMemsetAsync(D.bp, 0, (size_t)std::max(D.nP, 1) * 6 * 8, o->stream));
        CSLAM_CUDA(cudaMemsetAsync(D.Hll, 0, (size_t)D.nMP * 9 * 8, o->stream));
        CSLAM_CUDA(cudaMemsetAsync(D.bl, 0, (size_t)D.nMP * 3 * 8, o->stream));
        const int useSmem = D.nP > 0 && D.nP <= 160;
        k_ba_linearize<<<grid(D.nE), 256, useSmem ? (size_t)D.nP * 42 * 8 : 0, o->stream>>>(D, useSmem); o->launches++;
        CSLAM_CUDA(cudaGetLastError());
        // landmark sharding: every rank linearised only the edges of its own landmarks -> the pose blocks are partial sums
        int rc;
        if ((rc = allreduce(D.Hpp, (size_t)D.nP * 36, NCCL_SUM)) || (rc = allreduce(D.bp, (size_t)D.nP * 6, NCCL_SUM))) return rc;
        return 0;
    }
    int lambda_init(double* lam) {
        int rc = zero_scal(2, 1); if (rc) return rc;
        k_ba_maxdiag<<<grid(D.nP * 6 + D.nMP * 3), 256, 0, o->stream>>>(D); o->launches++;   // poses: full Hpp (see build_system); landmarks: own
        if ((rc = allreduce(D.scal + 2, 1, NCCL_MAX))) return rc;
        if ((rc = fetch_scal())) return rc;
        *lam = 1e-5 * o->h_scal[2];
        return 0;
    }
    // one LM trial's linear solve (BlockSolver::solve): returns ok flag
    int solve(bool* ok) {
        int rc;
        k_ba_dinv<<<grid(D.nMP), 256, 0, o->stream>>>(D, lambda); o->launches++;
        if (D.n > 0) {
            // Hpp / bp are already full sums on every rank: rank 0 alone seeds [S | g | bpr] with them, the others start from 0
            if (o->rank == 0) { k_ba_s_init<<<grid(D.n * D.n), 256, 0, o->stream>>>(D); o->launches++; }
            else CSLAM_CUDA(cudaMemsetAsync(D.S, 0, ((size_t)D.n * D.n + 2 * D.n) * 8, o->stream));
            const int chunks = 16;
            k_ba_schur<<<dim3(chunks, D.nKF), 256, (size_t)(6 * D.n + 6) * 8, o->stream>>>(D, chunks); o->launches++;
            if ((rc = allreduce(D.S, (size_t)D.n * D.n + 2 * D.n, NCCL_SUM))) return rc;
            k_ba_add_lambda<<<grid(D.n), 256, 0, o->stream>>>(D, lambda); o->launches++;
            if ((rc = zero_scal(3, 1))) return rc;
            const size_t smem = (2 * (size_t)(D.n + 1) * (LD_NB + 1) + 2 * D.n) * 8;
            if (smem > 200 * 1024) { set_error("reduced camera system too large for the single-CTA solver (n=%d)", D.n); return CSLAM_E_CAPACITY; }
            k_ba_solve<<<1, 1024, smem, o->stream>>>(D); o->launches++;
        }
        k_ba_backsub<<<grid(D.nMP), 256, 0, o->stream>>>(D); o->launches++;
        CSLAM_CUDA(cudaGetLastError());
        *ok = true;   // the flag itself is read together with chi2/scale after the update (one D2H per trial)
        return 0;
    }
    // OptimizationAlgorithmLevenberg::solve ; result 0 OK, 1 Terminate
    int lm_iteration(int iteration, int* result) {
        int rc; double currentChi = 0;
        if ((rc = errors_chi2(&currentChi))) return rc;
        const double iniChi = currentChi; double tempChi = currentChi;
        if ((rc = build_system())) return rc;
        if (iteration == 0) { if ((rc = lambda_init(&lambda))) return rc; ni = 2; nBad = 0; }
        double rho = 0; int qmax = 0; int accepted = 0;
        do {
            CSLAM_CUDA(cudaMemcpyAsync(D.poseBak, D.pose, D.nKF * sizeof(Pose), cudaMemcpyDeviceToDevice, o->stream));
            CSLAM_CUDA(cudaMemcpyAsync(D.Xbak, D.X, (size_t)D.nMP * 3 * 8, cudaMemcpyDeviceToDevice, o->stream));
            bool ok2 = true;
            if ((rc = solve(&ok2))) return rc;
            k_ba_update<<<grid(std::max(D.nKF, D.nMP)), 256, 0, o->stream>>>(D); o->launches++;
            if ((rc = zero_scal(0, 2))) return rc;
            k_ba_errors<<<grid(D.nE), 256, 0, o->stream>>>(D); o->launches++;
            k_ba_scale<<<grid(D.n + 3 * D.nMP), 256, 0, o->stream>>>(D, lambda); o->launches++;
            if ((rc = allreduce(D.scal, 2, NCCL_SUM))) return rc;
            if ((rc = fetch_scal())) return rc;
            tempChi = o->h_scal[0];
            ok2 = o->h_scal[3] == 0.0;
            trials++;
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            const double scale = o->h_scal[1] + 1e-3;
            rho /= scale;
            if (rho > 0 && std::isfinite(tempChi) && ok2) {
                double alpha = 1. - std::pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                lambda *= std::max(1. / 3., alpha); ni = 2; currentChi = tempChi; accepted = 1;
            } else {
                lambda *= ni; ni *= 2;
                CSLAM_CUDA(cudaMemcpyAsync(D.pose, D.poseBak, D.nKF * sizeof(Pose), cudaMemcpyDeviceToDevice, o->stream));
                CSLAM_CUDA(cudaMemcpyAsync(D.X, D.Xbak, (size_t)D.nMP * 3 * 8, cudaMemcpyDeviceToDevice, o->stream));
                if (!ok2) rho = -1;
            }
            qmax++;
        } while (rho < 0 && qmax < 10 && !terminate());
        if (res && res->lm_log && iterations < res->log_cap) {
            double* L = res->lm_log + 4 * iterations; L[0] = currentChi; L[1] = lambda; L[2] = qmax; L[3] = accepted;
        }
        iterations++;
        *result = 0;
        if (qmax == 10 || rho == 0) { *result = 1; return 0; }
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
        if (nBad >= 3) *result = 1;
        return 0;
    }
    int optimize(int its) {
        if (nActive == 0) return 0;   // g2o: "0 vertices to optimize"
        int result = 0;
        for (int i = 0; i < its && !terminate() && result == 0; i++) { int rc = lm_iteration(i, &result); if (rc) return rc; }
        return 0;
    }
    double* d_flag64 = nullptr;
    int classify(std::vector<uint8_t>& flags) {
        k_ba_classify<<<grid(D.nE), 256, 0, o->stream>>>(D, d_flag); o->launches++;
        flags.resize(D.nE);
        CSLAM_CUDA(cudaMemcpyAsync(flags.data(), d_flag, D.nE, cudaMemcpyDeviceToHost, o->stream));
        CSLAM_CUDA(cudaStreamSynchronize(o->stream));
        if (o->nranks > 1) {   // a rank only holds valid errors / points for the edges of its own landmarks: combine the owners' flags
            std::vector<double> fl(D.nE);
            for (int s = 0; s < D.nE; s++) fl[s] = ((h_eMP[s] % o->nranks) == o->rank && flags[s]) ? 1.0 : 0.0;
            CSLAM_CUDA(cudaMemcpyAsync(d_flag64, fl.data(), (size_t)D.nE * 8, cudaMemcpyHostToDevice, o->stream));
            int rc = allreduce(d_flag64, D.nE, NCCL_SUM);
            if (rc) return rc;
            CSLAM_CUDA(cudaMemcpyAsync(fl.data(), d_flag64, (size_t)D.nE * 8, cudaMemcpyDeviceToHost, o->stream));
            CSLAM_CUDA(cudaStreamSynchronize(o->stream));
            for (int s = 0; s < D.nE; s++) flags[s] = fl[s] != 0.0;
        }
        return 0;
    }
};
}  // namespace

extern "C" int cslam_local_ba(cslam_optimizer* o, cslam_ba_problem* p, const volatile uint8_t* stop_flag, int its1, int its2, cslam_ba_result* r) {
    if (!o || !p || p->n_kf <= 0 || p->n_mp < 0 || p->n_edges < 0 || !p->Tcw || !p->kf_fixed || (p->n_mp && !p->points)) { set_error("cslam_local_ba: bad problem"); return CSLAM_E_BADARG; }
    if (p->face_w != p->face_h || p->face_w <= 0) { set_error("cube faces must be square"); return CSLAM_E_BADARG; }
    CSLAM_CUDA(cudaSetDevice(o->device));
    free_pool(o);
    const int nKF = p->n_kf, nMP = p->n_mp, nE = p->n_edges;
    if (r) { r->iterations = 0; r->trials = 0; if (r->outlier) std::memset(r->outlier, 0, nE); }
    if (stop_flag && *stop_flag) return CSLAM_OK;   // src/Optimizer.cpp:359-361
    BAHost H; H.o = o; H.res = r; H.stop = stop_flag;
    BADev& D = H.D; std::memset(&D, 0, sizeof(D));
    D.nKF = nKF; D.nMP = nMP; D.nE = nE; D.f = p->face_w / 2.0; D.rank = o->rank; D.nranks = o->nranks;
    const float dlt = (float)std::sqrt(5.991); D.delta = (double)dlt; D.dsqr = D.delta * D.delta; D.robust = 1;
    // ---- edges sorted by landmark (stable), CSR by landmark and by keyframe
    H.perm.resize(nE);
    std::vector<int> cnt(nMP + 1, 0);
    for (int e = 0; e < nE; e++) {
        if (p->edge_mp[e] < 0 || p->edge_mp[e] >= nMP || p->edge_kf[e] < 0 || p->edge_kf[e] >= nKF) { set_error("edge %d references a vertex out of range", e); return CSLAM_E_BADARG; }
        cnt[p->edge_mp[e] + 1]++;
    }
    for (int l = 0; l < nMP; l++) cnt[l + 1] += cnt[l];
    std::vector<int> lmStart(cnt), fill(cnt.begin(), cnt.end() - 1);
    for (int e = 0; e < nE; e++) H.perm[fill[p->edge_mp[e]]++] = e;
    H.h_eMP.resize(nE); H.h_eKF.resize(nE);
    std::vector<double> obs((size_t)nE * 3); std::vector<int8_t> face(nE);
    for (int s = 0; s < nE; s++) {
        const int e = H.perm[s];
        H.h_eMP[s] = p->edge_mp[e]; H.h_eKF[s] = p->edge_kf[e];
        const float kx = p->kp_xy[2 * e], ky = p->kp_xy[2 * e + 1];
        const float fi = kx / (float)p->face_w, fj = ky / (float)p->face_h;   // FaceInCubemap(cv::Point2f)
        int f = -1;
        if (fi >= 0 && fi < 1 && fj >= 1 && fj < 2) f = 1;
        else if (fi >= 1 && fi < 2 && fj >= 0 && fj < 1) f = 3;
        else if (fi >= 1 && fi < 2 && fj >= 1 && fj < 2) f = 0;
        else if (fi >= 1 && fi < 2 && fj >= 2 && fj < 3) f = 4;
        else if (fi >= 2 && fi < 3 && fj >= 1 && fj < 2) f = 2;
        if (f < 0) { set_error("edge %d: keypoint (%g,%g) is on no cube face (the reference calls exit() here)", e, kx, ky); return CSLAM_E_BADARG; }
        face[s] = (int8_t)f;
        obs[3 * s] = (double)kx - std::floor((double)kx / p->face_w) * p->face_w;     // GetPosInFace<double>
        obs[3 * s + 1] = (double)ky - std::floor((double)ky / p->face_h) * p->face_h;
        obs[3 * s + 2] = (double)p->inv_sigma2[e];
    }
    std::vector<int> peStart(nKF + 1, 0), peList(nE);
    for (int s = 0; s < nE; s++) peStart[H.h_eKF[s] + 1]++;
    for (int k = 0; k < nKF; k++) peStart[k + 1] += peStart[k];
    { std::vector<int> f2(peStart.begin(), peStart.end() - 1); for (int s = 0; s < nE; s++) peList[f2[H.h_eKF[s]]++] = s; }
    std::vector<Pose> poses(nKF);
    for (int k = 0; k < nKF; k++) poses[k] = pose_from_Tcw32(p->Tcw + 16 * k);
    std::vector<double> X((size_t)nMP * 3);
    for (size_t i = 0; i < X.size(); i++) X[i] = (double)p->points[i];
    H.fixed.assign(p->kf_fixed, p->kf_fixed + nKF); H.level.assign(nE, 0); H.poseIdx.assign(nKF, -1); H.ptAct.assign(nMP, 0);
    int rc;
    const int nPmax = nKF, nmax = 6 * nPmax;
    const Pose* cpose = nullptr; const double* cX = nullptr;
    if ((rc = dupload(o, &cpose, poses)) || (rc = dupload(o, &cX, X)) || (rc = dupload(o, &D.eMP, H.h_eMP)) || (rc = dupload(o, &D.eKF, H.h_eKF)) ||
        (rc = dupload(o, &D.obs, obs)) || (rc = dupload(o, &D.face, face)) || (rc = dupload(o, &D.lmStart, lmStart)) || (rc = dupload(o, &D.peStart, peStart)) ||
        (rc = dupload(o, &D.peList, peList))) return rc;
    D.pose = const_cast<Pose*>(cpose); D.X = const_cast<double*>(cX);
    double* red = nullptr;
    if ((rc = dalloc(o, &D.poseBak, nKF)) || (rc = dalloc(o, &D.Xbak, (size_t)nMP * 3)) || (rc = dalloc(o, &D.err, (size_t)nE * 2, true)) || (rc = dalloc(o, &D.level, nE)) ||
        (rc = dalloc(o, &H.d_poseIdx, nKF)) || (rc = dalloc(o, &H.d_ptAct, nMP)) || (rc = dalloc(o, &H.d_flag, nE)) || (rc = dalloc(o, &D.Hpp, (size_t)nPmax * 36)) ||
        (rc = dalloc(o, &D.bp, (size_t)nPmax * 6)) || (rc = dalloc(o, &D.Hll, (size_t)nMP * 9)) || (rc = dalloc(o, &D.bl, (size_t)nMP * 3)) ||
        (rc = dalloc(o, &D.Hpl, (size_t)nE * 18)) || (rc = dalloc(o, &D.Dinv, (size_t)nMP * 9)) || (rc = dalloc(o, &D.db, (size_t)nMP * 3)) ||
        (rc = dalloc(o, &red, (size_t)nmax * nmax + 2 * nmax)) || (rc = dalloc(o, &D.xp, nmax, true)) || (rc = dalloc(o, &D.xl, (size_t)nMP * 3, true)) ||
        (rc = dalloc(o, &D.scal, 8, true)) || (rc = dalloc(o, &H.d_flag64, o->nranks > 1 ? nE : 1))) return rc;
    D.poseIdx = H.d_poseIdx; D.ptAct = H.d_ptAct;
    auto bind_reduce = [&]() { D.S = red; D.g = red + (size_t)D.n * D.n; D.bpr = D.g + D.n; };
    // ---- src/Optimizer.cpp:363-395
    if ((rc = H.initialize())) return rc;
    bind_reduce();
    if ((rc = H.optimize(its1))) return rc;
    std::vector<uint8_t> flags;
    if (!H.terminate()) {
        if ((rc = H.classify(flags))) return rc;
        for (int s = 0; s < nE; s++) if (flags[s]) H.level[s] = 1;
        D.robust = 0;
        if ((rc = H.initialize())) return rc;
        bind_reduce();
        if ((rc = H.optimize(its2))) return rc;
    }
    if ((rc = H.classify(flags))) return rc;
    // ---- write back (src/Optimizer.cpp:432-450): float32 poses / points; with landmark sharding every rank holds its own points
    if (o->nranks > 1) {
        // non-owned points were never touched; owners' values are combined with a sum of (owner ? X : 0)
        std::vector<double> Xh((size_t)nMP * 3);
        CSLAM_CUDA(cudaMemcpyAsync(Xh.data(), D.X, Xh.size() * 8, cudaMemcpyDeviceToHost, o->stream));
        CSLAM_CUDA(cudaStreamSynchronize(o->stream));
        for (int l = 0; l < nMP; l++) if ((l % o->nranks) != o->rank) { Xh[3 * l] = 0; Xh[3 * l + 1] = 0; Xh[3 * l + 2] = 0; }
        CSLAM_CUDA(cudaMemcpyAsync(D.X, Xh.data(), Xh.size() * 8, cudaMemcpyHostToDevice, o->stream));
        if ((rc = H.allreduce(D.X, (size_t)nMP * 3, NCCL_SUM))) return rc;
    }
    CSLAM_CUDA(cudaMemcpyAsync(poses.data(), D.pose, nKF * sizeof(Pose), cudaMemcpyDeviceToHost, o->stream));
    CSLAM_CUDA(cudaMemcpyAsync(X.data(), D.X, X.size() * 8, cudaMemcpyDeviceToHost, o->stream));
    CSLAM_CUDA(cudaStreamSynchronize(o->stream));
    for (int k = 0; k < nKF; k++) {
        pose_to_Tcw32(poses[k], p->Tcw + 16 * k);
        if (r && r->pose_fp64) { double* q = r->pose_fp64 + 7 * k; q[0] = poses[k].t[0]; q[1] = poses[k].t[1]; q[2] = poses[k].t[2]; q[3] = poses[k].q[0]; q[4] = poses[k].q[1]; q[5] = poses[k].q[2]; q[6] = poses[k].q[3]; }
    }
    for (size_t i = 0; i < X.size(); i++) { p->points[i] = (float)X[i]; if (r && r->points_fp64) r->points_fp64[i] = X[i]; }
    if (r) {
        if (r->outlier) for (int s = 0; s < nE; s++) r->outlier[H.perm[s]] = flags[s];
        r->iterations = H.iterations; r->trials = H.trials;
    }
    free_pool(o);
    return CSLAM_OK;
}

extern "C" int cslam_pose_optimization(cslam_optimizer* o, int nframes, const int32_t* offset, float* Tcw, const float* Xw, const float* kp_xy, const float* inv_sigma2,
                                       int face_w, int face_h, uint8_t* outlier, int32_t* inliers, double* pose_fp64) {
    if (!o || nframes <= 0 || !offset || !Tcw || !inliers) { set_error("cslam_pose_optimization: bad argument"); return CSLAM_E_BADARG; }
    if (face_w != face_h || face_w <= 0) { set_error("cube faces must be square"); return CSLAM_E_BADARG; }
    CSLAM_CUDA(cudaSetDevice(o->device));
    free_pool(o);
    const int n = offset[nframes];
    PoseOptArgs A; std::memset(&A, 0, sizeof(A));
    int* d_off; float *d_T, *d_X, *d_kp, *d_w; uint8_t *d_out, *d_lvl; int32_t* d_inl; double *d_p64, *d_err;
    int rc;
    if ((rc = dalloc(o, &d_off, nframes + 1)) || (rc = dalloc(o, &d_T, (size_t)nframes * 16)) || (rc = dalloc(o, &d_X, (size_t)n * 3)) || (rc = dalloc(o, &d_kp, (size_t)n * 2)) ||
        (rc = dalloc(o, &d_w, n)) || (rc = dalloc(o, &d_out, n)) || (rc = dalloc(o, &d_lvl, n)) || (rc = dalloc(o, &d_inl, nframes)) || (rc = dalloc(o, &d_p64, (size_t)nframes * 7)) ||
        (rc = dalloc(o, &d_err, (size_t)n * 2))) return rc;
    CSLAM_CUDA(cudaMemcpyAsync(d_off, offset, (nframes + 1) * 4, cudaMemcpyHostToDevice, o->stream));
    CSLAM_CUDA(cudaMemcpyAsync(d_T, Tcw, (size_t)nframes * 64, cudaMemcpyHostToDevice, o->stream));
    if (n) {
        CSLAM_CUDA(cudaMemcpyAsync(d_X, Xw, (size_t)n * 12, cudaMemcpyHostToDevice, o->stream));
        CSLAM_CUDA(cudaMemcpyAsync(d_kp, kp_xy, (size_t)n * 8, cudaMemcpyHostToDevice, o->stream));
        CSLAM_CUDA(cudaMemcpyAsync(d_w, inv_sigma2, (size_t)n * 4, cudaMemcpyHostToDevice, o->stream));
    }
    A.offset = d_off; A.Tcw = d_T; A.Xw = d_X; A.kpxy = d_kp; A.invSigma2 = d_w; A.faceW = face_w; A.faceH = face_h; A.outlier = d_out; A.inliers = d_inl;
    A.pose64 = d_p64; A.err = d_err; A.level = d_lvl;
    k_pose_opt<<<nframes, 256, 0, o->stream>>>(A); o->launches++;
    CSLAM_CUDA(cudaGetLastError());
    // frames with < 3 correspondences are returned untouched (src/Optimizer.cpp:133-134): copy back only the others
    std::vector<float> Tout((size_t)nframes * 16);
    CSLAM_CUDA(cudaMemcpyAsync(Tout.data(), d_T, Tout.size() * 4, cudaMemcpyDeviceToHost, o->stream));
    if (n && outlier) CSLAM_CUDA(cudaMemcpyAsync(outlier, d_out, n, cudaMemcpyDeviceToHost, o->stream));
    CSLAM_CUDA(cudaMemcpyAsync(inliers, d_inl, nframes * 4, cudaMemcpyDeviceToHost, o->stream));
    if (pose_fp64) CSLAM_CUDA(cudaMemcpyAsync(pose_fp64, d_p64, (size_t)nframes * 56, cudaMemcpyDeviceToHost, o->stream));
    CSLAM_CUDA(cudaStreamSynchronize(o->stream));
    for (int f = 0; f < nframes; f++) if (offset[f + 1] - offset[f] >= 3) std::memcpy(Tcw + 16 * f, Tout.data() + 16 * f, 64);
    free_pool(o);
    return CSLAM_OK;
}
