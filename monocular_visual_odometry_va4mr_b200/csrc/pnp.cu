// P3P-RANSAC on the GPU (replaces cv2.solvePnPRansac(..., flags=SOLVEPNP_P3P) at reference
// VisualOdometryPipeLine.py:343; spec SURVEY.md A.6-A.8 and oracle/pnp_oracle.c header).
//
// All `iters` hypotheses are drawn, solved and scored at once; a sequential replay of the
// per-hypothesis inlier counts then reproduces OpenCV's "first strictly better model wins,
// shrink niters" loop, so the winner (and therefore the inlier mask) is the one cv2 returns.
//   1. ransac_samples_kernel<4> (ransac.cuh) bit-exact cv::RNG((uint64)-1) 4-subsets (raw stream table + parallel mod)
//   2. pnp_solve_kernel     one thread per hypothesis: FP64 P3P + 4-point disambiguation
//   3. pnp_score_kernel     (point tile x hypothesis tile): FP64 projection -> f32 error,
//                           warp-aggregated inlier counts (ballot+popc, one atomicAdd per warp)
//   4. pnp_select_kernel    replay + winner mask + ordered inlier compaction
//   5. pnp_epnp_kernel      EPnP refit on the inliers (block reductions + small FP64 solves)
// Every kernel is batched over independent sequences (blockIdx.z / blockIdx.x = sequence).
#include "internal.cuh"
#include "mathdev.cuh"
#include "pnp.cuh"
#include "ransac.cuh"

using namespace vo;

// ------------------------------------------------------------------------------------------
// 2. minimal solver: one thread per hypothesis
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
pnp_solve_kernel(PnpArgs a)
{
    const int b = blockIdx.y;
    const int it = a.h_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= a.h_end) return;
    if (!a.head && it >= a.need[b]) return;      // the sequential loop can no longer reach this hypothesis
    const size_t hidx = (size_t)b * a.iters + it;
    a.counts[hidx] = 0;
    a.hyp_ok[hidx] = 0;
    const int* smp = a.samples + hidx * 4;
    const int N = a.n[b];
    if (smp[0] < 0 || N < 4) return;
    const float* obj = a.obj + (size_t)b * a.cap * 3;
    const float* img = a.img + (size_t)b * a.cap * 2;
    double X[12], xn[8], uv[8];
    const double ifx = 1. / a.fx, ify = 1. / a.fy;
    for (int k = 0; k < 4; ++k) {
        const int s = smp[k];
        X[3 * k] = obj[3 * s]; X[3 * k + 1] = obj[3 * s + 1]; X[3 * k + 2] = obj[3 * s + 2];
        uv[2 * k] = img[2 * s]; uv[2 * k + 1] = img[2 * s + 1];
        // cv2 normalises with undistortPoints on CV_32F input: double arithmetic, float32 result
        xn[2 * k] = (double)(float)((uv[2 * k] - a.cx) * ifx);
        xn[2 * k + 1] = (double)(float)((uv[2 * k + 1] - a.cy) * ify);
    }
    double R[4][9], t[4][3];
    const int ns = p3p(X, xn, R, t);
    int best = -1;
    double best_e = 0;
    for (int s = 0; s < ns; ++s) {
        double e = 0;
        for (int k = 0; k < 4; ++k) {
            double u, v;
            project_pt(R[s], t[s], a.fx, a.fy, a.cx, a.cy, X[3 * k], X[3 * k + 1], X[3 * k + 2], u, v);
            const double dx = uv[2 * k] - u, dy = uv[2 * k + 1] - v;
            e += dx * dx + dy * dy;
        }
        if (!isfinite(e)) continue;
        if (best < 0 || e < best_e) { best = s; best_e = e; }
    }
    if (best < 0) return;
    // the model is stored as rvec|tvec and re-expanded for scoring (as PnPRansacCallback does)
    double rv[3], Rr[9];
    R_to_rodrigues(R[best], rv);
    rodrigues_to_R(rv, Rr);
    double* h = a.hyp + hidx * 12;
    for (int k = 0; k < 9; ++k) h[k] = Rr[k];
    for (int k = 0; k < 3; ++k) h[9 + k] = t[best][k];
    double* hr = a.hyp_rvec + hidx * 3;
    hr[0] = rv[0]; hr[1] = rv[1]; hr[2] = rv[2];
    a.hyp_ok[hidx] = 1;
}

// ------------------------------------------------------------------------------------------
// 3. scoring
// ------------------------------------------------------------------------------------------
#define SCORE_HT 32   // hypotheses per block
#define PNP_HEAD 32   // hypotheses solved and scored before the first look at the adaptive stop

__device__ __forceinline__ bool pnp_is_inlier(const double* h, double fx, double fy, double cx, double cy,
                                              double X, double Y, double Z, float iu, float iv, float thr_sq)
{
    double u, v;
    project_pt(h, h + 9, fx, fy, cx, cy, X, Y, Z, u, v);
    const float pu = (float)u, pv = (float)v;      // projectPoints output is float32
    const float dx = __fsub_rn(iu, pu), dy = __fsub_rn(iv, pv);
    const float e = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    return e <= thr_sq;                              // NaN compares false
}

// Head chunk bookkeeping: the LAST CTA of a sequence to finish (ticket counter, no waiting) replays
// cv2's loop over the head hypotheses; what it leaves in `niters` bounds every later replay, so
// the tail launches skip hypotheses >= need[b] (typically all of them: at 10 % outliers cv2 itself
// stops after ~5 iterations).
__device__ __forceinline__ void pnp_head_done(const PnpArgs& a, int b, int N)
{
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(a.ticket + b, 1) == (int)(gridDim.x * gridDim.y) - 1;
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    int niters = a.iters > 1 ? a.iters : 1;
    if (N == 4) niters = 1;
    else if (N > 4) {
        const size_t hb = (size_t)b * a.iters;
        int max_good = 0;
        for (int it = a.h_begin; it < a.h_end && it < niters; ++it) {
            if (!((volatile int*)a.hyp_ok)[hb + it]) continue;
            const int good = ((volatile int*)a.counts)[hb + it];
            if (good > (max_good > 3 ? max_good : 3)) {
                max_good = good;
                niters = ransac_update_num_iters(a.conf, (double)(N - good) / N, 4, niters);
            }
        }
    } else niters = 0;
    a.need[b] = niters;
}

__global__ void __launch_bounds__(256)
pnp_score_kernel(PnpArgs a)
{
    __shared__ double s_h[SCORE_HT * 12];
    __shared__ int s_ok[SCORE_HT];
    const int b = blockIdx.z;
    const int N = a.n[b];
    const int p0 = blockIdx.x * blockDim.x;
    const int h0 = a.h_begin + blockIdx.y * SCORE_HT;
    if (!a.head && (p0 >= N || N <= 4 || h0 >= a.need[b])) return;
    if (a.head && (p0 >= N || N <= 4)) { pnp_head_done(a, b, N); return; }
    const int nh = min(SCORE_HT, a.h_end - h0);
    const size_t hbase = (size_t)b * a.iters + h0;
    for (int k = threadIdx.x; k < nh * 12; k += blockDim.x) s_h[k] = a.hyp[hbase * 12 + k];
    if (threadIdx.x < nh) s_ok[threadIdx.x] = a.hyp_ok[hbase + threadIdx.x];
    __syncthreads();
    const int i = p0 + threadIdx.x;
    const bool live = i < N;
    double X = 0, Y = 0, Z = 0;
    float iu = 0, iv = 0;
    if (live) {
        const float* o = a.obj + ((size_t)b * a.cap + i) * 3;
        const float* m = a.img + ((size_t)b * a.cap + i) * 2;
        X = o[0]; Y = o[1]; Z = o[2];
        iu = m[0]; iv = m[1];
    }
    const int lane = threadIdx.x & 31;
    for (int h = 0; h < nh; ++h) {
        if (!s_ok[h]) continue;   // block-uniform
        const bool in = live && pnp_is_inlier(s_h + h * 12, a.fx, a.fy, a.cx, a.cy, X, Y, Z, iu, iv, a.thr_sq);
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (lane == 0 && m) atomicAdd(a.counts + hbase + h, __popc(m));
    }
    if (a.head) pnp_head_done(a, b, N);
}

// ------------------------------------------------------------------------------------------
// 4. select: replay cv2's sequential loop, then winner mask + ordered compaction
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pnp_select_kernel(PnpArgs a)
{
    __shared__ int s_win, s_run;
    __shared__ double s_h[12];
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int b = blockIdx.x;
    const int N = a.n[b];
    const size_t hb = (size_t)b * a.iters;
    if (threadIdx.x == 0) {
        int win = -1, run = 0;
        if (N == 4) { win = a.hyp_ok[hb] ? 0 : -1; run = 1; }
        else if (N > 4) {
            int niters = a.iters > 1 ? a.iters : 1, max_good = 0, it;
            for (it = 0; it < niters; ++it) {
                if (!a.hyp_ok[hb + it]) continue;
                const int good = a.counts[hb + it];
                if (good > (max_good > 3 ? max_good : 3)) {
                    win = it; max_good = good;
                    niters = ransac_update_num_iters(a.conf, (double)(N - good) / N, 4, niters);
                }
            }
            run = it;
        }
        s_win = win; s_run = run; s_base = 0;
    }
    __syncthreads();
    const int win = s_win;
    uint8_t* mask = a.mask + (size_t)b * a.cap;
    int* inl = a.inliers + (size_t)b * a.cap;
    if (threadIdx.x == 0) { a.winner[b] = win; a.iters_run[b] = s_run; }
    if (win < 0) {
        for (int i = threadIdx.x; i < a.cap; i += blockDim.x) mask[i] = 0;
        if (threadIdx.x == 0) { a.n_inliers[b] = 0; a.ok[b] = 0; }
        return;
    }
    if (threadIdx.x < 12) s_h[threadIdx.x] = a.hyp[(hb + win) * 12 + threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < a.cap; base += blockDim.x) {
        const int i = base + threadIdx.x;
        bool in = false;
        if (i < N) {
            if (N == 4) in = true;
            else {
                const float* o = a.obj + ((size_t)b * a.cap + i) * 3;
                const float* m = a.img + ((size_t)b * a.cap + i) * 2;
                in = pnp_is_inlier(s_h, a.fx, a.fy, a.cx, a.cy, o[0], o[1], o[2], m[0], m[1], a.thr_sq);
            }
        }
        const unsigned bm = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_warp[warp] = __popc(bm);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (in) inl[off + __popc(bm & ((1u << lane) - 1))] = i;
        if (i < a.cap) mask[i] = in ? 1 : 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.n_inliers[b] = s_base;
        a.ok[b] = 1;
        // RANSAC model (what cv2 falls back to / returns for N == 4)
        double* pose = a.pose + (size_t)b * 6;
        const double* hr = a.hyp_rvec + (hb + win) * 3;
        pose[0] = hr[0]; pose[1] = hr[1]; pose[2] = hr[2];
        pose[3] = s_h[9]; pose[4] = s_h[10]; pose[5] = s_h[11];
    }
}

// ------------------------------------------------------------------------------------------
// 5. EPnP refit on the inliers (OpenCV epnp.cpp algorithm, float64 inputs)
// ------------------------------------------------------------------------------------------
#define EPNP_T 256

// Symmetric 12x12 eigen-decomposition by two-sided Jacobi rotations, executed by ONE WARP with the
// round-robin (tournament) ordering: 11 rounds per sweep, 6 disjoint (p,q) pairs per round, all six
// rotations of a round applied together (columns, then rows, then the eigenvector columns).
// A (in/out, destroyed) and V (out, eigenvectors in columns) live in shared memory, row-major 12x12.
// Only used for EPnP's M^T M, where the result is independent of eigenvector signs and of the
// rotation order (unlike the 3x3 control-point SVD, which replays OpenCV's order exactly).
__device__ inline void warp_jacobi_eig12(double* A, double* V, double* cs /* [12] + pair table */, int lane)
{
    // pair table of the tournament: pq[rnd][pr] = p | q << 8 (p < q), built once
    unsigned short* pq = reinterpret_cast<unsigned short*>(cs + 12);
    for (int k = lane; k < 66; k += 32) {
        const int rnd = k / 6, pr = k - rnd * 6;
        int p = pr == 0 ? 11 : (rnd + pr) % 11;
        int q = (rnd + 11 - pr) % 11;
        if (p > q) { const int t = p; p = q; q = t; }
        pq[k] = (unsigned short)(p | (q << 8));
    }
    for (int k = lane; k < 144; k += 32) V[k] = (k / 12 == k % 12) ? 1.0 : 0.0;
    // this lane's fixed element slots: columns pass (A and V: 12 rows x 6 pairs x 2 matrices = 144), rows pass (12 x 6 = 72)
    int c_row[5], c_pr[5], r_k[3], r_pr[3];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int e = lane + 32 * i, m = e / 72, r = (e % 72) / 6;
        c_pr[i] = e % 6;
        c_row[i] = e < 144 ? m * 144 + r * 12 : -1;       // V follows A in shared memory (A + 144)
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int e = lane + 32 * i;
        r_pr[i] = e % 6;
        r_k[i] = e < 72 ? e / 6 : -1;
    }
    __syncwarp();
    for (int sweep = 0; sweep < 30; ++sweep) {
        // convergence: sum of squared off-diagonal entries relative to the diagonal
        double off = 0, dia = 0;
        for (int k = lane; k < 144; k += 32) {
            const double v = A[k];
            if (k / 12 == k % 12) dia += v * v; else off += v * v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); dia += __shfl_xor_sync(0xffffffffu, dia, o); }
        if (off <= 1e-30 * dia || off == 0) break;
        for (int rnd = 0; rnd < 11; ++rnd) {
            if (lane < 6) {
                const int pqv = pq[rnd * 6 + lane], p = pqv & 0xff, q = pqv >> 8;
                const double apq = A[p * 12 + q];
                double c = 1.0, sn = 0.0;
                if (fabs(apq) > 1e-300) {
                    const double theta = (A[q * 12 + q] - A[p * 12 + p]) / (2 * apq);
                    const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                    c = 1 / sqrt(t * t + 1); sn = t * c;
                }
                cs[2 * lane] = c; cs[2 * lane + 1] = sn;
            }
            __syncwarp();
            // columns p,q of A and of V
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                if (c_row[i] >= 0) {
                    const int pqv = pq[rnd * 6 + c_pr[i]], p = pqv & 0xff, q = pqv >> 8;
                    const double c = cs[2 * c_pr[i]], sn = cs[2 * c_pr[i] + 1];
                    double* M = A + c_row[i];
                    const double x = M[p], y = M[q];
                    M[p] = c * x - sn * y;
                    M[q] = sn * x + c * y;
                }
            }
            __syncwarp();
            // rows p,q of A
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (r_k[i] >= 0) {
                    const int pqv = pq[rnd * 6 + r_pr[i]], p = pqv & 0xff, q = pqv >> 8;
                    const double c = cs[2 * r_pr[i]], sn = cs[2 * r_pr[i] + 1];
                    const double x = A[p * 12 + r_k[i]], y = A[q * 12 + r_k[i]];
                    A[p * 12 + r_k[i]] = c * x - sn * y;
                    A[q * 12 + r_k[i]] = sn * x + c * y;
                }
            }
            __syncwarp();
        }
    }
}

template <int NV>
__device__ __forceinline__ void block_reduce_sum(double* v, double* s_red /* [8][NV] */, double* s_out /* [NV] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_red[warp * NV + k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double x = 0;
        for (int w = 0; w < EPNP_T / 32; ++w) x += s_red[w * NV + threadIdx.x];
        s_out[threadIdx.x] = x;
    }
    __syncthreads();
}

struct EpnpShared {
    double red[8 * 40];
    double out[40];
    double cws[4][3];
    double ci[9];       // inverse of the control-point basis
    double ut[144];     // rows: singular vectors of MtM, descending singular value
    double L[60], rho[6];
    double Rs[3][9], ts[3][3];
    double errs[3];
    double mom_aX[12];  // sum_i alpha_ij X_ik
    double mom_a[4];    // sum_i alpha_ij
    double alpha0[4];   // barycentric coords of the first inlier (solve_for_sign)
    double work[3 * 144 + 12];
};

__device__ inline void epnp_pose_from_betas(EpnpShared& S, const double* betas, double n, const double* Xbar, double* R, double* t)
{
    const double* v[4] = {S.ut + 12 * 11, S.ut + 12 * 10, S.ut + 12 * 9, S.ut + 12 * 8};
    double ccs[4][3];
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 3; ++k) {
            double s = 0;
            for (int j = 0; j < 4; ++j) s += betas[j] * v[j][3 * i + k];
            ccs[i][k] = s;
        }
    // solve_for_sign: z of the first point in the camera frame
    const double z0 = S.alpha0[0] * ccs[0][2] + S.alpha0[1] * ccs[1][2] + S.alpha0[2] * ccs[2][2] + S.alpha0[3] * ccs[3][2];
    if (z0 < 0)
        for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) ccs[i][k] = -ccs[i][k];
    // pc_i = sum_j alpha_ij ccs_j is linear in alpha: the centroid and the cross-covariance follow
    // from the moments sum alpha, sum alpha X^T gathered in the reduction pass.
    double pc0[3];
    for (int k = 0; k < 3; ++k)
        pc0[k] = (S.mom_a[0] * ccs[0][k] + S.mom_a[1] * ccs[1][k] + S.mom_a[2] * ccs[2][k] + S.mom_a[3] * ccs[3][k]) / n;
    double abt[9];
    for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) {
            double s = 0;
            for (int q = 0; q < 4; ++q) s += ccs[q][j] * S.mom_aX[3 * q + k];
            abt[3 * j + k] = s - n * pc0[j] * Xbar[k];
        }
    double W[3], U[9], Vt[9], At[9];
    jacobi_svd<3>(abt, W, U, Vt, At);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = U[3 * i] * Vt[j] + U[3 * i + 1] * Vt[3 + j] + U[3 * i + 2] * Vt[6 + j];
    if (det3(R) < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
    for (int k = 0; k < 3; ++k) t[k] = pc0[k] - (R[3 * k] * Xbar[0] + R[3 * k + 1] * Xbar[1] + R[3 * k + 2] * Xbar[2]);
}

__device__ inline void epnp_gauss_newton(const double* L, const double* rho, double* b)
{
#pragma unroll 1
    for (int it = 0; it < 5; ++it) {
        double A[24], B[6], x[4];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double* r = L + 10 * i;
            A[4 * i + 0] = 2 * r[0] * b[0] + r[1] * b[1] + r[3] * b[2] + r[6] * b[3];
            A[4 * i + 1] = r[1] * b[0] + 2 * r[2] * b[1] + r[4] * b[2] + r[7] * b[3];
            A[4 * i + 2] = r[3] * b[0] + r[4] * b[1] + 2 * r[5] * b[2] + r[8] * b[3];
            A[4 * i + 3] = r[6] * b[0] + r[7] * b[1] + r[8] * b[2] + 2 * r[9] * b[3];
            B[i] = rho[i] - (r[0] * b[0] * b[0] + r[1] * b[0] * b[1] + r[2] * b[1] * b[1] + r[3] * b[0] * b[2] +
                             r[4] * b[1] * b[2] + r[5] * b[2] * b[2] + r[6] * b[0] * b[3] + r[7] * b[1] * b[3] +
                             r[8] * b[2] * b[3] + r[9] * b[3] * b[3]);
        }
        qr_lstsq6<4>(A, B, x);
#pragma unroll
        for (int k = 0; k < 4; ++k) b[k] += x[k];
    }
}

__global__ void __launch_bounds__(EPNP_T)
pnp_epnp_kernel(PnpArgs a)
{
    extern __shared__ __align__(16) uint8_t epnp_smem[];
    EpnpShared& S = *reinterpret_cast<EpnpShared*>(epnp_smem);
    const int b = blockIdx.x;
    if (!a.ok[b]) return;
    const int N = a.n[b];
    const int M = a.n_inliers[b];
    if (N == 4 || M < 4) return;   // N == 4: cv2 returns the direct P3P solve
    const float* obj = a.obj + (size_t)b * a.cap * 3;
    const float* img = a.img + (size_t)b * a.cap * 2;
    const int* inl = a.inliers + (size_t)b * a.cap;
    const double n = (double)M;
    const double ifx = 1. / a.fx, ify = 1. / a.fy;

    // pass 1: centroid
    double acc[40];
    acc[0] = acc[1] = acc[2] = 0;
    for (int k = threadIdx.x; k < M; k += EPNP_T) {
        const float* o = obj + 3 * inl[k];
        acc[0] += (double)o[0]; acc[1] += (double)o[1]; acc[2] += (double)o[2];
    }
    block_reduce_sum<3>(acc, S.red, S.out);
    const double c0[3] = {S.out[0] / n, S.out[1] / n, S.out[2] / n};
    __syncthreads();
    // pass 2: PW0^T PW0
    for (int k = 0; k < 6; ++k) acc[k] = 0;
    for (int k = threadIdx.x; k < M; k += EPNP_T) {
        const float* o = obj + 3 * inl[k];
        const double x = (double)o[0] - c0[0], y = (double)o[1] - c0[1], z = (double)o[2] - c0[2];
        acc[0] += x * x; acc[1] += x * y; acc[2] += x * z; acc[3] += y * y; acc[4] += y * z; acc[5] += z * z;
    }
    block_reduce_sum<6>(acc, S.red, S.out);
    if (threadIdx.x == 0) {
        const double Sm[9] = {S.out[0], S.out[1], S.out[2], S.out[1], S.out[3], S.out[4], S.out[2], S.out[4], S.out[5]};
        double W[3], U[9], Vt[9], At[9];
        jacobi_svd<3>(Sm, W, U, Vt, At);
        for (int k = 0; k < 3; ++k) S.cws[0][k] = c0[k];
        for (int i = 1; i < 4; ++i) {
            const double kk = sqrt((W[i - 1] > 0 ? W[i - 1] : 0) / n);
            for (int j = 0; j < 3; ++j) S.cws[i][j] = c0[j] + kk * U[3 * j + (i - 1)];
        }
        double cc[9];
        for (int i = 0; i < 3; ++i) for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = S.cws[j][i] - S.cws[0][i];
        if (!inv3(cc, S.ci)) for (int k = 0; k < 9; ++k) S.ci[k] = nan("");
    }
    __syncthreads();
    // pass 3: barycentric coordinates; M^T M in its 4x4-blocks-of-3x3 structure (40 sums) and moments
    double ci[9];
    for (int k = 0; k < 9; ++k) ci[k] = S.ci[k];
    for (int k = 0; k < 40; ++k) acc[k] = 0;
    double mo[16];
    for (int k = 0; k < 16; ++k) mo[k] = 0;
    for (int k = threadIdx.x; k < M; k += EPNP_T) {
        const int id = inl[k];
        const float* o = obj + 3 * id;
        const double X = o[0], Y = o[1], Z = o[2];
        const double dx = X - c0[0], dy = Y - c0[1], dz = Z - c0[2];
        double al[4];
        for (int j = 0; j < 3; ++j) al[1 + j] = ci[3 * j] * dx + ci[3 * j + 1] * dy + ci[3 * j + 2] * dz;
        al[0] = 1.0 - al[1] - al[2] - al[3];
        // undistortPoints (float64) then back to pixels, as cv2's EPnP front end does
        const double u = (((double)img[2 * id] - a.cx) * ifx) * a.fx + a.cx;
        const double v = (((double)img[2 * id + 1] - a.cy) * ify) * a.fy + a.cy;
        const double du = a.cx - u, dv = a.cy - v, dd = du * du + dv * dv;
        int q = 0;
        for (int j = 0; j < 4; ++j)
            for (int l = j; l < 4; ++l, ++q) {
                const double aa = al[j] * al[l];
                acc[4 * q] += aa; acc[4 * q + 1] += aa * du; acc[4 * q + 2] += aa * dv; acc[4 * q + 3] += aa * dd;
            }
        for (int j = 0; j < 4; ++j) {
            mo[3 * j] += al[j] * X; mo[3 * j + 1] += al[j] * Y; mo[3 * j + 2] += al[j] * Z;
            mo[12 + j] += al[j];
        }
        if (k == 0) for (int j = 0; j < 4; ++j) S.alpha0[j] = al[j];
    }
    block_reduce_sum<40>(acc, S.red, S.out);
    double Xbar[3];
    if (threadIdx.x == 0) {
        // assemble the 12x12 M^T M
        double* MtM = S.work;
        int q = 0;
        for (int j = 0; j < 4; ++j)
            for (int l = j; l < 4; ++l, ++q) {
                const double A = S.out[4 * q], B = S.out[4 * q + 1], Cc = S.out[4 * q + 2], D = S.out[4 * q + 3];
                const double blk[9] = {a.fx * a.fx * A, 0, a.fx * B, 0, a.fy * a.fy * A, a.fy * Cc, a.fx * B, a.fy * Cc, D};
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) {
                        MtM[(3 * j + r) * 12 + 3 * l + c] = blk[3 * r + c];
                        MtM[(3 * l + c) * 12 + 3 * j + r] = blk[3 * r + c];
                    }
            }
    }
    __syncthreads();
    block_reduce_sum<16>(mo, S.red, S.out);
    if (threadIdx.x < 12) S.mom_aX[threadIdx.x] = S.out[threadIdx.x];
    if (threadIdx.x < 4) S.mom_a[threadIdx.x] = S.out[12 + threadIdx.x];
    __syncthreads();
    Xbar[0] = c0[0]; Xbar[1] = c0[1]; Xbar[2] = c0[2];
    if (threadIdx.x < 32) {
        double* MtM = S.work;          // destroyed: its diagonal becomes the eigenvalues
        double* Vm = S.work + 144;     // eigenvectors in columns
        warp_jacobi_eig12(MtM, Vm, S.work + 288, threadIdx.x);
        __syncwarp();
        if (threadIdx.x == 0) {
            // rows of ut = eigenvectors by DESCENDING eigenvalue (what cvSVD(MtM, D, Ut) returns)
            int order[12];
            for (int i = 0; i < 12; ++i) order[i] = i;
            for (int i = 0; i < 12; ++i)
                for (int j = i + 1; j < 12; ++j)
                    if (MtM[order[j] * 13] > MtM[order[i] * 13]) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
            for (int i = 0; i < 12; ++i)
                for (int k = 0; k < 12; ++k) S.ut[12 * i + k] = Vm[12 * k + order[i]];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double* v[4] = {S.ut + 12 * 11, S.ut + 12 * 10, S.ut + 12 * 9, S.ut + 12 * 8};
        const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
        double dv[4][6][3];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 6; ++j)
                for (int k = 0; k < 3; ++k) dv[i][j][k] = v[i][3 * pa[j] + k] - v[i][3 * pb[j] + k];
#define VDOT(p, q) ((p)[0] * (q)[0] + (p)[1] * (q)[1] + (p)[2] * (q)[2])
        for (int i = 0; i < 6; ++i) {
            double* r = S.L + 10 * i;
            r[0] = VDOT(dv[0][i], dv[0][i]); r[1] = 2 * VDOT(dv[0][i], dv[1][i]); r[2] = VDOT(dv[1][i], dv[1][i]);
            r[3] = 2 * VDOT(dv[0][i], dv[2][i]); r[4] = 2 * VDOT(dv[1][i], dv[2][i]); r[5] = VDOT(dv[2][i], dv[2][i]);
            r[6] = 2 * VDOT(dv[0][i], dv[3][i]); r[7] = 2 * VDOT(dv[1][i], dv[3][i]); r[8] = 2 * VDOT(dv[2][i], dv[3][i]);
            r[9] = VDOT(dv[3][i], dv[3][i]);
            const double d[3] = {S.cws[pa[i]][0] - S.cws[pb[i]][0], S.cws[pa[i]][1] - S.cws[pb[i]][1], S.cws[pa[i]][2] - S.cws[pb[i]][2]};
            S.rho[i] = VDOT(d, d);
        }
    }
    __syncthreads();
    // three beta initialisations (one lane each), 5 Gauss-Newton steps, pose from the moments
    if (threadIdx.x < 3) {
        const int Nn = threadIdx.x + 1;
        const double* L = S.L;
        double be[4] = {0, 0, 0, 0};
        if (Nn == 1) {
            double A[24], b4[4];
            for (int i = 0; i < 6; ++i) { A[4 * i] = L[10 * i]; A[4 * i + 1] = L[10 * i + 1]; A[4 * i + 2] = L[10 * i + 3]; A[4 * i + 3] = L[10 * i + 6]; }
            qr_lstsq6<4>(A, S.rho, b4);
            if (b4[0] < 0) { be[0] = sqrt(-b4[0]); be[1] = -b4[1] / be[0]; be[2] = -b4[2] / be[0]; be[3] = -b4[3] / be[0]; }
            else { be[0] = sqrt(b4[0]); be[1] = b4[1] / be[0]; be[2] = b4[2] / be[0]; be[3] = b4[3] / be[0]; }
        } else if (Nn == 2) {
            double A[18], b3[3];
            for (int i = 0; i < 6; ++i) { A[3 * i] = L[10 * i]; A[3 * i + 1] = L[10 * i + 1]; A[3 * i + 2] = L[10 * i + 2]; }
            qr_lstsq6<3>(A, S.rho, b3);
            if (b3[0] < 0) { be[0] = sqrt(-b3[0]); be[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0; }
            else { be[0] = sqrt(b3[0]); be[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0; }
            if (b3[1] < 0) be[0] = -be[0];
        } else {
            double A[30], b5[5];
            for (int i = 0; i < 6; ++i) for (int k = 0; k < 5; ++k) A[5 * i + k] = L[10 * i + k];
            qr_lstsq6<5>(A, S.rho, b5);
            if (b5[0] < 0) { be[0] = sqrt(-b5[0]); be[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0; }
            else { be[0] = sqrt(b5[0]); be[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0; }
            if (b5[1] < 0) be[0] = -be[0];
            be[2] = b5[3] / be[0];
        }
        epnp_gauss_newton(S.L, S.rho, be);
        epnp_pose_from_betas(S, be, n, Xbar, S.Rs[threadIdx.x], S.ts[threadIdx.x]);
    }
    __syncthreads();
    // pass 4: mean reprojection error of the three candidates
    acc[0] = acc[1] = acc[2] = 0;
    for (int k = threadIdx.x; k < M; k += EPNP_T) {
        const int id = inl[k];
        const float* o = obj + 3 * id;
        const double X = o[0], Y = o[1], Z = o[2];
        const double u = (((double)img[2 * id] - a.cx) * ifx) * a.fx + a.cx;
        const double v = (((double)img[2 * id + 1] - a.cy) * ify) * a.fy + a.cy;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double* R = S.Rs[c];
            const double* t = S.ts[c];
            const double Xc = R[0] * X + R[1] * Y + R[2] * Z + t[0];
            const double Yc = R[3] * X + R[4] * Y + R[5] * Z + t[1];
            const double iZ = 1.0 / (R[6] * X + R[7] * Y + R[8] * Z + t[2]);
            const double ue = a.cx + a.fx * Xc * iZ, ve = a.cy + a.fy * Yc * iZ;
            const double du = u - ue, dv = v - ve;
            acc[c] += sqrt(du * du + dv * dv);
        }
    }
    block_reduce_sum<3>(acc, S.red, S.out);
    if (threadIdx.x == 0) {
        const double r1 = S.out[0] / n, r2 = S.out[1] / n, r3 = S.out[2] / n;
        int Nn = 0;
        double rb = r1;
        if (r2 < r1) { Nn = 1; rb = r2; }
        if (r3 < rb) Nn = 2;
        bool fin = true;
        for (int k = 0; k < 9; ++k) fin = fin && isfinite(S.Rs[Nn][k]);
        for (int k = 0; k < 3; ++k) fin = fin && isfinite(S.ts[Nn][k]);
        if (fin) {
            double* pose = a.pose + (size_t)b * 6;
            double rv[3];
            R_to_rodrigues(S.Rs[Nn], rv);
            pose[0] = rv[0]; pose[1] = rv[1]; pose[2] = rv[2];
            pose[3] = S.ts[Nn][0]; pose[4] = S.ts[Nn][1]; pose[5] = S.ts[Nn][2];
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
void vo_rng_raw_stream(uint32_t* out, int n)
{
    // cv::RNG((uint64)-1): state = (uint32)state * 4164903690 + (state >> 32)
    uint64_t st = 0xFFFFFFFFFFFFFFFFull;
    for (int i = 0; i < n; ++i) {
        st = (uint64_t)(uint32_t)st * 4164903690u + (st >> 32);
        out[i] = (uint32_t)st;
    }
}

int vo_rng_table(b200vo_ctx* ctx, int n, const uint32_t** d_table)
{
    if (n > ctx->n_rng) {
        const int cap = n + n / 2 + 1024;
        uint32_t* h = (uint32_t*)malloc((size_t)cap * sizeof(uint32_t));
        if (!h) return vo_set_err(ctx, B200VO_E_NOMEM, "host alloc");
        vo_rng_raw_stream(h, cap);
        int rc = vo_reserve(ctx, ctx->d_rng, (size_t)cap * sizeof(uint32_t));
        if (rc == 0) {
            cudaError_t e = cudaMemcpyAsync(ctx->d_rng.p, h, (size_t)cap * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) rc = vo_cuda_fail(ctx, e, "rng table upload");
        }
        free(h);
        if (rc) return rc;
        ctx->n_rng = cap;
    }
    *d_table = (const uint32_t*)ctx->d_rng.p;
    return 0;
}

size_t vo_pnp_workspace_bytes(int batch, int cap, int iters)
{
    size_t b = 0;
    b += vo_align((size_t)batch * iters * 4 * sizeof(int), 256);       // samples
    b += vo_align((size_t)batch * iters * 12 * sizeof(double), 256);   // hyp
    b += vo_align((size_t)batch * iters * 3 * sizeof(double), 256);    // hyp_rvec
    b += 2 * vo_align((size_t)batch * iters * sizeof(int), 256);       // hyp_ok, counts
    b += 4 * vo_align((size_t)batch * sizeof(int), 256);               // winner, iters_run, n_inliers, ok_ws
    b += vo_align((size_t)batch * 3 * sizeof(int), 256);               // flags | ticket | need
    return b;
}

void vo_pnp_carve_workspace(PnpArgs& a, void* ws)
{
    uint8_t* p = (uint8_t*)ws;
    auto take = [&](size_t bytes) { void* r = p; p += vo_align(bytes, 256); return r; };
    a.samples = (int*)take((size_t)a.batch * a.iters * 4 * sizeof(int));
    a.hyp = (double*)take((size_t)a.batch * a.iters * 12 * sizeof(double));
    a.hyp_rvec = (double*)take((size_t)a.batch * a.iters * 3 * sizeof(double));
    a.hyp_ok = (int*)take((size_t)a.batch * a.iters * sizeof(int));
    a.counts = (int*)take((size_t)a.batch * a.iters * sizeof(int));
    a.winner = (int*)take((size_t)a.batch * sizeof(int));
    a.iters_run = (int*)take((size_t)a.batch * sizeof(int));
    a.n_inliers = (int*)take((size_t)a.batch * sizeof(int));
    a.flags = (int*)take((size_t)a.batch * 3 * sizeof(int));
    a.ticket = a.flags + a.batch;
    a.need = a.flags + 2 * a.batch;
    a.ok_ws = (uint8_t*)take((size_t)a.batch * sizeof(int));
}

int vo_pnp_launch(b200vo_ctx* ctx, const PnpArgs& a, bool gen_samples)
{
    if (a.batch <= 0) return 0;
    VO_CUDA(ctx, cudaMemsetAsync(a.flags, 0, (size_t)a.batch * 3 * sizeof(int), ctx->stream));
    if (gen_samples) {
        const size_t smem = (size_t)a.n_raw * sizeof(int);
        if (smem > 48 * 1024)
            VO_CUDA(ctx, cudaFuncSetAttribute(ransac_samples_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ransac_samples_kernel<4><<<a.batch, 128, smem, ctx->stream>>>(a.rng_raw, a.n_raw, a.n, a.iters, a.samples, a.flags);
        ctx->launches++;
    }
    // head chunk: the first PNP_HEAD hypotheses; its last scoring CTA per sequence leaves need[b].
    // tail chunk: whatever the replay can still reach (CTAs beyond need[b] exit at once).
    const int head_n = (a.full_counts || a.iters < PNP_HEAD) ? a.iters : PNP_HEAD;
    for (int part = 0; part < 2; ++part) {
        PnpArgs c = a;
        c.head = part == 0;
        c.h_begin = part == 0 ? 0 : head_n;
        c.h_end = part == 0 ? head_n : a.iters;
        const int nh = c.h_end - c.h_begin;
        if (nh <= 0) continue;
        dim3 g1((nh + 63) / 64, a.batch);
        pnp_solve_kernel<<<g1, 64, 0, ctx->stream>>>(c);
        dim3 g2((a.cap + 255) / 256, (nh + SCORE_HT - 1) / SCORE_HT, a.batch);
        pnp_score_kernel<<<g2, 256, 0, ctx->stream>>>(c);
        ctx->launches += 2;
    }
    pnp_select_kernel<<<a.batch, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    {
        const size_t smem = sizeof(EpnpShared);
        pnp_epnp_kernel<<<a.batch, EPNP_T, smem, ctx->stream>>>(a);
        ctx->launches++;
    }
    VO_CUDA(ctx, cudaGetLastError());
    return 0;
}
