// The two components next to the hot path, downstream of the tracker and of
// goodFeaturesToTrack in the reference's per-frame loop:
//   f2  candidate min-distance filter, reference VisualOdometryPipeLine.py:258
//         valid[i] = np.all(np.linalg.norm(pts[i,:] - self.potential_keys, axis=1) > min_dist)
//       (float32: squares, sum and square root rounded as numpy rounds them);
//   f1  triangulate_landmarks, reference :107-206: age gate (:171-174), bearing-angle gate
//       (:117-147), cv2.triangulatePoints (:188-193: 4x4 DLT, OpenCV's small-matrix Jacobi SVD,
//       float32 homogeneous output), de-homogenisation (:194), depth window in both cameras
//       (:149-168); accepted landmarks and their keypoints appended in candidate order.
// The reference spends ~67 ms (f1) and ~48 ms (f2) per frame on these in Python loops
// (SURVEY.md 8f); here each is one or two small launches.
#include "internal.cuh"
#include "mathdev.cuh"

using namespace vo;

// ---- f2: one warp per new corner, lanes stride over the existing candidates ----
__global__ void __launch_bounds__(256)
min_distance_kernel(const float2* __restrict__ pts, int n, const float2* __restrict__ existing, int m, float min_dist,
                    uint8_t* __restrict__ valid)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const float2 p = pts[i];
    bool ok = true;
    for (int j0 = 0; j0 < m; j0 += 32) {
        const int j = j0 + lane;
        if (j < m) {
            const float2 q = existing[j];
            const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y);
            const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            ok = d > min_dist;           // NaN compares false, as in numpy
        }
        if (!__all_sync(0xffffffffu, ok)) { ok = false; break; }
    }
    if (lane == 0) valid[i] = ok ? 1 : 0;
}

struct TriArgs {
    double K[9], Ki[9];
    double cur[12];            // R_CW | t_CW of the current frame (as the reference stores them)
    double min_dist, max_dist, min_angle_deg;
    int min_frames, n, n_poses;
    const float* first_keys; const float* keys; const int* first_pose;
    const double* poses;       // [n_poses][12]
    uint8_t* keep;             // [n] too_short_baseline
    float* lm_slot;            // [n][3] per-candidate result
    float* out_lm; float* out_kp; int* n_new;   // compacted
    int* flags;                // bit0: first_pose out of range
};

__device__ __forceinline__ void invert_pose(const double* cw, double* R, double* t)
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[3 * i + j] = cw[3 * j + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = -(R[3 * i] * cw[9] + R[3 * i + 1] * cw[10] + R[3 * i + 2] * cw[11]);
}
__device__ __forceinline__ void proj_matrix(const double* K, const double* R, const double* t, double* P)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) P[4 * i + j] = K[3 * i] * R[j] + K[3 * i + 1] * R[3 + j] + K[3 * i + 2] * R[6 + j];
        P[4 * i + 3] = K[3 * i] * t[0] + K[3 * i + 1] * t[1] + K[3 * i + 2] * t[2];
    }
}

// ---- f1: one thread per candidate ----
__device__ __forceinline__ void triangulate_one(const TriArgs& a, int i)
{
    a.keep[i] = 1;
    const int fp = a.first_pose[i];
    if (a.n_poses > 1 && a.n_poses - fp <= a.min_frames) return;
    if (fp < 0 || fp >= a.n_poses) { atomicOr(a.flags, 1); return; }
    double past[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) past[k] = a.poses[12 * fp + k];
    const double u = a.keys[2 * i], v = a.keys[2 * i + 1], u0 = a.first_keys[2 * i], v0 = a.first_keys[2 * i + 1];
    {   // bearing angle between the two viewing rays (check_baseline)
        const double ax = a.Ki[0] * u + a.Ki[2], ay = a.Ki[4] * v + a.Ki[5], az = 1.0;
        double rel[9], M[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) rel[3 * r + c] = past[r] * a.cur[c] + past[3 + r] * a.cur[3 + c] + past[6 + r] * a.cur[6 + c];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[3 * r + c] = rel[3 * r] * a.Ki[c] + rel[3 * r + 1] * a.Ki[3 + c] + rel[3 * r + 2] * a.Ki[6 + c];
        const double bx = M[0] * u0 + M[1] * v0 + M[2], by = M[3] * u0 + M[4] * v0 + M[5], bz = M[6] * u0 + M[7] * v0 + M[8];
        double cs = (ax * bx + ay * by + az * bz) / (sqrt(ax * ax + ay * ay + az * az) * sqrt(bx * bx + by * by + bz * bz));
        cs = cs < -1.0 ? -1.0 : (cs > 1.0 ? 1.0 : cs);
        const double alpha = acos(cs) * (180.0 / 3.14159265358979323846);
        if (alpha < a.min_angle_deg) return;
    }
    double Rc[9], tc[3], Pc[12], Rp[9], tp[3], Pp[12];
    invert_pose(a.cur, Rc, tc);
    proj_matrix(a.K, Rc, tc, Pc);
    invert_pose(past, Rp, tp);
    proj_matrix(a.K, Rp, tp, Pp);
    double A[16], W[4], U[16], Vt[16], At[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        A[k] = u0 * Pp[8 + k] - Pp[k];
        A[4 + k] = v0 * Pp[8 + k] - Pp[4 + k];
        A[8 + k] = u * Pc[8 + k] - Pc[k];
        A[12 + k] = v * Pc[8 + k] - Pc[4 + k];
    }
    jacobi_svd<4>(A, W, U, Vt, At);
    const float X0 = (float)Vt[12], X1 = (float)Vt[13], X2 = (float)Vt[14], X3 = (float)Vt[15];
    const float L0 = __fdiv_rn(X0, X3), L1 = __fdiv_rn(X1, X3), L2 = __fdiv_rn(X2, X3);
    const double zc = Rc[6] * (double)L0 + Rc[7] * (double)L1 + Rc[8] * (double)L2 + tc[2];
    const double zp = Rp[6] * (double)L0 + Rp[7] * (double)L1 + Rp[8] * (double)L2 + tp[2];
    if (zc > a.min_dist && zp > a.min_dist && zc < a.max_dist && zp < a.max_dist) {
        a.keep[i] = 0;
        a.lm_slot[3 * i] = L0; a.lm_slot[3 * i + 1] = L1; a.lm_slot[3 * i + 2] = L2;
    }
}

__global__ void __launch_bounds__(64)
triangulate_kernel(TriArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) triangulate_one(a, i);
}

// ordered compaction of the accepted candidates (keep == 0): one CTA, ballot + running base
__device__ __forceinline__ void triangulate_compact_block(const TriArgs& a)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < a.n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const bool acc = i < a.n && a.keep[i] == 0;
        const unsigned bm = __ballot_sync(0xffffffffu, acc);
        if (lane == 0) s_warp[warp] = __popc(bm);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (acc) {
            const int o = off + __popc(bm & ((1u << lane) - 1));
            a.out_lm[3 * o] = a.lm_slot[3 * i]; a.out_lm[3 * o + 1] = a.lm_slot[3 * i + 1]; a.out_lm[3 * o + 2] = a.lm_slot[3 * i + 2];
            a.out_kp[2 * o] = a.keys[2 * i]; a.out_kp[2 * o + 1] = a.keys[2 * i + 1];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 32; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.n_new = s_base;
}

__global__ void __launch_bounds__(1024)
triangulate_compact_kernel(TriArgs a) { triangulate_compact_block(a); }

// ---- batched forms (SURVEY 8f "batched f1/f2 on the resident batch state"): the same per-sequence work for every
// sequence of a batch in one launch each; blockIdx.y = sequence, ragged counts per sequence ----
struct TriBatchArgs {
    TriArgs common;            // K, Ki, gates (the per-sequence fields are filled in by the kernels)
    int cap, pose_cap;
    const int* n;              // [batch] candidates per sequence
    const int* n_poses;        // [batch]
    const double* cur;         // [batch][12]
    const float* first_keys; const float* keys; const int* first_pose;   // [batch][cap]...
    const double* poses;       // [batch][pose_cap][12]
    uint8_t* keep; float* lm_slot; float* out_lm; float* out_kp; int* n_new; int* flags;   // [batch]...
};

__device__ __forceinline__ TriArgs tri_sequence(const TriBatchArgs& b, int s)
{
    TriArgs a = b.common;
    const size_t o = (size_t)s * b.cap;
    a.n = min(max(b.n[s], 0), b.cap);
    a.n_poses = min(max(b.n_poses[s], 0), b.pose_cap);
    for (int k = 0; k < 12; ++k) a.cur[k] = b.cur[12 * s + k];
    a.first_keys = b.first_keys + 2 * o; a.keys = b.keys + 2 * o; a.first_pose = b.first_pose + o;
    a.poses = b.poses + (size_t)s * b.pose_cap * 12;
    a.keep = b.keep + o; a.lm_slot = b.lm_slot + 3 * o; a.out_lm = b.out_lm + 3 * o; a.out_kp = b.out_kp + 2 * o;
    a.n_new = b.n_new + s; a.flags = b.flags + s;
    return a;
}

__global__ void __launch_bounds__(64)
triangulate_batch_kernel(TriBatchArgs b)
{
    const TriArgs a = tri_sequence(b, blockIdx.y);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.n_poses < 1) { if (i < a.n) a.keep[i] = 1; return; }
    if (i < a.n) triangulate_one(a, i);
}

__global__ void __launch_bounds__(1024)
triangulate_compact_batch_kernel(TriBatchArgs b)
{
    const TriArgs a = tri_sequence(b, blockIdx.x);
    triangulate_compact_block(a);
}

__global__ void __launch_bounds__(256)
min_distance_batch_kernel(const float2* __restrict__ pts, const int* __restrict__ n, int n_cap, const float2* __restrict__ existing,
                          const int* __restrict__ m, int m_cap, float min_dist, uint8_t* __restrict__ valid)
{
    const int s = blockIdx.y;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= min(max(n[s], 0), n_cap)) return;
    const float2 p = pts[(size_t)s * n_cap + i];
    const float2* ex = existing + (size_t)s * m_cap;
    const int ms = min(max(m[s], 0), m_cap);
    bool ok = true;
    for (int j0 = 0; j0 < ms; j0 += 32) {
        const int j = j0 + lane;
        if (j < ms) {
            const float2 q = ex[j];
            const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y);
            const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            ok = d > min_dist;           // NaN compares false, as in numpy
        }
        if (!__all_sync(0xffffffffu, ok)) { ok = false; break; }
    }
    if (lane == 0) valid[(size_t)s * n_cap + i] = ok ? 1 : 0;
}

extern "C" int b200vo_min_distance_mask(b200vo_ctx* ctx, const float* pts, int n, const float* existing, int m,
                                        float min_dist, uint8_t* valid)
{
    if (!ctx || n < 0 || m < 0 || (n > 0 && (!pts || !valid)) || (m > 0 && !existing)) return B200VO_E_BADARG;
    if (n == 0) return 0;
    VO_CUDA(ctx, cudaSetDevice(ctx->device));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    const size_t b_p = vo_align((size_t)n * 8, 256), b_e = vo_align((size_t)(m > 0 ? m : 1) * 8, 256), b_v = vo_align((size_t)n, 256);
    VO_TRY(vo_reserve(ctx, ctx->d_scratch[6], b_p + b_e + b_v));
    VO_TRY(vo_reserve_pinned(ctx, b_p + b_e + b_v));
    uint8_t* d = (uint8_t*)ctx->d_scratch[6].p;
    uint8_t* hp = (uint8_t*)ctx->h_pin;
    memcpy(hp, pts, (size_t)n * 8);
    if (m > 0) memcpy(hp + b_p, existing, (size_t)m * 8);
    VO_CUDA(ctx, cudaMemcpyAsync(d, hp, b_p + b_e, cudaMemcpyHostToDevice, ctx->stream));
    min_distance_kernel<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>((const float2*)d, n, (const float2*)(d + b_p), m, min_dist,
                                                                      d + b_p + b_e);
    ctx->launches++;
    VO_CUDA(ctx, cudaGetLastError());
    VO_CUDA(ctx, cudaMemcpyAsync(hp + b_p + b_e, d + b_p + b_e, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    memcpy(valid, hp + b_p + b_e, (size_t)n);
    return 0;
}

extern "C" int b200vo_triangulate_landmarks(b200vo_ctx* ctx, const double K[9], double min_dist, double max_dist,
                                            double min_baseline_angle_deg, int min_baseline_frames,
                                            const float* first_keys, const float* keys, const int32_t* first_pose, int n,
                                            const double* poses_cw, int n_poses, const double cur_pose_cw[12],
                                            uint8_t* too_short_baseline, float* new_landmarks, float* new_keypoints,
                                            int* n_new)
{
    if (!ctx || !K || !cur_pose_cw || !n_new || n < 0 || n_poses < 1 || !poses_cw) return B200VO_E_BADARG;
    *n_new = 0;
    if (n == 0) return 0;
    if (!first_keys || !keys || !first_pose || !too_short_baseline || !new_landmarks || !new_keypoints)
        return vo_set_err(ctx, B200VO_E_BADARG, "null pointer");
    VO_CUDA(ctx, cudaSetDevice(ctx->device));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    TriArgs a{};
    for (int k = 0; k < 9; ++k) a.K[k] = K[k];
    const double Ki[9] = {1.0 / K[0], 0, -K[2] / K[0], 0, 1.0 / K[4], -K[5] / K[4], 0, 0, 1};
    for (int k = 0; k < 9; ++k) a.Ki[k] = Ki[k];
    for (int k = 0; k < 12; ++k) a.cur[k] = cur_pose_cw[k];
    a.min_dist = min_dist; a.max_dist = max_dist; a.min_angle_deg = min_baseline_angle_deg;
    a.min_frames = min_baseline_frames; a.n = n; a.n_poses = n_poses;
    // device layout: inputs [first_keys | keys | first_pose | poses] then outputs [keep | out_lm | out_kp | n_new,flags] + lm_slot
    const size_t b_k = vo_align((size_t)n * 8, 256), b_fp = vo_align((size_t)n * 4, 256), b_ps = vo_align((size_t)n_poses * 96, 256);
    const size_t b_keep = vo_align((size_t)n, 256), b_lm = vo_align((size_t)n * 12, 256), b_small = 256;
    const size_t in_bytes = 2 * b_k + b_fp + b_ps, out_bytes = b_keep + b_lm + b_k + b_small;
    VO_TRY(vo_reserve(ctx, ctx->d_scratch[7], in_bytes + out_bytes + b_lm));
    VO_TRY(vo_reserve_pinned(ctx, in_bytes + out_bytes));
    uint8_t* d = (uint8_t*)ctx->d_scratch[7].p;
    uint8_t* hp = (uint8_t*)ctx->h_pin;
    memcpy(hp, first_keys, (size_t)n * 8);
    memcpy(hp + b_k, keys, (size_t)n * 8);
    memcpy(hp + 2 * b_k, first_pose, (size_t)n * 4);
    memcpy(hp + 2 * b_k + b_fp, poses_cw, (size_t)n_poses * 96);
    VO_CUDA(ctx, cudaMemcpyAsync(d, hp, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    a.first_keys = (const float*)d; a.keys = (const float*)(d + b_k); a.first_pose = (const int*)(d + 2 * b_k);
    a.poses = (const double*)(d + 2 * b_k + b_fp);
    uint8_t* dout = d + in_bytes;
    a.keep = dout; a.out_lm = (float*)(dout + b_keep); a.out_kp = (float*)(dout + b_keep + b_lm);
    a.n_new = (int*)(dout + b_keep + b_lm + b_k); a.flags = a.n_new + 1;
    a.lm_slot = (float*)(dout + out_bytes);
    VO_CUDA(ctx, cudaMemsetAsync(a.n_new, 0, b_small, ctx->stream));
    triangulate_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(a);
    triangulate_compact_kernel<<<1, 1024, 0, ctx->stream>>>(a);
    ctx->launches += 2;
    VO_CUDA(ctx, cudaGetLastError());
    uint8_t* ho = hp + in_bytes;
    VO_CUDA(ctx, cudaMemcpyAsync(ho, dout, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    const int* small = (const int*)(ho + b_keep + b_lm + b_k);
    if (small[1] & 1) return vo_set_err(ctx, B200VO_E_BADARG, "first_pose index outside [0, n_poses)");
    const int cnt = small[0];
    memcpy(too_short_baseline, ho, (size_t)n);
    memcpy(new_landmarks, ho + b_keep, (size_t)cnt * 12);
    memcpy(new_keypoints, ho + b_keep + b_lm, (size_t)cnt * 8);
    *n_new = cnt;
    return 0;
}

// Batched f2: the candidate min-distance filter (:258) for every sequence of a batch in one launch.
// pts float32 (batch, n_cap, 2) with n[batch] live rows, existing float32 (batch, m_cap, 2) with m[batch]; valid uint8 (batch, n_cap).
extern "C" int b200vo_batch_min_distance_mask(b200vo_ctx* ctx, int batch, const float* pts, const int32_t* n, int n_cap,
                                              const float* existing, const int32_t* m, int m_cap, float min_dist, uint8_t* valid)
{
    if (!ctx || batch < 1 || n_cap < 1 || m_cap < 0 || !pts || !n || !m || !valid || (m_cap > 0 && !existing)) return B200VO_E_BADARG;
    VO_CUDA(ctx, cudaSetDevice(ctx->device));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    const int mc = m_cap > 0 ? m_cap : 1;
    const size_t b_p = vo_align((size_t)batch * n_cap * 8, 256), b_e = vo_align((size_t)batch * mc * 8, 256);
    const size_t b_c = vo_align((size_t)batch * 4, 256), b_v = vo_align((size_t)batch * n_cap, 256);
    VO_TRY(vo_reserve(ctx, ctx->d_scratch[6], b_p + b_e + 2 * b_c + b_v));
    VO_TRY(vo_reserve_pinned(ctx, b_p + b_e + 2 * b_c + b_v));
    uint8_t* d = (uint8_t*)ctx->d_scratch[6].p;
    uint8_t* hp = (uint8_t*)ctx->h_pin;
    memcpy(hp, pts, (size_t)batch * n_cap * 8);
    if (m_cap > 0) memcpy(hp + b_p, existing, (size_t)batch * m_cap * 8);
    memcpy(hp + b_p + b_e, n, (size_t)batch * 4);
    memcpy(hp + b_p + b_e + b_c, m, (size_t)batch * 4);
    VO_CUDA(ctx, cudaMemcpyAsync(d, hp, b_p + b_e + 2 * b_c, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* dv = d + b_p + b_e + 2 * b_c;
    VO_CUDA(ctx, cudaMemsetAsync(dv, 0, b_v, ctx->stream));
    min_distance_batch_kernel<<<dim3((n_cap * 32 + 255) / 256, batch), 256, 0, ctx->stream>>>(
        (const float2*)d, (const int*)(d + b_p + b_e), n_cap, (const float2*)(d + b_p), (const int*)(d + b_p + b_e + b_c), m_cap, min_dist, dv);
    ctx->launches++;
    VO_CUDA(ctx, cudaGetLastError());
    uint8_t* ho = hp + b_p + b_e + 2 * b_c;
    VO_CUDA(ctx, cudaMemcpyAsync(ho, dv, (size_t)batch * n_cap, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    memcpy(valid, ho, (size_t)batch * n_cap);
    return 0;
}

// Batched f1: the candidate loop of triangulate_landmarks (:170-204) for every sequence of a batch: two launches.
// Per sequence s: n[s] candidates in rows [0, n[s]) of the (batch, cap, ...) arrays, n_poses[s] stored poses in
// poses_cw (batch, pose_cap, 12), the current pose cur_pose_cw (batch, 12).  Outputs: too_short_baseline uint8 (batch, cap),
// new_landmarks float32 (batch, cap, 3) / new_keypoints float32 (batch, cap, 2) compacted per sequence, n_new int32 (batch).
extern "C" int b200vo_batch_triangulate_landmarks(b200vo_ctx* ctx, const double K[9], double min_dist, double max_dist,
                                                  double min_baseline_angle_deg, int min_baseline_frames, int batch, int cap,
                                                  const float* first_keys, const float* keys, const int32_t* first_pose,
                                                  const int32_t* n, const double* poses_cw, const int32_t* n_poses, int pose_cap,
                                                  const double* cur_pose_cw, uint8_t* too_short_baseline, float* new_landmarks,
                                                  float* new_keypoints, int32_t* n_new)
{
    if (!ctx || !K || batch < 1 || cap < 1 || pose_cap < 1 || !first_keys || !keys || !first_pose || !n || !poses_cw || !n_poses ||
        !cur_pose_cw || !too_short_baseline || !new_landmarks || !new_keypoints || !n_new)
        return B200VO_E_BADARG;
    VO_CUDA(ctx, cudaSetDevice(ctx->device));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    TriBatchArgs b{};
    for (int k = 0; k < 9; ++k) b.common.K[k] = K[k];
    const double Ki[9] = {1.0 / K[0], 0, -K[2] / K[0], 0, 1.0 / K[4], -K[5] / K[4], 0, 0, 1};
    for (int k = 0; k < 9; ++k) b.common.Ki[k] = Ki[k];
    b.common.min_dist = min_dist; b.common.max_dist = max_dist; b.common.min_angle_deg = min_baseline_angle_deg;
    b.common.min_frames = min_baseline_frames;
    b.cap = cap; b.pose_cap = pose_cap;
    const size_t nc = (size_t)batch * cap;
    const size_t b_k = vo_align(nc * 8, 256), b_fp = vo_align(nc * 4, 256), b_ps = vo_align((size_t)batch * pose_cap * 96, 256);
    const size_t b_cnt = vo_align((size_t)batch * 4, 256), b_cur = vo_align((size_t)batch * 96, 256);
    const size_t b_keep = vo_align(nc, 256), b_lm = vo_align(nc * 12, 256);
    const size_t in_bytes = 2 * b_k + b_fp + b_ps + 2 * b_cnt + b_cur, out_bytes = b_keep + b_lm + b_k + 2 * b_cnt;
    VO_TRY(vo_reserve(ctx, ctx->d_scratch[7], in_bytes + out_bytes + b_lm));
    VO_TRY(vo_reserve_pinned(ctx, in_bytes + out_bytes));
    uint8_t* d = (uint8_t*)ctx->d_scratch[7].p;
    uint8_t* hp = (uint8_t*)ctx->h_pin;
    size_t o = 0;
    memcpy(hp + o, first_keys, nc * 8); b.first_keys = (const float*)(d + o); o += b_k;
    memcpy(hp + o, keys, nc * 8); b.keys = (const float*)(d + o); o += b_k;
    memcpy(hp + o, first_pose, nc * 4); b.first_pose = (const int*)(d + o); o += b_fp;
    memcpy(hp + o, poses_cw, (size_t)batch * pose_cap * 96); b.poses = (const double*)(d + o); o += b_ps;
    memcpy(hp + o, n, (size_t)batch * 4); b.n = (const int*)(d + o); o += b_cnt;
    memcpy(hp + o, n_poses, (size_t)batch * 4); b.n_poses = (const int*)(d + o); o += b_cnt;
    memcpy(hp + o, cur_pose_cw, (size_t)batch * 96); b.cur = (const double*)(d + o); o += b_cur;
    VO_CUDA(ctx, cudaMemcpyAsync(d, hp, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* dout = d + in_bytes;
    b.keep = dout; b.out_lm = (float*)(dout + b_keep); b.out_kp = (float*)(dout + b_keep + b_lm);
    b.n_new = (int*)(dout + b_keep + b_lm + b_k); b.flags = (int*)(dout + b_keep + b_lm + b_k + b_cnt);
    b.lm_slot = (float*)(dout + out_bytes);
    VO_CUDA(ctx, cudaMemsetAsync(dout, 0, out_bytes, ctx->stream));
    triangulate_batch_kernel<<<dim3((cap + 63) / 64, batch), 64, 0, ctx->stream>>>(b);
    triangulate_compact_batch_kernel<<<batch, 1024, 0, ctx->stream>>>(b);
    ctx->launches += 2;
    VO_CUDA(ctx, cudaGetLastError());
    uint8_t* ho = hp + in_bytes;
    VO_CUDA(ctx, cudaMemcpyAsync(ho, dout, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    const int* flg = (const int*)(ho + b_keep + b_lm + b_k + b_cnt);
    for (int s = 0; s < batch; ++s)
        if (flg[s] & 1) return vo_set_err(ctx, B200VO_E_BADARG, "sequence %d: first_pose index outside [0, n_poses)", s);
    memcpy(too_short_baseline, ho, nc);
    memcpy(new_landmarks, ho + b_keep, nc * 12);
    memcpy(new_keypoints, ho + b_keep + b_lm, nc * 8);
    memcpy(n_new, ho + b_keep + b_lm + b_k, (size_t)batch * 4);
    return 0;
}

// ------------------------------------------------------------------------------------------
// f3: cv2.recoverPose(E, points1, points2, K) at reference :315 (default distanceThresh = 50).
//   recover_decompose_kernel   one thread: SVD of E (OpenCV's small-matrix Jacobi), det sign fix,
//                              R1 = U W Vt, R2 = U W^T Vt, t = U[:,2]; the four [R | +-t]
//   recover_cheirality_kernel  one thread per (point, pose): 4x4 DLT with P0 = [I|0] on
//                              K-normalised float64 points, Jacobi SVD, the three sign / depth
//                              tests -> mask 0 / 255; warp-aggregated counts
// The winner (most passing points; ties in the order 1, 2, 3, 4) is picked by the host from the
// four counts -- one 16-byte read the synchronous call needs anyway.
// ------------------------------------------------------------------------------------------
struct RecArgs {
    double E[9];
    double fx, fy, cx, cy, dist;
    int n;
    const float* p1; const float* p2;
    double* poses;     // [4][12] R | t, then R1, R2, t (21 doubles) for the host
    uint8_t* masks;    // [4][n]
    int* good;         // [4]
};

__device__ __forceinline__ double det3d(const double* M)
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C)
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

__global__ void recover_decompose_kernel(RecArgs a)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double W[3], U[9], Vt[9], At[9], T[9], R1[9], R2[9];
    jacobi_svd<3>(a.E, W, U, Vt, At);
    if (det3d(U) < 0) for (int k = 0; k < 9; ++k) U[k] = -U[k];
    if (det3d(Vt) < 0) for (int k = 0; k < 9; ++k) Vt[k] = -Vt[k];
    const double Wm[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1}, Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
    mat3_mul(U, Wm, T); mat3_mul(T, Vt, R1);
    mat3_mul(U, Wt, T); mat3_mul(T, Vt, R2);
    const double t[3] = {U[2], U[5], U[8]};
    for (int h = 0; h < 4; ++h) {
        const double* R = (h & 1) ? R2 : R1;
        const double sg = h < 2 ? 1.0 : -1.0;
        for (int k = 0; k < 9; ++k) a.poses[12 * h + k] = R[k];
        for (int k = 0; k < 3; ++k) a.poses[12 * h + 9 + k] = sg * t[k];
    }
}

__global__ void __launch_bounds__(128)
recover_cheirality_kernel(RecArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = blockIdx.y;
    bool ok = false;
    if (i < a.n) {
        double P[12];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) P[4 * r + c] = a.poses[12 * h + 3 * r + c];
            P[4 * r + 3] = a.poses[12 * h + 9 + r];
        }
        const double x1 = ((double)a.p1[2 * i] - a.cx) / a.fx, y1 = ((double)a.p1[2 * i + 1] - a.cy) / a.fy;
        const double x2 = ((double)a.p2[2 * i] - a.cx) / a.fx, y2 = ((double)a.p2[2 * i + 1] - a.cy) / a.fy;
        double A[16] = {-1, 0, x1, 0, 0, -1, y1, 0, 0, 0, 0, 0, 0, 0, 0, 0}, W[4], U[16], Vt[16], At[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) { A[8 + k] = x2 * P[8 + k] - P[k]; A[12 + k] = y2 * P[8 + k] - P[4 + k]; }
        jacobi_svd<4>(A, W, U, Vt, At);
        double Q0 = Vt[12], Q1 = Vt[13], Q2 = Vt[14], Q3 = Vt[15];
        ok = Q2 * Q3 > 0;
        Q0 /= Q3; Q1 /= Q3; Q2 /= Q3; Q3 /= Q3;
        ok = ok && (Q2 < a.dist);
        const double z = P[8] * Q0 + P[9] * Q1 + P[10] * Q2 + P[11] * Q3;
        ok = ok && (z > 0) && (z < a.dist);
        a.masks[(size_t)h * a.n + i] = ok ? 255 : 0;
    }
    const unsigned bm = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && bm) atomicAdd(a.good + h, __popc(bm));
}

extern "C" int b200vo_recover_pose(b200vo_ctx* ctx, const double E[9], const float* p1, const float* p2, int n,
                                   const double K[9], double distance_thresh, double R[9], double t[3], uint8_t* mask,
                                   int* n_good)
{
    if (!ctx || !E || !K || !R || !t || !n_good || n < 0 || (n > 0 && (!p1 || !p2 || !mask))) return B200VO_E_BADARG;
    VO_CUDA(ctx, cudaSetDevice(ctx->device));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    RecArgs a{};
    for (int k = 0; k < 9; ++k) a.E[k] = E[k];
    a.fx = K[0]; a.fy = K[4]; a.cx = K[2]; a.cy = K[5]; a.dist = distance_thresh; a.n = n;
    const int nn = n > 0 ? n : 1;
    const size_t b_p = vo_align((size_t)nn * 8, 256), b_m = vo_align((size_t)4 * nn, 256), b_small = 512;
    VO_TRY(vo_reserve(ctx, ctx->d_scratch[7], 2 * b_p + b_m + b_small));
    VO_TRY(vo_reserve_pinned(ctx, 2 * b_p + b_m + b_small));
    uint8_t* d = (uint8_t*)ctx->d_scratch[7].p;
    uint8_t* hp = (uint8_t*)ctx->h_pin;
    if (n > 0) { memcpy(hp, p1, (size_t)n * 8); memcpy(hp + b_p, p2, (size_t)n * 8); }
    VO_CUDA(ctx, cudaMemcpyAsync(d, hp, 2 * b_p, cudaMemcpyHostToDevice, ctx->stream));
    a.p1 = (const float*)d; a.p2 = (const float*)(d + b_p);
    a.masks = d + 2 * b_p;
    a.poses = (double*)(d + 2 * b_p + b_m); a.good = (int*)(d + 2 * b_p + b_m + 384);
    VO_CUDA(ctx, cudaMemsetAsync(a.poses, 0, b_small, ctx->stream));
    recover_decompose_kernel<<<1, 32, 0, ctx->stream>>>(a);
    ctx->launches++;
    if (n > 0) {
        recover_cheirality_kernel<<<dim3((n + 127) / 128, 4), 128, 0, ctx->stream>>>(a);
        ctx->launches++;
    }
    VO_CUDA(ctx, cudaGetLastError());
    VO_CUDA(ctx, cudaMemcpyAsync(hp + 2 * b_p, a.masks, b_m + b_small, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    const double* poses = (const double*)(hp + 2 * b_p + b_m);
    const int* good = (const int*)(hp + 2 * b_p + b_m + 384);
    int win;
    if (good[0] >= good[1] && good[0] >= good[2] && good[0] >= good[3]) win = 0;
    else if (good[1] >= good[0] && good[1] >= good[2] && good[1] >= good[3]) win = 1;
    else if (good[2] >= good[0] && good[2] >= good[1] && good[2] >= good[3]) win = 2;
    else win = 3;
    for (int k = 0; k < 9; ++k) R[k] = poses[12 * win + k];
    for (int k = 0; k < 3; ++k) t[k] = poses[12 * win + 9 + k];
    if (n > 0) memcpy(mask, hp + 2 * b_p + (size_t)win * n, (size_t)n);
    *n_good = good[win];
    return 0;
}
