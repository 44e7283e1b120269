// Bit-exact cv::RNG((uint64)-1) minimal-sample generator shared by the PnP and essential-matrix
// RANSACs (OpenCV RANSACPointSetRegistrator::getSubset): the raw 32-bit MWC stream is
// a fixed table; sample i draws idx = raw[k] % N with redraws on duplicates, so the whole sample
// sequence is a pure function of (N, model points).  One CTA per sequence: all residues in
// parallel, then a warp walks the samples 32 at a time and falls back to the sequential redraw loop
// only for the (rare) sample that contains a duplicate.
#pragma once
#include <cstdint>

template <int MP>
__global__ void __launch_bounds__(128)
ransac_samples_kernel(const uint32_t* __restrict__ rng_raw, int n_raw, const int* __restrict__ n_pts, int iters,
                      int* __restrict__ samples, int* __restrict__ flags)
{
    extern __shared__ int s_mod[];   // raw[k] % N
    const int b = blockIdx.x;
    const int N = n_pts[b];
    int* out = samples + (size_t)b * iters * MP;
    if (N < MP) {
        for (int k = threadIdx.x; k < iters * MP; k += blockDim.x) out[k] = -1;
        return;
    }
    if (N == MP) {   // count == modelPoints: a single direct solve on all points
        for (int k = threadIdx.x; k < iters * MP; k += blockDim.x) out[k] = k < MP ? k : -1;
        return;
    }
    for (int k = threadIdx.x; k < n_raw; k += blockDim.x) s_mod[k] = (int)(rng_raw[k] % (uint32_t)N);
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    int pos = 0, i0 = 0;
    while (i0 < iters) {
        const int i = i0 + lane, p = pos + MP * lane;
        const bool live = i < iters;
        const bool inb = p + MP - 1 < n_raw;
        int sv[MP];
        bool dup = false;
        if (live && inb) {
#pragma unroll
            for (int j = 0; j < MP; ++j) sv[j] = s_mod[p + j];
#pragma unroll
            for (int j = 1; j < MP; ++j)
#pragma unroll
                for (int m = 0; m < j; ++m) dup = dup || (sv[j] == sv[m]);
        }
        const unsigned stop = __ballot_sync(0xffffffffu, live && (dup || !inb));
        const int first = stop ? __ffs(stop) - 1 : 32;
        if (live && lane < first) {
#pragma unroll
            for (int j = 0; j < MP; ++j) out[MP * i + j] = sv[j];
        }
        if (first == 32) { i0 += 32; pos += 32 * MP; continue; }
        int npos = 0;
        if (lane == first) {   // getSubset's redraw loop, sequentially, for this one sample
            int q = p, idx[MP];
            bool okk = true;
#pragma unroll
            for (int j = 0; j < MP; ++j) idx[j] = -1;
            for (int j = 0; j < MP && okk; ++j) {
                for (;;) {
                    if (q >= n_raw) { okk = false; break; }
                    const int v = s_mod[q++];
                    bool d = false;
                    for (int m = 0; m < j; ++m) d = d || (idx[m] == v);
                    if (!d) { idx[j] = v; break; }
                }
            }
            if (!okk) {
#pragma unroll
                for (int j = 0; j < MP; ++j) idx[j] = -1;
                flags[b] |= 1;
                q = n_raw;
            }
#pragma unroll
            for (int j = 0; j < MP; ++j) out[MP * i + j] = idx[j];
            npos = q;
        }
        pos = __shfl_sync(0xffffffffu, npos, first);
        i0 += first + 1;
    }
}
