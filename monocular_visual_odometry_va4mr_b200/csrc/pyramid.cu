// Image pyramid construction for the pyramidal LK tracker (replaces OpenCV's
// buildOpticalFlowPyramid / pyrDown behind cv2.calcOpticalFlowPyrLK, reference
// VisualOdometryPipeLine.py:281,287; spec SURVEY.md A.1).
//
// HBM layout: one "slab" per frame holding every level with a materialised REFLECT_101
// border of VO_BORDER pixels on all four sides (cv2 does the same with copyMakeBorder), so
// that the tracker's window reads and the 5x5 / 3x3 stencils never branch on the image edge.
// Row pitch is a multiple of 128 B and pixel (0,0) of every level is 32 B aligned.
#include "internal.cuh"

int vo_pyr_levels(int w, int h, int win_w, int win_h, int max_level)
{
    int levels = 1;
    for (int l = 1; l <= max_level && levels < VO_MAX_LEVELS; ++l) {
        w = (w + 1) / 2;
        h = (h + 1) / 2;
        if (w <= win_w || h <= win_h) break;
        ++levels;
    }
    return levels;
}

void vo_pyr_geom(int rows, int cols, int levels, PyrGeom* g)
{
    g->levels = levels;
    size_t off = 256;  // front pad: word-aligned staging may touch a few bytes before row -VO_BORDER
    int w = cols, h = rows;
    for (int l = 0; l < levels; ++l) {
        g->w[l] = w;
        g->h[l] = h;
        g->pitch[l] = (int)vo_align((size_t)w + 2 * VO_BORDER, 128);
        off = vo_align(off, 256);
        g->off[l] = off + (size_t)VO_BORDER * g->pitch[l] + VO_BORDER;
        off += (size_t)g->pitch[l] * (h + 2 * VO_BORDER);
        w = (w + 1) / 2;
        h = (h + 1) / 2;
    }
    g->slab_bytes = vo_align(off, 256) + 256;
}

Pyramid vo_pyramid_at(const PyrGeom& g, uint8_t* slab)
{
    Pyramid p;
    p.levels = g.levels;
    for (int l = 0; l < g.levels; ++l) {
        p.lv[l].base = slab + g.off[l];
        p.lv[l].w = g.w[l];
        p.lv[l].h = g.h[l];
        p.lv[l].pitch = g.pitch[l];
    }
    return p;
}

__device__ __forceinline__ int reflect101(int p, int len)
{
    // general (multi-bounce) REFLECT_101; len >= 2
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

// raw tight image -> bordered level 0.  One thread = one 16-byte store of the bordered row.
__global__ void __launch_bounds__(256)
pad_level0_kernel(const uint8_t* __restrict__ raw, size_t raw_stride, int w, int h,
                  uint8_t* __restrict__ slab, size_t slab_stride, size_t off0, int pitch)
{
    const int seq = blockIdx.z;
    const uint8_t* src = raw + seq * raw_stride;
    uint8_t* dst = slab + seq * slab_stride + off0;
    const int gx = (blockIdx.x * blockDim.x + threadIdx.x) * 16 - VO_BORDER;  // first of 16 px
    const int gy = blockIdx.y * blockDim.y + threadIdx.y - VO_BORDER;
    if (gx >= w + VO_BORDER || gy >= h + VO_BORDER) return;
    const int sy = reflect101(gy, h);
    const uint8_t* srow = src + (size_t)sy * w;
    uint32_t v[4];
    if (gx >= 0 && gx + 15 < w) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint8_t* s = srow + gx + 4 * k;
            v[k] = (uint32_t)__ldg(s) | ((uint32_t)__ldg(s + 1) << 8) | ((uint32_t)__ldg(s + 2) << 16) |
                   ((uint32_t)__ldg(s + 3) << 24);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t acc = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int x = gx + 4 * k + b;
                uint32_t px = (x < w + VO_BORDER) ? __ldg(srow + reflect101(x, w)) : 0u;
                acc |= px << (8 * b);
            }
            v[k] = acc;
        }
    }
    // pitch >= w + 2*VO_BORDER rounded up to 128, so a 16-byte store never leaves the row
    *reinterpret_cast<uint4*>(dst + (long long)gy * pitch + gx) = make_uint4(v[0], v[1], v[2], v[3]);
}

// level l (bordered) -> level l+1 (bordered).  One thread = 4 consecutive pixels of the
// bordered destination row; 5x5 binomial [1 4 6 4 1]^2, (sum+128)>>8.
__global__ void __launch_bounds__(256)
pyr_down_kernel(const uint8_t* __restrict__ slab_src, uint8_t* __restrict__ slab_dst, size_t slab_stride,
                size_t off_s, int sw, int sh, int spitch, size_t off_d, int dw, int dh, int dpitch)
{
    const int seq = blockIdx.z;
    const uint8_t* src = slab_src + seq * slab_stride + off_s;
    uint8_t* dst = slab_dst + seq * slab_stride + off_d;
    const int gx = (blockIdx.x * blockDim.x + threadIdx.x) * 4 - VO_BORDER;
    const int gy = blockIdx.y * blockDim.y + threadIdx.y - VO_BORDER;
    if (gx >= dw + VO_BORDER || gy >= dh + VO_BORDER) return;
    const int ry = reflect101(gy, dh);
    uint32_t out = 0;
    if (gx >= 0 && gx + 3 < dw) {
        // interior fast path: 4 outputs share an 11-byte source footprint per row
        int acc[4] = {0, 0, 0, 0};
        const int wy[5] = {1, 4, 6, 4, 1};
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const uint8_t* s = src + (long long)(2 * ry - 2 + r) * spitch + (2 * gx - 2);
            int p[11];
#pragma unroll
            for (int k = 0; k < 11; ++k) p[k] = __ldg(s + k);
#pragma unroll
            for (int o = 0; o < 4; ++o)
                acc[o] += wy[r] * (p[2 * o] + 4 * p[2 * o + 1] + 6 * p[2 * o + 2] + 4 * p[2 * o + 3] + p[2 * o + 4]);
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) out |= (uint32_t)((acc[o] + 128) >> 8) << (8 * o);
    } else {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            int x = gx + o;
            if (x >= dw + VO_BORDER) break;
            int rx = reflect101(x, dw);
            int acc = 0;
            const int wy[5] = {1, 4, 6, 4, 1};
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const uint8_t* s = src + (long long)(2 * ry - 2 + r) * spitch + (2 * rx - 2);
                acc += wy[r] * (__ldg(s) + 4 * __ldg(s + 1) + 6 * __ldg(s + 2) + 4 * __ldg(s + 3) + __ldg(s + 4));
            }
            out |= (uint32_t)((acc + 128) >> 8) << (8 * o);
        }
    }
    *reinterpret_cast<uint32_t*>(dst + (long long)gy * dpitch + gx) = out;
}

int vo_build_pyramids(b200vo_ctx* ctx, const uint8_t* d_raw, size_t raw_stride, int rows, int cols,
                      const PyrGeom& g, uint8_t* d_slab, size_t slab_stride, int batch)
{
    {
        dim3 block(32, 8);
        int gxn = (cols + 2 * VO_BORDER + 15) / 16;
        dim3 grid((gxn + block.x - 1) / block.x, (rows + 2 * VO_BORDER + block.y - 1) / block.y, batch);
        pad_level0_kernel<<<grid, block, 0, ctx->stream>>>(d_raw, raw_stride, cols, rows, d_slab, slab_stride,
                                                            g.off[0], g.pitch[0]);
        ctx->launches++;
    }
    for (int l = 1; l < g.levels; ++l) {
        dim3 block(32, 8);
        int gxn = (g.w[l] + 2 * VO_BORDER + 3) / 4;
        dim3 grid((gxn + block.x - 1) / block.x, (g.h[l] + 2 * VO_BORDER + block.y - 1) / block.y, batch);
        pyr_down_kernel<<<grid, block, 0, ctx->stream>>>(d_slab, d_slab, slab_stride, g.off[l - 1], g.w[l - 1],
                                                          g.h[l - 1], g.pitch[l - 1], g.off[l], g.w[l], g.h[l],
                                                          g.pitch[l]);
        ctx->launches++;
    }
    VO_CUDA(ctx, cudaGetLastError());
    return 0;
}
