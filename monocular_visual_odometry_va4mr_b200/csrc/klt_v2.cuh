// klt_kernel_v2<WW,WH>: the tracker for the two window sizes the workloads use (15x15: the
// reference's options, main.py:36; 21x21: cv2's default), restructured around what the first ncu
// capture showed (profiles/r1a_klt_kernel_full.txt: the v1 kernel is instruction-issue bound,
// 16 k warp-instructions per point, DRAM and L2 idle).  Same arithmetic, bit-identical results:
//   * a lane owns a run of 8 horizontally adjacent window pixels (row y, x0 = 0, 8, 16) instead of
//     every 32nd pixel: 9 taps per row instead of 16, fetched as three aligned 32-bit shared-memory
//     words and re-aligned with funnel shifts;
//   * the bilinear interpolation I00*w00 + I01*w01 (+ next row) is two dp2a (16-bit weights x 8-bit
//     pixels) instead of four IMAD;
//   * the window of the NEXT image is staged in shared memory per level (window + 4 px margin,
//     restaged only if the point leaves it), so iterations never touch L1/L2;
//   * the template gradients live in registers; sum(Iwin*dI) is folded into two constants, so an
//     iteration needs neither Iwin nor a subtraction per pixel;
//   * the Scharr derivative of the previous image is 5 dp4a per pixel on the staged patch.
#pragma once

template <int WW, int WH>
struct KV2 {
    static constexpr int NSEG = (WW + 7) / 8;            // 8-pixel runs per window row
    static constexpr int NTASK = WH * NSEG;
    static constexpr int NROUND = (NTASK + 31) / 32;
    static constexpr int DW = WW + 1, DH = WH + 1;       // derivative tap grid
    static constexpr int DSEG = (DW + 7) / 8;
    static constexpr int DTASK = DH * DSEG;
    static constexpr int DROUND = (DTASK + 31) / 32;
    static constexpr int DS = ((NSEG * 8 + 1 + 3) / 4) * 4 > DSEG * 8 ? ((NSEG * 8 + 1 + 3) / 4) * 4 : DSEG * 8;  // ints per der row
    static constexpr int PS0 = ((WW + 3 + 3 + 3) / 4) * 4 + 8;  // patch row stride (bytes), room for 4-word reads
    static constexpr int PS = (PS0 / 4) % 2 == 0 ? PS0 + 4 : PS0; // odd number of words: rows spread over the banks
    static constexpr int PROWS = WH + 3;
    static constexpr int MARGIN = 4;
    static constexpr int JSV = ((WW + 1 + 2 * MARGIN + 3 + 3) / 4) * 4;   // valid staged bytes per row
    static constexpr int JS0 = JSV + 8;
    static constexpr int JS = (JS0 / 4) % 2 == 0 ? JS0 + 4 : JS0;   // row stride (bytes), odd number of words
    static constexpr int JR = WH + 1 + 2 * MARGIN;
    static constexpr int IS = NSEG * 8;                   // Iwin row stride (shorts)
    static constexpr int B_PATCH = ((PS * PROWS + 15) / 16) * 16;
    static constexpr int B_J = ((JS * JR + 15) / 16) * 16;
    static constexpr int B_DER = DS * DH * 4;             // per plane
    static constexpr int B_IWIN = ((IS * WH * 2 + 15) / 16) * 16;
    static constexpr int PER_WARP = 2 * B_PATCH + B_J + 2 * B_DER + B_IWIN;   // two patch buffers (prefetch)
    // resident warps per SM the shared memory allows (227 KB): 32 (64 registers) or 24 (80); the register budget follows it
    static constexpr int MIN_CTAS = ((227 * 1024) / (KLT_WARPS * PER_WARP + 64 + 1024) * KLT_WARPS >= 32 ? 32 : 24) / KLT_WARPS;
};

// 16-bit weights x 8-bit pixels.  The weights are SIGNED halves: iw11 = 2^14 - iw00 - iw01 - iw10 is -1
// when the three rounded weights add up to 2^14 + 1 (a fractional part of ~0 in x or y, about one
// position in a thousand) -- cv2 carries that -1 through its integer arithmetic, and so must this.
__device__ __forceinline__ uint32_t dp2a_lo_u(uint32_t a, uint32_t b, uint32_t c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"((int)c));
    return (uint32_t)d;
}
__device__ __forceinline__ uint32_t dp2a_hi_u(uint32_t a, uint32_t b, uint32_t c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"((int)c));
    return (uint32_t)d;
}
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c)   // unsigned pixels x signed coefficients
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// asynchronous global -> shared staging (LDGSTS): the copy proceeds while the warp computes
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int n)   // wait until at most n groups are pending
{
    if (n <= 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else if (n == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 2;" ::: "memory");
}
// Each lane owns a fixed (row-in-round, word) slot: 32 / WPR rows per round, so the loop is a
// pointer increment per copy instead of a division per copy.
template <int WPR, int ROWS>
__device__ __forceinline__ void stage_rows_async(uint8_t* dst, const uint8_t* src_aligned, int pitch, int lane)
{
    constexpr int RPR = 32 / WPR;
    const int r0 = lane / WPR, c = lane - r0 * WPR;
    if (r0 >= RPR) return;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(src_aligned) + (long long)r0 * (pitch >> 2) + c;
    uint32_t* d = reinterpret_cast<uint32_t*>(dst) + r0 * WPR + c;
    const long long sstep = (long long)RPR * (pitch >> 2);
#pragma unroll
    for (int r = 0; r < (ROWS + RPR - 1) / RPR; ++r) {
        if (r * RPR + r0 < ROWS) cp_async4(d, src);
        src += sstep;
        d += RPR * WPR;
    }
}

// 8 interpolated intensities (13-bit, = 32 x grey level) of the run starting at byte offset `off`
// of `row0` (and the row below it) in a shared-memory image with row stride RS.
// wt = iw00 | iw01 << 16, wb = iw10 | iw11 << 16 (signed 16-bit halves).  off & 3 == sh for every lane of the warp.
template <int RS>
__device__ __forceinline__ void interp_run8(const uint8_t* img, int off, int sh8, uint32_t wt, uint32_t wb, int* I)
{
    const uint32_t* t = reinterpret_cast<const uint32_t*>(img + (off & ~3));
    const uint32_t* b = reinterpret_cast<const uint32_t*>(img + (off & ~3) + RS);
    const uint32_t t0 = t[0], t1 = t[1], t2 = t[2], b0 = b[0], b1 = b[1], b2 = b[2];
    const uint32_t ta0 = __funnelshift_r(t0, t1, sh8), ta1 = __funnelshift_r(t1, t2, sh8), ta2 = t2 >> sh8;
    const uint32_t ba0 = __funnelshift_r(b0, b1, sh8), ba1 = __funnelshift_r(b1, b2, sh8), ba2 = b2 >> sh8;
    const uint32_t ts0 = __funnelshift_r(ta0, ta1, 8), ts1 = __funnelshift_r(ta1, ta2, 8);
    const uint32_t bs0 = __funnelshift_r(ba0, ba1, 8), bs1 = __funnelshift_r(ba1, ba2, 8);
    const uint32_t R = 1u << (W_BITS - 5 - 1);
    I[0] = (int)(dp2a_lo_u(wb, ba0, dp2a_lo_u(wt, ta0, R)) >> (W_BITS - 5));
    I[1] = (int)(dp2a_lo_u(wb, bs0, dp2a_lo_u(wt, ts0, R)) >> (W_BITS - 5));
    I[2] = (int)(dp2a_hi_u(wb, ba0, dp2a_hi_u(wt, ta0, R)) >> (W_BITS - 5));
    I[3] = (int)(dp2a_hi_u(wb, bs0, dp2a_hi_u(wt, ts0, R)) >> (W_BITS - 5));
    I[4] = (int)(dp2a_lo_u(wb, ba1, dp2a_lo_u(wt, ta1, R)) >> (W_BITS - 5));
    I[5] = (int)(dp2a_lo_u(wb, bs1, dp2a_lo_u(wt, ts1, R)) >> (W_BITS - 5));
    I[6] = (int)(dp2a_hi_u(wb, ba1, dp2a_hi_u(wt, ta1, R)) >> (W_BITS - 5));
    I[7] = (int)(dp2a_hi_u(wb, bs1, dp2a_hi_u(wt, ts1, R)) >> (W_BITS - 5));
}

// Same with the image given as a 32-bit shared-memory address that lives in a register (the iteration loop:
// the compiler otherwise rematerialises the warp's shared-memory base from %tid / %cluster_ctaid every iteration).
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int RS>
__device__ __forceinline__ void interp_run8_s(uint32_t img_s, int off, int sh8, uint32_t wt, uint32_t wb, int* I)
{
    const uint32_t t = img_s + (uint32_t)(off & ~3), b = t + RS;
    const uint32_t t0 = lds_u32(t), t1 = lds_u32(t + 4), t2 = lds_u32(t + 8), b0 = lds_u32(b), b1 = lds_u32(b + 4), b2 = lds_u32(b + 8);
    const uint32_t ta0 = __funnelshift_r(t0, t1, sh8), ta1 = __funnelshift_r(t1, t2, sh8), ta2 = t2 >> sh8;
    const uint32_t ba0 = __funnelshift_r(b0, b1, sh8), ba1 = __funnelshift_r(b1, b2, sh8), ba2 = b2 >> sh8;
    const uint32_t ts0 = __funnelshift_r(ta0, ta1, 8), ts1 = __funnelshift_r(ta1, ta2, 8);
    const uint32_t bs0 = __funnelshift_r(ba0, ba1, 8), bs1 = __funnelshift_r(ba1, ba2, 8);
    const uint32_t R = 1u << (W_BITS - 5 - 1);
    I[0] = (int)(dp2a_lo_u(wb, ba0, dp2a_lo_u(wt, ta0, R)) >> (W_BITS - 5));
    I[1] = (int)(dp2a_lo_u(wb, bs0, dp2a_lo_u(wt, ts0, R)) >> (W_BITS - 5));
    I[2] = (int)(dp2a_hi_u(wb, ba0, dp2a_hi_u(wt, ta0, R)) >> (W_BITS - 5));
    I[3] = (int)(dp2a_hi_u(wb, bs0, dp2a_hi_u(wt, ts0, R)) >> (W_BITS - 5));
    I[4] = (int)(dp2a_lo_u(wb, ba1, dp2a_lo_u(wt, ta1, R)) >> (W_BITS - 5));
    I[5] = (int)(dp2a_lo_u(wb, bs1, dp2a_lo_u(wt, ts1, R)) >> (W_BITS - 5));
    I[6] = (int)(dp2a_hi_u(wb, ba1, dp2a_hi_u(wt, ta1, R)) >> (W_BITS - 5));
    I[7] = (int)(dp2a_hi_u(wb, bs1, dp2a_hi_u(wt, ts1, R)) >> (W_BITS - 5));
}

template <int WW, int WH>
__global__ void __launch_bounds__(KLT_WARPS * 32, (KV2<WW, WH>::MIN_CTAS))
klt_kernel_v2(const KltArgs a)
{
    using C = KV2<WW, WH>;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * KLT_WARPS + warp;
    const int cap_all = a.cap[0] + a.cap[1];
    const int seq = (int)(gw / cap_all);
    int pi = (int)(gw - (long long)seq * cap_all);
    if (seq >= a.batch) return;
    const int seg = pi >= a.cap[0] ? 1 : 0;
    pi -= seg ? a.cap[0] : 0;
    const int n_here = a.n_pts[seg] ? a.n_pts[seg][seq] : a.n_fixed;
    if (pi >= n_here) return;

    uint8_t* wbase = smem + (size_t)warp * C::PER_WARP;
    uint8_t* patch_buf = wbase;                                  // two buffers of B_PATCH
    uint8_t* jreg = wbase + 2 * C::B_PATCH;
    uint32_t jreg_s;      // opaque to the compiler: stays in a register instead of being rebuilt per iteration
    asm volatile("mov.u32 %0, %1;" : "=r"(jreg_s) : "r"((uint32_t)__cvta_generic_to_shared(jreg)));
    int* derx = reinterpret_cast<int*>(wbase + 2 * C::B_PATCH + C::B_J);
    int* dery = derx + C::DS * C::DH;
    short* Iwin = reinterpret_cast<short*>(wbase + 2 * C::B_PATCH + C::B_J + 2 * C::B_DER);

    const uint8_t* prev = a.prev + (size_t)seq * a.prev_stride;
    const uint8_t* next = a.next + (size_t)seq * a.next_stride;
    const size_t pidx = (size_t)seq * a.cap[seg] + pi;
    const float px0 = a.pts[seg][2 * pidx], py0 = a.pts[seg][2 * pidx + 1];
    const float hwx = (WW - 1) * 0.5f, hwy = (WH - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);

    // this lane's template runs: (row ty[r], first column tx[r], live pixels tl[r]) per round
    int ty[C::NROUND], tx[C::NROUND], tl[C::NROUND];
#pragma unroll
    for (int r = 0; r < C::NROUND; ++r) {
        const int t = r * 32 + lane;
        ty[r] = t / C::NSEG;
        tx[r] = (t - ty[r] * C::NSEG) * 8;
        tl[r] = t < C::NTASK ? min(8, WW - tx[r]) : 0;
        if (t >= C::NTASK) { ty[r] = 0; tx[r] = 0; }
    }

    // this lane's byte offsets inside the staged J window, pinned to registers (see jreg_s)
    int joffl[C::NROUND];
#pragma unroll
    for (int r = 0; r < C::NROUND; ++r) asm volatile("mov.s32 %0, %1;" : "=r"(joffl[r]) : "r"(ty[r] * C::JS + tx[r]));
    float outx = 0.f, outy = 0.f;
    int st = 1;
    float e = 0.f;
    int pb = 0, pf_level = -1;     // patch buffer in use; level whose patch was prefetched into it
    const float eps_lo = a.eps_lo, eps_hi = a.eps_hi;

    for (int level = a.levels - 1; level >= 0; --level) {
        int lw, lh;
        asm volatile("mov.s32 %0, %1;" : "=r"(lw) : "r"(a.w[level]));
        asm volatile("mov.s32 %0, %1;" : "=r"(lh) : "r"(a.h[level]));
        const int pitch = a.pitch[level];
        const uint8_t* I = prev + a.off[level];
        const uint8_t* J = next + a.off[level];
        const float sc = (float)(1.0 / (double)(1 << level));
        float ppx = __fmul_rn(px0, sc), ppy = __fmul_rn(py0, sc);
        float nx, ny;
        if (level == a.levels - 1) { nx = ppx; ny = ppy; }
        else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
        outx = nx; outy = ny;
        ppx = __fsub_rn(ppx, hwx); ppy = __fsub_rn(ppy, hwy);
        const int ipx = floor_to_int(ppx), ipy = floor_to_int(ppy);
        if (ipx < -WW || ipx >= lw || ipy < -WH || ipy >= lh) {
            if (level == 0) { st = 0; e = 0.f; }
            continue;
        }
        int iw00, iw01, iw10, iw11;
        bilinear_weights(__fsub_rn(ppx, (float)ipx), __fsub_rn(ppy, (float)ipy), iw00, iw01, iw10, iw11);

        // ---- staging pipeline (cp.async): this level's patch of I was prefetched during the previous
        // level; now start (1) the window of J around the initial position and (2) the NEXT level's
        // patch of I (its position depends only on the input point), and overlap both with the
        // derivative / template arithmetic of this level ----
        __syncwarp();
        uint8_t* patch = patch_buf + pb * C::B_PATCH;
        const uint8_t* src0 = I + (long long)(ipy - 1) * pitch + (ipx - 1);
        const int mis = (int)(reinterpret_cast<uintptr_t>(src0) & 3);
        if (pf_level != level) {   // not prefetched (top level, or the previous level was skipped)
            stage_rows_async<C::PS / 4, C::PROWS>(patch, src0 - mis, pitch, lane);
            cp_async_commit();
        }
        const float jx0 = __fsub_rn(nx, hwx), jy0 = __fsub_rn(ny, hwy);
        int rx0 = 0x40000000, ry0 = 0;   // staged J region origin (none)
        {
            const int inx = floor_to_int(jx0), iny = floor_to_int(jy0);
            if (!(inx < -WW || inx >= lw || iny < -WH || iny >= lh)) {
                ry0 = iny - C::MARGIN;
                const uint8_t* s0 = J + (long long)ry0 * pitch + (inx - C::MARGIN);
                const int m2 = (int)(reinterpret_cast<uintptr_t>(s0) & 3);
                rx0 = inx - C::MARGIN - m2;
                stage_rows_async<C::JS / 4, C::JR>(jreg, s0 - m2, pitch, lane);
            }
            cp_async_commit();
        }
        bool pf_next = false;
        if (level > 0) {
            const int nl = level - 1;
            const float sc2 = (float)(1.0 / (double)(1 << nl));
            const int qx = floor_to_int(__fsub_rn(__fmul_rn(px0, sc2), hwx)), qy = floor_to_int(__fsub_rn(__fmul_rn(py0, sc2), hwy));
            if (!(qx < -WW || qx >= a.w[nl] || qy < -WH || qy >= a.h[nl])) {
                const uint8_t* n0 = prev + a.off[nl] + (long long)(qy - 1) * a.pitch[nl] + (qx - 1);
                const int m3 = (int)(reinterpret_cast<uintptr_t>(n0) & 3);
                stage_rows_async<C::PS / 4, C::PROWS>(patch_buf + (pb ^ 1) * C::B_PATCH, n0 - m3, a.pitch[nl], lane);
                pf_next = true;
            }
            cp_async_commit();
        }
        cp_async_wait_pending(level > 0 ? 2 : 1);   // everything older than (J, next patch): this level's patch
        __syncwarp();
        // patch pixel (x, y), x in [-1, WW+1], y in [-1, WH+1], lives at byte (y+1)*PS + mis + 1 + x

        // ---- Scharr on the tap grid: 5 dp4a per pixel, zero outside the image ----
#pragma unroll
        for (int r = 0; r < C::DROUND; ++r) {
            const int t = r * 32 + lane;
            if (t < C::DTASK) {
                const int gy = t / C::DSEG, gx0 = (t - gy * C::DSEG) * 8;
                const int off = gy * C::PS + mis + gx0;      // byte of pixel (gx0-1, gy-1)
                const int sh8 = (off & 3) * 8;
                int ix[8], iy[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { ix[k] = 0; iy[k] = 0; }
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                    const uint32_t* w = reinterpret_cast<const uint32_t*>(patch + (off & ~3) + rr * C::PS);
                    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
                    const uint32_t a0 = __funnelshift_r(w0, w1, sh8), a1 = __funnelshift_r(w1, w2, sh8), a2 = __funnelshift_r(w2, w3, sh8);
                    uint32_t q[8];
                    q[0] = a0; q[1] = __funnelshift_r(a0, a1, 8); q[2] = __funnelshift_r(a0, a1, 16); q[3] = __funnelshift_r(a0, a1, 24);
                    q[4] = a1; q[5] = __funnelshift_r(a1, a2, 8); q[6] = __funnelshift_r(a1, a2, 16); q[7] = __funnelshift_r(a1, a2, 24);
                    // q[k] = pixels (x-1, x, x+1, x+2) of row gy-1+rr for x = gx0+k
                    const int cx = rr == 1 ? 0x000A00F6 : 0x000300FD;                  // (-10,0,10,0) / (-3,0,3,0)
                    const int cy = rr == 0 ? 0x00FDF6FD : 0x00030A03;                  // (-3,-10,-3,0) / (3,10,3,0)
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        ix[k] = dp4a_us(q[k], cx, ix[k]);
                        if (rr != 1) iy[k] = dp4a_us(q[k], cy, iy[k]);
                    }
                }
                const int X0 = ipx + gx0, Y = ipy + gy;
                const bool rowok = Y >= 0 && Y < lh;
                if (!(rowok && X0 >= 0 && X0 + 7 < lw)) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const bool ok = rowok && X0 + k >= 0 && X0 + k < lw;
                        ix[k] = ok ? ix[k] : 0; iy[k] = ok ? iy[k] : 0;
                    }
                }
                int4* dxp = reinterpret_cast<int4*>(derx + gy * C::DS + gx0);
                int4* dyp = reinterpret_cast<int4*>(dery + gy * C::DS + gx0);
                dxp[0] = make_int4(ix[0], ix[1], ix[2], ix[3]); dxp[1] = make_int4(ix[4], ix[5], ix[6], ix[7]);
                dyp[0] = make_int4(iy[0], iy[1], iy[2], iy[3]); dyp[1] = make_int4(iy[4], iy[5], iy[6], iy[7]);
            }
        }
        __syncwarp();

        // ---- template: Iwin (smem), dIx/dIy (registers), normal matrix, sum(Iwin*dI) ----
        const uint32_t wt = (uint32_t)iw00 | ((uint32_t)iw01 << 16), wb = (uint32_t)iw10 | ((uint32_t)iw11 << 16);
        int gxr[C::NROUND][8], gyr[C::NROUND][8];
        int sA11 = 0, sA12 = 0, sA22 = 0, c1 = 0, c2 = 0;
#pragma unroll
        for (int r = 0; r < C::NROUND; ++r) {
            const int off = (ty[r] + 1) * C::PS + mis + 1 + tx[r];
            int Iv[8];
            interp_run8<C::PS>(patch, off, (off & 3) * 8, wt, wb, Iv);
            const int* d0 = derx + ty[r] * C::DS + tx[r];
            const int* e0 = dery + ty[r] * C::DS + tx[r];
            int dx0[9], dx1[9], dy0[9], dy1[9];
            {
                const int4 p0 = *reinterpret_cast<const int4*>(d0), p1 = *reinterpret_cast<const int4*>(d0 + 4);
                const int4 q0 = *reinterpret_cast<const int4*>(d0 + C::DS), q1 = *reinterpret_cast<const int4*>(d0 + C::DS + 4);
                dx0[0] = p0.x; dx0[1] = p0.y; dx0[2] = p0.z; dx0[3] = p0.w; dx0[4] = p1.x; dx0[5] = p1.y; dx0[6] = p1.z; dx0[7] = p1.w; dx0[8] = d0[8];
                dx1[0] = q0.x; dx1[1] = q0.y; dx1[2] = q0.z; dx1[3] = q0.w; dx1[4] = q1.x; dx1[5] = q1.y; dx1[6] = q1.z; dx1[7] = q1.w; dx1[8] = d0[C::DS + 8];
                const int4 r0 = *reinterpret_cast<const int4*>(e0), r1 = *reinterpret_cast<const int4*>(e0 + 4);
                const int4 s0 = *reinterpret_cast<const int4*>(e0 + C::DS), s1 = *reinterpret_cast<const int4*>(e0 + C::DS + 4);
                dy0[0] = r0.x; dy0[1] = r0.y; dy0[2] = r0.z; dy0[3] = r0.w; dy0[4] = r1.x; dy0[5] = r1.y; dy0[6] = r1.z; dy0[7] = r1.w; dy0[8] = e0[8];
                dy1[0] = s0.x; dy1[1] = s0.y; dy1[2] = s0.z; dy1[3] = s0.w; dy1[4] = s1.x; dy1[5] = s1.y; dy1[6] = s1.z; dy1[7] = s1.w; dy1[8] = e0[C::DS + 8];
            }
            short iws[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool livek = k < tl[r];
                int ixv = (dx0[k] * iw00 + dx0[k + 1] * iw01 + dx1[k] * iw10 + dx1[k + 1] * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
                int iyv = (dy0[k] * iw00 + dy0[k + 1] * iw01 + dy1[k] * iw10 + dy1[k + 1] * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
                ixv = livek ? ixv : 0; iyv = livek ? iyv : 0;
                gxr[r][k] = ixv; gyr[r][k] = iyv;
                iws[k] = (short)Iv[k];
                sA11 += ixv * ixv; sA12 += ixv * iyv; sA22 += iyv * iyv;
                c1 += Iv[k] * ixv; c2 += Iv[k] * iyv;
            }
            if (tl[r] > 0) {
                uint4 pk;
                pk.x = (uint16_t)iws[0] | ((uint32_t)(uint16_t)iws[1] << 16); pk.y = (uint16_t)iws[2] | ((uint32_t)(uint16_t)iws[3] << 16);
                pk.z = (uint16_t)iws[4] | ((uint32_t)(uint16_t)iws[5] << 16); pk.w = (uint16_t)iws[6] | ((uint32_t)(uint16_t)iws[7] << 16);
                *reinterpret_cast<uint4*>(Iwin + ty[r] * C::IS + tx[r]) = pk;
            }
        }
        const float A11 = __fmul_rn(__ll2float_rn(warp_sum_i64(sA11)), FLT_SCALE);
        const float A12 = __fmul_rn(__ll2float_rn(warp_sum_i64(sA12)), FLT_SCALE);
        const float A22 = __fmul_rn(__ll2float_rn(warp_sum_i64(sA22)), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float minEig = __fdiv_rn(
            __fsub_rn(__fadd_rn(A22, A11),
                      __fsqrt_rn(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
            (float)(2 * WW * WH));
        // the prefetched patch becomes the current one at the next level, whatever happens below
        if (pf_next) { pb ^= 1; pf_level = level - 1; }
        if (minEig < a.min_eig_thr || D < 1.192092896e-07f) {
            if (level == 0) st = 0;
            cp_async_wait_pending(level > 0 ? 1 : 0);   // drain the J window copy before jreg is reused
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = jx0; ny = jy0;
        float pdx = 0.f, pdy = 0.f;
        cp_async_wait_pending(level > 0 ? 1 : 0);   // the window of J has landed (the next patch may still fly)
        __syncwarp();

        for (int j = 0; j < a.max_count; ++j) {
            const int inx = floor_to_int(nx), iny = floor_to_int(ny);
            if (inx < -WW || inx >= lw || iny < -WH || iny >= lh) {
                if (level == 0) st = 0;
                break;
            }
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), iw00, iw01, iw10, iw11);
            if (inx < rx0 || inx + WW + 2 > rx0 + C::JSV || iny < ry0 || iny + WH + 1 >= ry0 + C::JR) {
                // (re)stage the window of J with a margin around the current position
                __syncwarp();
                ry0 = iny - C::MARGIN;
                const uint8_t* s0 = J + (long long)ry0 * pitch + (inx - C::MARGIN);
                const int m2 = (int)(reinterpret_cast<uintptr_t>(s0) & 3);
                rx0 = inx - C::MARGIN - m2;
                stage_rows_async<C::JS / 4, C::JR>(jreg, s0 - m2, pitch, lane);
                cp_async_commit();
                cp_async_wait_pending(0);
                __syncwarp();
            }
            const uint32_t jt = (uint32_t)iw00 | ((uint32_t)iw01 << 16), jb = (uint32_t)iw10 | ((uint32_t)iw11 << 16);
            const int joff = (iny - ry0) * C::JS + (inx - rx0);
            int sb1 = -c1, sb2 = -c2;
#pragma unroll
            for (int r = 0; r < C::NROUND; ++r) {
                const int off = joff + joffl[r];
                int Iv[8];
                interp_run8_s<C::JS>(jreg_s, off, (off & 3) * 8, jt, jb, Iv);
#pragma unroll
                for (int k = 0; k < 8; ++k) { sb1 += Iv[k] * gxr[r][k]; sb2 += Iv[k] * gyr[r][k]; }
            }
            const float b1 = __fmul_rn(__ll2float_rn(warp_sum_i64(sb1)), FLT_SCALE);
            const float b2 = __fmul_rn(__ll2float_rn(warp_sum_i64(sb2)), FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            outx = __fadd_rn(nx, hwx); outy = __fadd_rn(ny, hwy);
            // cv2: delta.ddot(delta) <= eps in double; decided in float unless within 1e-6 of the threshold
            const float s2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
            bool conv = s2 < eps_lo;
            if (!conv && !(s2 > eps_hi))
                conv = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= a.eps_sq;
            if (conv) break;
            if (j > 0 && fabsf(__fadd_rn(dx, pdx)) < 0.01f && fabsf(__fadd_rn(dy, pdy)) < 0.01f) {
                outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f));
                outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }

        if (st && level == 0) {
            const float qx = __fsub_rn(outx, hwx), qy = __fsub_rn(outy, hwy);
            const int inx = floor_to_int(qx), iny = floor_to_int(qy);
            if (inx < -WW || inx >= lw || iny < -WH || iny >= lh) {
                st = 0;
                continue;
            }
            bilinear_weights(__fsub_rn(qx, (float)inx), __fsub_rn(qy, (float)iny), iw00, iw01, iw10, iw11);
            if (inx < rx0 || inx + WW + 2 > rx0 + C::JSV || iny < ry0 || iny + WH + 1 >= ry0 + C::JR) {
                __syncwarp();
                ry0 = iny - C::MARGIN;
                const uint8_t* s0 = J + (long long)ry0 * pitch + (inx - C::MARGIN);
                const int m2 = (int)(reinterpret_cast<uintptr_t>(s0) & 3);
                rx0 = inx - C::MARGIN - m2;
                stage_rows_async<C::JS / 4, C::JR>(jreg, s0 - m2, pitch, lane);
                cp_async_commit();
                cp_async_wait_pending(0);
                __syncwarp();
            }
            const uint32_t jt = (uint32_t)iw00 | ((uint32_t)iw01 << 16), jb = (uint32_t)iw10 | ((uint32_t)iw11 << 16);
            const int joff = (iny - ry0) * C::JS + (inx - rx0);
            int se = 0;
#pragma unroll
            for (int r = 0; r < C::NROUND; ++r) {
                const int off = joff + ty[r] * C::JS + tx[r];
                int Iv[8];
                interp_run8<C::JS>(jreg, off, (off & 3) * 8, jt, jb, Iv);
                const uint4 pk = *reinterpret_cast<const uint4*>(Iwin + ty[r] * C::IS + tx[r]);
                const uint32_t pw4[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int iwv = (int)(short)((pw4[k >> 1] >> ((k & 1) * 16)) & 0xFFFF);
                    se += (k < tl[r]) ? abs(Iv[k] - iwv) : 0;
                }
            }
            e = __fdiv_rn(__ll2float_rn(warp_sum_i64(se)), (float)(32 * WW * WH));
        }
    }
    if (lane == 0) {
        a.out[seg][2 * pidx] = outx;
        a.out[seg][2 * pidx + 1] = outy;
        a.status[seg][pidx] = (uint8_t)st;
        if (a.err[seg]) a.err[seg][pidx] = e;
    }
}
