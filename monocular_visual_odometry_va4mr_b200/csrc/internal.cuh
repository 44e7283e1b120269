// Internal declarations shared by the sm_100a kernels and the C-ABI glue of libb200vo.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include "../../include/b200vo.h"

#define VO_MAX_LEVELS 12
#define VO_BORDER 32  // materialised REFLECT_101 border around every pyramid level (pixels)

// One pyramid level of one frame: `base` points at pixel (0,0); rows/cols in
// [-VO_BORDER, dim+VO_BORDER) are valid memory holding the REFLECT_101 extension.
struct PyrLevel {
    uint8_t* base;
    int w, h, pitch;
};
struct Pyramid {
    int levels;
    PyrLevel lv[VO_MAX_LEVELS];
};
// Geometry of a pyramid slab (one frame incl. all levels and borders) -- shared by every
// frame of the same (rows, cols, levels).
struct PyrGeom {
    int levels;
    int w[VO_MAX_LEVELS], h[VO_MAX_LEVELS], pitch[VO_MAX_LEVELS];
    size_t off[VO_MAX_LEVELS];  // byte offset of pixel (0,0) of level l inside the slab
    size_t slab_bytes;
};

struct KltParams {
    int win_w, win_h;
    int max_count;
    double eps_sq;     // already squared
    float min_eig_thr;
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct FrameSlot {
    DevBuf slab;
    PyrGeom geom{};
    int rows = 0, cols = 0;
    bool valid = false;
};

struct b200vo_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    char err[1024] = {0};
    long long launches = 0;
    float last_ms = 0.f;
    int num_sms = 148;
    // staging
    DevBuf d_stage_img[2];   // raw uploaded image (tight rows)
    void* h_pin = nullptr;   // pinned host staging
    size_t h_pin_cap = 0;
    DevBuf d_scratch[8];     // generic device scratch (per-API use)
    FrameSlot slots[B200VO_MAX_SLOTS + 2];  // +2 internal slots for the stateless cv2-style call
    int* knn_bad_flag = nullptr;   // device flag of the last kNN call (inside d_scratch[3])
    DevBuf d_klt_queue;      // ring of work-queue counters for the persistent tracker kernel
    unsigned klt_queue_next = 0;
    DevBuf d_sift;           // SIFT scale-space workspace (sift.cu)
    DevBuf d_rng;            // raw cv::RNG stream (uint32)
    int n_rng = 0;
    // tensor-map encode entry point (driver API, fetched lazily)
    void* encode_tiled = nullptr;
};

int vo_set_err(b200vo_ctx* ctx, int code, const char* fmt, ...);
int vo_cuda_fail(b200vo_ctx* ctx, cudaError_t e, const char* what);
int vo_reserve(b200vo_ctx* ctx, DevBuf& b, size_t bytes);
int vo_reserve_pinned(b200vo_ctx* ctx, size_t bytes);

#define VO_CUDA(ctx, call)                                         \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return vo_cuda_fail(ctx, _e, #call); \
    } while (0)

#define VO_TRY(expr)            \
    do {                        \
        int _rc = (expr);       \
        if (_rc != 0) return _rc; \
    } while (0)

static inline size_t vo_align(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- pyramid.cu ----
int vo_pyr_levels(int w, int h, int win_w, int win_h, int max_level);
void vo_pyr_geom(int rows, int cols, int levels, PyrGeom* g);
Pyramid vo_pyramid_at(const PyrGeom& g, uint8_t* slab);
// raw (tight, device) -> bordered level 0, then levels 1..L-1; `batch` slabs/raws strided.
int vo_build_pyramids(b200vo_ctx* ctx, const uint8_t* d_raw, size_t raw_stride, int rows, int cols,
                      const PyrGeom& g, uint8_t* d_slab, size_t slab_stride, int batch);

// ---- gftt.cu: detection on device-resident bordered level-0 images, one per batch entry ----
#define VO_GFTT_SMALL 4          // ints per image in the result block: max bits | n candidates | n corners | spare
#define VO_GFTT_BATCH_CAP 32768  // candidates per image the batched path can rank
size_t vo_gftt_batch_workspace(int rows, int cols, int batch, int max_corners);
int vo_gftt_batch_launch(b200vo_ctx* ctx, const uint8_t* d_img0, size_t img_stride, int pitch, int rows, int cols, int batch,
                         int max_corners, double quality, double min_dist, void* ws, float** d_corners_out, int** d_small_out);

// ---- klt.cu ----
struct KltPointSet {
    int cap;            // slots per sequence
    const int* n;       // [batch] live points per sequence (device) or nullptr -> n_fixed
    const float* pts;   // [batch][cap][2]
    float* out;         // [batch][cap][2]
    uint8_t* status;    // [batch][cap]
    float* err;         // [batch][cap] or nullptr
};
void vo_klt_trace_read(b200vo_ctx* ctx, std::vector<unsigned long long>& out);   // B200VO_TRACE_FILE aid
int vo_klt_launch2(b200vo_ctx* ctx, const PyrGeom& g, const uint8_t* d_prev_slab, size_t prev_stride,
                   const uint8_t* d_next_slab, size_t next_stride, int batch, const KltPointSet* sets, int n_sets,
                   int n_fixed, const KltParams& kp);
// Tracks points of `batch` independent frame pairs.  pts/next/status/err are
// [batch][cap] arrays; n_pts (device) gives the live count per pair (or nullptr -> n_fixed).
int vo_klt_launch(b200vo_ctx* ctx, const PyrGeom& g, const uint8_t* d_prev_slab, size_t prev_stride,
                  const uint8_t* d_next_slab, size_t next_stride, int batch, int cap,
                  const int* d_n_pts, int n_fixed, const float* d_pts, float* d_next,
                  uint8_t* d_status, float* d_err, const KltParams& kp);
