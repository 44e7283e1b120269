"""Batched per-frame hot path over independent sequences (BASELINE config 5).

``SequenceBatch`` wraps ``b200vo_batch_*`` (include/b200vo.h): per step and per sequence it
does what the reference's ``continuous_operation`` does on its hot path -- KLT on the landmark
keypoints and on the candidate keypoints (``VisualOdometryPipeLine.py:281,:287``), keep
``status == 1`` (``:282-284``), P3P-RANSAC + EPnP on the tracked landmarks (``:343``) -- for
``batch`` sequences in one set of kernel launches, with the previous frames' pyramids resident
in HBM.  torch is used only as a device-buffer allocator for the ``*_dev`` form.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BatchCfg, c_f32p, c_f64p, c_i32p, c_u8p


def _p(a, t):
    return a.ctypes.data_as(t)


class _PinnedBlock:
    """Owner of one b200vo_host_alloc block.  The numpy arrays carved from it keep it alive (it hangs off the
    ctypes buffer that is the arrays' base), so a result array that outlives its SequenceBatch -- e.g.
    ``SequenceBatch(...).step(...)['pose']`` -- never points into freed page-locked memory: the block is
    released when the LAST array referring to it is collected."""

    def __init__(self, ctx: _lib.Context, nbytes: int):
        self.ctx, self.nbytes = ctx, nbytes
        self.ptr = ctx.lib.b200vo_host_alloc(ctx.h, nbytes)
        if not self.ptr:
            raise MemoryError("b200vo_host_alloc")

    def array(self, shape, dtype) -> np.ndarray:
        dt = np.dtype(dtype)
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        buf._owner = self          # ndarray.base -> memoryview -> buf -> this block
        return np.frombuffer(buf, dt, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if self.ptr:
                # the context may already be closed: page-locked memory is freed all the same (ctx == NULL)
                self.ctx.lib.b200vo_host_free(self.ctx.h if getattr(self.ctx, "h", None) else None, self.ptr)
                self.ptr = None
        except Exception:
            pass


class SequenceBatch:
    def __init__(self, batch, rows, cols, K, win=(15, 15), max_level=5, criteria=(3, 50, 0.01),
                 min_eig_thr=1e-4, pnp_iters=500, pnp_reproj_err=8.0, pnp_conf=0.99,
                 max_landmarks=1024, max_candidates=1024, ctx: _lib.Context | None = None):
        self.ctx = ctx or _lib.default_context(0)
        self.batch, self.rows, self.cols = int(batch), int(rows), int(cols)
        self.L, self.Cn = int(max_landmarks), int(max_candidates)
        cfg = BatchCfg()
        cfg.rows, cfg.cols = self.rows, self.cols
        cfg.win_w, cfg.win_h, cfg.max_level = int(win[0]), int(win[1]), int(max_level)
        cfg.crit_type, cfg.crit_max_count, cfg.crit_eps = int(criteria[0]), int(criteria[1]), float(criteria[2])
        cfg.min_eig_thr = float(min_eig_thr)
        cfg.pnp_iters, cfg.pnp_reproj_err, cfg.pnp_conf = int(pnp_iters), float(pnp_reproj_err), float(pnp_conf)
        Kf = np.asarray(K, np.float64).reshape(9)
        for i in range(9):
            cfg.K[i] = float(Kf[i])
        cfg.max_landmarks, cfg.max_candidates = self.L, self.Cn
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.ctx.lib.b200vo_batch_create(self.ctx.h, self.batch, C.byref(cfg), C.byref(h))
        if rc != 0:
            raise _lib.B200VOError(f"b200vo_batch_create failed ({rc}): {self.ctx.last_error()}")
        self.h = h
        # host outputs (reused, page-locked so that step() reads them back without a staging copy)
        b, L, Cn = self.batch, self.L, self.Cn
        self.lm_next = self.pinned_empty((b, L, 2), np.float32)
        self.lm_status = self.pinned_empty((b, L), np.uint8)
        self.cand_next = self.pinned_empty((b, max(Cn, 1), 2), np.float32)
        self.cand_status = self.pinned_empty((b, max(Cn, 1)), np.uint8)
        self.pose = self.pinned_empty((b, 6), np.float64)
        self.pnp_ok = self.pinned_empty((b,), np.uint8)
        self.inlier_mask = self.pinned_empty((b, L), np.uint8)
        self.n_inliers = self.pinned_empty((b,), np.int32)
        # ctypes views of the reused result arrays, built once (step() is called thousands of times per second)
        self._out_ptrs = (_p(self.lm_next, c_f32p), _p(self.lm_status, c_u8p), _p(self.cand_next, c_f32p), _p(self.cand_status, c_u8p),
                          _p(self.pose, c_f64p), _p(self.pnp_ok, c_u8p), _p(self.inlier_mask, c_u8p), _p(self.n_inliers, c_i32p))
        self._out_dict = dict(lm_next=self.lm_next, lm_status=self.lm_status, cand_next=self.cand_next, cand_status=self.cand_status,
                              pose=self.pose, pnp_ok=self.pnp_ok, inlier_mask=self.inlier_mask, n_inliers=self.n_inliers)

    def close(self):
        """Destroys the device-side batch.  Page-locked arrays handed out by this object (step() results,
        pinned_empty / pinned_like / pinned_frames) stay valid for as long as they are referenced: each block is
        owned by its arrays (_PinnedBlock) and released with the last of them."""
        if getattr(self, "h", None):
            self.ctx.lib.b200vo_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc != 0:
            raise _lib.B200VOError(f"{what} failed ({rc}): {self.ctx.last_error()}")

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array in page-locked host memory (b200vo_host_alloc): DMA'd in place by step()."""
        dt = np.dtype(dtype)
        nbytes = max(int(np.prod(shape)) * dt.itemsize, 1)
        return _PinnedBlock(self.ctx, nbytes).array(shape, dt)

    def pinned_like(self, arr: np.ndarray) -> np.ndarray:
        out = self.pinned_empty(arr.shape, arr.dtype)
        out[...] = arr
        return out

    def pinned_frames(self, n_sets: int = 1) -> np.ndarray:
        """(n_sets, batch, rows, cols) uint8 array in page-locked memory (b200vo_host_alloc)."""
        return self.pinned_empty((n_sets, self.batch, self.rows, self.cols), np.uint8)

    def prime(self, frames: np.ndarray):
        frames = np.ascontiguousarray(frames, np.uint8).reshape(self.batch, self.rows, self.cols)
        self._chk(self.ctx.lib.b200vo_batch_prime(self.h, _p(frames, c_u8p)), "b200vo_batch_prime")

    def submit_frames(self, frames: np.ndarray):
        """Hand over the frames of a FUTURE step (page-locked array from pinned_frames()): their upload and
        pyramid build overlap the step in flight; the step called with frames=None consumes them."""
        assert frames.dtype == np.uint8 and frames.flags.c_contiguous and frames.size == self.batch * self.rows * self.cols
        self._chk(self.ctx.lib.b200vo_batch_submit_frames(self.h, _p(frames, c_u8p)), "b200vo_batch_submit_frames")

    def submit_frames_dev(self, frames_dev: int):
        """The same look-ahead for frames already resident in device memory (int from tensor.data_ptr()): the
        pyramids are built beside the step in flight; step_dev(frames_dev=None, ...) consumes them."""
        self._chk(self.ctx.lib.b200vo_batch_submit_frames_dev(self.h, frames_dev), "b200vo_batch_submit_frames_dev")

    def step(self, frames, lm_pts, lm_obj, n_lm, cand_pts=None, n_cand=None):
        """Host buffers in, host buffers out (synchronous).  Returns a dict of views on REUSED page-locked arrays:
        the next step() overwrites them (copy what must survive it); they remain valid memory after close().
        frames=None: use the oldest frame set given to submit_frames()."""
        b, L, Cn = self.batch, self.L, self.Cn
        assert frames is None or (frames.dtype == np.uint8 and frames.flags.c_contiguous and frames.size == b * self.rows * self.cols)
        assert lm_pts.dtype == np.float32 and lm_pts.shape == (b, L, 2) and lm_pts.flags.c_contiguous
        assert lm_obj.dtype == np.float32 and lm_obj.shape == (b, L, 3) and lm_obj.flags.c_contiguous
        assert n_lm.dtype == np.int32 and n_lm.shape == (b,)
        if Cn > 0:
            assert cand_pts is not None and cand_pts.dtype == np.float32 and cand_pts.shape == (b, Cn, 2)
            assert n_cand is not None and n_cand.dtype == np.int32 and n_cand.shape == (b,)
        rc = self.ctx.lib.b200vo_batch_step(
            self.h, _p(frames, c_u8p) if frames is not None else None, _p(lm_pts, c_f32p), _p(lm_obj, c_f32p), _p(n_lm, c_i32p),
            _p(cand_pts, c_f32p) if Cn > 0 else None, _p(n_cand, c_i32p) if Cn > 0 else None, *self._out_ptrs)
        if rc != 0:
            # the counts are validated once, inside the call (a single sequence steps thousands of times per second)
            msg = self.ctx.last_error()
            if "is outside [0," in msg:
                raise ValueError(msg)
            self._chk(rc, "b200vo_batch_step")
        return self._out_dict

    def good_features(self, max_corners=1400, quality=0.1, min_dist=10.0):
        """cv2.goodFeaturesToTrack (reference :256) on the current frame of every sequence (the frames the last
        prime()/step() left resident).  -> (corners float32 (batch, max_corners, 2), n int32 (batch,)); row s holds
        n[s] corners in cv2's order."""
        corners = np.zeros((self.batch, int(max_corners), 2), np.float32)
        n = np.zeros(self.batch, np.int32)
        self._chk(self.ctx.lib.b200vo_batch_good_features(self.h, int(max_corners), float(quality), float(min_dist),
                                                          _p(corners, c_f32p), _p(n, c_i32p)), "b200vo_batch_good_features")
        return corners, n

    def step_dev(self, frames_dev, lm_pts_dev, lm_obj_dev, n_lm_dev, cand_pts_dev, n_cand_dev, out_dev: dict):
        """Device pointers in/out (ints from tensor.data_ptr()); asynchronous on the ctx stream.
        frames_dev=None: use the oldest frame set given to submit_frames_dev() / submit_frames()."""
        o = out_dev
        rc = self.ctx.lib.b200vo_batch_step_dev(
            self.h, frames_dev or None, lm_pts_dev, lm_obj_dev, n_lm_dev, cand_pts_dev, n_cand_dev,
            o["lm_next"], o["lm_status"], o["cand_next"], o["cand_status"], o["pose"], o["pnp_ok"],
            o["inlier_mask"], o["n_inliers"])
        self._chk(rc, "b200vo_batch_step_dev")
