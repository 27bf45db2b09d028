"""Host-side mirror of the reference RAFT band interface (bands/flow_raft.py) on top of the C ABI.

`RaftFlowEngine(state_dict)` ≙ `init_model(args)` (:38-48); `infer_pair(prev, curr)` ≙ the loop body of
`process_video` (:100-113): resize x`scale`, `infer(args, image1, image2)` with image1=[prev,curr], image2=[curr,prev]
(forward and backward flow in one pass), `process_flow` of both.  All arithmetic happens in libprisma_b200.so.
"""
import ctypes as C

import numpy as np

from ._lib import PrismaError, check, fptr, lib, u8ptr, c_i64_p


def consistency_masks(fwd, bwd, want_u16=False, device=0):
    """compute_fwdbwd_mask (bands/common/flow.py:28-40) and, with want_u16, encode_flow (common/encode.py:105-110) of both
    directions: fwd/bwd HxWx2 f32 -> (fwd_mask, bwd_mask) bool HxW [, fwd_u16, bwd_u16 HxWx3 uint16]."""
    fwd, bwd = np.ascontiguousarray(fwd, np.float32), np.ascontiguousarray(bwd, np.float32)
    if fwd.shape != bwd.shape or fwd.ndim != 3 or fwd.shape[2] != 2:
        raise PrismaError("expected two HxWx2 float32 flow fields of the same size")
    h, w = fwd.shape[:2]
    fm, bm = np.empty((h, w), np.uint8), np.empty((h, w), np.uint8)
    fu = np.empty((h, w, 3), np.uint16) if want_u16 else None
    bu = np.empty((h, w, 3), np.uint16) if want_u16 else None
    u16 = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_uint16))
    check(lib().prisma_flow_masks(device, fptr(fwd), fptr(bwd), h, w, u8ptr(fm), u8ptr(bm), u16(fu), u16(bu)))
    return (fm.view(bool), bm.view(bool), fu, bu) if want_u16 else (fm.view(bool), bm.view(bool))


def write_flo(path, flow):
    """Middlebury .flo (what common/io.py:175-198 writes): f32 magic 202021.25, i32 width, i32 height, f32 HxWx2."""
    flow = np.ascontiguousarray(flow, np.float32)
    with open(path, "wb") as f:
        np.array([202021.25], np.float32).tofile(f)
        np.array([flow.shape[1], flow.shape[0]], np.int32).tofile(f)
        flow.tofile(f)


class RaftFlowEngine:
    def __init__(self, state_dict=None, device=0, iterations=20, scale=0.75):
        self._h = C.c_void_p()
        self.iterations, self.scale = iterations, scale  # defaults of flow_raft.py:32,183
        check(lib().prisma_flow_create(device, C.byref(self._h)))
        if state_dict is not None:
            self.load_state_dict(state_dict)

    def load_state_dict(self, state_dict):
        l = lib()
        for name, t in state_dict.items():
            a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
            if a.dtype.kind != "f":
                continue  # num_batches_tracked
            a = np.ascontiguousarray(a, dtype=np.float32)
            shape = (C.c_int64 * max(a.ndim, 1))(*(a.shape if a.ndim else (1,)))
            check(l.prisma_flow_load_tensor(self._h, name.encode(), fptr(a), C.cast(shape, c_i64_p), max(a.ndim, 1)))
        check(l.prisma_flow_finalize(self._h))

    def out_size(self, h, w):
        """cv::resize(None, fx, fy): dsize = cvRound(src * fx) in double, half to even -- the same expression the engine
        evaluates (the C ABI takes `scale` as a double for exactly this reason)."""
        return int(np.rint(h * float(self.scale))), int(np.rint(w * float(self.scale)))

    def infer_clip(self, frames, continue_clip=False, want_flow=True, want_rgb=True, out=None):
        """A chunk of the band's video loop: frames [n,H,W,3] u8 (consecutive frames of one clip) -> dict of pair-major
        arrays: fwd, bwd [p,hs,ws,2] f32, fwd_rgb, bwd_rgb [p,hs,ws,3] u8, max_fwd, max_bwd [p].  p = n-1 for a new clip,
        n when the chunk continues the previous call's clip.  Uploads / downloads overlap the compute
        (prisma_flow_infer_stream); `frames` and the arrays in `out` (same keys) may be pinned_empty() buffers."""
        x = frames if isinstance(frames, np.ndarray) else np.stack(frames)
        if x.dtype != np.uint8 or x.ndim != 4 or x.shape[3] != 3 or not x.flags.c_contiguous:
            raise PrismaError("expected a C-contiguous [n,H,W,3] uint8 RGB array")
        n, h, w = x.shape[:3]
        hs, ws = self.out_size(h, w)
        out = out or {}
        cap = n  # upper bound of the number of pairs

        def buf(key, shape, dtype, want):
            if not want:
                return None
            a = out.get(key)
            if a is None:
                a = np.empty((cap,) + shape, dtype)
            if a.shape[0] < cap or a.shape[1:] != shape or a.dtype != dtype or not a.flags.c_contiguous:
                raise PrismaError(f"output buffer '{key}' has the wrong shape / dtype")
            return a
        fwd, bwd = buf("fwd", (hs, ws, 2), np.float32, want_flow), buf("bwd", (hs, ws, 2), np.float32, want_flow)
        frgb, brgb = buf("fwd_rgb", (hs, ws, 3), np.uint8, want_rgb), buf("bwd_rgb", (hs, ws, 3), np.uint8, want_rgb)
        mf, mb = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        pairs = C.c_int()
        check(lib().prisma_flow_infer_stream(self._h, u8ptr(x), n, h, w, float(self.scale), self.iterations, int(continue_clip),
                                             fptr(fwd), fptr(bwd), u8ptr(frgb), u8ptr(brgb), fptr(mf), fptr(mb), C.byref(pairs)))
        p = pairs.value
        cut = lambda a: None if a is None else a[:p]
        return dict(pairs=p, fwd=cut(fwd), bwd=cut(bwd), fwd_rgb=cut(frgb), bwd_rgb=cut(brgb), max_fwd=mf[:p], max_bwd=mb[:p])

    @property
    def pairs_per_pass(self):
        """Frame pairs one pass of the clip path covers at most (prisma_flow_set_pairs_per_pass; 4 by default)."""
        return check(lib().prisma_flow_pairs_per_pass(self._h))

    @property
    def plan_pairs(self):
        """Pairs per pass of the plan the last call built (the setting, lowered for very large frames; 1 after infer_pair)."""
        return check(lib().prisma_flow_plan_pairs(self._h))

    @pairs_per_pass.setter
    def pairs_per_pass(self, n):
        check(lib().prisma_flow_set_pairs_per_pass(self._h, int(n)))

    def time_resident(self, h, w, reps):
        """ms per PAIR of `reps` passes over the frames resident on the device (CUDA events inside the C ABI)."""
        ms = C.c_float()
        check(lib().prisma_flow_infer_resident(self._h, h, w, float(self.scale), self.iterations, reps, C.byref(ms)))
        return ms.value / self.plan_pairs

    def infer_pair(self, prev, curr, want_rgb=False, reuse_prev=False):
        """prev/curr: HxWx3 u8 RGB -> dict(fwd, bwd [hs,ws,2] f32, max_fwd, max_bwd[, fwd_rgb, bwd_rgb], ms).
        reuse_prev: `prev` is the `curr` of the previous call (video loop): its encoder features are reused (same results)."""
        prev, curr = np.ascontiguousarray(prev), np.ascontiguousarray(curr)
        if prev.shape != curr.shape or prev.dtype != np.uint8 or prev.ndim != 3:
            raise PrismaError("expected two HxWx3 uint8 RGB frames of the same size")
        h, w = prev.shape[:2]
        hs, ws = self.out_size(h, w)
        fwd = np.empty((hs, ws, 2), np.float32)
        bwd = np.empty((hs, ws, 2), np.float32)
        frgb = np.empty((hs, ws, 3), np.uint8) if want_rgb else None
        brgb = np.empty((hs, ws, 3), np.uint8) if want_rgb else None
        mf, mb, ms = C.c_float(), C.c_float(), C.c_float()
        check(lib().prisma_flow_infer_video(self._h, u8ptr(prev), u8ptr(curr), h, w, self.scale, self.iterations, int(reuse_prev),
                                            fptr(fwd), fptr(bwd), u8ptr(frgb), u8ptr(brgb), C.byref(mf), C.byref(mb), C.byref(ms)))
        return dict(fwd=fwd, bwd=bwd, max_fwd=mf.value, max_bwd=mb.value, fwd_rgb=frgb, bwd_rgb=brgb, ms=ms.value)

    def read_tap(self, name, shape):
        out = np.empty(int(np.prod(shape)), np.float32)
        n = check(lib().prisma_flow_read_tap(self._h, name.encode(), fptr(out), out.size))
        assert n == out.size, (name, n, out.size)
        return out.reshape(shape)

    def work(self, h, w):
        out = (C.c_double * 4)()
        check(lib().prisma_flow_work(self._h, h, w, self.scale, self.iterations, out))
        return dict(flop=out[0], launches=int(out[1]), hs=int(out[2]), ws=int(out[3]))

    def work_detail(self, h, w):
        out = (C.c_double * 8)()
        check(lib().prisma_flow_work_detail(self._h, h, w, float(self.scale), self.iterations, out))
        np_ = self.plan_pairs  # FLOP / bytes per PAIR (both directions), kernel steps per PASS of np_ pairs
        return dict(conv_flop_full=out[0] / np_, conv_flop_video=out[1] / np_, corr_flop=out[2] / np_, corr_bytes=out[3] / np_,
                    launches_full=int(out[4]), launches_video=int(out[5]), hs=int(out[6]), ws=int(out[7]), pairs_per_pass=np_)

    def profile(self, h, w):
        """ms per PAIR by kernel group of one pass (CUDA events, ungraphed): see prisma_flow_profile."""
        out = (C.c_float * 8)()
        check(lib().prisma_flow_profile(self._h, h, w, float(self.scale), self.iterations, out))
        keys = ["pre", "conv_gemm", "corr_build", "corr_lookup", "instnorm", "pointwise", "post", "total"]
        np_ = self.plan_pairs
        return dict(zip(keys, [float(v) / np_ for v in out]))

    def close(self):
        if self._h:
            lib().prisma_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
