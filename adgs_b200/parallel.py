"""Multi-GPU data path: the views of a multi-camera / multi-timestep batch are sharded across
ranks (one process per GPU), every rank renders its own views against a full parameter replica,
and the parameter gradients are summed with ONE all-reduce over a flat gradient buffer
(NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests).

The reference has no distributed code at all: it renders one view per
iteration (train.py:55-61). The oracle for this module is therefore "sum over views of the
single-GPU gradients" (SURVEY.md section 8e), plus the three non-linear densification statistics
that a plain gradient sum does not cover (gaussian_model.py:863-867, train.py:151):
    xyz_gradient_accum += || grad_means2D[:, :2] ||   per view   -> all-reduce SUM of per-view norms
    denom              += visible                      per view   -> all-reduce SUM
    max_radii2D         = max(max_radii2D, radii)      per view   -> all-reduce MAX
"""
from typing import Dict, List, Sequence

import os

import torch
import torch.distributed as dist


def shard_views(views: Sequence, rank: int, world_size: int) -> List:
    """Rank r renders views r, r+G, r+2G, ... (round-robin keeps cameras of one timestep apart)."""
    return list(views[rank::world_size])


class FlatGradBucket:
    """All hot-path parameters' gradients as views of one contiguous fp32 buffer, so the
    cross-rank reduction is a single collective launch (no per-tensor latency, full NVLink
    message size)."""

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.names = list(params)
        self.shapes = {k: tuple(v.shape) for k, v in params.items()}
        self.offsets, off = {}, 0
        for k, v in params.items():
            self.offsets[k] = off
            off += (v.numel() + 3) // 4 * 4      # keep every view 16-byte aligned
        self.numel = off
        any_p = next(iter(params.values()))
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=any_p.device)

    def views(self) -> Dict[str, torch.Tensor]:
        out = {}
        for k in self.names:
            n = 1
            for s in self.shapes[k]:
                n *= s
            out[k] = self.flat[self.offsets[k]: self.offsets[k] + n].view(self.shapes[k])
        return out

    def zero_(self):
        self.flat.zero_()

    def accumulate(self, grads: Dict[str, torch.Tensor]):
        v = self.views()
        for k, g in grads.items():
            if g is not None:
                v[k].add_(g)

    def all_reduce(self, group=None, async_op=False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None


class DensifyStats:
    """Per-view densification statistics and their cross-rank reductions."""

    def __init__(self, n: int, device):
        self.grad_norm_sum = torch.zeros(n, 1, device=device)
        self.visible_count = torch.zeros(n, 1, device=device)
        self.max_radii = torch.zeros(n, device=device)

    def add_view(self, viewspace_grad: torch.Tensor, visibility_filter: torch.Tensor, radii: torch.Tensor):
        if self.grad_norm_sum.is_cuda:
            # one launch (adgs_densify_stats; visibility_filter is radii > 0, gaussian_renderer/__init__.py:104):
            # 0.015 ms at 1 M Gaussians instead of 2.7 ms for the four indexed torch ops below
            from . import _lib as L
            dev = self.grad_norm_sum.device
            g = viewspace_grad.contiguous()
            r = radii.to(torch.int32).contiguous()
            with torch.cuda.device(dev):
                st = L.load().adgs_densify_stats(r.shape[0], L.ptr(g), L.ptr(r), L.ptr(self.grad_norm_sum),
                                                 L.ptr(self.visible_count), L.ptr(self.max_radii),
                                                 torch.cuda.current_stream(dev).cuda_stream)
            L.check(st, "densify_stats")
            return
        # host tensors (the gloo tests of the reduction logic)
        vis = visibility_filter
        self.grad_norm_sum[vis] += torch.norm(viewspace_grad[vis, :2], dim=-1, keepdim=True)
        self.visible_count[vis] += 1
        self.max_radii[vis] = torch.max(self.max_radii[vis], radii[vis].to(self.max_radii.dtype))

    def all_reduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad_norm_sum, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.visible_count, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=group)


class MultiViewStep:
    """One multi-view iteration: render + backward the local shard of `views`, then reduce.

    `render_fn(view)` must return the result dict of gaussian_renderer.render();
    `cotangent_fn(view, result)` returns (outputs, cotangents) for torch.autograd.backward
    (in training: the loss; in the benchmark: fixed random cotangents)."""

    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        from .gaussian_model import PARAM_NAMES
        self.names = PARAM_NAMES
        self._sized_for = None
        self._resize()

    def _resize(self):
        """(Re)build the flat bucket for the model's current arrays (densification changes their sizes)."""
        shapes = tuple(tuple(getattr(self.model, k).shape) for k in self.names)
        if shapes != self._sized_for:
            self.bucket = FlatGradBucket({k: getattr(self.model, k) for k in self.names})
            self._sized_for = shapes

    def run(self, views, render_fn, cotangent_fn, reduce_stats=True):
        """Returns the flat bucket's gradient views. `self.stats` afterwards holds THIS run's densification
        statistics summed / maxed over all ranks' views (fresh buffers every run: all-reducing a buffer that already
        holds earlier, already-global sums would count them world_size times); the caller folds them into its
        running accumulators (model.xyz_gradient_accum / denom / max_radii2D, scene/gaussian_model.py:863-867)."""
        self._resize()
        self.stats = DensifyStats(self.model.get_pts_num, self.model.xyz.device)
        local = shard_views(views, self.rank, self.world)
        params = {k: getattr(self.model, k) for k in self.names}
        first = True
        for view in local:
            for p in params.values():
                p.grad = None
            # first local view: the backward writes straight into the flat bucket (no copy)
            self.model._grad_sink = self.bucket.views if first else None
            try:
                res = render_fn(view)
                outs, cots = cotangent_fn(view, res)
                torch.autograd.backward(outs, cots)
            finally:
                self.model._grad_sink = None
            grads = {k: p.grad for k, p in params.items()}
            v = self.bucket.views()
            if first:
                for k, g in grads.items():
                    if g is None:
                        v[k].zero_()
                    elif g.data_ptr() != v[k].data_ptr():   # autograd cloned it (should not happen)
                        v[k].copy_(g)
                first = False
            else:
                self.bucket.accumulate(grads)
            if reduce_stats:
                self.stats.add_view(res["viewspace_points"].grad, res["visibility_filter"], res["radii"])
        if first:
            self.bucket.zero_()
        self.bucket.all_reduce(self.group)
        if reduce_stats:
            self.stats.all_reduce(self.group)
        v = self.bucket.views()
        for k, p in params.items():
            p.grad = v[k]
        return v


class _RawCudaArray:
    """Zero-copy torch view of raw device memory (torch.as_tensor reads __cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerBuffers:
    """This rank's exchange buffers, mapped into every rank of the box (adgs_peer_alloc / adgs_peer_open: CUDA IPC
    over NVLink / NVSwitch), and the flag barrier over them. One allocation per rank:

        flags  (8) u32        barrier flags, one word per peer
        rec    (G, n, 16) f32 slot r = rank r's 64-byte blend records of THIS rank's view
        meta   (5, G, n) i32  plane-major (depth key, tiles touched, radius, mean x, mean y), rank-major inside a
                              plane: exactly the concatenated arrays adgs_splats_bin reads, no permute
        grec   (G, n, 16) f32 gradient records of this rank's view, written by its own blend backward
    """

    def __init__(self, L, lib, group, world, rank, n, device):
        import ctypes as C
        self.L, self.lib, self.world, self.rank, self.n, self.device = L, lib, world, rank, n, device
        P = world * n
        self.off_bg = 512                       # this rank's partial of the background-trajectory gradient (<= 3.5 KB)
        self.off_rec = 4096
        self.off_meta = self.off_rec + P * 64
        self.off_grec = self.off_meta + 5 * P * 4
        self.bytes = self.off_grec + P * 64
        with torch.cuda.device(device):
            # every step below is collective: a rank that cannot allocate / map still takes part, and ALL ranks then
            # raise together (SplatExchangeStep falls back to the NCCL exchange) instead of one of them hanging the rest
            ptr = C.c_void_p()
            handle = C.create_string_buffer(64)
            ok = lib.adgs_peer_alloc(self.bytes, C.byref(ptr), handle) == 0
            self.local = int(ptr.value) if ok else 0
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle.raw) if ok else None, group=group)
            self.base = []
            ok = all(h is not None for h in handles)
            if ok:
                for p in range(world):
                    if p == rank:
                        self.base.append(self.local)
                        continue
                    q = C.c_void_p()
                    if lib.adgs_peer_open(handles[p], C.byref(q)) != 0:
                        ok = False
                        break
                    self.base.append(int(q.value))
            flag = torch.tensor([1 if ok else 0], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                for p, b in enumerate(self.base):
                    if p != rank:
                        lib.adgs_peer_close(b)
                if self.local:
                    lib.adgs_peer_free(self.local)
                raise RuntimeError("peer memory (CUDA IPC) is not available between the ranks: "
                                   + lib.adgs_last_cuda_error().decode())
        self.flag_arr = (C.c_void_p * world)(*self.base)        # the flag words sit at offset 0
        self.status_ptr = self.local + 256
        self.epoch = 0

    def rec(self, p, slot=0):
        return self.base[p] + self.off_rec + slot * self.n * 64

    def meta(self, p, plane, slot=0):
        return self.base[p] + self.off_meta + ((plane * self.world + slot) * self.n) * 4

    def grec(self, p, slot=0):
        return self.base[p] + self.off_grec + slot * self.n * 64

    def bg(self, p):
        return self.base[p] + self.off_bg

    def barrier(self, stream):
        self.epoch += 1
        self.L.check(self.lib.adgs_peer_barrier(self.world, self.rank, self.flag_arr, self.epoch, self.status_ptr,
                                                stream), "peer_barrier")

    def local_radii(self):
        if "_local_radii" not in self.__dict__:
            self._local_radii = torch.as_tensor(_RawCudaArray(self.meta(self.rank, 2), (self.world * self.n,), "<i4"),
                                                device=self.device)
        return self._local_radii

    def timed_out(self):
        return bool(torch.as_tensor(_RawCudaArray(self.status_ptr, (1,), "<u4"), device=self.device).item())

    def close(self):
        for p, b in enumerate(self.base):
            if p != self.rank:
                self.lib.adgs_peer_close(b)
        self.lib.adgs_peer_free(self.local)
        self.base = []


# =====================================================================================================
# Splat exchange: Gaussians sharded across ranks, every view blended on one rank.
# =====================================================================================================
class SplatExchangeStep:
    """Multi-view iteration WITHOUT a parameter-gradient all-reduce.

    The all-reduce of `MultiViewStep` moves every parameter gradient (~640 B per Gaussian) through
    NVLink once per iteration, which costs as much as the whole single-GPU step. Here the model is
    sharded by Gaussian instead (rank s owns N/G Gaussians and their optimizer state) and only the
    per-view *splats* travel: each rank runs the trajectory + projection front end of ITS Gaussians for
    ALL G views of the round (`adgs_shard_forward`), an all-to-all hands view v's splats (76 B per
    Gaussian: 64-byte blend record + depth key + tile count + radius) to rank v, which bins and blends
    that one view (`adgs_splats_forward`); the backward mirrors it: blend backward on rank v
    (`adgs_splats_backward`), all-to-all of the 64-byte gradient records back to the owners, per-Gaussian
    backward per view accumulated into the local gradient shard (`adgs_shard_backward`). Per rank and
    round that is ~140 B per Gaussian over NVLink instead of ~1.3 KB, no gradient all-reduce at all
    (only the 3xC_bg background-trajectory gradient, which every Gaussian shares, is all-reduced), and
    the same arithmetic per Gaussian and per pixel as the single-GPU path.

    `views`: G*m entries (camera, flow_time); in round i rank r blends views[i*G + r].
    `cotangent_fn(view, images) -> dict(color, depth, opacity, flow, semantic)` (None = zero).
    """

    def __init__(self, shard, group=None, render_objmask=True, exchange=None):
        """exchange: "peer" (default when world > 1) = splats and gradient records travel through peer memory:
        the front end stores straight into the blending rank's buffers over NVLink, the per-Gaussian backward loads
        its gradient records from the blending ranks, two flag barriers per step and no collective on the data
        path; "nccl" = two all_to_all_single per step (also what views_per_rank > 1 uses)."""
        from . import _lib as L
        self.L = L
        self.lib = L.load()
        self.model = shard
        self.group = group
        init = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if init else 0
        self.world = dist.get_world_size(group) if init else 1
        self.render_objmask = bool(render_objmask)
        from .gaussian_model import PARAM_NAMES
        self.names = PARAM_NAMES
        self.grads = {k: torch.zeros_like(getattr(shard, k)) for k in PARAM_NAMES}
        self._capacity = 0
        self.exchange = exchange or os.environ.get("ADGS_EXCHANGE", "peer")
        assert self.exchange in ("peer", "nccl")
        self._peer = None
        self._checks = []   # deferred {num_rendered, overflow} read-backs of sync-free steps

    # -- helpers -------------------------------------------------------------------------------------
    def _camera(self, cam, pipe, keep):
        import math
        from .rasterizer import GaussianRasterizationSettings, _camera
        dev = self.model.xyz.device
        s = GaussianRasterizationSettings(
            image_height=int(cam.image_height), image_width=int(cam.image_width),
            tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
            bg=self._zero_bg(dev), scale_modifier=1.0, viewmatrix=cam.world_view_transform.to(dev),
            projmatrix=cam.full_proj_transform.to(dev), sh_degree=self.model.active_sh_degree,
            campos=cam.camera_center.to(dev), prefiltered=False, inv_depth=getattr(pipe, "inv_depth", False),
            debug=getattr(pipe, "debug", False))
        return _camera(s, keep)

    def _zero_bg(self, dev):
        if "_bg" not in self.__dict__:
            self._bg = torch.zeros(3, device=dev)
        return self._bg

    def _all_to_all(self, send):
        """send: (G, ...) -> recv: (G, ...), chunk g goes to rank g."""
        if self.world == 1:
            return send
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv

    def run(self, views, cotangent_fn, pipe, views_per_rank=1):
        """One iteration over `views` (a multiple of world_size * views_per_rank entries). In every round
        rank r blends views [r*k, (r+1)*k) of the round's world_size*k views (k = views_per_rank <= 8/G)."""
        import ctypes as C
        L, lib, m = self.L, self.lib, self.model
        G, k, n, dev = self.world, int(views_per_rank), m.get_pts_num, m.xyz.device
        V = G * k
        assert 1 <= V <= 8, "at most 8 views per round (ADGS_MAX_VIEWS)"
        assert len(views) % V == 0, "the batch must hold a multiple of world_size * views_per_rank views"
        stream = torch.cuda.current_stream(dev).cuda_stream
        results, stats = [], []
        first = True
        timing = self.__dict__.setdefault("_timing", None)
        if timing is None and os.environ.get("ADGS_EXCHANGE_TIMING"):
            timing = self.__dict__["_timing"] = {"marks": [], "sums": {}, "count": 0}

        def mark(name):
            if timing is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(torch.cuda.current_stream(dev))
                timing["marks"].append((name, ev))

        self._resolve_checks()
        if G > 1 and k == 1 and self.exchange == "peer":
            return self._run_peer(views, cotangent_fn, pipe)
        mark("start")
        o = dict(dtype=torch.float32, device=dev)
        D_S = 1 if self.render_objmask else 0
        for rnd in range(len(views) // V):
            batch = views[rnd * V:(rnd + 1) * V]
            keep = []
            with torch.cuda.device(dev):
                # ---- front end of my shard for every view of the round: ONE launch ---------------------
                rec = torch.empty((V, n, 16), **o)
                # small per-splat state in one buffer: planes 0 depth key, 1 tiles touched, 2 radius,
                # 3 pixel mean x, 4 pixel mean y (20 B per splat: all that binning needs)
                meta = torch.empty((V, 5, n), dtype=torch.int32, device=dev)
                state = torch.empty((V, lib.adgs_shard_state_bytes(n)), dtype=torch.uint8, device=dev)
                cams = [self._camera(cam, pipe, keep) for cam, _ in batch]
                bases = [m.time_basis(cam.time, flow_t) for cam, flow_t in batch]
                cam_arr = (L.Camera * V)(*cams)
                basis_arr = (L.TimeBasis * V)(*bases)
                splat_arr = (L.Splats * V)(*[L.Splats(P=n, _pad=0, record=rec[v].data_ptr(),
                                                      depth_keys=meta[v, 0].data_ptr(),
                                                      tiles_touched=meta[v, 1].data_ptr(), radii=meta[v, 2].data_ptr(),
                                                      mean_x=meta[v, 3].data_ptr(), mean_y=meta[v, 4].data_ptr())
                                             for v in range(V)])
                state_arr = (C.c_void_p * V)(*[state[v].data_ptr() for v in range(V)])
                cmodel = m.c_model()
                L.check(lib.adgs_shard_forward_multi(V, cam_arr, C.byref(cmodel), basis_arr, int(self.render_objmask),
                                                     splat_arr, state_arr, stream), "shard_forward_multi")
                radii = meta[:, 2]
                zeroed = self._zero_planes_async(dev) if first else False
                mark("shard_forward")
                # ---- splats travel to the rank that blends their view (chunk d of dim 0 -> rank d): the
                #      20-byte binning state first, the 64-byte records behind it on the NCCL stream, so
                #      that sorting and binning overlap the bulk of the transfer ---------------------------
                if G > 1:
                    r_meta = torch.empty_like(meta)
                    r_rec = torch.empty_like(rec)
                    w_meta = dist.all_to_all_single(r_meta, meta, group=self.group, async_op=True)
                    w_rec = dist.all_to_all_single(r_rec, rec, group=self.group, async_op=True)
                    w_meta.wait()
                    mark("meta_all_to_all")
                else:
                    r_meta, r_rec, w_rec = meta, rec, None
                # (G, k, 5, n) -> per local view and plane, rank-major over the shards
                r_meta = r_meta.view(G, k, 5, n).permute(1, 2, 0, 3).contiguous()      # (k, 5, G, n)
                r_rec = r_rec.view(G, k, n, 16)
                gback = torch.empty((G, k, n, 16), **o)
                P = G * n
                for i in range(k):
                    v_glob = self.rank * k + i
                    cam, flow_t = batch[v_glob]
                    cc = cams[v_glob]
                    H, W = int(cam.image_height), int(cam.image_width)
                    # all shards of my i-th view, rank-major
                    s_keys, s_tiles, s_radii = r_meta[i, 0].view(-1), r_meta[i, 1].view(-1), r_meta[i, 2].view(-1)
                    s_mx, s_my = r_meta[i, 3].view(-1), r_meta[i, 4].view(-1)
                    img = dict(color=torch.empty((3, H, W), **o), depth=torch.empty((1, H, W), **o),
                               opacity=torch.empty((1, H, W), **o), flow=torch.empty((3, H, W), **o),
                               semantic=torch.empty((D_S, H, W), **o))
                    images = L.Images(color=L.ptr(img["color"]), depth=L.ptr(img["depth"]),
                                      opacity=L.ptr(img["opacity"]), flow=L.ptr(img["flow"]),
                                      semantic=L.ptr(img["semantic"]), radii=None)
                    splats = L.Splats(P=P, _pad=0, record=None, depth_keys=s_keys.data_ptr(),
                                      tiles_touched=s_tiles.data_ptr(), radii=s_radii.data_ptr(),
                                      mean_x=s_mx.data_ptr(), mean_y=s_my.data_ptr())
                    geom = torch.empty((lib.adgs_geometry_bytes(P),), dtype=torch.uint8, device=dev)
                    imgbuf = torch.empty((lib.adgs_image_bytes(W, H),), dtype=torch.uint8, device=dev)
                    has_flow = int(flow_t is not None)
                    sync_free = bool(getattr(pipe, "sync_free", True)) and self._capacity > 0
                    if sync_free:
                        capacity = self._capacity
                        binning = torch.empty((lib.adgs_binning_bytes(capacity),), dtype=torch.uint8, device=dev)
                        L.check(lib.adgs_splats_bin(C.byref(cc), C.byref(splats), L.ptr(geom), L.ptr(binning), capacity,
                                                    L.ALLOC_FN(), None, L.ptr(imgbuf), stream), "splats_bin")
                        counters = m._pinned_counters()
                        L.check(lib.adgs_read_counters(L.ptr(geom), P, counters.data_ptr(), stream), "read_counters")
                        ev = torch.cuda.Event()
                        ev.record(torch.cuda.current_stream(dev))
                    else:
                        holder = {}

                        def _alloc(nbytes, _u, holder=holder):
                            holder["b"] = torch.empty((int(nbytes),), dtype=torch.uint8, device=dev)
                            return holder["b"].data_ptr()

                        cb = L.ALLOC_FN(_alloc)
                        R = L.check(lib.adgs_splats_bin(C.byref(cc), C.byref(splats), L.ptr(geom), None, 0, cb, None,
                                                        L.ptr(imgbuf), stream), "splats_bin")
                        binning, capacity = holder["b"], int(R)
                        self._capacity = max(self._capacity, int(1.3 * R) + 65536)
                        counters = ev = None
                    mark("binning")
                    # ---- the records have (by now, mostly) arrived: blend ---------------------------------
                    if w_rec is not None:
                        w_rec.wait()
                        w_rec = None
                    mark("record_all_to_all_wait")
                    s_rec = r_rec[:, i].reshape(P, 16)          # a plain view when k == 1
                    splats.record = s_rec.data_ptr()
                    L.check(lib.adgs_splats_blend(C.byref(cc), C.byref(splats), D_S, has_flow, C.byref(images),
                                                  L.ptr(geom), L.ptr(binning), int(capacity), L.ptr(imgbuf), stream),
                            "splats_blend")
                    res = {"render": img["color"], "depth": img["depth"][0], "img_opacity": img["opacity"][0],
                           "img_flow": img["flow"] if has_flow else None,
                           "img_semantic": img["semantic"] if self.render_objmask else None, "radii": s_radii}
                    results.append(res)
                    mark("blend_forward")
                    # ---- blend backward of my view -> gradient records for every shard ------------------
                    cot = cotangent_fn((cam, flow_t), res)
                    ct = {kk: (None if cot.get(kk) is None else cot[kk].contiguous()) for kk in
                          ("color", "depth", "opacity", "flow", "semantic")}
                    ig = L.ImageGrads(dL_dcolor=L.ptr(ct["color"]), dL_ddepth=L.ptr(ct["depth"]),
                                      dL_dflow=L.ptr(ct["flow"]), dL_dsemantic=L.ptr(ct["semantic"]),
                                      dL_dopacity=L.ptr(ct["opacity"]))
                    if ev is not None:
                        # no host block inside the step: the blend kernels skip an overflowed frame on the device
                        # (its gradient records stay zero); the counters are looked at before the next step
                        self._checks.append((counters, ev, capacity))
                    grec = gback[:, 0].view(P, 16) if k == 1 else torch.empty((P, 16), **o)
                    L.check(lib.adgs_splats_backward(C.byref(cc), C.byref(splats), D_S, has_flow, L.ptr(binning),
                                                     int(capacity), L.ptr(imgbuf), L.ptr(img["opacity"]), C.byref(ig),
                                                     grec.data_ptr(), stream), "splats_backward")
                    if k > 1:
                        gback[:, i] = grec.view(G, n, 16)
                    mark("cotangents+blend_backward")
                # ---- gradient records back to the owners; per-Gaussian backward of my shard: ONE launch ----
                r_grec = self._all_to_all(gback.view(V, n, 16))
                mark("gradient_all_to_all")
                del r_meta
                scratch = torch.empty((lib.adgs_shard_scratch_bytes(V, m.n_obj),), dtype=torch.uint8, device=dev)
                gm = m.c_model_from(self.grads, with_time=False)
                d2 = torch.empty((V, n, 3), **o)
                radii_arr = (C.c_void_p * V)(*[radii[v].data_ptr() for v in range(V)])
                grec_arr = (C.c_void_p * V)(*[r_grec[v].data_ptr() for v in range(V)])
                d2_arr = (C.c_void_p * V)(*[d2[v].data_ptr() for v in range(V)])
                if zeroed:
                    torch.cuda.current_stream(dev).wait_stream(self._fill_stream)
                L.check(lib.adgs_shard_backward_multi(V, cam_arr, C.byref(cmodel), basis_arr, radii_arr, state_arr,
                                                      grec_arr, C.byref(gm), int(not first) | (2 if zeroed else 0),
                                                      d2_arr, L.ptr(scratch), stream), "shard_backward_multi")
                first, zeroed = False, False
                mark("shard_backward")
                for v in range(V):
                    stats.append((d2[v], radii[v]))
        # the background trajectory is shared by every Gaussian: its gradient sums over the shards
        if self.world > 1 and self.grads["background_deform"].numel():
            dist.all_reduce(self.grads["background_deform"], op=dist.ReduceOp.SUM, group=self.group)
        for name in self.names:
            getattr(m, name).grad = self.grads[name]
        self._finish_timing(dev)
        return results, stats

    def _finish_timing(self, dev):
        timing = self.__dict__.get("_timing")
        if timing is None:
            return
        torch.cuda.synchronize(dev)
        marks = timing["marks"]
        timing["calls"] = timing.get("calls", 0) + 1
        if timing["calls"] > 10:  # skip the warm-up calls (NCCL channel set-up, allocator growth)
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                timing["sums"][n1] = timing["sums"].get(n1, 0.0) + e0.elapsed_time(e1)
            timing["count"] += 1
        timing["marks"] = []

    def _zero_planes_async(self, dev):
        """Zero-fill of the control-point gradient planes (their B-spline windows differ between views, so the
        per-Gaussian backward accumulates into them) on a side stream: pure HBM writes that run underneath the
        issue-bound blend kernels instead of in front of the per-Gaussian backward."""
        if "_fill_stream" not in self.__dict__:
            self._fill_stream = torch.cuda.Stream(device=dev)
        fs = self._fill_stream
        fs.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(fs):
            self.grads["xyz_deform"].zero_()
            self.grads["rot_deform"].zero_()
        return True

    def _resolve_checks(self):
        """Counters of earlier sync-free steps: grow the arena; an overflowed step was dropped on the device."""
        rest = self._checks
        self._checks = []
        for counters, ev, capacity in rest:
            ev.synchronize()
            R, ovf = int(counters[0]), bool(counters[1])
            self._capacity = max(self._capacity, int(1.3 * R) + 65536)
            if ovf or R > capacity:
                import warnings
                warnings.warn("adgs_b200: the binning arena overflowed in a sync-free splat-exchange step; that "
                              "view's images are invalid and its gradient records were left zero on the device; the "
                              "arena has been enlarged")

    def _bin_view(self, cc, splats, P, W, H, pipe, dev, stream):
        """Depth sort, scan, instance emission, tile sort, ranges of one view; exact (host reads num_rendered) the
        first time, sync-free (arena from the running bound, counters read back asynchronously) afterwards."""
        import ctypes as C
        L, lib, m = self.L, self.lib, self.model
        geom = torch.empty((lib.adgs_geometry_bytes(P),), dtype=torch.uint8, device=dev)
        imgbuf = torch.empty((lib.adgs_image_bytes(W, H),), dtype=torch.uint8, device=dev)
        if bool(getattr(pipe, "sync_free", True)) and self._capacity > 0:
            capacity = self._capacity
            binning = torch.empty((lib.adgs_binning_bytes(capacity),), dtype=torch.uint8, device=dev)
            L.check(lib.adgs_splats_bin(C.byref(cc), C.byref(splats), L.ptr(geom), L.ptr(binning), capacity,
                                        L.ALLOC_FN(), None, L.ptr(imgbuf), stream), "splats_bin")
            counters = m._pinned_counters()
            L.check(lib.adgs_read_counters(L.ptr(geom), P, counters.data_ptr(), stream), "read_counters")
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self._checks.append((counters, ev, capacity))
        else:
            holder = {}

            def _alloc(nbytes, _u, holder=holder):
                holder["b"] = torch.empty((int(nbytes),), dtype=torch.uint8, device=dev)
                return holder["b"].data_ptr()

            cb = L.ALLOC_FN(_alloc)
            R = L.check(lib.adgs_splats_bin(C.byref(cc), C.byref(splats), L.ptr(geom), None, 0, cb, None,
                                            L.ptr(imgbuf), stream), "splats_bin")
            binning, capacity = holder["b"], int(R)
            self._capacity = max(self._capacity, int(1.3 * R) + 65536)
        return geom, imgbuf, binning, int(capacity)

    def _run_peer(self, views, cotangent_fn, pipe):
        """run() for one view per rank with the exchange through peer memory (PeerBuffers): per round
        front end of my shard for all G views, stores landing in the blending ranks -> barrier -> bin + blend my
        view forward and backward from / into my own buffers -> barrier -> per-Gaussian backward of my shard, the
        gradient records of view v loaded from rank v."""
        import ctypes as C
        L, lib, m = self.L, self.lib, self.model
        G, r, n, dev = self.world, self.rank, m.get_pts_num, m.xyz.device
        assert len(views) % G == 0 and G <= 8
        if self._peer is None or self._peer.n != n:
            if self._peer is not None:
                self._peer.close()
                self._peer = None
            try:
                self._peer = PeerBuffers(L, lib, self.group, G, r, n, dev)
            except RuntimeError as exc:       # raised on every rank together
                import warnings
                warnings.warn(f"adgs_b200: {exc}; the splat exchange falls back to NCCL all-to-all")
                self.exchange = "nccl"
                return self.run(views, cotangent_fn, pipe)
        pb = self._peer
        stream = torch.cuda.current_stream(dev).cuda_stream
        results, stats = [], []
        first = True
        o = dict(dtype=torch.float32, device=dev)
        D_S = 1 if self.render_objmask else 0
        P = G * n
        timing = self.__dict__.setdefault("_timing", None)
        if timing is None and os.environ.get("ADGS_EXCHANGE_TIMING"):
            timing = self.__dict__["_timing"] = {"marks": [], "sums": {}, "count": 0}

        def mark(name):
            if timing is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(torch.cuda.current_stream(dev))
                timing["marks"].append((name, ev))

        mark("start")
        for rnd in range(len(views) // G):
            batch = views[rnd * G:(rnd + 1) * G]
            keep = []
            with torch.cuda.device(dev):
                state = torch.empty((G, lib.adgs_shard_state_bytes(n)), dtype=torch.uint8, device=dev)
                cams = [self._camera(cam, pipe, keep) for cam, _ in batch]
                bases = [m.time_basis(cam.time, flow_t) for cam, flow_t in batch]
                cam_arr = (L.Camera * G)(*cams)
                basis_arr = (L.TimeBasis * G)(*bases)
                # view v's outputs of MY Gaussians go to slot r of rank v's buffers
                splat_arr = (L.Splats * G)(*[L.Splats(P=n, _pad=0, record=pb.rec(v, r), depth_keys=pb.meta(v, 0, r),
                                                      tiles_touched=pb.meta(v, 1, r), radii=pb.meta(v, 2, r),
                                                      mean_x=pb.meta(v, 3, r), mean_y=pb.meta(v, 4, r))
                                             for v in range(G)])
                state_arr = (C.c_void_p * G)(*[state[v].data_ptr() for v in range(G)])
                cmodel = m.c_model()
                L.check(lib.adgs_shard_forward_multi(G, cam_arr, C.byref(cmodel), basis_arr, int(self.render_objmask),
                                                     splat_arr, state_arr, stream), "shard_forward_multi")
                zeroed = self._zero_planes_async(dev) if first else False
                mark("shard_forward")
                pb.barrier(stream)          # every rank's splats of my view have landed
                mark("barrier_1")
                cam, flow_t = batch[r]
                cc = cams[r]
                H, W = int(cam.image_height), int(cam.image_width)
                splats = L.Splats(P=P, _pad=0, record=pb.rec(r), depth_keys=pb.meta(r, 0), tiles_touched=pb.meta(r, 1),
                                  radii=pb.meta(r, 2), mean_x=pb.meta(r, 3), mean_y=pb.meta(r, 4))
                geom, imgbuf, binning, capacity = self._bin_view(cc, splats, P, W, H, pipe, dev, stream)
                mark("binning")
                img = dict(color=torch.empty((3, H, W), **o), depth=torch.empty((1, H, W), **o),
                           opacity=torch.empty((1, H, W), **o), flow=torch.empty((3, H, W), **o),
                           semantic=torch.empty((D_S, H, W), **o))
                images = L.Images(color=L.ptr(img["color"]), depth=L.ptr(img["depth"]), opacity=L.ptr(img["opacity"]),
                                  flow=L.ptr(img["flow"]), semantic=L.ptr(img["semantic"]), radii=None)
                has_flow = int(flow_t is not None)
                L.check(lib.adgs_splats_blend(C.byref(cc), C.byref(splats), D_S, has_flow, C.byref(images), L.ptr(geom),
                                              L.ptr(binning), capacity, L.ptr(imgbuf), stream), "splats_blend")
                res = {"render": img["color"], "depth": img["depth"][0], "img_opacity": img["opacity"][0],
                       "img_flow": img["flow"] if has_flow else None,
                       "img_semantic": img["semantic"] if self.render_objmask else None,
                       "radii": pb.local_radii()}      # a view of the exchange buffer: valid until the next round
                results.append(res)
                mark("blend_forward")
                cot = cotangent_fn((cam, flow_t), res)
                ct = {kk: (None if cot.get(kk) is None else cot[kk].contiguous()) for kk in
                      ("color", "depth", "opacity", "flow", "semantic")}
                ig = L.ImageGrads(dL_dcolor=L.ptr(ct["color"]), dL_ddepth=L.ptr(ct["depth"]), dL_dflow=L.ptr(ct["flow"]),
                                  dL_dsemantic=L.ptr(ct["semantic"]), dL_dopacity=L.ptr(ct["opacity"]))
                L.check(lib.adgs_splats_backward(C.byref(cc), C.byref(splats), D_S, has_flow, L.ptr(binning), capacity,
                                                 L.ptr(imgbuf), L.ptr(img["opacity"]), C.byref(ig), pb.grec(r), stream),
                        "splats_backward")
                mark("cotangents+blend_backward")
                pb.barrier(stream)          # every rank's gradient records are complete
                mark("barrier_2")
                scratch = torch.empty((lib.adgs_shard_scratch_bytes(G, m.n_obj),), dtype=torch.uint8, device=dev)
                gm = m.c_model_from(self.grads, with_time=False)
                n_bg = self.grads["background_deform"].numel()
                assert n_bg * 4 <= pb.off_rec - pb.off_bg
                if n_bg:
                    gm.background_deform = pb.bg(r)     # my partial sum goes where the peers can read it
                d2 = torch.empty((G, n, 3), **o)
                radii_arr = (C.c_void_p * G)()      # null: the radii copy the forward kept in the shard state
                grec_arr = (C.c_void_p * G)(*[pb.grec(v, r) for v in range(G)])         # peer loads (cp.async ring)
                d2_arr = (C.c_void_p * G)(*[d2[v].data_ptr() for v in range(G)])
                if zeroed:
                    torch.cuda.current_stream(dev).wait_stream(self._fill_stream)
                L.check(lib.adgs_shard_backward_multi(G, cam_arr, C.byref(cmodel), basis_arr, radii_arr, state_arr,
                                                      grec_arr, C.byref(gm), int(not first) | (2 if zeroed else 0),
                                                      d2_arr, L.ptr(scratch), stream), "shard_backward_multi")
                first, zeroed = False, False
                mark("shard_backward")
                roff = lib.adgs_shard_state_radii_offset(n)
                for v in range(G):
                    # radii of my Gaussians in view v: the owner-side copy inside the (128-byte aligned) shard state
                    off = (-state[v].data_ptr()) % 128 + roff
                    stats.append((d2[v], state[v, off:off + 4 * n].view(torch.int32)))
        # the background trajectory is shared by every Gaussian: its gradient sums over the shards (3 x C_bg floats,
        # summed with peer loads in rank order after a barrier: identical bits on every rank, no NCCL latency)
        n_bg = self.grads["background_deform"].numel()
        if n_bg:
            pb.barrier(stream)
            parts = (C.c_void_p * G)(*[pb.bg(p) for p in range(G)])
            L.check(lib.adgs_peer_sum(G, parts, n_bg, self.grads["background_deform"].data_ptr(), stream), "peer_sum")
        mark("background_all_reduce")
        for name in self.names:
            getattr(m, name).grad = self.grads[name]
        self._finish_timing(dev)
        return results, stats

    def timing_report(self):
        """Mean milliseconds per phase of run() (set ADGS_EXCHANGE_TIMING=1; adds a device sync per step)."""
        t = self.__dict__.get("_timing")
        if not t or not t["count"]:
            return None
        return {k: round(v / t["count"], 4) for k, v in t["sums"].items()}
