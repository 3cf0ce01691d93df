"""Parity of the multi-GPU data path that bench.py --gpus N measures (adgs_b200.parallel.SplatExchangeStep at
world size > 1, one process per GPU, NCCL): the reference renders one view per iteration on one GPU
(train.py:55-61), so the oracle is the SEQUENTIAL PER-VIEW SUM on one GPU:
  * the image rank v blends == the single-GPU render of view v, bit for bit (same splats, same order);
  * every rank's gradient shard == its slice of the summed single-GPU gradients (<= 1e-5: only the
    order of the floating-point summation over views differs);
  * the per-view densification inputs (||grad means2D||, radii: scene/gaussian_model.py:863-867, train.py:151).
Both exchange modes are covered: peer memory (default) and the NCCL all-to-all fallback."""
import os
import socket
import sys
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rounds):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import helpers as Hh
        import test_fused_gpu as TF
        from adgs_b200.gaussian_model import GaussianModel, PARAM_NAMES
        from adgs_b200.gaussian_renderer import render
        from adgs_b200.parallel import SplatExchangeStep

        order_args, ref, c = TF._scene(3000, 1000, "kitti75")          # identical on every rank (seeded)
        model = GaussianModel.from_reference({f: getattr(ref, f) for f in ref.FIELDS}, order_args)
        cam = c["cam"]
        cot = Hh.cotangents(c)
        pipe = SimpleNamespace(inv_depth=True, debug=False, sync_free=True)

        def vcam(t):
            return SimpleNamespace(image_height=c["H"], image_width=c["W"], FoVx=cam.FoVx, FoVy=cam.FoVy,
                                   world_view_transform=cam.world_view_transform,
                                   full_proj_transform=cam.full_proj_transform, camera_center=cam.camera_center, time=t)

        V = world * rounds
        views = [(vcam(0.1 + 0.8 * i / max(V - 1, 1)), 0.12 + 0.8 * i / max(V - 1, 1)) for i in range(V)]
        render_fn = lambda v: render(v[0], model, None, pipe, flow_pkg=[v[1], None, None, None, None, None],
                                     render_objmask=True)
        cot_fn = lambda v, r: ((r["render"], r["depth"], r["img_opacity"], r["img_flow"], r["img_semantic"]),
                               (cot["color"], cot["depth"][0], cot["opacity"][0], cot["flow"], cot["semantic"]))
        # oracle: every view on this one GPU, one after the other, gradients summed over views
        want, want_imgs, want_d2, want_radii = None, [], [], []
        for v in views:
            model.zero_grad()
            res = render_fn(v)
            outs, cots = cot_fn(v, res)
            torch.autograd.backward(outs, cots)
            want_imgs.append({k: res[k].detach().clone() for k in ("render", "depth", "img_opacity", "img_flow",
                                                                     "img_semantic")})
            want_d2.append(res["viewspace_points"].grad.clone())
            want_radii.append(res["radii"].clone())
            g = {k: getattr(model, k).grad.clone() for k in PARAM_NAMES}
            want = g if want is None else {k: want[k] + g[k] for k in g}

        shard = model.shard(rank, world)
        ns, no = shard.n_scene, shard.n_obj
        full = model.to_reference()
        N_s, N_o = model.n_scene, model.n_obj

        for mode in ("peer", "nccl"):
            ex = SplatExchangeStep(shard, exchange=mode)
            for rep in range(3):       # rep 0 sizes the arena (exact path), 1 and 2 run sync-free
                results, stats = ex.run(views, lambda v, r: cot, pipe)
                torch.cuda.synchronize()
                assert len(results) == rounds and len(stats) == V
                for rnd in range(rounds):
                    wi = want_imgs[rnd * world + rank]
                    got = results[rnd]
                    for k in wi:
                        assert torch.equal(got[k], wi[k]), (mode, rep, rnd, k, Hh.rel_err(got[k], wi[k]))
                # gradient shards against slices of the summed single-GPU gradients (reference layout)
                got_ref = shard.to_reference(grads=True)
                for name in PARAM_NAMES:
                    getattr(model, name).grad = want[name]
                want_ref = model.to_reference(grads=True)
                for k, g in got_ref.items():
                    if k == "gs_time" or not g.numel():
                        continue
                    w = want_ref[k]
                    if k == "background_deform_param":
                        assert Hh.rel_err(g, w) <= 1e-5, (mode, k, Hh.rel_err(g, w))
                        continue
                    part = w[rank::world]           # the strided cut of GaussianModel.shard
                    assert Hh.rel_err(g[:part.shape[0]], part) <= 1e-5, (mode, rep, k, Hh.rel_err(g[:part.shape[0]], part))
                    frac, worst = Hh.elementwise_err(g[:part.shape[0]], part, rtol=1e-4, atol_frac=1e-5)
                    assert frac <= 1e-4, (mode, rep, k, frac, worst)
                    assert g[part.shape[0]:].abs().max().item() == 0.0 if g.shape[0] > part.shape[0] else True, (k, "padding")
                # per-view statistics of my Gaussians: [scene block ; object block] of the shard
                for vi in range(V):
                    d2, radii = stats[vi]
                    full_d2, full_r = want_d2[vi], want_radii[vi]
                    d2_s, d2_o = full_d2[:N_s][rank::world], full_d2[N_s:][rank::world]
                    r_s, r_o = full_r[:N_s][rank::world], full_r[N_s:][rank::world]
                    n_s, n_o = d2_s.shape[0], d2_o.shape[0]
                    assert Hh.rel_err(d2[:n_s], d2_s) <= 1e-5 and Hh.rel_err(d2[ns:ns + n_o], d2_o) <= 1e-5
                    if radii is not None:
                        assert torch.equal(radii[:n_s], r_s) and torch.equal(radii[ns:ns + n_o], r_o)
            if ex._peer is not None:
                assert not ex._peer.timed_out()
            dist.barrier()
        print(f"rank {rank}/{world}: splat exchange == sequential per-view sum ({V} views, peer + nccl)", flush=True)
    finally:
        dist.destroy_process_group()


def _run(world, rounds=2):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), rounds), nprocs=world, join=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_splat_exchange_world_2_matches_sequential_per_view_sum():
    _run(2)


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs")
def test_splat_exchange_world_4_matches_sequential_per_view_sum():
    _run(4)


@pytest.mark.skipif(torch.cuda.device_count() < 8, reason="needs 8 GPUs")
def test_splat_exchange_world_8_matches_sequential_per_view_sum():
    _run(8, rounds=1)
