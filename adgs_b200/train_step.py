"""One training iteration of the hot path with its two neighbours: the inner part of `train.py:74-167`
(render -> losses -> backward -> optimizer step) on the B200-native pieces -- `gaussian_renderer.render`
(fused trajectory + rasterizer), `losses.image_loss` / `losses.pixel_losses` (fused loss front-end) and
`optimizer.FusedAdam` (one-launch Adam over the reference's 18 parameter groups).

Not reproduced here (out of scope): densification / pruning, the near-index
regularisers (lambda_reg, lambda_sigma_reg: they need `obj_near_idx` from pytorch3d knn_points), logging,
checkpointing. An `adgs_b200.env.EnvironmentMap` passed as `env_map` is composited by render() and stepped here. `lambda_sigma`'s own term (train.py:105-107) is included because it
only touches `gs_time_sigma`.
"""
import torch

from . import losses as LS
from .gaussian_renderer import render


def training_iteration(model, viewpoint_cam, opt, pipe, iteration, env_map=None, flow_pkg=None, frame_gap=None):
    """`viewpoint_cam` needs the reference Camera's fields render() reads plus the targets train.py reads:
    `original_image` (3,H,W), and optionally `depth` (H,W), `semantic` (H,W), `sky` (H,W).
    `opt`: lambda_dssim, lambda_l1, lambda_depth, lambda_flow, lambda_obj, lambda_sky, lambda_sigma.
    Returns the dict of loss values (device tensors, no host synchronisation) and the render package."""
    model.update_learning_rate(iteration)
    lam = lambda k: float(getattr(opt, k, 0.0))
    use_flow = lam("lambda_flow") > 0.0 and flow_pkg is not None
    render_pkg = render(viewpoint_cam, model, env_map, pipe, flow_pkg=flow_pkg if use_flow else None,
                        render_objmask=lam("lambda_obj") > 0.0)
    image = render_pkg["render"]
    loss = LS.image_loss(image, viewpoint_cam.original_image, lam("lambda_dssim"), lam("lambda_l1") or 1.0)
    px = LS.pixel_losses(
        depth=render_pkg["depth"], img_semantic=render_pkg["img_semantic"], img_opacity=render_pkg["img_opacity"],
        img_flow=render_pkg["img_flow"] if use_flow else None,
        gt_depth=getattr(viewpoint_cam, "depth", None), gt_semantic=getattr(viewpoint_cam, "semantic", None),
        gt_sky=getattr(viewpoint_cam, "sky", None), flow_pkg=flow_pkg if use_flow else None,
        flow_dist=getattr(model, "scene_extent", 1.0) * 1e-3, lambda_depth=lam("lambda_depth"),
        lambda_obj=lam("lambda_obj"), lambda_sky=lam("lambda_sky"), lambda_flow=lam("lambda_flow"))
    loss = loss + px["weighted"]
    sigma_loss = None
    if lam("lambda_sigma") > 0.0 and model.n_obj > 0 and frame_gap is not None:
        time_sigma = torch.exp(model.gs_time_sigma)                                    # train.py:105-107
        sigma_loss = torch.mean(torch.abs(frame_gap / torch.mean(time_sigma, dim=-1)))
        loss = loss + lam("lambda_sigma") * sigma_loss
    loss.backward()
    model.optimizer.step()
    model.optimizer.zero_grad(set_to_none=True)
    if env_map is not None and getattr(env_map, "optimizer", None) is not None:   # train.py:165,167
        env_map.optimizer.step()
        env_map.optimizer.zero_grad(set_to_none=True)
    logs = {"total_loss": loss.detach(), "depth_loss": px["depth_loss"], "flow_loss": px["flow_loss"],
            "obj_loss": px["obj_loss"], "sky_loss": px["sky_loss"], "sigma_loss": sigma_loss}
    return logs, render_pkg
