"""One training iteration of the hot path with its two neighbours: the inner part of `train.py:74-167`
(render -> losses -> backward -> optimizer step) on the B200-native pieces -- `gaussian_renderer.render`
(fused trajectory + rasterizer), `losses.image_loss` / `losses.pixel_losses` (fused loss front-end) and
`optimizer.FusedAdam` (one-launch Adam over the reference's 18 parameter groups).

`densification_step` is the block of train.py:149-160 (statistics, densify_and_prune every
densification_interval iterations, near-index refresh, opacity reset) on adgs_b200/densify.py. The near-index
regularisers (lambda_reg, lambda_sigma_reg, train.py:101-110) are torch gathers over `obj_near_idx`
(adgs_b200/losses.py); their gradient is dense over all control-point columns, so they cannot be combined with
the window-aware optimizer mode. An `adgs_b200.env.EnvironmentMap` passed as `env_map` is composited by render()
and stepped here. Not reproduced: logging, evaluation, checkpoint scheduling.
"""
import torch

from . import losses as LS
from .gaussian_renderer import render


def densification_step(model, opt, iteration, render_pkg, white_background=False):
    """train.py:149-160. `opt`: densify_until_iter, densify_from_iter, densification_interval,
    opacity_reset_interval, densify_scene_grad_threshold, densify_obj_grad_threshold, near_idx_reset_interval.
    Returns True when the Gaussian set changed (the gradients of this iteration then no longer match the
    parameters: like the reference, which steps the freshly concatenated tensors with no .grad, the caller's
    optimizer step is a no-op for them -- here the gradients are dropped)."""
    if iteration >= opt.densify_until_iter:
        return False
    model.add_densification_stats(render_pkg)            # + the max_radii2D update of train.py:151
    changed = False
    if iteration > opt.densify_from_iter and iteration % opt.densification_interval == 0:
        model.densify_and_prune(opt.densify_scene_grad_threshold, opt.densify_obj_grad_threshold, 0.005,
                                iteration > opt.opacity_reset_interval)
        changed = True
    elif getattr(model, "use_near_idx", False) and iteration % opt.near_idx_reset_interval == 0:
        model.set_obj_near_idx()
    if iteration % opt.opacity_reset_interval == 0 or (white_background and iteration == opt.densify_from_iter):
        model.reset_opacity()
    return changed


def training_iteration(model, viewpoint_cam, opt, pipe, iteration, env_map=None, flow_pkg=None, frame_gap=None,
                       densify=None):
    """`viewpoint_cam` needs the reference Camera's fields render() reads plus the targets train.py reads:
    `original_image` (3,H,W), and optionally `depth` (H,W), `semantic` (H,W), `sky` (H,W).
    `opt`: lambda_dssim, lambda_l1, lambda_depth, lambda_flow, lambda_obj, lambda_sky, lambda_sigma.
    `densify`: the optimization arguments of densification_step (None = no densification).
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
    reg_loss = reg_sigma_loss = None
    near = getattr(model, "use_near_idx", False) and model.n_obj > 0 and \
        getattr(model, "obj_near_idx", None) is not None and model.obj_near_idx.numel() > 0
    if near and lam("lambda_reg") > 0.0:                                               # train.py:101-103
        if model.sparse_deform_grads:
            raise RuntimeError("lambda_reg > 0 needs dense control-point gradients: use training_setup(window_aware=False)")
        reg_loss = LS.near_reg_loss(model)
        loss = loss + lam("lambda_reg") * reg_loss
    if near and lam("lambda_sigma") > 0.0 and lam("lambda_sigma_reg") > 0.0:           # train.py:108-110
        reg_sigma_loss = LS.near_sigma_reg_loss(model)
        loss = loss + lam("lambda_sigma_reg") * reg_sigma_loss
    loss.backward()
    if densify is not None:                                                            # train.py:149-160
        densification_step(model, densify, iteration, render_pkg)
    model.optimizer.step()
    model.optimizer.zero_grad(set_to_none=True)
    if env_map is not None and getattr(env_map, "optimizer", None) is not None:   # train.py:165,167
        env_map.optimizer.step()
        env_map.optimizer.zero_grad(set_to_none=True)
    logs = {"total_loss": loss.detach(), "depth_loss": px["depth_loss"], "flow_loss": px["flow_loss"],
            "obj_loss": px["obj_loss"], "sky_loss": px["sky_loss"], "sigma_loss": sigma_loss, "reg_loss": reg_loss, "reg_sigma_loss": reg_sigma_loss}
    return logs, render_pkg
