"""adgs_b200 -- B200-native (sm_100a) implementation of the AD-GS per-iteration hot path.

Public surface (mirrors the reference's plugin API):
  adgs_b200.rasterizer   GaussianRasterizationSettings, GaussianRasterizer, _C   (diff_gaussian_rasterization)
  adgs_b200.simple_knn   distCUDA2                                               (simple_knn._C)
  adgs_b200.gaussian_model / gaussian_renderer   fused trajectory + render path  (scene.gaussian_model, gaussian_renderer)
  adgs_b200.optimizer    FusedAdam, training_setup, update_learning_rate         (GaussianModel.training_setup, optimizer.step)
  adgs_b200.losses       l1_loss, ssim, image_loss, pixel_losses, near_reg_loss  (utils.loss_utils, train.py:79-115)
  adgs_b200.env          EnvironmentMap                                          (scene.env)
  adgs_b200.densify      densify_and_prune, prune_points, reset_opacity, add_densification_stats, set_obj_near_idx,
                         knn_points                                              (GaussianModel densification, pytorch3d knn_points)
  adgs_b200.checkpoint   save_ply / load_ply (point_cloud.ply + deform.pth)      (GaussianModel.save_ply / load_ply)
  adgs_b200.train_step   training_iteration, densification_step                  (train.py:74-167)
  adgs_b200.parallel     MultiViewStep, SplatExchangeStep                        (multi-GPU; no reference counterpart)
All compute goes through libadgs_b200.so (include/adgs_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
