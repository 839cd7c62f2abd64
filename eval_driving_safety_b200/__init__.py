"""B200-native (sm_100a) hot path of the eval_driving_safety attack loops.

Public surface:
  attack   -- pgd_step / pgd_step_pair / stereo_rcnn_pgd_step / patch_apply /
              patch_update / pgd_attack / patch_attack_step
  ops      -- autograd Functions over the C ABI (cost volume, grid_sample lifting,
              conv3d/deconv3d + GroupNorm, RoIAlign)
  modules  -- nn.Module drop-ins (BuildCostVolume, Conv3dSm100, ...) + swap_modules
  dsgn     -- StereoNet with the reference's call signature
  parallel -- pair sharding over ranks + NCCL gather of per-pair statistics
There is no CPU fallback: importing works anywhere, calling needs libb2attack.so
and CUDA tensors.
"""
__version__ = "0.1.0"
