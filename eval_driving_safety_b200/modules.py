"""nn.Module drop-ins with the stock constructor signatures and parameter
names/shapes (SURVEY 8b "module seams"), so upstream checkpoints load and the
reference's scripts run unchanged after ``swap_modules(model)``."""
import torch
import torch.nn as nn

from . import ops


class BuildCostVolume(nn.Module):
    """Upstream ``dsgn.layers.BuildCostVolume(downsample)``:
    forward(left, right, shift) with shift [N,D] disparities in feature px."""

    def __init__(self, downsample=4, channels_last=False):
        super().__init__()
        self.downsample = downsample
        self.channels_last = channels_last

    def forward(self, left, right, shift):
        return ops.build_cost_volume(left, right, shift, self.channels_last)


class Conv3dSm100(nn.Conv3d):
    """nn.Conv3d(in, out, 3, stride 1|2, 1, bias=False) on the sm_100a kernels."""

    def forward(self, x):
        if self.kernel_size != (3, 3, 3) or self.padding != (1, 1, 1) or self.bias is not None \
                or self.dilation != (1, 1, 1) or self.groups != 1:
            raise RuntimeError("Conv3dSm100 supports k3/p1/bias-free/dense convolutions only")
        if self.out_channels == 1:
            return ops.conv3d_c1(x, self.weight)
        return ops.conv3d(x, self.weight, stride=self.stride[0], transposed=False)


class ConvTranspose3dSm100(nn.ConvTranspose3d):
    """nn.ConvTranspose3d(in, out, 3, 2, 1, output_padding=1, bias=False)."""

    def forward(self, x):
        if self.kernel_size != (3, 3, 3) or self.padding != (1, 1, 1) or self.stride != (2, 2, 2) \
                or self.output_padding != (1, 1, 1) or self.bias is not None:
            raise RuntimeError("ConvTranspose3dSm100 supports k3/s2/p1/op1/bias-free only")
        return ops.conv3d(x, self.weight, stride=2, transposed=True)


class GroupNorm3dSm100(nn.GroupNorm):
    """nn.GroupNorm on 5-D volumes; ``relu=True`` fuses the following ReLU."""

    def __init__(self, num_groups, num_channels, eps=1e-5, relu=False):
        super().__init__(num_groups, num_channels, eps)
        self.relu = relu

    def forward(self, x, res=None):
        return ops.groupnorm_act(x, self.weight, self.bias, self.num_groups, self.eps, self.relu, res)


def swap_modules(model):
    """Replace stock Conv3d / ConvTranspose3d / (5-D) GroupNorm children of
    ``model`` by the sm_100a modules in place, keeping parameters (and therefore
    state_dict keys); parameters are frozen (attack path)."""
    for name, child in list(model.named_children()):
        new = None
        if type(child) is nn.Conv3d:
            new = Conv3dSm100(child.in_channels, child.out_channels, 3, child.stride, 1, bias=False)
        elif type(child) is nn.ConvTranspose3d:
            new = ConvTranspose3dSm100(child.in_channels, child.out_channels, 3, 2, 1, output_padding=1, bias=False)
        if new is not None:
            new.weight = child.weight
            setattr(model, name, new)
        else:
            swap_modules(child)
    for p in model.parameters():
        p.requires_grad_(False)
    return model
