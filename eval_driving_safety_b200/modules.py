"""nn.Module drop-ins with the stock constructor signatures and parameter
names/shapes (SURVEY 8b "module seams"), so upstream checkpoints load and the
reference's scripts run unchanged after ``swap_modules(model)``.

Every drop-in validates ITS OWN configuration at call time and ``swap_modules``
validates the layer it replaces before replacing it: a layer the sm_100a kernels
cannot reproduce exactly (bias on a 3-D conv, other padding / dilation / groups /
output_padding ...) is never silently recomputed as something else."""
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


def _conv3d_problem(m):
    """None if the sm_100a kernels compute exactly this nn.Conv3d, else the reason."""
    if tuple(m.kernel_size) != (3, 3, 3):
        return "kernel_size %s (only 3x3x3)" % (tuple(m.kernel_size),)
    if tuple(m.padding) != (1, 1, 1) or m.padding_mode != 'zeros':
        return "padding %s / %s (only zero padding 1)" % (m.padding, m.padding_mode)
    if tuple(m.dilation) != (1, 1, 1) or m.groups != 1:
        return "dilation %s groups %d (only dense, undilated)" % (tuple(m.dilation), m.groups)
    if tuple(m.stride) not in ((1, 1, 1), (2, 2, 2)):
        return "stride %s (only 1 or 2)" % (tuple(m.stride),)
    if m.bias is not None:
        return "bias (the hourglass convs are bias-free)"
    if m.out_channels == 1 and tuple(m.stride) != (1, 1, 1):
        return "Cout = 1 with stride 2"
    return None


def _deconv3d_problem(m):
    if tuple(m.kernel_size) != (3, 3, 3) or tuple(m.stride) != (2, 2, 2) or tuple(m.padding) != (1, 1, 1) \
            or tuple(m.output_padding) != (1, 1, 1):
        return "k%s s%s p%s op%s (only k3 s2 p1 op1)" % (tuple(m.kernel_size), tuple(m.stride), tuple(m.padding),
                                                       tuple(m.output_padding))
    if tuple(m.dilation) != (1, 1, 1) or m.groups != 1:
        return "dilation %s groups %d" % (tuple(m.dilation), m.groups)
    if m.bias is not None:
        return "bias"
    return None


def _conv2d_problem(m):
    k = m.kernel_size[0]
    if m.kernel_size[0] != m.kernel_size[1] or k not in (1, 3):
        return "kernel_size %s (only 1x1 / 3x3)" % (tuple(m.kernel_size),)
    if m.stride[0] != m.stride[1] or m.stride[0] not in (1, 2):
        return "stride %s" % (tuple(m.stride),)
    if m.dilation[0] != m.dilation[1] or m.dilation[0] not in (1, 2) or (m.dilation[0] == 2 and m.stride[0] != 1):
        return "dilation %s" % (tuple(m.dilation),)
    if isinstance(m.padding, str) or tuple(m.padding) != (m.dilation[0] * (k // 2),) * 2 or m.padding_mode != 'zeros':
        return "padding %s (only dilation * (k // 2), zeros)" % (m.padding,)
    if m.groups != 1:
        return "groups %d" % m.groups
    if m.in_channels == 3:
        if not (k == 3 and m.stride[0] == 2 and m.bias is None and m.out_channels % 8 == 0 and m.out_channels <= 64):
            return "3-channel layer other than k3 s2 p1 bias-free"
    elif m.in_channels % 32:
        return "%d input channels (multiples of 32, or the 3-channel first layer)" % m.in_channels
    return None


class Conv3dSm100(nn.Conv3d):
    """nn.Conv3d(in, out, 3, stride 1|2, 1, bias=False) on the sm_100a kernels."""

    def forward(self, x):
        why = _conv3d_problem(self)
        if why:
            raise RuntimeError("Conv3dSm100: unsupported configuration: " + why)
        if self.out_channels == 1:
            return ops.conv3d_c1(x, self.weight)
        return ops.conv3d(x, self.weight, stride=self.stride[0], transposed=False)


class ConvTranspose3dSm100(nn.ConvTranspose3d):
    """nn.ConvTranspose3d(in, out, 3, 2, 1, output_padding=1, bias=False)."""

    def forward(self, x, output_size=None):
        why = _deconv3d_problem(self)
        if why or output_size is not None:
            raise RuntimeError("ConvTranspose3dSm100: unsupported configuration: " + (why or "output_size"))
        return ops.conv3d(x, self.weight, stride=2, transposed=True)


class Conv2dSm100(nn.Conv2d):
    """nn.Conv2d (k 1|3, stride 1|2, padding = dilation * (k // 2), optional bias) on the sm_100a kernels
    (3xTF32 tensor-core path; the 3-channel first layer in exact fp32)."""

    def forward(self, x):
        why = _conv2d_problem(self)
        if why:
            raise RuntimeError("Conv2dSm100: unsupported configuration: " + why)
        return ops.conv2d(x, self.weight, self.bias, self.stride[0], self.dilation[0])


class GroupNormSm100(nn.GroupNorm):
    """nn.GroupNorm on 4-D maps / 5-D volumes; ``relu=True`` fuses a following ReLU, ``res`` a residual."""

    def __init__(self, num_groups, num_channels, eps=1e-5, affine=True, relu=False):
        super().__init__(num_groups, num_channels, eps, affine)
        self.relu = relu

    def forward(self, x, res=None):
        if x.dim() not in (4, 5) or not self.affine:
            raise RuntimeError("GroupNormSm100: needs an affine norm on [N,C,H,W] or [N,C,D,H,W] input")
        return ops.groupnorm_act(x, self.weight, self.bias, self.num_groups, self.eps, self.relu, res)


GroupNorm3dSm100 = GroupNormSm100      # round-1 name


def _clone_as(cls, child, *args, **kw):
    new = cls.__new__(cls)
    nn.Module.__init__(new)
    new.__dict__.update({k: v for k, v in child.__dict__.items() if not k.startswith('_')})   # hyper-parameters
    new._parameters = child._parameters           # the very same Parameter objects: state_dict keys unchanged
    new._buffers = child._buffers
    new.training = child.training
    for k, v in kw.items():
        setattr(new, k, v)
    return new


def swap_modules(model, strict=True):
    """Replace the stock Conv3d / ConvTranspose3d / Conv2d / GroupNorm children of ``model`` by the sm_100a
    modules in place.  The replacement keeps the child's own hyper-parameters and its Parameter objects (so
    state_dict keys and values are untouched).  A layer the kernels cannot reproduce exactly raises
    (``strict``) or is left as it is (``strict=False``).  Parameters are frozen (attack path).
    Returns the model."""
    for name, child in list(model.named_children()):
        new = None
        if type(child) is nn.Conv3d:
            why = _conv3d_problem(child)
            new = None if why else _clone_as(Conv3dSm100, child)
        elif type(child) is nn.ConvTranspose3d:
            why = _deconv3d_problem(child)
            new = None if why else _clone_as(ConvTranspose3dSm100, child)
        elif type(child) is nn.Conv2d:
            why = _conv2d_problem(child)
            new = None if why else _clone_as(Conv2dSm100, child)
        elif type(child) is nn.GroupNorm:
            why = None if child.affine else "GroupNorm without affine parameters"
            new = None if why else _clone_as(GroupNormSm100, child, relu=False)
        else:
            swap_modules(child, strict)
            continue
        if new is not None:
            setattr(model, name, new)
        elif strict:
            raise ValueError("swap_modules: %s (%r) is not supported by the sm_100a kernels: %s" % (name, child, why))
    for p in model.parameters():
        p.requires_grad_(False)
    return model
