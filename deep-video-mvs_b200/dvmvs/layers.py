"""Parameter containers with the reference's names (dvmvs/layers.py:39-65).  The returned nn.Sequential objects
only HOLD parameters / BatchNorm buffers under the reference's state-dict keys ('0.weight', '1.running_mean', ...);
the forward pass of the owning module never calls them -- it runs the folded weights through the sm_100a kernels."""
import torch


def conv_layer(input_channels, output_channels, kernel_size, stride, apply_bn_relu):
    mods = [torch.nn.Conv2d(input_channels, output_channels, kernel_size, padding=(kernel_size - 1) // 2, stride=stride, bias=False)]
    if apply_bn_relu:
        mods += [torch.nn.BatchNorm2d(output_channels), torch.nn.ReLU(inplace=True)]
    return torch.nn.Sequential(*mods)


def depth_layer_3x3(input_channels):
    return torch.nn.Sequential(torch.nn.Conv2d(input_channels, 1, 3, padding=1), torch.nn.Sigmoid())
