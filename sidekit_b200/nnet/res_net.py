"""The ResNet trunks with the reference's module tree and state_dict keys (sidekit/nnet/res_net.py:186-320, :430-610).

Inside an ``Xtractor`` these modules are parameter containers: the fused engine (csrc/engine.cu) runs the whole trunk
without leaving its plane layout.  On their own -- ``BasicBlock(...)(x)``, ``SELayer``, ``ResBlock``,
``PreHalfResNet34()(feats)`` -- their ``forward`` composes the stand-alone operators of ``nnet/functional.py`` (the same
tcgen05 convolution kernel behind ``skb_conv2d_bn_act``), eval-mode semantics (BatchNorm running statistics), dense fp32
CUDA tensors in and out.  Training is out of scope.
"""
import torch

from . import functional as Fn


class SELayer(torch.nn.Module):
    def __init__(self, channel, reduction=16):
        super().__init__()
        self.fc = torch.nn.Sequential(
            torch.nn.Linear(channel, channel // reduction, bias=False),
            torch.nn.ReLU(inplace=True),
            torch.nn.Linear(channel // reduction, channel, bias=False),
            torch.nn.Sigmoid())

    def gate(self, mean):
        """(B, C) channel means -> (B, C) gates (res_net.py:265-270, :280)."""
        return Fn.se_gate(mean, self.fc[0].weight, self.fc[2].weight)

    def forward(self, x):
        """res_net.py:272-281: ``x * sigmoid(fc(mean_{h,w}(x)))``."""
        return Fn.scale_residual_act(x, self.gate(Fn.channel_mean(x)), None, 1.0)


class ResBlock(torch.nn.Module):
    """Pre-activation residual block (sidekit/nnet/res_net.py:186-255): [BN + LeakyReLU unless ``is_first``] -> conv1 ->
    BN2 -> LeakyReLU -> conv2 -> the SAME BN2 again -> + identity (1x1 conv + BN when the width changes) -> LeakyReLU.
    Both convolutions carry a bias and take ``stride``; as in the reference only ``stride=1`` yields shapes that can be
    added to the identity."""

    def __init__(self, in_channels, out_channels, stride, is_first=False):
        super().__init__()
        self.is_first = is_first
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.expansion = self.out_channels // self.in_channels
        self.resample = None
        if not self.in_channels == self.out_channels:
            self.resample = torch.nn.Sequential(
                torch.nn.Conv2d(in_channels=self.in_channels, out_channels=self.out_channels, kernel_size=1),
                torch.nn.BatchNorm2d(self.in_channels * self.expansion))
        if not self.is_first:
            self.batch_norm1 = torch.nn.BatchNorm2d(num_features=self.in_channels)
        self.activation = torch.nn.LeakyReLU()
        self.conv1 = torch.nn.Conv2d(in_channels=self.in_channels, out_channels=self.out_channels, kernel_size=(3, 3), stride=stride,
                                     padding=1, padding_mode='zeros', dilation=1)
        self.conv2 = torch.nn.Conv2d(in_channels=self.out_channels, out_channels=self.out_channels, stride=stride, kernel_size=(3, 3),
                                     padding=1, padding_mode='zeros', dilation=1)
        self.batch_norm2 = torch.nn.BatchNorm2d(num_features=self.out_channels)

    def forward(self, x, compute_dtype="fp16"):
        slope = self.activation.negative_slope
        stride = self.conv1.stride
        pre = None if self.is_first else Fn.bn_affine(self.batch_norm1, x.device)
        w1, b1 = Fn.fold_conv_bn(self.conv1, self.batch_norm2)
        w2, b2 = Fn.fold_conv_bn(self.conv2, self.batch_norm2)
        out = Fn.conv2d_bn_act(x, w1, b1, stride, slope, pre=pre, pre_slope=slope, compute_dtype=compute_dtype)
        identity = x
        if not self.expansion == 1:
            wr, br = Fn.fold_conv_bn(self.resample[0], self.resample[1])
            identity = Fn.conv2d_bn_act(x, wr, br, 1, 1.0, compute_dtype=compute_dtype)
        out = Fn.conv2d_bn_act(out, w2, b2, stride, slope, residual=identity, compute_dtype=compute_dtype)
        Fn.check_overflow(x.device)
        return out


class BasicBlock(torch.nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = torch.nn.Conv2d(in_planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = torch.nn.BatchNorm2d(planes)
        self.conv2 = torch.nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = torch.nn.BatchNorm2d(planes)
        self.se = SELayer(planes)
        self.shortcut = torch.nn.Sequential()
        # the reference compares the *tuple* stride with 1, so layer1.0 gets a 1x1 shortcut too (res_net.py:302)
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = torch.nn.Sequential(
                torch.nn.Conv2d(in_planes, self.expansion * planes, kernel_size=1, stride=stride, bias=False),
                torch.nn.BatchNorm2d(self.expansion * planes))

    def forward(self, x, compute_dtype="fp16", check=True):
        """res_net.py:309-320: ``relu(se(bn2(conv2(relu(bn1(conv1(x)))))) + shortcut(x))``."""
        stride = self.conv1.stride
        w1, b1 = Fn.fold_conv_bn(self.conv1, self.bn1)
        w2, b2 = Fn.fold_conv_bn(self.conv2, self.bn2)
        y1 = Fn.conv2d_bn_act(x, w1, b1, stride, 0.0, compute_dtype=compute_dtype)
        y2 = Fn.conv2d_bn_act(y1, w2, b2, 1, 1.0, compute_dtype=compute_dtype)
        scale = self.se.gate(Fn.channel_mean(y2))
        res = x
        if len(self.shortcut):
            ws, bs = Fn.fold_conv_bn(self.shortcut[0], self.shortcut[1])
            res = Fn.conv2d_bn_act(x, ws, bs, stride, 1.0, compute_dtype=compute_dtype)
        out = Fn.scale_residual_act(y2, scale, res, 0.0)
        if check:
            Fn.check_overflow(x.device)
        return out


class PreHalfResNet34(torch.nn.Module):
    def __init__(self, block=BasicBlock, num_blocks=(3, 4, 6, 3), speaker_number=10):
        super().__init__()
        if tuple(num_blocks) != (3, 4, 6, 3) or block is not BasicBlock:
            raise NotImplementedError("the CUDA engine implements the (3, 4, 6, 3) BasicBlock HalfResNet34")
        self.in_planes = 32
        self.speaker_number = speaker_number
        self.conv1 = torch.nn.Conv2d(1, 32, kernel_size=3, stride=(1, 1), padding=1, bias=False)
        self.bn1 = torch.nn.BatchNorm2d(32)
        self.layer1 = self._make_layer(block, 32, num_blocks[0], stride=(1, 1))
        self.layer2 = self._make_layer(block, 64, num_blocks[1], stride=(2, 2))
        self.layer3 = self._make_layer(block, 128, num_blocks[2], stride=(2, 2))
        self.layer4 = self._make_layer(block, 256, num_blocks[3], stride=(2, 2))

    def _make_layer(self, block, planes, num_blocks, stride):
        strides = [stride] + [1] * (num_blocks - 1)
        layers = []
        for stride in strides:
            layers.append(block(self.in_planes, planes, stride))
            self.in_planes = planes * block.expansion
        return torch.nn.Sequential(*layers)

    def forward(self, x, compute_dtype="fp16"):
        """res_net.py:539-554 (eval mode): (B, F, T) features -> (B, 1, T, F) -> stem -> layers; returns (B, C, T', F')."""
        if len(x.shape) == 3:
            x = x.unsqueeze(1).permute(0, 1, 3, 2)
        w, b = Fn.fold_conv_bn(self.conv1, self.bn1)
        x = Fn.conv2d_bn_act(x.contiguous(), w, b, 1, 0.0, compute_dtype=compute_dtype)
        for name, layer in self.named_children():
            if name.startswith("layer"):
                for block in layer:
                    x = block(x, compute_dtype=compute_dtype, check=False)
        Fn.check_overflow(x.device)
        return x


class PreResNet34(torch.nn.Module):
    """sidekit/nnet/res_net.py:430-498: 128-channel stem, seven layers (3, 1, 3, 1, 5, 1, 2) at 128 / 128 / 128 / 256 /
    256 / 256 / 256 channels with INT strides 1, 2, 1, 2, 1, 2, 1 (so only the three stride-2 / widening blocks have a
    1x1 shortcut).  As in the reference, layer7 is built with ``num_blocks[5]`` and therefore has ONE block."""

    def __init__(self, block=BasicBlock, num_blocks=(3, 1, 3, 1, 5, 1, 2), speaker_number=10):
        super().__init__()
        if tuple(num_blocks) != (3, 1, 3, 1, 5, 1, 2) or block is not BasicBlock:
            raise NotImplementedError("the CUDA engine implements the (3, 1, 3, 1, 5, 1, 2) BasicBlock ResNet34")
        self.in_planes = 128
        self.speaker_number = speaker_number
        self.conv1 = torch.nn.Conv2d(1, 128, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = torch.nn.BatchNorm2d(128)
        self.layer1 = self._make_layer(block, 128, num_blocks[0], stride=1)
        self.layer2 = self._make_layer(block, 128, num_blocks[1], stride=2)
        self.layer3 = self._make_layer(block, 128, num_blocks[2], stride=1)
        self.layer4 = self._make_layer(block, 256, num_blocks[3], stride=2)
        self.layer5 = self._make_layer(block, 256, num_blocks[4], stride=1)
        self.layer6 = self._make_layer(block, 256, num_blocks[5], stride=2)
        self.layer7 = self._make_layer(block, 256, num_blocks[5], stride=1)     # num_blocks[5] as in the reference (:455)

    _make_layer = PreHalfResNet34._make_layer
    forward = PreHalfResNet34.forward


class PreFastResNet34(torch.nn.Module):
    """sidekit/nnet/res_net.py:557-610: 7x7 stem with stride (1, 2) to 16 channels, layers (3, 4, 6, 3) at 16 / 32 / 64 /
    128 channels with strides 1 (an int: layer1.0 has no shortcut), (2, 2), (2, 2), (1, 1) -- layer4 widens at the
    resolution of layer3 and, its stride being a tuple, its first block has a 1x1 shortcut."""

    def __init__(self, block=BasicBlock, num_blocks=(3, 4, 6, 3), speaker_number=10):
        super().__init__()
        if tuple(num_blocks) != (3, 4, 6, 3) or block is not BasicBlock:
            raise NotImplementedError("the CUDA engine implements the (3, 4, 6, 3) BasicBlock FastResNet34")
        self.in_planes = 16
        self.speaker_number = speaker_number
        self.conv1 = torch.nn.Conv2d(1, 16, kernel_size=7, stride=(1, 2), padding=3, bias=False)
        self.bn1 = torch.nn.BatchNorm2d(16)
        self.layer1 = self._make_layer(block, 16, num_blocks[0], stride=1)
        self.layer2 = self._make_layer(block, 32, num_blocks[1], stride=(2, 2))
        self.layer3 = self._make_layer(block, 64, num_blocks[2], stride=(2, 2))
        self.layer4 = self._make_layer(block, 128, num_blocks[3], stride=(1, 1))

    _make_layer = PreHalfResNet34._make_layer

    def forward(self, x):
        raise RuntimeError("the 7x7 stride-(1, 2) stem only exists inside the fused CUDA engine: call "
                           "Xtractor('fastresnet34').forward(x, is_eval=True)")
