"""Parameter containers for the HalfResNet34 trunk with the reference's module tree and state_dict
keys (sidekit/nnet/res_net.py:258-320, :500-554).  They hold torch Parameters / BatchNorm statistics
only; the arithmetic (conv + folded BN + ReLU + SE + residual) runs in csrc/conv_umma.cuh and
csrc/layers.cu, driven by the extractor engine.
"""
import torch


class SELayer(torch.nn.Module):
    def __init__(self, channel, reduction=16):
        super().__init__()
        self.fc = torch.nn.Sequential(
            torch.nn.Linear(channel, channel // reduction, bias=False),
            torch.nn.ReLU(inplace=True),
            torch.nn.Linear(channel // reduction, channel, bias=False),
            torch.nn.Sigmoid())


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

    def forward(self, x):
        raise RuntimeError("the trunk runs inside the fused CUDA engine: call Xtractor.forward(x, is_eval=True) "
                           "(Xtractor.debug_stage exposes per-block activations for tests)")


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

    def forward(self, x):
        raise RuntimeError("the trunk runs inside the fused CUDA engine: call Xtractor.forward(x, is_eval=True)")


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
        raise RuntimeError("the trunk runs inside the fused CUDA engine: call Xtractor.forward(x, is_eval=True)")
