"""Backbone for the end-to-end harness: mirror of /root/reference/common/nets/resnet.py (ResNetBackbone).

Off the accelerated path (stays stock PyTorch / cuDNN); it exists so that BASELINE.json's
"ResNet-50 + deconv head + integral loss" configuration can be trained end to end.  Parameter names match the
reference (`conv1, bn1, layer1..layer4`), so its checkpoints load.  No pretrained download (no network): weights
are initialised like resnet.py:27-33 (normal(0, 0.001) convs, unit BatchNorm).
"""
import torch.nn as nn
from torchvision.models.resnet import BasicBlock, Bottleneck, ResNet

_SPEC = {18: (BasicBlock, [2, 2, 2, 2]), 34: (BasicBlock, [3, 4, 6, 3]), 50: (Bottleneck, [3, 4, 6, 3]),
         101: (Bottleneck, [3, 4, 23, 3]), 152: (Bottleneck, [3, 8, 36, 3])}


class ResNetBackbone(ResNet):
    def __init__(self, resnet_type):
        block, layers = _SPEC[resnet_type]
        super().__init__(block, layers)
        del self.fc, self.avgpool                      # the reference keeps only the convolutional trunk
        self.name = "resnet%d" % resnet_type
        self.out_channels = 512 * block.expansion
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, mean=0, std=0.001)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):                              # resnet.py:54-65
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))

    def init_weights(self):
        """The reference downloads ImageNet weights here (resnet.py:67-73); there is no network in this setting,
        so the random initialisation above stays."""
        return None
