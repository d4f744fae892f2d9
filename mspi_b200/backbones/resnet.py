"""ResNet18-VGGSound audio encoder (parameter container), mirrors backbones/resnet.py:57-154:
1-channel conv1, four stages of two BasicBlocks, no pool/fc on the output path.  Forward arithmetic:
mspi_b200.engine.ForwardPlan.resnet18."""
import torch

from ..params import ParamNode, conv_bn, linear


class ResNet18Audio(ParamNode):
    def __init__(self, num_classes: int = 1000):
        super().__init__()
        conv_bn(self, "conv1", "bn1", 64, 1, (7, 7), init="kaiming_out")
        cin = 64
        for li, c in enumerate((64, 128, 256, 512), 1):
            for bi in range(2):
                q = f"layer{li}.{bi}"
                conv_bn(self, q + ".conv1", q + ".bn1", c, cin if bi == 0 else c, (3, 3), init="kaiming_out")
                conv_bn(self, q + ".conv2", q + ".bn2", c, c, (3, 3), init="kaiming_out")
                if bi == 0 and li > 1:
                    conv_bn(self, q + ".downsample.0", q + ".downsample.1", c, cin, (1, 1), init="kaiming_out")
            cin = c


def get_resnet18(pretrained=True, path=None, **kwargs):
    """backbones/resnet.py:149-154"""
    model = ResNet18Audio(**kwargs)
    if pretrained:
        model.load_state_dict(torch.load(path, map_location="cpu"))
    return model
