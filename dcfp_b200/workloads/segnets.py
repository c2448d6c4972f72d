"""Calibration workloads: the segmentation networks whose conv/BN feature maps are scored.

These are *workload generators* for bench.py / tests on the GPU box, where
`/root/reference` does not exist.  They are written from scratch as one table-driven
builder, but reproduce the reference's architectures exactly -- same module names
(so `ignore_prune_layer`, `score.pth` keys and `channel_cfg` keys agree), same
parameter-creation order (so `torch.manual_seed(s)` + default init yields bit-identical
weights) and the same op order in `forward` (bit-identical CPU outputs).
`tests/test_reference_live_cpu.py::test_workload_nets_equal_reference` checks all three against the unmodified reference.

Architectures (reference file:line):
  dilated ResNet-50/101/152, 3-conv stem, multi-grid layer4   networks/backbone/resnet.py:60-187
  ASPP head                                                    networks/tools/aspp.py:37-84
  DeepLabV3  `Seg_Model`                                       networks/deeplabv3.py:12-59
  DeepLabV3+ `Seg_Model` + decoder                             networks/deeplabv3p.py:12-99
  PSPNet `Seg_Model` + pyramid pooling                         networks/psp.py:12-49, networks/tools/ppm.py:10-38
  CE + 0.4 * deep-supervision CE                               loss/criterion.py:48-74
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

_DEPTHS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}
_STRIDES = {8: ((1, 2, 1, 1), (1, 1, 2, 4)), 16: ((1, 2, 2, 1), (1, 1, 1, 2)), 32: ((1, 2, 2, 2), (1, 1, 1, 1))}
_ASPP_RATES = {8: (1, 12, 24, 36), 16: (1, 6, 12, 18), 32: (1, 3, 6, 9)}


def _cbr(cin, cout, k, **kw):
    """conv (no bias) -> BN -> in-place ReLU, as a list."""
    return [nn.Conv2d(cin, cout, k, bias=False, **kw), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]


class Bottleneck(nn.Module):
    def __init__(self, cin, width, stride, dilation, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, dilation=dilation, padding=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, width * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(width * 4)
        self.relu = nn.ReLU(inplace=True)
        self.relu_inplace = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        shortcut = x if self.downsample is None else self.downsample(x)
        return self.relu_inplace(y + shortcut)


class DilatedResNet(nn.Module):
    def __init__(self, depths, output_stride=8, inplanes=128, mg_unit=(1, 2, 4), out_index=(3, 4)):
        super().__init__()
        strides, dilations = _STRIDES[output_stride]
        self.out_index = tuple(out_index)
        stem = _cbr(3, 64, 3, stride=2, padding=1) + _cbr(64, 64, 3, stride=1, padding=1)
        stem.append(nn.Conv2d(64, inplanes, 3, 1, 1, bias=False))
        self.conv1 = nn.Sequential(*stem)
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.relu1 = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        cin = inplanes
        for i, width in enumerate((64, 128, 256, 512)):
            if i < 3:
                rates = [dilations[i]] * depths[i]
            else:  # multi-grid unit replaces the plain stage (resnet.py:93,117-134)
                rates = [m * dilations[i] for m in mg_unit]
            blocks = []
            for j, rate in enumerate(rates):
                stride = strides[i] if j == 0 else 1
                down = None
                if j == 0 and (stride != 1 or cin != width * 4):
                    down = nn.Sequential(nn.Conv2d(cin, width * 4, 1, stride=stride, bias=False),
                                         nn.BatchNorm2d(width * 4))
                blocks.append(Bottleneck(cin, width, stride, rate, down))
                cin = width * 4
            setattr(self, "layer%d" % (i + 1), nn.Sequential(*blocks))

    def forward(self, x):
        x = self.maxpool(self.relu1(self.bn1(self.conv1(x))))
        outs = []
        for i in range(1, 5):
            x = getattr(self, "layer%d" % i)(x)
            if i in self.out_index:
                outs.append(x)
        return tuple(outs)


class _ASPPBranch(nn.Module):
    def __init__(self, cin, cout, k, rate):
        super().__init__()
        self.atrous_conv = nn.Conv2d(cin, cout, k, stride=1, padding=0 if k == 1 else rate, dilation=rate, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.relu(self.bn(self.atrous_conv(x)))


class ASPP(nn.Module):
    def __init__(self, output_stride, align_corner, inplanes=2048, outplanes=512):
        super().__init__()
        r = _ASPP_RATES[output_stride]
        self.align_corner = align_corner
        self.aspp1 = _ASPPBranch(inplanes, 256, 1, r[0])
        self.aspp2 = _ASPPBranch(inplanes, 256, 3, r[1])
        self.aspp3 = _ASPPBranch(inplanes, 256, 3, r[2])
        self.aspp4 = _ASPPBranch(inplanes, 256, 3, r[3])
        self.global_avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), *_cbr(inplanes, 256, 1, stride=1))
        self.conv1 = nn.Conv2d(1280, outplanes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(outplanes)
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout2d(0.1)  # declared, never applied (aspp.py:67,83)

    def forward(self, x):
        branches = [self.aspp1(x), self.aspp2(x), self.aspp3(x), self.aspp4(x)]
        pooled = self.global_avg_pool(x)
        pooled = F.interpolate(pooled, size=branches[-1].shape[2:], mode="bilinear", align_corners=self.align_corner)
        x = torch.cat(branches + [pooled], dim=1)
        return self.relu(self.bn1(self.conv1(x)))


class PPMModule(nn.Module):
    def __init__(self, features, out_features=512, sizes=(1, 2, 3, 6), align_corners=True):
        super().__init__()
        self.align_corners = align_corners
        self.stages = nn.ModuleList(
            [nn.Sequential(nn.AdaptiveAvgPool2d((s, s)), *_cbr(features, out_features, 1)) for s in sizes])
        self.bottleneck = nn.Sequential(*_cbr(features + len(sizes) * out_features, out_features, 3, padding=1, dilation=1))

    def forward(self, feats):
        hw = feats.shape[2:]
        priors = [F.interpolate(s(feats), size=hw, mode="bilinear", align_corners=self.align_corners)
                  for s in self.stages] + [feats]
        return self.bottleneck(torch.cat(priors, 1))


class Decoder(nn.Module):
    def __init__(self, num_classes, align_corner, high_level_inplanes=512, low_level_inplanes=256):
        super().__init__()
        self.align_corner = align_corner
        self.conv1 = nn.Conv2d(low_level_inplanes, 48, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(48)
        self.relu = nn.ReLU(inplace=True)
        self.last_conv = nn.Sequential(*_cbr(high_level_inplanes + 48, 256, 3, stride=1, padding=1),
                                       *_cbr(256, 256, 3, stride=1, padding=1),
                                       nn.Conv2d(256, num_classes, 1, stride=1))

    def forward(self, x, low):
        low = self.relu(self.bn1(self.conv1(low)))
        x = F.interpolate(x, size=low.shape[2:], mode="bilinear", align_corners=self.align_corner)
        return self.last_conv(torch.cat((x, low), dim=1))


def _deepsup_head(cin, num_classes):
    return nn.Sequential(*_cbr(cin, 512, 3, stride=1, padding=1), nn.Dropout2d(0.1), nn.Conv2d(512, num_classes, 1, stride=1))


class CalibrationLoss(nn.Module):
    """CE(main) + ds_weight * CE(deep supervision), ignore label 255 (loss/criterion.py:48-74)."""

    def __init__(self, ignore_index=255, ds_weight=0.4):
        super().__init__()
        self.ds_weight = ds_weight
        self.criterion = nn.CrossEntropyLoss(ignore_index=ignore_index, reduction="mean")

    def forward(self, preds, target):
        loss = self.criterion(preds[0], target)
        if len(preds) >= 2:
            loss = loss + self.criterion(preds[1], target) * self.ds_weight
        return {"loss": loss}


class SegNet(nn.Module):
    """`arch` in {'deeplabv3', 'deeplabv3p', 'psp'}; same ctor/forward protocol as the reference Seg_Model."""

    def __init__(self, arch="deeplabv3", backbone="resnet101", backbone_para=None, model_para=None, num_classes=19,
                 align_corner=True, criterion=None, deepsup=True):
        super().__init__()
        bp = dict(backbone_para or {})
        mp = dict(model_para or {})
        os_ = bp.get("os", 8)
        self.arch = arch
        self.align_corner = align_corner
        no_prune_backbone = bp.get("no_prune", ["backbone.layer4.2.bn3"])
        if arch == "deeplabv3":
            self.ignore_prune_layer = mp.get("no_prune", ["aspp.bn1"]) + no_prune_backbone
            out_index, ds_in = (3, 4), 1024
        elif arch == "deeplabv3p":
            self.ignore_prune_layer = mp.get("no_prune", ["decoder.bn1", "aspp.bn1"]) + no_prune_backbone
            out_index, ds_in = (1, 3, 4), 1024
        elif arch == "psp":
            self.ignore_prune_layer = list(no_prune_backbone)
            out_index, ds_in = (3, 4), 1024
        else:
            raise ValueError(arch)
        self.backbone = DilatedResNet(_DEPTHS[backbone], os_, bp.get("inplanes", 128), bp.get("mg_unit", [1, 2, 4]), out_index)
        if arch == "deeplabv3":
            self.aspp = ASPP(os_, align_corner, inplanes=2048)
            self.last_conv = nn.Sequential(*_cbr(512, 256, 3, stride=1, padding=1), *_cbr(256, 256, 3, stride=1, padding=1),
                                           nn.Conv2d(256, num_classes, 1, stride=1))
        elif arch == "deeplabv3p":
            self.aspp = ASPP(os_, align_corner, inplanes=2048)
            self.decoder = Decoder(num_classes, align_corner, low_level_inplanes=256)
        else:
            self.ppm = PPMModule(2048, out_features=512, align_corners=align_corner)
            self.last_conv = nn.Conv2d(512, num_classes, 1, stride=1)
        self.criterion = criterion
        self.deepsup = deepsup
        if deepsup:
            self.conv_deepsup = _deepsup_head(ds_in, num_classes)

    def forward(self, input, labels=None, deepsup=False):
        size = input.shape[2:]
        feats = self.backbone(input)
        if self.arch == "deeplabv3":
            x = self.last_conv(self.aspp(feats[-1]))
        elif self.arch == "deeplabv3p":
            x = self.decoder(self.aspp(feats[-1]), feats[0])
        else:
            x = self.last_conv(self.ppm(feats[-1]))
        outs = [F.interpolate(x, size=size, mode="bilinear", align_corners=self.align_corner)]
        if self.deepsup and deepsup:
            d = self.conv_deepsup(feats[-2])
            outs.append(F.interpolate(d, size=size, mode="bilinear", align_corners=self.align_corner))
        if self.criterion is not None and labels is not None:
            return self.criterion(outs, labels)
        return outs


BACKBONE_PARA = {"os": 8, "mg_unit": [1, 2, 4], "inplanes": 128, "pretrained": False}

#: BASELINE.json configs c1..c4: arch, backbone, classes, H, W
CONFIGS = {
    "c1": dict(arch="deeplabv3", backbone="resnet50", num_classes=19, height=512, width=1024),
    "c2": dict(arch="deeplabv3", backbone="resnet101", num_classes=19, height=512, width=1024),
    "c3": dict(arch="psp", backbone="resnet101", num_classes=150, height=512, width=512),
    "c4": dict(arch="deeplabv3p", backbone="resnet101", num_classes=171, height=512, width=512),
}


def build_segnet(arch, backbone, num_classes, seed=0, with_loss=True, deepsup=True):
    """Random-init network exactly as the reference builds it for scoring (prune.py:81-88)."""
    torch.manual_seed(seed)
    return SegNet(arch=arch, backbone=backbone, backbone_para=dict(BACKBONE_PARA), model_para={}, num_classes=num_classes,
                  align_corner=True, criterion=CalibrationLoss() if with_loss else None, deepsup=deepsup)
