"""ResNet-FPN backbone (1/8 + 1/2 outputs), same parameter names as the reference so its checkpoints load
(mp3d_loftr/src/loftr/backbone/resnet_fpn.py:43-119).  Inside the forward() boundary but NOT a hand-kernel
target: it stays on cuDNN; we run it channels_last so the 1/2-res map comes
out NHWC, which is the layout the fine-window gather kernel reads coalesced."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _c1(i, o, stride=1):
    return nn.Conv2d(i, o, kernel_size=1, stride=stride, padding=0, bias=False)


def _c3(i, o, stride=1):
    return nn.Conv2d(i, o, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = _c3(in_planes, planes, stride)
        self.conv2 = _c3(planes, planes)
        self.bn1 = nn.BatchNorm2d(planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None if stride == 1 else nn.Sequential(_c1(in_planes, planes, stride=stride),
                                                                 nn.BatchNorm2d(planes))

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)


class ResNetFPN_8_2(nn.Module):
    def __init__(self, config):
        super().__init__()
        d0 = config['initial_dim']
        b1, b2, b3 = config['block_dims']
        self.config = config
        self.conv1 = nn.Conv2d(1, d0, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(d0)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = nn.Sequential(BasicBlock(d0, b1, 1), BasicBlock(b1, b1, 1))   # 1/2
        self.layer2 = nn.Sequential(BasicBlock(b1, b2, 2), BasicBlock(b2, b2, 1))   # 1/4
        self.layer3 = nn.Sequential(BasicBlock(b2, b3, 2), BasicBlock(b3, b3, 1))   # 1/8
        self.layer3_outconv = _c1(b3, b3)
        self.layer2_outconv = _c1(b2, b3)
        self.layer2_outconv2 = nn.Sequential(_c3(b3, b3), nn.BatchNorm2d(b3), nn.LeakyReLU(), _c3(b3, b2))
        self.layer1_outconv = _c1(b1, b2)
        self.layer1_outconv2 = nn.Sequential(_c3(b2, b2), nn.BatchNorm2d(b2), nn.LeakyReLU(), _c3(b2, b1))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    memory_format = torch.channels_last
    fused_eval = True   # eval + CUDA: BN folded into the convs, cuDNN conv+bias(+residual)+ReLU, fused FPN glue kernels

    def forward(self, x):
        if self.memory_format == torch.channels_last and x.is_cuda and \
                not self.layer1[0].conv1.weight.is_contiguous(memory_format=torch.channels_last):
            # a 1-channel input is layout-agnostic, so the NHWC request has to come from the weights: without this
            # cuDNN runs NCHW and inserts nchw<->nhwc conversion kernels around every TF32 conv (18 ms / step measured)
            self.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=self.memory_format)
        if self.fused_eval and not self.training and x.is_cuda and self.memory_format == torch.channels_last \
                and not torch.is_grad_enabled():
            return self._forward_fused(x)
        x0 = self.relu(self.bn1(self.conv1(x)))
        x1 = self.layer1(x0)
        x2 = self.layer2(x1)
        x3 = self.layer3(x2)
        x3_out = self.layer3_outconv(x3)
        x3_up = F.interpolate(x3_out, scale_factor=2., mode='bilinear', align_corners=True)
        x2_out = self.layer2_outconv2(self.layer2_outconv(x2) + x3_up)
        x2_up = F.interpolate(x2_out, scale_factor=2., mode='bilinear', align_corners=True)
        x1_out = self.layer1_outconv2(self.layer1_outconv(x1) + x2_up)
        return [x3_out, x1_out]

    # ---- eval-time fused path -------------------------------------------------------------------------------
    # Same arithmetic as above with eval-mode BatchNorm folded into the preceding convolution (w' = w*g/sqrt(var+eps),
    # b' = beta - mean*g/sqrt(var+eps)); every conv+BN+ReLU and conv+BN+residual+ReLU is ONE cuDNN fused call, the
    # two `1x1 conv + upsample + add` joins and the two BN+LeakyReLU passes are one far_* kernel each.  Removes ~30
    # elementwise passes over the 1/2- and 1/4-resolution maps (2.5-3.9 GB each for a 32-pair batch).
    def _folded(self):
        key = tuple(int(t._version) for t in list(self.parameters()) + list(self.buffers())) + \
            (self.conv1.weight.data_ptr(),)
        cache = getattr(self, "_fold_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1]

        def fold(conv, bn):
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            w = (conv.weight * scale[:, None, None, None]).contiguous(memory_format=torch.channels_last)
            return w, (bn.bias - bn.running_mean * scale).contiguous()

        def affine(bn):
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            return scale.contiguous(), (bn.bias - bn.running_mean * scale).contiguous()

        with torch.no_grad():
            f = {"stem": fold(self.conv1, self.bn1)}
            for li, layer in enumerate((self.layer1, self.layer2, self.layer3)):
                for bi, blk in enumerate(layer):
                    f[(li, bi, 1)] = fold(blk.conv1, blk.bn1)
                    f[(li, bi, 2)] = fold(blk.conv2, blk.bn2)
                    if blk.downsample is not None:
                        f[(li, bi, "d")] = fold(blk.downsample[0], blk.downsample[1])
                        f[(li, bi, "2d")] = (f[(li, bi, 2)][1] + f[(li, bi, "d")][1]).contiguous()
            f["o2"] = affine(self.layer2_outconv2[1])
            f["o1"] = affine(self.layer1_outconv2[1])
        self._fold_cache = (key, f)
        return f

    def _forward_fused(self, x):
        from .. import ops
        f = self._folded()
        one, pad1, pad0 = (1, 1), (1, 1), (0, 0)

        def block(x, li, bi, blk):
            stride = blk.conv1.stride
            w1, b1 = f[(li, bi, 1)]
            w2, b2 = f[(li, bi, 2)]
            y = torch.cudnn_convolution_relu(x, w1, b1, stride, pad1, one, 1)
            if blk.downsample is not None:
                # relu(conv2(y) + b2 + (conv_d(x) + b_d)): the downsample bias rides on conv2's bias, so the 1x1
                # stride-2 conv needs no separate bias pass over its output
                wd, bd = f[(li, bi, "d")]
                x = F.conv2d(x, wd, None, stride=stride)
                b2 = f[(li, bi, "2d")]
            return torch.cudnn_convolution_add_relu(y, w2, x, 1.0, b2, one, pad1, one, 1)

        w, b = f["stem"]
        if w.shape[0] == 128 and x.shape[1] == 1:
            x0 = ops.stem_conv7x7s2_relu(x, w, b)      # hand kernel: conv + bias + ReLU in one pass, NHWC out
        else:
            x0 = torch.relu_(F.conv2d(x, w, b, stride=2, padding=3))   # C_in = 1: not a fused-engine shape
        feats = []
        cur = x0
        for li, layer in enumerate((self.layer1, self.layer2, self.layer3)):
            for bi, blk in enumerate(layer):
                cur = block(cur, li, bi, blk)
            feats.append(cur)
        x1, x2, x3 = feats
        x3_out = self.layer3_outconv(x3)
        j2 = ops.upsample2x_add(x3_out, self.layer2_outconv(x2))
        t2 = self.layer2_outconv2[0](j2)
        ops.scale_shift_act_(t2, f["o2"][0], f["o2"][1], self.layer2_outconv2[2].negative_slope)
        x2_out = self.layer2_outconv2[3](t2)
        j1 = ops.upsample2x_add(x2_out, self.layer1_outconv(x1))
        t1 = self.layer1_outconv2[0](j1)
        ops.scale_shift_act_(t1, f["o1"][0], f["o1"][1], self.layer1_outconv2[2].negative_slope)
        x1_out = self.layer1_outconv2[3](t1)
        return [x3_out, x1_out]


def fold_batchnorm(module):
    """Eval-time rewrite: fold every (Conv2d -> BatchNorm2d) pair into the convolution (w' = w * g/sqrt(var+eps),
    b' = beta - mean * g/sqrt(var+eps)).  Legal because BN is in eval mode on the whole path;
    removes one elementwise pass per conv.  Returns a deep copy; the original keeps the checkpoint's parameter names."""
    import copy
    m = copy.deepcopy(module).eval()

    def fuse(conv, bn):
        w = conv.weight
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        fused = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, bias=True)
        fused = fused.to(w.device, w.dtype)
        with torch.no_grad():
            fused.weight.copy_(w * scale[:, None, None, None])
            b0 = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
            fused.bias.copy_((b0 - bn.running_mean) * scale + bn.bias)
        return fused

    m.conv1, m.bn1 = fuse(m.conv1, m.bn1), nn.Identity()
    for layer in (m.layer1, m.layer2, m.layer3):
        for blk in layer:
            blk.conv1, blk.bn1 = fuse(blk.conv1, blk.bn1), nn.Identity()
            blk.conv2, blk.bn2 = fuse(blk.conv2, blk.bn2), nn.Identity()
            if blk.downsample is not None:
                blk.downsample = nn.Sequential(fuse(blk.downsample[0], blk.downsample[1]))
    for seq in (m.layer2_outconv2, m.layer1_outconv2):
        seq[0], seq[1] = fuse(seq[0], seq[1]), nn.Identity()
    return m


def build_backbone(config):
    """mp3d_loftr/src/loftr/backbone/__init__.py:4-11 (only the (8, 2) resolution is on the path)."""
    if config['backbone_type'] == 'ResNetFPN' and tuple(config['resolution']) == (8, 2):
        return ResNetFPN_8_2(config['resnetfpn'])
    raise ValueError(f"LOFTR.BACKBONE_TYPE {config['backbone_type']} / resolution {config['resolution']} not supported.")
