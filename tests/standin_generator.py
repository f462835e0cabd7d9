"""A structural stand-in for the reference generator, for tests and timing on the GPU box (the reference's own
models/generator_resnet_attn.py cannot travel there).  Same attribute layout -- ``initial``, ``downsample``,
``res_blocks``, ``upsample``, ``output`` -- and the same logical feature numbering as
``ResNetGenerator.get_feature_layers`` (generator_resnet_attn.py:190-235): 0 = after ``initial``, then one index per
ReLU of ``downsample``, one per residual block, one per ReLU of ``upsample``.  ngf=64, n_blocks=9 gives the
ResNet-9 shapes of SURVEY.md section 8 (64x256x256, 128x128x128, 256x64x64 x10, 128x128x128, 64x256x256)."""
import torch
import torch.nn as nn


class _Res(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv_block = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(ch, ch, 3), nn.InstanceNorm2d(ch), nn.ReLU(True),
                                        nn.ReflectionPad2d(1), nn.Conv2d(ch, ch, 3), nn.InstanceNorm2d(ch))

    def forward(self, x):
        return x + self.conv_block(x)


class StandInGenerator(nn.Module):
    def __init__(self, ngf=8, n_blocks=3, n_down=2):
        super().__init__()
        self.initial = nn.Sequential(nn.ReflectionPad2d(3), nn.Conv2d(3, ngf, 7), nn.InstanceNorm2d(ngf), nn.ReLU(True))
        down, up, ch = [], [], ngf
        for _ in range(n_down):
            down += [nn.Conv2d(ch, ch * 2, 3, stride=2, padding=1), nn.InstanceNorm2d(ch * 2), nn.ReLU(True)]
            ch *= 2
        self.downsample = nn.Sequential(*down)
        self.res_blocks = nn.ModuleList(_Res(ch) for _ in range(n_blocks))
        for _ in range(n_down):
            up += [nn.ConvTranspose2d(ch, ch // 2, 3, stride=2, padding=1, output_padding=1), nn.InstanceNorm2d(ch // 2),
                   nn.ReLU(True)]
            ch //= 2
        self.upsample = nn.Sequential(*up)
        self.output = nn.Sequential(nn.ReflectionPad2d(3), nn.Conv2d(ngf, 3, 7), nn.Tanh())
        self.feature_passes = 0

    def _stages(self, x):
        """Yields (tensor, is_logical_layer) after every stage, in forward order."""
        x = self.initial(x)
        yield x, True
        for m in self.downsample:
            x = m(x)
            yield x, isinstance(m, nn.ReLU)
        for blk in self.res_blocks:
            x = blk(x)
            yield x, True
        for m in self.upsample:
            x = m(x)
            yield x, isinstance(m, nn.ReLU)

    def forward(self, x):
        for x, _ in self._stages(x):
            pass
        return self.output(x)

    def get_feature_layers(self, x, layer_ids=None):
        self.feature_passes += 1
        wanted = set([0, 4, 8, 12, 16] if layer_ids is None else layer_ids)
        feats, logical = [], 0
        for t, counts in self._stages(x):
            if counts:
                if logical in wanted:
                    feats.append(t)
                logical += 1
        return feats
