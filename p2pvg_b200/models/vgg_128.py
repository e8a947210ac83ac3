"""``models.vgg_128`` drop-in (reference models/vgg_128.py:4-120): VGG-style 128x128 frame encoder / decoder with the
reference's constructor arguments and ``state_dict`` keys.  The nn layers hold parameters and BatchNorm buffers only
(created in the reference's order, so the torch RNG stream is consumed identically); arithmetic runs in the sm_100a
kernels (p2pvg_b200/engine_vgg.py for training, p2pvg_b200/infer_vgg.py for stand-alone calls)."""
import torch.nn as nn

from .vgg_64 import vgg_layer


class encoder(nn.Module):
    backbone = "vgg"
    image_width = 128
    nstage = 5

    def __init__(self, dim, nc=1):
        super().__init__()
        self.dim, self.nc = dim, nc
        self.c1 = nn.Sequential(vgg_layer(nc, 64), vgg_layer(64, 64))
        self.c2 = nn.Sequential(vgg_layer(64, 128), vgg_layer(128, 128))
        self.c3 = nn.Sequential(vgg_layer(128, 256), vgg_layer(256, 256), vgg_layer(256, 256))
        self.c4 = nn.Sequential(vgg_layer(256, 512), vgg_layer(512, 512), vgg_layer(512, 512))
        self.c5 = nn.Sequential(vgg_layer(512, 512), vgg_layer(512, 512), vgg_layer(512, 512))
        self.c6 = nn.Sequential(nn.Conv2d(512, dim, 4, 1, 0), nn.BatchNorm2d(dim), nn.Tanh())
        self.mp = nn.MaxPool2d(kernel_size=2, stride=2, padding=0)

    def forward(self, input):
        from ..infer_vgg import vgg_encoder_forward
        return vgg_encoder_forward(self, input)


class decoder(nn.Module):
    backbone = "vgg"
    image_width = 128
    nstage = 5

    def __init__(self, dim, nc=1):
        super().__init__()
        self.dim, self.nc = dim, nc
        self.upc1 = nn.Sequential(nn.ConvTranspose2d(dim, 512, 4, 1, 0), nn.BatchNorm2d(512), nn.LeakyReLU(0.2, inplace=True))
        self.upc2 = nn.Sequential(vgg_layer(512 * 2, 512), vgg_layer(512, 512), vgg_layer(512, 512))
        self.upc3 = nn.Sequential(vgg_layer(512 * 2, 512), vgg_layer(512, 512), vgg_layer(512, 256))
        self.upc4 = nn.Sequential(vgg_layer(256 * 2, 256), vgg_layer(256, 256), vgg_layer(256, 128))
        self.upc5 = nn.Sequential(vgg_layer(128 * 2, 128), vgg_layer(128, 64))
        self.upc6 = nn.Sequential(vgg_layer(64 * 2, 64), nn.ConvTranspose2d(64, nc, 3, 1, 1), nn.Sigmoid())
        self.up = nn.UpsamplingNearest2d(scale_factor=2)

    def forward(self, input):
        from ..infer_vgg import vgg_decoder_forward
        vec, skip = input
        return vgg_decoder_forward(self, vec, skip)
