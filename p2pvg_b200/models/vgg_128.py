"""``models.vgg_128`` drop-in: ``encoder(dim, nc=1)`` / ``decoder(dim, nc=1)`` (reference models/vgg_128.py:16,66)."""
from .vgg import VggDecoder, VggEncoder, vgg_layer  # noqa: F401


class encoder(VggEncoder):
    image_width = 128


class decoder(VggDecoder):
    image_width = 128
