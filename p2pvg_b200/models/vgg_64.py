"""``models.vgg_64`` drop-in: ``encoder(dim, nc=1)`` / ``decoder(dim, nc=1)`` (reference models/vgg_64.py:16,59)."""
from .vgg import VggDecoder, VggEncoder, vgg_layer  # noqa: F401


class encoder(VggEncoder):
    image_width = 64


class decoder(VggDecoder):
    image_width = 64
