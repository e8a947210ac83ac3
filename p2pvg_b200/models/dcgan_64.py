"""``models.dcgan_64`` drop-in: ``encoder(dim, nc=1)`` / ``decoder(dim, nc=1)`` (reference models/dcgan_64.py:28,57)."""
from .backbone import DcganDecoder, DcganEncoder


class encoder(DcganEncoder):
    image_width = 64


class decoder(DcganDecoder):
    image_width = 64
