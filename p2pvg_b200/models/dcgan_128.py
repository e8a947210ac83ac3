"""``models.dcgan_128`` drop-in: ``encoder(dim, nc=1)`` / ``decoder(dim, nc=1)`` (reference models/dcgan_128.py:28,60)."""
from .backbone import DcganDecoder, DcganEncoder


class encoder(DcganEncoder):
    image_width = 128


class decoder(DcganDecoder):
    image_width = 128
