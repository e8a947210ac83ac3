"""``init_weights`` with the reference's behaviour (misc/utils.py:157-164): class-name matching, so that the
same torch seed yields the same initial parameters as the reference."""


def init_weights(m):
    cls = type(m).__name__
    if "Conv" in cls or "Linear" in cls:
        m.weight.data.normal_(0.0, 0.02)
        m.bias.data.fill_(0)
    elif "BatchNorm" in cls:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)
