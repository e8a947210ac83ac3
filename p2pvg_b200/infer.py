"""Stand-alone (no-autograd) forward passes of the drop-in modules on the sm_100a kernels: what
``p2p_generate`` / ``generate.py`` / ``misc/visualize.py`` of the reference call.
Filled in after the training path."""


def _todo(name):
    raise NotImplementedError(f"{name}: stand-alone inference forward is not built yet (training path only)")


def encoder_forward(mod, x):
    _todo("encoder.forward")


def decoder_forward(mod, vec, skip):
    _todo("decoder.forward")


def lstm_forward(mod, inp):
    _todo("lstm.forward")


def gaussian_lstm_forward(mod, inp):
    _todo("gaussian_lstm.forward")


def p2p_generate(model, x, len_output, eval_cp_ix, model_mode="full", skip_frame=False, init_hidden=True):
    _todo("P2PModel.p2p_generate")
