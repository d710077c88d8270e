"""Module-tree mirrors of the reference networks (networks.py, spade_networks.py) and of its teacher-training models
(models/__init__.py:7-51 in the reference: name -> class lookup)."""


def find_model_using_name(model_name):
    if model_name == 'pix2pix':
        from .pix2pix_model import Pix2PixModel
        return Pix2PixModel
    if model_name == 'cycle_gan':
        from .cycle_gan_model import CycleGANModel
        return CycleGANModel
    if model_name == 'spade':
        from .spade_model import SPADEModel
        return SPADEModel
    raise NotImplementedError('model [%s] is not a CAT training model (pix2pix | cycle_gan | spade)' % model_name)


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt, verbose=True):
    model = find_model_using_name(opt.model)(opt)
    if verbose:
        print('model [%s] was created' % type(model).__name__)
    return model
