"""Reference distiller protocol on the CUDA engine (distillers/__init__.py:13-41: name -> class lookup)."""


def find_distiller_using_name(distiller_name):
    if distiller_name == 'inception':
        from .inception_distiller import InceptionDistiller
        return InceptionDistiller
    if distiller_name == 'spade':
        from .spade_distiller import SPADEDistiller
        return SPADEDistiller
    raise NotImplementedError('distiller [%s] is not a CAT distiller (inception | spade)' % distiller_name)


def get_option_setter(distiller_name):
    return find_distiller_using_name(distiller_name).modify_commandline_options


def create_distiller(opt, verbose=True):
    distiller = find_distiller_using_name(opt.distiller)(opt)
    if verbose:
        print('distiller [%s] was created' % type(distiller).__name__)
    return distiller
