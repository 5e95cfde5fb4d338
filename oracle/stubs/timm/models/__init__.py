def create_model(*a, **k):
    from timm import create_model as _cm

    return _cm(*a, **k)
