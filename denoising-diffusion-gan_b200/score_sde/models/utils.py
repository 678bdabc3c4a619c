"""Model registry (reference: score_sde/models/utils.py:33-57)."""
_MODELS = {}


def register_model(cls=None, *, name=None):
    def _register(cls):
        local_name = cls.__name__ if name is None else name
        if local_name in _MODELS:
            raise ValueError(f'Already registered model with name: {local_name}')
        _MODELS[local_name] = cls
        return cls
    return _register if cls is None else _register(cls)


def get_model(name):
    return _MODELS[name]
