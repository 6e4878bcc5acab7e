"""Renderer plugin registry: same surface as `gaustudio/renderers/__init__.py:1-28` of the reference
(`register(name)` decorator, `make(config)` factory with the same ValueErrors)."""
renderers = {}


def register(name):
    def decorator(cls):
        renderers[name] = cls
        return cls
    return decorator


def make(config):
    if isinstance(config, str):
        name, config = config, {}
    else:
        name = config.get('name')
    if not name:
        raise ValueError('Renderer name is required')
    if name not in renderers:
        raise ValueError(f'Unknown renderer: {name}')
    return renderers[name](config)


from . import vanilla_renderer  # noqa: E402,F401  (the only renderer on the hot path)
