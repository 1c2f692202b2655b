class _R:
    env_specs = {}
registry = _R()
def register(id, entry_point): registry.env_specs[id] = entry_point
