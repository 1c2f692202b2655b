# PROBE-ONLY minimal functional gin: bind_parameter + configurable injection by (scope name, param).
import functools, inspect
REQUIRED = object()
_BIND = {}
def bind_parameter(name, value):
    scope, param = name.rsplit('.', 1)
    _BIND[(scope.split('.')[-1], param)] = value
def clear_config(): _BIND.clear()
def _wrap(f, scope):
    if inspect.isclass(f):
        orig = f.__init__
        params = inspect.signature(orig).parameters
        names = list(params)
        @functools.wraps(orig)
        def init(self, *a, **k):
            if type(self).__name__ == scope or True:
                given = set(names[1:1 + len(a)])
                for (s, p), v in _BIND.items():
                    if s == scope and p in params and p not in k and p not in given:
                        k[p] = v
            return orig(self, *a, **k)
        f.__init__ = init
        return f
    params = inspect.signature(f).parameters
    names = list(params)
    @functools.wraps(f)
    def g(*a, **k):
        given = set(names[:len(a)])
        for (s, p), v in _BIND.items():
            if s == scope and p in params and p not in k and p not in given:
                k[p] = v
        return f(*a, **k)
    return g
def configurable(fn_or_name=None, **kw):
    if callable(fn_or_name):
        return _wrap(fn_or_name, fn_or_name.__name__)
    return lambda f: _wrap(f, fn_or_name if isinstance(fn_or_name, str) else f.__name__)
def add_config_file_search_path(*a): pass
def parse_config_files_and_bindings(*a, **k): pass
def parse_config_file(*a, **k): pass
