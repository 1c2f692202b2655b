class errors:
    class NotFoundError(Exception): pass
    class PermissionDeniedError(Exception): pass
class io:
    class gfile:
        @staticmethod
        def GFile(p, m='r'):
            try: return open(p, m)
            except FileNotFoundError as e: raise errors.NotFoundError(str(e))
