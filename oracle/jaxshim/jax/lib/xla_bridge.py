class _Backend:
    platform = 'cpu'


def get_backend():
    return _Backend()
