"""No-op pyplot stub (oracle-only): every attribute is a callable returning
another stub, so the reference's plotting helpers run without a display."""


class _Stub:
    def __call__(self, *a, **k):
        return _Stub()

    def __getattr__(self, name):
        return _Stub()

    def __iter__(self):
        return iter((_Stub(), _Stub()))


def __getattr__(name):
    return _Stub()
