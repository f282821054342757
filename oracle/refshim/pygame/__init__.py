"""TEST INFRASTRUCTURE: a no-op stand-in for pygame so /root/reference/src imports headless (constants/ui.py:3).
Only used by oracle/gen_golden.py; never imported by the product."""


class _Dummy:
    def __getattr__(self, name):
        return _Dummy()

    def __call__(self, *a, **k):
        return _Dummy()

    def __iter__(self):
        return iter(())

    def __bool__(self):
        return False


def get_init():
    return False


def __getattr__(name):
    if name.startswith("K_") or name.isupper():
        return hash(name) & 0xFFFF
    return _Dummy()
