"""mmh3 stand-in: imported by graphax `primitives.py:5`, never called."""


def hash(*a, **k):   # noqa: A001
    raise NotImplementedError("jaxlite: mmh3.hash")
