"""equinox stand-in: `graphax/__init__.py:3` imports the equinox adapter, which is outside the jacve hot path
(SURVEY.md section 2.1).  Names exist so the import succeeds; none is functional."""


def _absent(*a, **k):
    raise NotImplementedError("jaxlite: equinox is not available")


is_array = filter_make_jaxpr = _absent
