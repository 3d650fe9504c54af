from . import core, util, pjit, ad_util  # noqa: F401
