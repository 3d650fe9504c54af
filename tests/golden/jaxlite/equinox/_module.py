from . import _absent


class Module:
    pass


class Partial:
    pass


module_update_wrapper = _absent
