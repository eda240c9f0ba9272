from . import constants, utils  # noqa: F401
