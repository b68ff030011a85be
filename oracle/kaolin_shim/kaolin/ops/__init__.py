from . import spc  # noqa: F401
from . import conversions  # noqa: F401
