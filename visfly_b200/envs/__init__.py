from .HoverEnv import HoverEnv  # noqa: F401
from .NavigationEnv import NavigationEnv  # noqa: F401
from .RacingEnv import RacingEnv, RacingEnv2  # noqa: F401
