from .modules import *  # noqa: F401,F403
from .networks import *  # noqa: F401,F403
