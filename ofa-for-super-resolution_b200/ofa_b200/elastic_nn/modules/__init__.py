from .dynamic_op import *  # noqa: F401,F403
from .dynamic_layers import *  # noqa: F401,F403
