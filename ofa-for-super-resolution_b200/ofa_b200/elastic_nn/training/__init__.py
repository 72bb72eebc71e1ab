from .progressive_shrinking import *  # noqa: F401,F403
