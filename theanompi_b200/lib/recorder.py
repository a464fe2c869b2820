from ..utils.recorder import *  # noqa: F401,F403
