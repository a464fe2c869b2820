from ..utils.helper_funcs import *  # noqa: F401,F403
