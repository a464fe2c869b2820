from ..parallel.exchanger import *  # noqa: F401,F403
