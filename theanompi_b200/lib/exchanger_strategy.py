from ..parallel.exchanger_strategy import *  # noqa: F401,F403
