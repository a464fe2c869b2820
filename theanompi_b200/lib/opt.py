from ..utils.opt import *  # noqa: F401,F403
from ..utils.opt import _BSP_MSGD, _clip_paramlist  # noqa: F401,E402
