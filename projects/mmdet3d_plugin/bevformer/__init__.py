from .dense_heads import *   # noqa: F401,F403
from .detectors import *     # noqa: F401,F403
from .modules import *       # noqa: F401,F403
