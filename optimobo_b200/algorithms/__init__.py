from .optimisers import MonoSurrogateOptimiser, MultiSurrogateOptimiser  # noqa: F401
from .parego import ParEGO  # noqa: F401
from .cparego import ParEGO_C1, ParEGO_C2  # noqa: F401
from .keep import KEEP  # noqa: F401
from .emo import EMO  # noqa: F401
from .turbo import TuRBO_1, TuRBO_M  # noqa: F401
