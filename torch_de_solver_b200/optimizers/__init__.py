from .optimizer import Optimizer
from .closure import Closure
