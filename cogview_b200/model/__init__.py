"""model/__init__.py:17-19 of the reference."""
from .distributed import *  # noqa: F401,F403
from .gpt2_modeling import gpt2_get_params_for_weight_decay_optimization
from .gpt2_modeling import GPT2Model
