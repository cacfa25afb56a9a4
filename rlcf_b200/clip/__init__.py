from .clip import *  # noqa: F401,F403
from .clip import available_models, load, tokenize  # noqa: F401
