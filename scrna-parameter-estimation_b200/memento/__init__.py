"""Drop-in alias: ``import memento`` resolves to the B200 implementation when
``scrna-parameter-estimation_b200/`` is on ``sys.path`` ahead of the reference package."""
from memento_b200 import *  # noqa: F401,F403
from memento_b200 import main, getters  # noqa: F401
