"""Same import surface as the reference's transformer/__init__.py (`from transformer import GPTEncoder, GPTDecoder`)."""
from .encoders import *   # noqa: F401,F403
from .decoders import *   # noqa: F401,F403
from .utils import *      # noqa: F401,F403
from .minGPT import *     # noqa: F401,F403
