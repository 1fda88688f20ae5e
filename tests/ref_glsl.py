"""The binding lives next to the libraries it binds (oracle/pyrefglsl.py); tests import it under this name."""
from oracle.pyrefglsl import *  # noqa: F401,F403
from oracle.pyrefglsl import available  # noqa: F401
