"""Drop-in for the reference's src/GraphGAN/discriminator.py (``import discriminator``; graph_gan.py:10)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from graphgan_b200.discriminator import Discriminator  # noqa: E402,F401
